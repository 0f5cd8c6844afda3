"""Drop-in module path of the reference (`from Models.Models_RESNET50_TRUNCATE_GRAM_with_Attention import ...`,
train_best_RESNET50_Truncate_gram_attention.py:11, test_RESNET50_Truncate_gram_attention.py:10). The classes are the
B200 implementations in heuristique_style_transfer_code_b200/modules.py."""
from heuristique_style_transfer_code_b200.modules import TruncatedResNet50, TruncatedResNet50_for_test  # noqa: F401

__all__ = ["TruncatedResNet50", "TruncatedResNet50_for_test"]
