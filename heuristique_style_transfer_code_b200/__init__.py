"""B200-native Gram + attention style-feature head (drop-in for the reference's TruncatedResNet50 classes).

    from heuristique_style_transfer_code_b200 import TruncatedResNet50, TruncatedResNet50_for_test

The same classes are importable under the reference's module paths (Models/..., functions/...) from the repo root, so
the reference's train_best_/test_ scripts run unchanged (tools/run_ref_script.py).
"""
from .modules import TruncatedResNet50, TruncatedResNet50_for_test, GramAttentionHead  # noqa: F401

__all__ = ["TruncatedResNet50", "TruncatedResNet50_for_test", "GramAttentionHead"]
