"""Drop-in module path of the reference (`from Models.Models_Multi_PatchGAN import MultiScaleDiscriminator_test`,
test_Multi_PatchGAN.py:11; `... import MultiScaleDiscriminator`, train_best_Multi_PatchGAN.py:11). The classes are the
B200 implementations in heuristique_style_transfer_code_b200/patchgan.py (the Gram head of the *_test
classes runs on the C-ABI kernels; the convolution stacks stay on cuDNN)."""
from heuristique_style_transfer_code_b200.patchgan import (  # noqa: F401
    PATCH_TYPES, MultiScaleDiscriminator, MultiScaleDiscriminator_test, VariablePatchesNLayerDiscriminator,
    VariablePatchesNLayerDiscriminator_test)

__all__ = ["PATCH_TYPES", "MultiScaleDiscriminator", "MultiScaleDiscriminator_test",
           "VariablePatchesNLayerDiscriminator", "VariablePatchesNLayerDiscriminator_test"]
