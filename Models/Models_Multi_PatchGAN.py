"""Drop-in module path of the reference (`from Models.Models_Multi_PatchGAN import MultiScaleDiscriminator_test`,
test_Multi_PatchGAN.py:11; `... import MultiScaleDiscriminator`, train_best_Multi_PatchGAN.py:11). The *_test classes are
the B200 implementations in heuristique_style_transfer_code_b200/patchgan.py (their Gram head runs on the
C-ABI kernels; the convolution stacks stay on cuDNN). The head-less training classes are the reference's own, loaded on
first use (heuristique_style_transfer_code_b200/_reference.py)."""
from heuristique_style_transfer_code_b200 import patchgan as _patchgan
from heuristique_style_transfer_code_b200.patchgan import (  # noqa: F401
    PATCH_TYPES, MultiScaleDiscriminator_test, VariablePatchesNLayerDiscriminator_test)

__all__ = ["PATCH_TYPES", "MultiScaleDiscriminator", "MultiScaleDiscriminator_test",
           "VariablePatchesNLayerDiscriminator", "VariablePatchesNLayerDiscriminator_test"]


def __getattr__(name):
    if name in ("MultiScaleDiscriminator", "VariablePatchesNLayerDiscriminator"):
        return getattr(_patchgan, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
