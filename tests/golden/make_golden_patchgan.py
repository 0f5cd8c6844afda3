"""Generates tests/golden/patchgan_head.npz from the UNMODIFIED reference class
Models/Models_Multi_PatchGAN.py::VariablePatchesNLayerDiscriminator_test (run in the build container, where
/root/reference exists):   python tests/golden/make_golden_patchgan.py

One small discriminator (ndf 16, gram_matrix_dim 16, batch norm in eval mode, patch_size 30 -> five collected layers),
batch 2, 24x40 input, so the maps are 12x20, 6x10, 3x5, 2x4 and 1x3 (adaptive bins that overlap and repeat). Stored:
the raw 1x1 projections x_proj of every collected layer (captured with forward hooks, i.e. BEFORE the layer norm of
:198), the twelve head tensors of the state_dict, and the reference's gram_norms, embeddings and output.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.patchgan_fp64 import HEAD_KEYS  # noqa: E402
from oracle.ref_loader import load_reference_patchgan  # noqa: E402


def main():
    ref = load_reference_patchgan()
    torch.manual_seed(1234)
    model = ref.VariablePatchesNLayerDiscriminator_test(ndf=16, norm='batch', patch_size=30, num_classes=5,
                                                        gram_matrix_dim=16).eval()
    with torch.no_grad():                       # non-trivial BN statistics and biases
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
        for name in ("attention_per_layer", "attention_per_patch"):
            getattr(model, name).in_proj_bias.normal_(0, 0.1)
            getattr(model, name).out_proj.bias.normal_(0, 0.1)
    x = torch.randn(2, 3, 24, 40)
    raw = []
    hooks = [p.register_forward_hook(lambda mod, inp, out: raw.append(out.detach().clone()))
             for p in model.projection_layers]
    with torch.no_grad():
        emb, out = model(x)
    for h in hooks:
        h.remove()
    sd = model.state_dict()
    blob = {f"x_proj_{i}": t.numpy() for i, t in enumerate(raw)}
    blob.update({"param/" + k: sd[k].numpy() for k in HEAD_KEYS})
    blob["gram_norms"] = torch.stack(model.get_gram_norms(), 0).numpy()
    blob["embeddings"] = emb.numpy()
    blob["output"] = out.numpy()
    blob["input"] = x.numpy()
    path = os.path.join(HERE, "patchgan_head.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes;", len(raw), "layers", [tuple(t.shape) for t in raw])


if __name__ == "__main__":
    main()
