"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the build container, where /root/reference
exists):   python tests/golden/make_golden.py

  head_small_g8.npz   tiny conv encoder (children 0-3 stem, 4-6 three conv stages: C = 64/128/256 -> k = 8/16/32 at
                      g = 8), batch 2, 32x32 input. Everything is stored: stage activations, the six head parameters,
                      embeddings, logits, CE loss, and the reference's autograd gradients w.r.t. stage activations and
                      head parameters (train class and _for_test class share weights).
  resnet_g32.npz      torchvision resnet50(weights=None) truncated at 7, g = 32, batch 2, 64x64 input, train-mode BN:
                      stage activations (256x256, 512x64, 1024x16), pooled descriptors, embeddings, logits, CE-loss
                      gradients w.r.t. the stage activations. The 4.2 M attention weights are not stored; they are
                      re-created from the seed and guarded by a checksum.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_loader import load_reference_models  # noqa: E402


def tiny_encoder():
    return nn.Sequential(
        nn.Conv2d(3, 16, 3, padding=1), nn.BatchNorm2d(16), nn.ReLU(), nn.MaxPool2d(2),
        nn.Sequential(nn.Conv2d(16, 64, 3, padding=1), nn.ReLU()),
        nn.Sequential(nn.Conv2d(64, 128, 3, stride=2, padding=1), nn.ReLU()),
        nn.Sequential(nn.Conv2d(128, 256, 3, stride=2, padding=1), nn.ReLU()),
        nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(256, 10))


def head_params(model):
    return dict(in_proj_weight=model.attention.in_proj_weight, in_proj_bias=model.attention.in_proj_bias,
                out_proj_weight=model.attention.out_proj.weight, out_proj_bias=model.attention.out_proj.bias,
                classifier_weight=model.classifier.weight, classifier_bias=model.classifier.bias)


def run_reference(model, x, labels):
    """Unmodified reference forward; stage activations captured with hooks, gradients from its own autograd.
    A pre-hook detaches the input of every block after the first stage, so that d_stage{i} holds only what the HEAD
    sends back to stage i (without it, autograd adds the gradient arriving through the later encoder blocks, which is
    backbone work and not part of the path under test). The reference's code itself runs unmodified."""
    acts = []
    cuts = [blk.register_forward_pre_hook(lambda module, inputs: (inputs[0].detach().requires_grad_(True),))
            for blk in list(model.truncated_encoder.children())[5:]]

    def capture(module, inputs, output):
        output.retain_grad()
        acts.append(output)

    hooks = [blk.register_forward_hook(capture) for blk in list(model.truncated_encoder.children())[4:]]
    out = model(x)
    for h in hooks + cuts:
        h.remove()
    emb, logits = out if isinstance(out, tuple) else (None, out)
    loss = nn.functional.cross_entropy(logits, labels)
    model.zero_grad()
    loss.backward()
    return acts, emb, logits, loss


def params_checksum(params):
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(params[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    ref = load_reference_models()

    # ---- head_small_g8 -------------------------------------------------------------------------------------------
    torch.manual_seed(1234)
    model = ref.TruncatedResNet50_for_test(tiny_encoder(), 7, 5, 8)
    with torch.no_grad():   # non-zero biases so every term of the backward is exercised
        model.attention.in_proj_bias.normal_(0, 0.2)
        model.attention.out_proj.bias.normal_(0, 0.2)
    model.train()
    x = torch.randn(2, 3, 32, 32)
    labels = torch.tensor([3, 0])
    acts, emb, logits, loss = run_reference(model, x, labels)
    trainer = ref.TruncatedResNet50(tiny_encoder(), 7, 5, 8)
    trainer.load_state_dict(model.state_dict())
    trainer.train()
    logits_train = trainer(x)
    out = {f"stage{i}": a.detach().numpy() for i, a in enumerate(acts)}
    out.update({f"d_stage{i}": a.grad.numpy() for i, a in enumerate(acts)})
    for k, p in head_params(model).items():
        out[k] = p.detach().numpy()
        out["grad_" + k] = p.grad.numpy()
    out.update(embeddings=emb.detach().numpy(), logits=logits.detach().numpy(), logits_train_class=logits_train.detach().numpy(),
               loss=np.float64(loss.item()), labels=labels.numpy(), g=np.int64(8), x=x.numpy())
    np.savez_compressed(os.path.join(HERE, "head_small_g8.npz"), **out)
    print("head_small_g8:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})

    # ---- resnet_g32 ----------------------------------------------------------------------------------------------
    from torchvision import models
    torch.manual_seed(0)
    model = ref.TruncatedResNet50_for_test(models.resnet50(weights=None), 7, 4, 32)
    model.train()
    torch.manual_seed(1)
    x = torch.randn(2, 3, 64, 64)
    labels = torch.tensor([1, 2])
    acts, emb, logits, loss = run_reference(model, x, labels)
    grams = [nn.functional.adaptive_avg_pool2d(model.gram_matrix(a.detach()), (32, 32)) for a in acts]
    desc = torch.stack(grams, dim=1).flatten(2)
    out = {f"stage{i}": a.detach().numpy() for i, a in enumerate(acts)}
    out.update({f"d_stage{i}": a.grad.numpy() for i, a in enumerate(acts)})
    out.update(descriptors=desc.numpy(), embeddings=emb.detach().numpy(), logits=logits.detach().numpy(),
               loss=np.float64(loss.item()), labels=labels.numpy(), g=np.int64(32),
               params_sha256=np.array(params_checksum(head_params(model))),
               torch_version=np.array(torch.__version__))
    np.savez_compressed(os.path.join(HERE, "resnet_g32.npz"), **out)
    print("resnet_g32:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
