"""Checkpoint converters for `load_model`'s bare-encoder format.

The reference's `load_model(model, path, device)` (functions/functions_RESNET50_Truncate_Gram_Attention.py:30-59) tries
every checkpoint key `k` as `truncated_encoder.k` and silently drops what does not exist in the model. The encoder is
`nn.Sequential(*list(resnet50.children())[:n])`, whose keys are NUMERIC (`0.weight`, `1.running_mean`,
`4.0.conv1.weight`, ...), so the two kinds of file a user is most likely to pass -- a plain torchvision ResNet50
state_dict (`conv1.weight`, `layer1.0.conv1.weight`, ...: what the README's pre-trained encoder is) and the three-section
checkpoint written by `save_model_weights` (README.md:101 passes exactly that file) -- match NOTHING and the call
"succeeds" while loading nothing. `load_model` keeps that behaviour (it is the reference's); these converters produce the
file it does load:

    to_bare_encoder(state)            any of the formats below -> {"0.weight": ..., "4.0.conv1.weight": ...}
    convert_file(src, dst)            the same on files; returns the number of tensors written
    coverage(model, state)            how many encoder tensors of `model` a state_dict would fill through load_model

Accepted inputs: torchvision ResNet names; a three-section checkpoint (its `truncated_encoder` section is already
bare); a flat model state_dict with `truncated_encoder.` prefixes; `module.`-prefixed (DataParallel / DDP) variants of
all of them; an already bare dict (returned as is). `fc.*` is dropped, as load_model does.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Mapping

import torch

# children of torchvision.models.resnet50 in order: the index a child gets inside nn.Sequential(*children[:n])
RESNET_CHILDREN = ("conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4", "avgpool", "fc")
_CHILD_INDEX = {name: i for i, name in enumerate(RESNET_CHILDREN)}


def _strip_module(key: str) -> str:
    while key.startswith("module."):
        key = key[len("module."):]
    return key


def to_bare_encoder(state: Mapping[str, object]) -> "OrderedDict[str, torch.Tensor]":
    """-> state_dict of the bare nn.Sequential encoder (numeric child indices), whatever the input format."""
    if not isinstance(state, Mapping):
        raise TypeError(f"expected a state_dict (mapping), got {type(state).__name__}")
    if "truncated_encoder" in state and isinstance(state["truncated_encoder"], Mapping):
        state = state["truncated_encoder"]                       # three-section checkpoint: the section is already bare
    elif "state_dict" in state and isinstance(state["state_dict"], Mapping):
        state = state["state_dict"]                              # {'state_dict': ..., 'epoch': ...} training checkpoints
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, value in state.items():
        if not torch.is_tensor(value):
            continue
        key = _strip_module(key)
        if key.startswith("truncated_encoder."):
            key = key[len("truncated_encoder."):]
        head, _, rest = key.partition(".")
        if head in ("classifier", "attention", "fc"):
            continue
        if head.isdigit():
            out[key] = value
        elif head in _CHILD_INDEX and rest:
            out[f"{_CHILD_INDEX[head]}.{rest}"] = value
    return out


def convert_file(src: str, dst: str) -> int:
    bare = to_bare_encoder(torch.load(src, map_location="cpu"))
    torch.save(bare, dst)
    return len(bare)


def coverage(model: torch.nn.Module, state: Mapping[str, object]) -> Dict[str, int]:
    """What load_model would do with `state`: {'matched': tensors that land in the model's encoder, 'encoder': tensors the
    encoder has, 'dropped': checkpoint tensors load_model would silently ignore}."""
    want = {k for k in model.state_dict() if k.startswith("truncated_encoder.")}
    keys = [k for k, v in state.items() if torch.is_tensor(v) and not k.startswith("fc.")] if isinstance(state, Mapping) else []
    matched = sum(1 for k in keys if f"truncated_encoder.{k}" in want)
    return {"matched": matched, "encoder": len(want), "dropped": len(keys) - matched}


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="Convert a ResNet50 / three-section checkpoint to load_model's bare-encoder format")
    ap.add_argument("src")
    ap.add_argument("dst")
    args = ap.parse_args(argv)
    n = convert_file(args.src, args.dst)
    print(f"wrote {n} encoder tensors to {args.dst}")
    return 0 if n else 1


if __name__ == "__main__":
    raise SystemExit(main())
