"""Locates the UNMODIFIED reference sources and loads single files from them by path, for the pieces of the reference's
API that are outside the accelerated path (t-SNE / Tk viewers, the head-less Multi-PatchGAN training classes): those are
delegated to the reference's own code instead of being restated here. Nothing on the Gram + attention path uses this.

Search order: $GRAMHEAD_REFERENCE_ROOT, <repo>/baseline/_ref (staged by tools/stage_reference.py, git-ignored),
/root/reference."""
from __future__ import annotations

import importlib.util
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LOADED = {}


def reference_root(required_file: str):
    for root in (os.environ.get("GRAMHEAD_REFERENCE_ROOT"), os.path.join(_REPO, "baseline", "_ref"), "/root/reference"):
        if root and os.path.isfile(os.path.join(root, required_file)):
            return root
    return None


def load_reference_file(rel_path: str):
    """Imports <reference root>/<rel_path> under a private module name (the repo's own `Models` / `functions` packages
    shadow the reference's). Raises NotImplementedError when the reference tree is not available. The reference files
    switch torch's anomaly detection on at import; the previous setting is restored."""
    if rel_path in _LOADED:
        return _LOADED[rel_path]
    root = reference_root(rel_path)
    if root is None:
        raise NotImplementedError(
            f"{rel_path} belongs to the reference project and is outside the Gram + attention path this package "
            "implements; point GRAMHEAD_REFERENCE_ROOT at a checkout of Hamedkiri/heuristique_style_transfer_code "
            "(or run tools/stage_reference.py) to use it.")
    import torch
    anomaly = torch.is_anomaly_enabled()
    name = "_gramhead_reference_" + rel_path.replace(os.sep, "_").replace(".py", "")
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel_path))
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    try:
        spec.loader.exec_module(module)
    except BaseException:
        sys.modules.pop(name, None)
        raise
    finally:
        torch.autograd.set_detect_anomaly(anomaly)
    _LOADED[rel_path] = module
    return module
