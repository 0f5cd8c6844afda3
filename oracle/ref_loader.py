"""Imports the UNMODIFIED reference from /root/reference (build container only) -- TEST INFRASTRUCTURE.

/root/reference does not exist on the GPU box: nothing under `-m gpu`, smoke() or bench.py may call this. It is used by
tests/golden/make_golden.py to generate the committed fixtures and by CPU-side tests (skipped when the tree is absent)
that pin oracle/ and the drop-in module against the real thing.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GH_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Models", "Models_RESNET50_TRUNCATE_GRAM_with_Attention.py"))


def _stub_gui_modules():
    """matplotlib / tkinter are absent in the image; the reference functions file imports them at module scope
    (functions/functions_RESNET50_Truncate_Gram_Attention.py:10-17). Empty stand-ins let the non-GUI functions load."""
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    try:
        import matplotlib  # noqa: F401
    except Exception:
        mp = mod("matplotlib")
        mp.pyplot = mod("matplotlib.pyplot")
        mp.backends = mod("matplotlib.backends")
        mod("matplotlib.backends.backend_tkagg", FigureCanvasTkAgg=object)
        mod("matplotlib.widgets", PolygonSelector=object)
        mod("matplotlib.path", Path=object)
    try:
        import tkinter  # noqa: F401
    except Exception:
        tk = mod("tkinter")
        tk.ttk = mod("tkinter.ttk")
    try:
        from PIL import ImageTk  # noqa: F401
    except Exception:
        import PIL
        PIL.ImageTk = mod("PIL.ImageTk")


def _load(relpath: str, alias: str):
    """Load a reference file under a private module name so it never shadows this repo's drop-in `Models`/`functions`."""
    import importlib.util
    import torch

    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(alias, path)
    m = importlib.util.module_from_spec(spec)
    anomaly = torch.is_anomaly_enabled()
    spec.loader.exec_module(m)
    # the reference flips the global anomaly mode on at import (Models/...:9, functions/...:23); undo that side effect
    torch.autograd.set_detect_anomaly(anomaly)
    return m


def load_reference_models():
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    return _load("Models/Models_RESNET50_TRUNCATE_GRAM_with_Attention.py", "_ref_models_gram_attention")


def load_reference_functions():
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    _stub_gui_modules()
    return _load("functions/functions_RESNET50_Truncate_Gram_Attention.py", "_ref_functions_gram_attention")


def load_reference_patchgan():
    """Models/Models_Multi_PatchGAN.py (imports only torch); the *_test classes carry the Gram head."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    return _load("Models/Models_Multi_PatchGAN.py", "_ref_models_multi_patchgan")
