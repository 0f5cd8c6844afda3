"""Camera-mode streaming on the GPU.

The reference handles a camera frame on the host: cv2.cvtColor(BGR2RGB) -> PIL -> torchvision transform (Resize
[+ CenterCrop] + ToTensor + Normalize) -> .to(device) -> model -> softmax -> .cpu()
(functions/functions_RESNET50_Truncate_Gram_Attention.py:494-507). At 1080p the host-side resize alone costs ~9 ms per
frame, four times the forward pass. `CameraPipeline` keeps the same arithmetic but moves it:

  frame (uint8 HWC, BGR) --memcpy--> pinned buffer --H2D (6 MB)--> gh_preprocess_frame (one kernel, bit-identical to
  Pillow's antialiased bilinear resize + ToTensor + Normalize) --> encoder + Gram/attention head --> softmax
  --D2H (num_classes floats)--> pinned buffer

Everything between the two host buffers is captured once in a CUDA graph and replayed per frame, so a batch-1 forward
(~150 kernel launches) costs one graph launch. The coefficient tables of the resize depend only on the geometry and are
computed once here (the same rule as Pillow's precompute_coeffs: triangle filter, support scaled by the down-scale
factor, 22-bit fixed point); a CenterCrop is the sub-range of output coordinates the tables are built for.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import GramHeadError, check

_PRECISION_BITS = 32 - 8 - 2


def resample_tables(in_size: int, out_size: int, first: int = 0, count: Optional[int] = None):
    """Integer tables of the antialiased bilinear resample along one axis, for output coordinates [first, first+count):
    (min[count], size[count], coeff[count, kmax]) as int32 arrays."""
    count = out_size - first if count is None else count
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale                       # bilinear: support 1.0 * filterscale
    kmax = int(math.ceil(support)) * 2 + 1
    lo_a = np.zeros(count, dtype=np.int32)
    n_a = np.zeros(count, dtype=np.int32)
    kk = np.zeros((count, kmax), dtype=np.int32)
    inv = 1.0 / filterscale
    for t in range(count):
        center = (first + t + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = [max(0.0, 1.0 - abs((x + lo - center + 0.5) * inv)) for x in range(n)]
        tot = sum(w)
        if tot != 0.0:
            w = [v / tot for v in w]
        lo_a[t], n_a[t] = lo, n
        for x, v in enumerate(w):
            kk[t, x] = int(v * (1 << _PRECISION_BITS) + 0.5)
    return lo_a, n_a, kk


def _resize_output_size(h: int, w: int, size) -> Tuple[int, int]:
    if isinstance(size, (tuple, list)):
        if len(size) == 2:
            return int(size[0]), int(size[1])
        size = size[0]
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = int(size), int(size * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def parse_transform(transform):
    """Recognises torchvision Compose([Resize(bilinear), [CenterCrop], ToTensor, Normalize]) and returns
    dict(resize=, crop=, mean=, std=) -- or None for anything else (the caller then keeps the host path)."""
    try:
        from torchvision import transforms as T
        from torchvision.transforms import InterpolationMode
    except Exception:
        return None
    steps = list(getattr(transform, "transforms", []))
    if len(steps) not in (3, 4) or not isinstance(steps[0], T.Resize):
        return None
    rz = steps[0]
    if rz.interpolation != InterpolationMode.BILINEAR or getattr(rz, "max_size", None) is not None:
        return None
    crop = None
    rest = steps[1:]
    if len(rest) == 3:
        if not isinstance(rest[0], T.CenterCrop):
            return None
        crop = rest[0].size
        rest = rest[1:]
    if not (isinstance(rest[0], T.ToTensor) and isinstance(rest[1], T.Normalize)):
        return None
    mean, std = [float(v) for v in rest[1].mean], [float(v) for v in rest[1].std]
    if len(mean) != 3 or len(std) != 3:
        return None
    return dict(resize=rz.size, crop=crop, mean=mean, std=std)


class CameraPipeline:
    """frame (H, W, 3) uint8 numpy -> probabilities (num_classes,) numpy; model must be a TruncatedResNet50_for_test
    (forward -> (embeddings, logits)) on a CUDA device."""

    def __init__(self, model, frame_shape: Sequence[int], resize, crop=None, mean=(0.485, 0.456, 0.406),
                 std=(0.229, 0.224, 0.225), bgr: bool = True, use_graph: bool = True):
        self.model = model
        self.device = torch.device(model.device)
        if self.device.type != "cuda":
            raise GramHeadError("gramhead: CameraPipeline needs the model on a CUDA device (there is no CPU path)")
        h, w = int(frame_shape[0]), int(frame_shape[1])
        if len(frame_shape) != 3 or frame_shape[2] != 3:
            raise GramHeadError(f"gramhead: frames must be (H, W, 3) uint8, got shape {tuple(frame_shape)}")
        rh, rw = _resize_output_size(h, w, resize)
        top, left, oh, ow = 0, 0, rh, rw
        if crop is not None:
            oh, ow = (int(crop), int(crop)) if not isinstance(crop, (tuple, list)) else (int(crop[0]), int(crop[-1]))
            if oh > rh or ow > rw:
                raise GramHeadError("gramhead: CenterCrop larger than the resized frame is not supported on the GPU path")
            top, left = int(round((rh - oh) / 2.0)), int(round((rw - ow) / 2.0))
        self.frame_shape, self.out_hw, self.bgr = (h, w, 3), (oh, ow), bool(bgr)
        dev = self.device
        hx = resample_tables(w, rw, left, ow)
        vy = resample_tables(h, rh, top, oh)
        self._tables = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (*hx, *vy)]
        self._hkmax, self._vkmax = int(hx[2].shape[1]), int(vy[2].shape[1])
        self._mean = (torch.tensor(mean, dtype=torch.float32).numpy().copy())
        self._std = (torch.tensor(std, dtype=torch.float32).numpy().copy())
        self._frame_pin = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
        self._frame_np = self._frame_pin.numpy()
        self._frame_dev = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
        self._input = torch.empty((1, 3, oh, ow), dtype=torch.float32, device=dev)
        self._probs_dev = None
        self._probs_pin = None
        self._graph = None
        self._done = torch.cuda.Event()
        model.eval()
        with torch.cuda.device(dev), torch.no_grad():
            self._body()                                  # allocates the output buffers, loads every kernel
            torch.cuda.synchronize(dev)
            if use_graph:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(2):
                        self._body()
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._body()
                self._graph = graph
                # the graph holds raw addresses of the model's folded inference plan: keep that plan (and its tensors)
                # alive for as long as this pipeline exists, whatever the model does with its own reference later
                self._captured_plan = getattr(model, "_plan", None)

    # the device-side work for one frame; captured into the graph
    def preprocess_(self):
        """Runs the preprocessing kernel on the uploaded frame; result in self._input (1, 3, oh, ow)."""
        h, w, _ = self.frame_shape
        oh, ow = self.out_hw
        t = self._tables
        rc = _lib.lib().gh_preprocess_frame(
            self._frame_dev.data_ptr(), w * 3, h, w, 1 if self.bgr else 0, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(),
            self._hkmax, t[3].data_ptr(), t[4].data_ptr(), t[5].data_ptr(), self._vkmax, self._mean.ctypes.data,
            self._std.ctypes.data, self._input.data_ptr(), oh, ow, torch.cuda.current_stream(self.device).cuda_stream)
        check(rc, "gh_preprocess_frame")
        return self._input

    def _body(self):
        self._frame_dev.copy_(self._frame_pin, non_blocking=True)
        self.preprocess_()
        out = self.model(self._input)
        logits = out[1] if isinstance(out, (tuple, list)) else out
        probs = torch.softmax(logits, dim=1)
        if self._probs_dev is None:
            self._probs_dev = torch.empty_like(probs)
            self._probs_pin = torch.empty(probs.shape, dtype=probs.dtype).pin_memory()
        self._probs_dev.copy_(probs)
        self._probs_pin.copy_(self._probs_dev, non_blocking=True)

    def __call__(self, frame: np.ndarray) -> np.ndarray:
        if frame.shape != self.frame_shape or frame.dtype != np.uint8:
            raise GramHeadError(f"gramhead: CameraPipeline was built for uint8 frames of shape {self.frame_shape}, "
                                f"got {frame.dtype} {tuple(frame.shape)}")
        np.copyto(self._frame_np, frame)
        with torch.cuda.device(self.device), torch.no_grad():
            if self._graph is not None:
                self._graph.replay()
            else:
                self._body()
            self._done.record(torch.cuda.current_stream(self.device))
        self._done.synchronize()
        return self._probs_pin.numpy()[0].copy()

    @classmethod
    def from_transform(cls, model, transform, frame_shape, bgr: bool = True, use_graph: bool = True):
        """Pipeline equivalent to `transform` applied to a PIL image of the frame, or None when the transform is not the
        Resize [+ CenterCrop] + ToTensor + Normalize chain the GPU kernel reproduces."""
        spec = parse_transform(transform)
        if spec is None:
            return None
        return cls(model, frame_shape, spec["resize"], spec["crop"], spec["mean"], spec["std"], bgr=bgr, use_graph=use_graph)
