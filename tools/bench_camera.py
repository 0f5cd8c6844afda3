"""BASELINE.json configs[3]: camera-mode streaming inference on synthetic 1080p frames resized to 448x448, through the
reference-shaped run_camera() loop (functions.py; upstream functions/functions_RESNET50_Truncate_Gram_Attention.py:477-536)
with a synthetic capture object instead of cv2.VideoCapture(0). Reports per-frame latency (p50/p99) and frames/s, as
measured by run_camera itself (wall clock around preprocess + forward + D2H, like --measure_time upstream).
    python tools/bench_camera.py [--frames 200]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class SyntheticCapture:
    def __init__(self, frames, h=1080, w=1920, pool=8):
        rng = np.random.default_rng(0)
        self.pool = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(pool)]
        self.left = frames
        self.i = 0

    def isOpened(self):
        return True

    def read(self):
        if self.left <= 0:
            return False, None
        self.left -= 1
        self.i += 1
        return True, self.pool[self.i % len(self.pool)].copy()

    def release(self):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--size", type=int, default=448)
    args = ap.parse_args()
    import torch
    from torchvision import models, transforms
    from heuristique_style_transfer_code_b200 import TruncatedResNet50_for_test
    from heuristique_style_transfer_code_b200.functions import run_camera
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = TruncatedResNet50_for_test(models.resnet50(weights=None), 7, 4, 32, device=dev)
    tf = transforms.Compose([transforms.Resize((args.size, args.size)), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    out = tempfile.mkdtemp()
    res = {}
    for pipe in ("host", "gpu"):
        run_camera(model, tf, ["fog", "rain", "snow", "sun"], False, out, 0.5, False, capture=SyntheticCapture(20),
                   display=False, pipeline=pipe)
        times = run_camera(model, tf, ["fog", "rain", "snow", "sun"], False, out, 0.5, True,
                           capture=SyntheticCapture(args.frames), display=False, pipeline=pipe)
        tt = np.array(times) * 1e3
        res[pipe] = {"latency_ms_p50": round(float(np.percentile(tt, 50)), 3), "latency_ms_p99": round(float(np.percentile(tt, 99)), 3),
                     "frames_per_s": round(1e3 / float(tt.mean()), 1)}
    t = tt
    # forward only (frame already a normalised tensor on the device), CUDA events
    x = torch.randn(1, 3, args.size, args.size, device=dev)
    with torch.no_grad():
        for _ in range(10):
            model(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            model(x)
        b.record()
        torch.cuda.synchronize()
    print(json.dumps({"workload": f"configs[3]: camera mode, synthetic 1080p frames -> {args.size}x{args.size}, batch 1",
                      "frames": len(t), "latency_ms_p50": round(float(np.percentile(t, 50)), 3),
                      "latency_ms_p99": round(float(np.percentile(t, 99)), 3), "frames_per_s": round(1e3 / float(t.mean()), 1),
                      "forward_only_eager_ms": round(a.elapsed_time(b) / 50, 3),
                      "host_pipeline": res["host"], "gpu_pipeline": res["gpu"],
                      "note": "top-level latency/fps = run_camera's default (GPU pipeline: frame memcpy to pinned memory, "
                              "H2D, preprocessing kernel + forward + softmax as one CUDA graph, D2H); host_pipeline = PIL "
                              "resize + ToTensor + normalise on the host, H2D, eager forward, D2H (as upstream)"}))


if __name__ == "__main__":
    main()
