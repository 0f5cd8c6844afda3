"""Shortest visit: the GPU tests of the uint8 upload against the rebuilt gh_normalize_u8, then its timing alone."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
t0 = time.time()
import pytest  # noqa: E402

with open(os.path.join(OUT, "r4c_pytest_u8.log"), "w") as log:
    stdout, sys.stdout = sys.stdout, log
    try:
        rc = pytest.main(["tests/test_gpu_module.py", "-q", "-m", "gpu", "-k", "normalize_u8 or uint8", "-p", "no:cacheprovider"])
    finally:
        sys.stdout = stdout
print(f"pytest rc={int(rc)} after {time.time() - t0:.1f} s", flush=True)
import torch  # noqa: E402
from heuristique_style_transfer_code_b200 import ops  # noqa: E402
lines = []
for shape in ((256, 3, 224, 224), (64, 3, 448, 448)):
    u8 = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda")
    dst = torch.empty(shape, device="cuda")
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    for _ in range(3):
        ops.normalize_u8(u8, mean, std, out=dst)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ops.normalize_u8(u8, mean, std, out=dst)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    gbs = u8.numel() * 5 / us / 1e3
    lines.append(f"normalize_u8[{'x'.join(map(str, shape))}]: {us:.1f} us per launch, {gbs:.0f} GB/s of algorithmic bytes (1 B read + "
                 f"4 B written per value) = {gbs / 6548.8:.2f} of the measured HBM peak (6548.8 GB/s)")
    del u8, dst
open(os.path.join(OUT, "r4c_normalize_u8_timing.log"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines), flush=True)
