"""One short GPU visit (the round's last two GPU-minutes): the three GPU tests of the opt-in uint8 upload, then the
inference-only bench line (e2e with its three passes, e2e_uint8), in ONE process so that torch is imported once."""
import io
import os
import runpy
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
t0 = time.time()
import pytest  # noqa: E402

if os.environ.get("VISIT_SKIP_PYTEST") != "1":
    with open(os.path.join(OUT, "r4b_pytest_u8.log"), "w") as log:
        stdout = sys.stdout
        sys.stdout = log
        try:
            rc = pytest.main(["tests/test_gpu_module.py", "-q", "-x", "-m", "gpu", "-k", "normalize_u8 or uint8", "-p", "no:cacheprovider"])
        finally:
            sys.stdout = stdout
    print(f"pytest rc={int(rc)} after {time.time() - t0:.1f} s", flush=True)
sys.argv = ["bench.py", "--steps", "20", "--warmup", "5", "--skip-train", "--skip-cpu", "--skip-handoff", "--skip-patchgan",
            "--skip-reference-gpu"]
fd = os.open(os.path.join(OUT, "r4b_bench_infer.json"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
os.dup2(fd, 1)
runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
sys.stderr.write(f"bench done after {time.time() - t0:.1f} s\n")
# gh_normalize_u8 alone: CUDA events around 20 launches on a 256-image batch (38.5 MB in, 154 MB out: larger than L2)
import torch  # noqa: E402
from heuristique_style_transfer_code_b200 import ops  # noqa: E402
u8 = torch.randint(0, 256, (256, 3, 224, 224), dtype=torch.uint8, device="cuda")
dst = torch.empty((256, 3, 224, 224), device="cuda")
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
with open(os.path.join(OUT, "r4b_normalize_u8_timing.log"), "w") as f:
    f.write(f"normalize_u8[256x3x224x224]: {us:.1f} us per launch, {u8.numel() * 5 / us / 1e3:.0f} GB/s of algorithmic bytes "
            f"(1 B read + 4 B written per value; measured HBM peak 6548.8 GB/s -> {u8.numel() * 5 / us / 1e3 / 6548.8:.2f})\n")
del u8, dst
# best effort with whatever time is left: the longer copy/compute overlap probe (informative on a box whose e2e is slow)
os.dup2(os.open(os.path.join(OUT, "r4b_copy_overlap.log"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644), 1)
sys.argv = ["diag_copy_overlap.py"]
runpy.run_path(os.path.join(ROOT, "tools", "diag_copy_overlap.py"), run_name="__main__")
