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

with open(os.path.join(OUT, "r4a_pytest_u8.log"), "w") as log:
    stdout = sys.stdout
    sys.stdout = log
    try:
        rc = pytest.main(["tests/test_gpu_module.py", "-q", "-x", "-m", "gpu", "-k", "normalize_u8 or uint8", "-p", "no:cacheprovider"])
    finally:
        sys.stdout = stdout
print(f"pytest rc={int(rc)} after {time.time() - t0:.1f} s", flush=True)
sys.argv = ["bench.py", "--steps", "20", "--warmup", "5", "--skip-train", "--skip-cpu", "--skip-handoff", "--skip-patchgan",
            "--skip-reference-gpu"]
fd = os.open(os.path.join(OUT, "r4a_bench_infer.json"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
os.dup2(fd, 1)
runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
sys.stderr.write(f"bench done after {time.time() - t0:.1f} s\n")
# best effort with whatever time is left: the longer copy/compute overlap probe (informative on a box whose e2e is slow)
os.dup2(os.open(os.path.join(OUT, "r4a_copy_overlap.log"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644), 1)
sys.argv = ["diag_copy_overlap.py"]
runpy.run_path(os.path.join(ROOT, "tools", "diag_copy_overlap.py"), run_name="__main__")
