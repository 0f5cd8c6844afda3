"""Where does the end-to-end (host buffers in, host results out) step lose time against the device-resident step when
several ranks share one box?  torchrun --nproc-per-node N tools/diag_e2e.py [--affinity 0|1]
Times, per rank: H2D bandwidth alone, the device-resident forward, the prefetch loop with / without the gather and the
D2H, and prints one line per rank."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torchvision import models  # noqa: E402

from heuristique_style_transfer_code_b200 import TruncatedResNet50_for_test, distributed as D  # noqa: E402
from heuristique_style_transfer_code_b200.functions import cuda_prefetch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--affinity", type=int, default=-1)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
if args.affinity >= 0:
    os.environ["GRAMHEAD_CPU_AFFINITY"] = str(args.affinity)
rank, world, local, device = D.init_from_env()
B = 256
total = B * world
torch.manual_seed(0)
model = TruncatedResNet50_for_test(models.resnet50(weights=None), 7, 4, 32, device=str(device)).to(device).eval()
x_host = torch.randn(B, 3, 224, 224).pin_memory()
x_dev = x_host.to(device)


def timed(fn, n):
    D.barrier(device); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def h2d():
    x_dev.copy_(x_host, non_blocking=True)


def fwd():
    with torch.no_grad():
        model(x_dev)


def loop(gather, d2h):
    def run():
        with torch.no_grad():
            for (xb,) in cuda_prefetch(((x_host,) for _ in range(args.steps)), device):
                emb, logits = model(xb)
                if gather and world > 1:
                    logits = D.gather_rows(logits, total, world)
                    emb = D.gather_rows(emb, total, world)
                if d2h:
                    emb.cpu(), logits.cpu()
    return run


for _ in range(3):
    fwd(); h2d()
loop(True, True)()
res = {
    "h2d_ms": timed(h2d, args.steps), "fwd_ms": timed(fwd, args.steps),
    "loop_nogather_nod2h": timed(loop(False, False), 1) / args.steps,
    "loop_nogather_d2h": timed(loop(False, True), 1) / args.steps,
    "loop_gather_d2h": timed(loop(True, True), 1) / args.steps,
}
res["h2d_GBps"] = x_host.numel() * 4 / res["h2d_ms"] / 1e6
aff = sorted(os.sched_getaffinity(0))
print(f"rank {rank}/{world} cpus={aff[0]}..{aff[-1]} ({len(aff)}) omp={os.environ.get('OMP_NUM_THREADS')} "
      + " ".join(f"{k}={v:.2f}" for k, v in res.items()), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
