"""End-to-end loop variants on one GPU: per-step .cpu() (the reference's shape) vs HostCollector at several depths."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torchvision import models  # noqa: E402

from heuristique_style_transfer_code_b200 import TruncatedResNet50_for_test  # noqa: E402
from heuristique_style_transfer_code_b200.functions import HostCollector, cuda_prefetch  # noqa: E402

device = torch.device("cuda:0")
B, steps = 256, 20
torch.manual_seed(0)
model = TruncatedResNet50_for_test(models.resnet50(weights=None), 7, 4, 32, device=str(device)).to(device).eval()
x_host = torch.randn(B, 3, 224, 224).pin_memory()
x_dev = x_host.to(device)


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


def resident():
    with torch.no_grad():
        for _ in range(steps):
            model(x_dev)


def sync_loop():
    with torch.no_grad():
        for (xb,) in cuda_prefetch(((x_host,) for _ in range(steps)), device, reuse_buffers=True):
            emb, logits = model(xb)
            emb.cpu(), logits.cpu()


def collector_loop(depth):
    def run():
        c = HostCollector(depth)
        with torch.no_grad():
            for (xb,) in cuda_prefetch(((x_host,) for _ in range(steps)), device, reuse_buffers=True):
                emb, logits = model(xb)
                c.push(emb, logits)
        c.finish()
    return run


def noprefetch_collector():
    c = HostCollector(2)
    with torch.no_grad():
        for _ in range(steps):
            xb = x_host.to(device, non_blocking=True)
            emb, logits = model(xb)
            c.push(emb, logits)
    c.finish()


resident()
for name, fn in [("resident", resident), ("sync", sync_loop), ("collector2", collector_loop(2)), ("collector4", collector_loop(4)),
                 ("collector8", collector_loop(8)), ("noprefetch_collector2", noprefetch_collector), ("sync", sync_loop),
                 ("collector2", collector_loop(2))]:
    a = timed(fn)
    b = timed(fn)
    print(f"{name:24s} first {a:7.2f} ms/step   second {b:7.2f} ms/step   mem reserved {torch.cuda.memory_reserved() / 2**30:.2f} GiB", flush=True)
