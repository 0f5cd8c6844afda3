"""Does a host->device copy overlap the forward on this box? (end-to-end numbers of bench.py differ box to box: 6.4 ms vs
8.8-9.7 ms per 256-image step with the same code.)  python tools/diag_copy_overlap.py > gpurun_out/copy_overlap.log
Times, with CUDA events on both streams and the host clock: the copy alone, the forward alone, both issued together in
either order, the copy cut in chunks, and the bench's end-to-end loop."""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torchvision import models  # noqa: E402
from heuristique_style_transfer_code_b200 import TruncatedResNet50_for_test  # noqa: E402
from heuristique_style_transfer_code_b200.functions import HostCollector, cuda_prefetch  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    print("CUDA_DEVICE_MAX_CONNECTIONS =", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), flush=True)
    q = subprocess.run(["nvidia-smi", "--query-gpu=name,compute_mode,mig.mode.current,persistence_mode,pcie.link.gen.current,"
                        "pcie.link.width.current", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    print("nvidia-smi:", q, flush=True)
    torch.manual_seed(0)
    model = TruncatedResNet50_for_test(models.resnet50(weights=None), 7, 4, 32, device=dev).eval()
    B = 256
    x_host = torch.randn(B, 3, 224, 224).pin_memory()
    x = x_host.to(dev)
    x2 = torch.empty_like(x)
    side = torch.cuda.Stream(dev)
    main_s = torch.cuda.current_stream(dev)

    def fwd():
        with torch.no_grad():
            model(x)

    def copy(chunks=1):
        n = B // chunks
        for c in range(chunks):
            x2[c * n:(c + 1) * n].copy_(x_host[c * n:(c + 1) * n], non_blocking=True)

    for _ in range(3):
        fwd()
        copy()
    torch.cuda.synchronize()

    def wall(fn, reps=10):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / reps * 1e3

    print(f"forward alone {wall(fwd):.2f} ms   copy alone {wall(copy):.2f} ms", flush=True)

    def both(copy_first, chunks=1):
        def run():
            if copy_first:
                with torch.cuda.stream(side):
                    copy(chunks)
                fwd()
            else:
                fwd()
                with torch.cuda.stream(side):
                    copy(chunks)
        return run

    for cf in (True, False):
        for chunks in (1, 8, 32):
            print(f"copy on a side stream issued {'before' if cf else 'after'} the forward, {chunks} chunk(s): "
                  f"{wall(both(cf, chunks)):.2f} ms per (forward + copy)", flush=True)

    def e2e(n=10):
        res = HostCollector()
        with torch.no_grad():
            for (xb,) in cuda_prefetch(((x_host,) for _ in range(n)), dev, reuse_buffers=True):
                emb, logits = model(xb)
                res.push(emb, logits)
        res.finish()

    e2e(3)
    print(f"end-to-end loop (cuda_prefetch + HostCollector): {wall(lambda: e2e(10), reps=2) / 10:.2f} ms per step", flush=True)

    # the same double-buffered loop written out, events on the device
    bufs = [torch.empty_like(x), torch.empty_like(x)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    used = [None, None]

    def manual(n=10):
        with torch.no_grad():
            with torch.cuda.stream(side):
                bufs[0].copy_(x_host, non_blocking=True)
                done[0].record(side)
            for i in range(n):
                s = i % 2
                if i + 1 < n:
                    with torch.cuda.stream(side):
                        if used[1 - s] is not None:
                            side.wait_event(used[1 - s])
                        bufs[1 - s].copy_(x_host, non_blocking=True)
                        done[1 - s].record(side)
                main_s.wait_event(done[s])
                model(bufs[s])
                ev = torch.cuda.Event()
                ev.record(main_s)
                used[s] = ev

    manual(3)
    print(f"hand-written double buffering, no result read-back: {wall(lambda: manual(10), reps=2) / 10:.2f} ms per step", flush=True)

    # Pinned host allocations inside a timed loop: HostCollector's ring needs 4 x (1 MB + 4 KB) of them; cudaHostAlloc maps
    # the pages for every GPU of the box and its cost differs widely between hosts.
    def pinned(nbytes):
        t = time.perf_counter()
        buf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        ms = (time.perf_counter() - t) * 1e3
        del buf
        return ms

    print("fresh pinned allocations: " + ", ".join(f"{n >> 10} KB {pinned(n):.2f} ms" for n in (5 << 10, 3 << 20, 5 << 20, 9 << 20)),
          flush=True)
    e2e(6)
    warm = wall(lambda: e2e(10), reps=1) / 10
    if hasattr(torch._C, "_host_emptyCache"):
        torch.cuda.synchronize()
        torch._C._host_emptyCache()                 # every ring slot of the next loop is a fresh cudaHostAlloc
        cold = wall(lambda: e2e(10), reps=1) / 10
        print(f"end-to-end loop, ring slots warm: {warm:.2f} ms per step; all 8 pinned buffers allocated inside the loop: "
              f"{cold:.2f} ms per step (10 steps)", flush=True)
    else:
        print(f"end-to-end loop, ring slots warm: {warm:.2f} ms per step", flush=True)


if __name__ == "__main__":
    main()
