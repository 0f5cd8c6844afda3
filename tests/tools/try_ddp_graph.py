"""Experiment: capture the DDP training step (NCCL all-reduce included) in a CUDA graph. Prints a marker before each phase so
that a hang can be located; run under torchrun with a short outer timeout.
    GH_PDL=0|1  GH_WARM=11  python -m torch.distributed.run --nproc-per-node 2 ... tests/tools/try_ddp_graph.py"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"
os.environ["NCCL_ASYNC_ERROR_HANDLING"] = "0"

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def mark(rank, msg):
    print(f"[rank {rank} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def main():
    from torchvision import models
    from heuristique_style_transfer_code_b200 import TruncatedResNet50, _lib
    from heuristique_style_transfer_code_b200 import distributed as D
    rank, world, local, device = D.init_from_env()
    _lib.lib().gh_set_option(b"pdl", int(os.environ.get("GH_PDL", "1")))
    gb = int(os.environ.get("GH_GLOBAL_BATCH", "512"))
    batch = gb // world
    warm = int(os.environ.get("GH_WARM", "11"))
    torch.manual_seed(0)
    model = TruncatedResNet50(models.resnet50(weights=None), 7, 4, 32, device=device).train()
    torch.manual_seed(100 + rank)
    x = torch.randn(batch, 3, 224, 224, device=device)
    y = torch.randint(0, 4, (batch,), device=device)
    crit = torch.nn.CrossEntropyLoss()
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    mark(rank, "constructing DDP on a side stream")
    with torch.cuda.stream(side):
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], bucket_cap_mb=16,
                                                        gradient_as_bucket_view=True)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True, capturable=True)

        def step():
            loss = crit(ddp(x), y)
            loss.backward()
            opt.step()
            return loss
        mark(rank, f"{warm} eager warm-up iterations")
        for _ in range(warm):
            opt.zero_grad(set_to_none=True)
            step()
    torch.cuda.current_stream(device).wait_stream(side)
    torch.cuda.synchronize(device)
    mark(rank, "eager timing")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(device_ids=[device.index])
    s.record()
    for _ in range(8):
        opt.zero_grad(set_to_none=True)
        step()
    e.record()
    torch.cuda.synchronize(device)
    eager_ms = s.elapsed_time(e) / 8
    mark(rank, f"eager {eager_ms:.2f} ms; capturing")
    opt.zero_grad(set_to_none=True)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = step()
    mark(rank, "captured; replaying")
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(device)
    mark(rank, "replayed 3x; timing")
    dist.barrier(device_ids=[device.index])
    s.record()
    for _ in range(8):
        graph.replay()
    e.record()
    torch.cuda.synchronize(device)
    graph_ms = s.elapsed_time(e) / 8
    t = torch.tensor([eager_ms, graph_ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "per_gpu_batch": batch, "eager_ms": round(float(t[0]), 3),
                          "graph_ms": round(float(t[1]), 3), "eager_img_s": round(gb / float(t[0]) * 1e3, 1),
                          "graph_img_s": round(gb / float(t[1]) * 1e3, 1), "loss": float(loss.item()),
                          "pdl": os.environ.get("GH_PDL", "1")}), flush=True)
    mark(rank, "releasing the graph")
    del graph, loss
    torch.cuda.synchronize(device)
    mark(rank, "destroying the process group")
    dist.destroy_process_group()
    mark(rank, "done")


if __name__ == "__main__":
    main()
