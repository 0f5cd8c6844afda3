"""Training step of the drop-in model, eager vs captured in a CUDA graph (functions.GraphedTrainStep), at a given batch
on one GPU:
    python tools/time_train_graph.py [--batch 64]"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
os.environ.setdefault("NCCL_ASYNC_ERROR_HANDLING", "0")

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--global-batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from torchvision import models
    from heuristique_style_transfer_code_b200 import TruncatedResNet50
    from heuristique_style_transfer_code_b200 import distributed as D
    from heuristique_style_transfer_code_b200.functions import GraphedTrainStep
    rank, world, local, device = D.init_from_env()
    batch = args.global_batch // world if args.global_batch else args.batch
    torch.manual_seed(0)
    model = TruncatedResNet50(models.resnet50(weights=None), 7, 4, 32, device=device).train()
    ref = TruncatedResNet50(models.resnet50(weights=None), 7, 4, 32, device=device).train()
    ref.load_state_dict(model.state_dict())
    torch.manual_seed(100 + rank)
    x = torch.randn(batch, 3, 224, 224, device=device)
    y = torch.randint(0, 4, (batch,), device=device)
    crit = torch.nn.CrossEntropyLoss()

    def timed(fn, steps):
        D.barrier(device)
        torch.cuda.synchronize(device)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        torch.cuda.synchronize(device)
        return D.max_over_ranks(s.elapsed_time(e) / steps, device)

    # eager
    side = torch.cuda.Stream(device=device)
    with torch.cuda.stream(side):
        ddp_e = D.wrap_ddp(ref, device, bucket_cap_mb=16, static_graph=True)
    torch.cuda.current_stream(device).wait_stream(side)
    opt_e = torch.optim.AdamW(ref.parameters(), lr=1e-3, fused=True)

    def eager():
        opt_e.zero_grad(set_to_none=True)
        loss = crit(ddp_e(x), y)
        loss.backward()
        opt_e.step()
        return loss
    for _ in range(3):
        eager()
    ems = timed(eager, args.steps)

    if world > 1:
        raise SystemExit("time_train_graph.py: the graphed step is single-process (a captured DDP step hung in testing)")
    ddp_g = model
    opt_g = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True, capturable=True)
    step = GraphedTrainStep(ddp_g, crit, opt_g, x, y)
    for _ in range(3):
        step()
    gms = timed(lambda: step(), args.steps)
    # same arithmetic: both models started from the same weights; compare after the same number of updates is not
    # possible (different step counts), so check one more step's loss is finite and the parameters stayed finite
    loss = step().item()
    finite = all(bool(torch.isfinite(p).all()) for p in model.parameters())
    if rank == 0:
        print(json.dumps({"world": world, "per_gpu_batch": batch, "eager_ms": round(ems, 3), "graph_ms": round(gms, 3),
                          "eager_img_s": round(batch * world / ems * 1e3, 1), "graph_img_s": round(batch * world / gms * 1e3, 1),
                          "loss": loss, "finite": finite}), flush=True)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
