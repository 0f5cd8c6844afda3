"""DDP training-step sweep for BASELINE.json configs[2] (global batch 512, strong scaling), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 \
        tools/sweep_ddp.py [--global-batch 512] [--out gpurun_out/ddp_sweep.json] [--timeline]

Times the step (CUDA events, max over ranks) for combinations of DDP bucket size, static_graph, bf16 gradient
compression, fused AdamW and cudnn.benchmark, and -- with --timeline -- profiles three steps of the best variant on
rank 0 with torch.profiler and reports how much of the NCCL all-reduce time is NOT covered by compute kernels
(the exposed communication) together with the busiest kernels of the step."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build_step(device, world, rank, gb, bucket, static_graph, compression, fused, image=224):
    from torchvision import models
    from heuristique_style_transfer_code_b200 import TruncatedResNet50
    from heuristique_style_transfer_code_b200 import distributed as D
    lo, hi = D.shard_bounds(gb, rank, world)
    torch.manual_seed(0)
    model = TruncatedResNet50(models.resnet50(weights=None), 7, 4, 32, device=device).train()
    ddp = D.wrap_ddp(model, device, bucket_cap_mb=bucket, static_graph=static_graph, grad_compression=compression)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=fused)
    crit = torch.nn.CrossEntropyLoss()
    torch.manual_seed(100 + rank)
    x = torch.randn(hi - lo, 3, image, image, device=device)
    y = torch.randint(0, 4, (hi - lo,), device=device)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(ddp(x), y)
        loss.backward()
        opt.step()
        return loss
    return step


def time_step(step, device, D, warmup=3, steps=8):
    for _ in range(warmup):
        step()
    D.barrier(device)
    torch.cuda.synchronize(device)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        step()
    e.record()
    torch.cuda.synchronize(device)
    D.barrier(device)
    return D.max_over_ranks(s.elapsed_time(e) / steps, device)


def union_length(intervals):
    total, cur_s, cur_e = 0.0, None, None
    for s, e in sorted(intervals):
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                total += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        total += cur_e - cur_s
    return total


def overlap_length(a, b):
    """Length of (union of a) intersected with (union of b)."""
    return union_length(a) + union_length(b) - union_length(list(a) + list(b))


def timeline(step, device, nsteps=3):
    from torch.profiler import ProfilerActivity, profile
    for _ in range(2):
        step()
    torch.cuda.synchronize(device)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(nsteps):
            step()
        torch.cuda.synchronize(device)
    nccl, compute, by_name = [], [], {}
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA or getattr(ev, "is_user_annotation", False):
            continue
        if ev.name.startswith(("DistributedDataParallel", "autograd::", "Optimizer.", "ProfilerStep")):
            continue
        s, d = ev.time_range.start, ev.time_range.end - ev.time_range.start
        if d <= 0:
            continue
        name = ev.name
        if "Memcpy" in name or "Memset" in name:
            continue
        if "nccl" in name.lower():
            nccl.append((s, s + d))
        else:
            compute.append((s, s + d))
            a = by_name.setdefault(name[:90], [0.0, 0])
            a[0] += d
            a[1] += 1
    span = (max(e for _, e in compute + nccl) - min(s for s, _ in compute + nccl)) if compute or nccl else 0.0
    nccl_len = union_length(nccl)
    hidden = overlap_length(nccl, compute)
    top = sorted(by_name.items(), key=lambda kv: -kv[1][0])[:14]
    return {"steps": nsteps, "span_us_per_step": round(span / nsteps, 1),
            "compute_busy_us_per_step": round(union_length(compute) / nsteps, 1),
            "nccl_us_per_step": round(nccl_len / nsteps, 1), "nccl_hidden_us_per_step": round(hidden / nsteps, 1),
            "nccl_exposed_us_per_step": round((nccl_len - hidden) / nsteps, 1),
            "gpu_idle_us_per_step": round((span - union_length(compute + nccl)) / nsteps, 1),
            "top_compute_kernels_us_per_step": [{"name": k, "us": round(v[0] / nsteps, 1), "launches": v[1] // nsteps}
                                                for k, v in top]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-batch", type=int, default=512)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ddp_sweep.json"))
    ap.add_argument("--timeline", action="store_true")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    from heuristique_style_transfer_code_b200 import distributed as D
    rank, world, local, device = D.init_from_env()
    gb = args.global_batch
    variants = [dict(bucket=25, static_graph=False, compression="none", fused=False, benchmark=False)]      # round-1 setting
    for bucket in ((8, 16) if args.quick else (4, 8, 16, 50)):
        variants.append(dict(bucket=bucket, static_graph=True, compression="none", fused=True, benchmark=False))
    variants.append(dict(bucket=25, static_graph=True, compression="none", fused=True, benchmark=False))
    variants.append(dict(bucket=8, static_graph=True, compression="bf16", fused=True, benchmark=False))
    variants.append(dict(bucket=16, static_graph=True, compression="bf16", fused=True, benchmark=False))
    variants.append(dict(bucket=8, static_graph=True, compression="none", fused=True, benchmark=True))
    results = []
    for v in variants:
        torch.backends.cudnn.benchmark = v["benchmark"]
        step = build_step(device, world, rank, gb, v["bucket"], v["static_graph"], v["compression"], v["fused"])
        ms = time_step(step, device, D)
        results.append(dict(v, ms_per_step=round(ms, 3), images_per_s=round(gb / ms * 1e3, 1)))
        if rank == 0:
            print(results[-1], flush=True)
        del step
        torch.cuda.empty_cache()
    torch.backends.cudnn.benchmark = False
    best = min(results, key=lambda r: r["ms_per_step"])
    out = {"world": world, "global_batch": gb, "per_gpu_batch": gb // world, "variants": results, "best": best}
    if args.timeline:
        for tag, v in (("round1_setting", variants[0]), ("best", best)):
            torch.backends.cudnn.benchmark = v["benchmark"]
            step = build_step(device, world, rank, gb, v["bucket"], v["static_graph"], v["compression"], v["fused"])
            tl = timeline(step, device) if rank == 0 else [step() for _ in range(5)] and None
            if rank == 0:
                out["timeline_" + tag] = tl
                print(tag, json.dumps(tl)[:600], flush=True)
            D.barrier(device)
            del step
            torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
        print("best:", best, flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
