"""Worker of tests/test_gpu_distributed.py (one process per GPU, NCCL): a DDP training step of the drop-in model on a
batch shard must equal the single-GPU step on the concatenated batch -- with the encoder's batch
norm in eval mode, so that shard statistics do not enter. Rank 0 prints one JSON line with the worst relative errors."""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")      # required for capturing a DDP step in a CUDA graph
os.environ.setdefault("NCCL_ASYNC_ERROR_HANDLING", "0")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build(device):
    from torchvision import models
    from heuristique_style_transfer_code_b200 import TruncatedResNet50
    torch.manual_seed(0)
    m = TruncatedResNet50(models.resnet50(weights=None), 7, 4, 32, device=device)
    m.train()
    return m


def freeze_batchnorm(m, x):
    """Eval-mode batch norm (the step then depends on the images only through sums over the batch), with running
    statistics calibrated on the full batch first: with the constructor's statistics (mean 0, variance 1) a random-init
    encoder's activations grow by orders of magnitude per stage and the attention softmax saturates, which
    makes every comparison of gradients a comparison of rounding noise."""
    bns = [mod for mod in m.modules() if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm)]
    for mod in bns:
        mod.momentum = 1.0                       # running statistics := this batch's statistics
    with torch.no_grad():
        m(x)
    for mod in bns:
        mod.eval()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def main():
    from heuristique_style_transfer_code_b200 import distributed as D
    rank, world, local, device = D.init_from_env()
    torch.backends.cudnn.allow_tf32 = False
    batch = 4 * world
    torch.manual_seed(1)
    x = torch.randn(batch, 3, 128, 128, device=device)
    y = torch.randint(0, 4, (batch,), device=device)
    lo, hi = D.shard_bounds(batch, rank, world)

    model = build(device)
    freeze_batchnorm(model, x)                   # every rank calibrates on the same full batch (same seed)
    ddp = D.wrap_ddp(model, device)
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(ddp(x[lo:hi]), y[lo:hi])
    loss.backward()
    opt.step()
    torch.cuda.synchronize(device)

    if rank == 0:
        single = build(device)
        freeze_batchnorm(single, x)
        sopt = torch.optim.SGD(single.parameters(), lr=0.05, momentum=0.9)
        sopt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(single(x), y).backward()
        sopt.step()
        torch.cuda.synchronize(device)
        worst_g, worst_w, head_g, worst_name = 0.0, 0.0, 0.0, ""
        for (n, p), (_, q) in zip(model.named_parameters(), single.named_parameters()):
            e = rel(p.grad, q.grad)
            if e > worst_g:
                worst_g, worst_name = e, n
            if n.startswith(("attention", "classifier")):
                head_g = max(head_g, e)
            worst_w = max(worst_w, rel(p.detach(), q.detach()))
        print(json.dumps({"world": world, "worst_grad_rel": worst_g, "worst_grad_name": worst_name,
                          "worst_head_grad_rel": head_g, "worst_weight_rel": worst_w,
                          "backend": dist.get_backend() if dist.is_initialized() else None}), flush=True)
    if "--graph" in sys.argv:
        # functions.GraphedTrainStep on a DDP model: the replayed step (NCCL all-reduce inside the graph) follows the eager
        # DDP step from the same weights on the same shard
        from heuristique_style_transfer_code_b200.functions import GraphedTrainStep
        crit = torch.nn.CrossEntropyLoss()
        losses = {}
        warm, k = 11, 3
        for kind in ("eager", "graph"):
            m = build(device)
            m.load_state_dict(model.state_dict())
            ddp2 = D.wrap_ddp(m, device, bucket_cap_mb=16, for_graph_capture=True)
            o = torch.optim.AdamW(m.parameters(), lr=1e-4, fused=True, capturable=True)
            xs, ys = x[lo:hi].contiguous(), y[lo:hi].contiguous()
            if kind == "graph":
                step = GraphedTrainStep(ddp2, crit, o, xs, ys, warmup=warm)
                losses[kind] = [step().item() for _ in range(k)]
                step.release()
                del step
            else:
                out = []
                for i in range(warm + k):
                    o.zero_grad(set_to_none=True)
                    loss = crit(ddp2(xs), ys)
                    loss.backward()
                    o.step()
                    if i >= warm:
                        out.append(loss.item())
                losses[kind] = out
            del ddp2, o, m
            torch.cuda.synchronize(device)
        if rank == 0:
            print(json.dumps({"graph_losses": losses["graph"], "eager_losses": losses["eager"]}), flush=True)
    if dist.is_initialized():
        dist.barrier(device_ids=[device.index])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
