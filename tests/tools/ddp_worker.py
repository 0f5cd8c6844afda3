"""Worker of tests/test_gpu_distributed.py (one process per GPU, NCCL): a DDP training step of the drop-in model on a
batch shard must equal the single-GPU step on the concatenated batch -- with the encoder's batch
norm in eval mode, so that shard statistics do not enter. Rank 0 prints one JSON line with the worst relative errors."""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build(device):
    from torchvision import models
    from heuristique_style_transfer_code_b200 import TruncatedResNet50
    torch.manual_seed(0)
    m = TruncatedResNet50(models.resnet50(weights=None), 7, 4, 32, device=device)
    m.train()
    for mod in m.modules():                      # eval-mode batch norm: the step depends on the images only through sums
        if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
            mod.eval()
    return m


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
    ddp = D.wrap_ddp(model, device)
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(ddp(x[lo:hi]), y[lo:hi])
    loss.backward()
    opt.step()
    torch.cuda.synchronize(device)

    if rank == 0:
        single = build(device)
        sopt = torch.optim.SGD(single.parameters(), lr=0.05, momentum=0.9)
        sopt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(single(x), y).backward()
        sopt.step()
        torch.cuda.synchronize(device)
        worst_g, worst_w, head_g = 0.0, 0.0, 0.0
        for (n, p), (_, q) in zip(model.named_parameters(), single.named_parameters()):
            e = rel(p.grad, q.grad)
            worst_g = max(worst_g, e)
            if n.startswith(("attention", "classifier")):
                head_g = max(head_g, e)
            worst_w = max(worst_w, rel(p.detach(), q.detach()))
        print(json.dumps({"world": world, "worst_grad_rel": worst_g, "worst_head_grad_rel": head_g,
                          "worst_weight_rel": worst_w, "backend": dist.get_backend() if dist.is_initialized() else None}),
              flush=True)
    if dist.is_initialized():
        dist.barrier(device_ids=[device.index])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
