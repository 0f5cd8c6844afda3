"""Diagnostic: loss trajectories of a few training steps -- eager vs GraphedTrainStep, TMA vs ld.global attention kernels, with
and without programmatic dependent launch. All six lines must agree (found the stale weight-plane cache under fused AdamW)."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from torchvision import models
from heuristique_style_transfer_code_b200 import TruncatedResNet50, _lib, ops
from heuristique_style_transfer_code_b200.functions import GraphedTrainStep
torch.backends.cudnn.allow_tf32 = False
def run(kind, pdl, attn_impl="tma"):
    _lib.lib().gh_set_option(b"pdl", pdl)
    ops.ATTN_IMPL = attn_impl
    ops.clear_weight_planes()
    torch.manual_seed(0)
    m = TruncatedResNet50(models.resnet50(weights=None), 6, 4, 32, device="cuda").train()
    torch.manual_seed(1)
    x = torch.randn(8, 3, 96, 96, device="cuda"); y = torch.randint(0, 4, (8,), device="cuda")
    crit = torch.nn.CrossEntropyLoss()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, fused=True, capturable=True)
    warm, k = 3, 4
    if kind == "graph":
        step = GraphedTrainStep(m, crit, opt, x, y, warmup=warm)
        return [round(step().item(), 5) for _ in range(k)]
    out = []
    for i in range(warm + k):
        opt.zero_grad(set_to_none=True)
        loss = crit(m(x), y); loss.backward(); opt.step()
        if i >= warm: out.append(round(loss.item(), 5))
    return out
for kind, pdl, impl in (("eager", 1, "tma"), ("eager", 1, "tma"), ("eager", 0, "ldg"), ("graph", 0, "ldg"), ("graph", 0, "tma"), ("graph", 1, "tma")):
    print(kind, "pdl", pdl, impl, run(kind, pdl, impl), flush=True)
