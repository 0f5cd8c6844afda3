"""Tiny single-kernel drivers for `ncu --set full` captures (a few launches, nothing else on the GPU).
    python tools/prof_one.py bwd|fwd|gemm [C HW B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from heuristique_style_transfer_code_b200 import ops  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "bwd"
C, HW, B = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (256, 3136, 256)
g = 32
torch.manual_seed(0)
if kind in ("bwd", "fwd"):
    x = torch.relu(torch.randn(B, C, HW, device="cuda"))
    desc = torch.empty(B, 1, g * g, device="cuda")
    dd = torch.randn(B, 1, g * g, device="cuda")
    for _ in range(4):
        if kind == "bwd":
            ops.gram_pool_bwd(x, g, dd, 0)
        else:
            ops.gram_pool_fwd_(x, g, desc, 0)
else:
    a = torch.randn(768, 1024, device="cuda")
    w = torch.randn(3072, 1024, device="cuda")
    for _ in range(4):
        ops.gemm_f32(a, w.t())
torch.cuda.synchronize()
print("done", kind, C, HW, B)
