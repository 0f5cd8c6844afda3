"""Tiny single-kernel drivers for `ncu --set full` captures (a few launches, nothing else on the GPU).
    python tools/prof_one.py bwd|fwd|pool|gemm [C HW B [f32|bf16 [pair(-1|0|1) [nchw|nhwc]]]]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from heuristique_style_transfer_code_b200 import ops, _lib  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "bwd"
C, HW, B = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (256, 3136, 256)
dt = sys.argv[5] if len(sys.argv) > 5 else "f32"
pair = int(sys.argv[6]) if len(sys.argv) > 6 else -1
layout = sys.argv[7] if len(sys.argv) > 7 else "nchw"
g = 32
torch.manual_seed(0)
_lib.lib().gh_set_option(b"gram_fwd_pair", pair)
_lib.lib().gh_set_option(b"gram_bwd_pair", pair)
if kind in ("bwd", "fwd"):
    x = torch.relu(torch.randn(B, C, HW, device="cuda"))
    if dt == "bf16":
        x = x.bfloat16()
    if layout == "nhwc":                       # what a channels_last backbone hands over
        side = int(round(HW ** 0.5))
        x = x.view(B, C, side, side).contiguous(memory_format=torch.channels_last)
    desc = torch.empty(B, 1, g * g, device="cuda")
    dd = torch.randn(B, 1, g * g, device="cuda")
    for _ in range(4):
        if kind == "bwd":
            ops.gram_pool_bwd(x, g, dd, 0)
        else:
            ops.gram_pool_fwd_(x, g, desc, 0)
elif kind == "pool":                          # stem max pool of the inference plan: (B, 64, 112, 112) channels_last
    y = torch.relu(torch.randn(B, 64, 112, 112, device="cuda")).contiguous(memory_format=torch.channels_last)
    if dt == "bf16":
        y = y.bfloat16()
    for _ in range(4):
        ops.maxpool2d_nhwc(y, 3, 2, 1)
else:
    a = torch.randn(768, 1024, device="cuda")
    w = torch.randn(3072, 1024, device="cuda")
    for _ in range(4):
        ops.gemm_f32(a, w.t())
torch.cuda.synchronize()
print("done", kind, C, HW, B, dt, pair, layout)
