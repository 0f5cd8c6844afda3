"""Gram kernels on channels_last (NHWC) features vs NCHW, kernel-only, batch 256 (python tools/time_nhwc.py)."""
import sys
import torch
sys.path.insert(0, '.')
from heuristique_style_transfer_code_b200 import ops


def t_us(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


g = 32
for (B, C, H) in [(256, 256, 56), (256, 512, 28), (256, 1024, 14), (512, 256, 56)]:
    for dt in (torch.float32, torch.bfloat16):
        x = torch.relu(torch.randn(B, C, H, H, device="cuda")).to(dt)
        xcl = x.contiguous(memory_format=torch.channels_last)
        desc = torch.empty(B, 1, g * g, device="cuda")
        dd = torch.randn(B, 1, g * g, device="cuda")
        row = f"{str(dt)[6:]:9s} B={B} C={C} HW={H*H}:"
        for name, t in (("nchw", x), ("nhwc", xcl)):
            f = t_us(lambda: ops.gram_pool_fwd_(t, g, desc, 0))
            bw = t_us(lambda: ops.gram_pool_bwd(t, g, dd, 0))
            row += f"  {name} fwd {f:7.1f} us bwd {bw:7.1f} us |"
        tr = t_us(lambda: ops.nhwc_to_nchw(xcl))
        print(row + f"  transpose {tr:6.1f} us", flush=True)
        del x, xcl, desc, dd
