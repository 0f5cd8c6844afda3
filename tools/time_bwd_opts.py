"""Gram backward timing experiments on a B200 (kernel alone, CUDA events, inputs > L2):
ring depths (gh_set_option gram_bwd_stages = a*16+b), x-tile widths (gram_bwd_nt) and the bf16 gradient output.
    python tools/time_bwd_opts.py > gpurun_out/bwd_opts.log"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from heuristique_style_transfer_code_b200 import _lib, ops  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    lib = _lib.lib()
    B, g = 256, 32
    for dtype in (torch.float32, torch.bfloat16):
        for C, side in ((256, 56), (512, 28), (1024, 14)):
            x = torch.relu(torch.randn(B, C, side, side, device="cuda")).to(dtype).contiguous(memory_format=torch.channels_last)
            dd = torch.randn(B, 1, g * g, device="cuda")
            row = [f"{str(dtype)[6:]} C={C} HW={side * side}:"]
            for a, b in ((4, 12), (5, 12)):
                assert lib.gh_set_option(b"gram_bwd_stages", a * 16 + b) == 0
                row.append(f"A{a}/B{b} {timeit(lambda: ops.gram_pool_bwd(x, g, dd, 0)):.1f}")
            lib.gh_set_option(b"gram_bwd_stages", 0)
            if side == 14:
                for nt in (208, 224):
                    lib.gh_set_option(b"gram_bwd_nt", nt)
                    row.append(f"NT{nt} {timeit(lambda: ops.gram_pool_bwd(x, g, dd, 0)):.1f}")
                lib.gh_set_option(b"gram_bwd_nt", 0)
            if dtype is torch.bfloat16:
                ops.BF16_GRADIENTS = False
                row.append(f"fp32-out {timeit(lambda: ops.gram_pool_bwd(x, g, dd, 0)):.1f}")
                ops.BF16_GRADIENTS = True
                row.append(f"bf16-out {timeit(lambda: ops.gram_pool_bwd(x, g, dd, 0)):.1f}")
            row.append(f"fwd {timeit(lambda: ops.gram_pool_fwd_(x, g, torch.empty(B, 1, g * g, device='cuda'), 0)):.1f}")
            print("  ".join(row), flush=True)


if __name__ == "__main__":
    main()
