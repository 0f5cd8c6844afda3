"""BASELINE.json configs[4]: Gram kernel sweep across truncation depths and batch sizes against the measured rooflines.
Kernel-only, features resident in HBM (relu(randn): ~50 % zeros like post-ReLU activations), CUDA events, >= 3 warm-ups,
each timed loop touches > 126 MB (or is flagged "L2-resident"). Features are channels_last (NHWC) 4-D tensors, the layout
the default backbone execution hands over (--nchw: (B, C, HW) rows, the reference's execution). Prints a markdown table;
run on a B200:
    python tools/bench_sweep.py [--nchw] > gpurun_out/sweep.md"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from heuristique_style_transfer_code_b200 import ops  # noqa: E402


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
    return a.elapsed_time(b) / n


def main():
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    hbm, tc = pk["hbm_gbs"], pk["bf16_tflops"]
    g = 32
    nchw = "--nchw" in sys.argv
    shapes = [(256, 3136), (512, 784), (1024, 196), (2048, 49), (256, 12544), (512, 3136), (1024, 784), (2048, 196)]
    batches = [64, 128, 256, 512, 1024, 2048]
    print(f"peaks: HBM {hbm} GB/s, bf16 {tc} TFLOP/s (burst: kernels timed alone), {torch.cuda.get_device_name(0)}; "
          f"features: {'NCHW rows (B, C, HW)' if nchw else 'channels_last (NHWC), the default hand-off'}; roofline us = "
          "max(algorithmic bytes / HBM, algorithmic FLOPs / bf16 peak) -- fp32 features run on tf32 operands (half the bf16 "
          "rate), so their 'frac of roofline' is against a peak that path cannot reach\n")
    print("| dir | dtype | C | HW | B | us | GB/s | of HBM | TFLOP/s (sym fwd / dense bwd) | of tensor | roofline us | frac of roofline | note |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for dtype in (torch.float32, torch.bfloat16):
        for (C, HW) in shapes:
            for B in batches:
                elems = B * C * HW
                if elems * 4 > 6e9 or (dtype is torch.bfloat16 and B not in (256, 1024)):
                    continue
                x = torch.relu(torch.randn(B, C, HW, device="cuda")).to(dtype)
                if not nchw:
                    side = int(round(HW ** 0.5))
                    x = x.view(B, C, side, side).contiguous(memory_format=torch.channels_last)
                desc = torch.empty(B, 1, g * g, device="cuda")
                dd = torch.randn(B, 1, g * g, device="cuda")
                es = x.element_size()
                note = "L2-resident" if elems * es < 126e6 else ""
                ms = timeit(lambda: ops.gram_pool_fwd_(x, g, desc, 0))
                by, fl = elems * es + B * g * g * 4, B * C * (C + 1) * HW
                roof = max(by / (hbm * 1e9), fl / (tc * 1e12)) * 1e3
                print(f"| fwd | {str(dtype)[6:]} | {C} | {HW} | {B} | {ms*1e3:.1f} | {by/ms/1e6:.0f} | {by/ms/1e6/hbm:.2f} | {fl/ms/1e9:.0f} | {fl/ms/1e9/tc:.2f} | {roof*1e3:.1f} | {roof/ms:.2f} | {note} |")
                if C % 16 == 0:
                    ms = timeit(lambda: ops.gram_pool_bwd(x, g, dd, 0))
                    by, fl = elems * (es + 4) + B * g * g * 4, 2 * B * C * C * HW
                    roof = max(by / (hbm * 1e9), fl / (tc * 1e12)) * 1e3
                    print(f"| bwd | {str(dtype)[6:]} | {C} | {HW} | {B} | {ms*1e3:.1f} | {by/ms/1e6:.0f} | {by/ms/1e6/hbm:.2f} | {fl/ms/1e9:.0f} | {fl/ms/1e9/tc:.2f} | {roof*1e3:.1f} | {roof/ms:.2f} | {note} |")
                del x, desc, dd
                sys.stdout.flush()


if __name__ == "__main__":
    main()
