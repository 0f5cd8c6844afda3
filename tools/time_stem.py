"""Stem of the folded encoder on a B200: the 7x7 stride-2 convolution with the input zero-padded to 3 / 4 / 8 channels
(channels_last, fp32 with TF32 allowed, and bf16), ATen's NHWC max pool against gh_maxpool2d_nhwc."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from heuristique_style_transfer_code_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


torch.manual_seed(0)
x3 = torch.randn(B, 3, 224, 224, device="cuda")
w3 = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
bias = torch.randn(64, device="cuda")
for dtype in (torch.float32, torch.bfloat16):
    ref = None
    for cin in (3, 4, 8):
        x = torch.empty(B, cin, 224, 224, device="cuda", dtype=dtype, memory_format=torch.channels_last).zero_()
        x[:, :3] = x3
        w = torch.zeros(64, cin, 7, 7, device="cuda", dtype=dtype)
        w[:, :3] = w3
        w = w.contiguous(memory_format=torch.channels_last)
        b = bias.to(dtype)
        f = lambda: torch.cudnn_convolution_relu(x, w, b, (2, 2), (3, 3), (1, 1), 1)
        y = f()
        if ref is None:
            ref = y
        def pad():
            xp = torch.empty(B, cin, 224, 224, device="cuda", dtype=dtype, memory_format=torch.channels_last).zero_()
            xp[:, :3] = x3
            return xp
        conv3 = lambda: x3.to(dtype).contiguous(memory_format=torch.channels_last)
        print(f"{dtype} cin={cin}: conv+bias+relu {timed(f):8.1f} us   input prep {timed(pad if cin > 3 else conv3):7.1f} us   "
              f"rel diff vs cin=3 {float((y.float() - ref.float()).norm() / ref.float().norm()):.2e}", flush=True)
    y = ref
    print(f"{dtype} max pool: ATen {timed(lambda: torch.nn.functional.max_pool2d(y, 3, 2, 1)):8.1f} us   "
          f"gh_maxpool2d_nhwc {timed(lambda: ops.maxpool2d_nhwc(y, 3, 2, 1)):8.1f} us   "
          f"roofline {(y.numel() * 1.25 * y.element_size()) / 6548.8e3:6.1f} us", flush=True)
