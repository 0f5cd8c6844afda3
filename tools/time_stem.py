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

# ---- space-to-depth formulation: 7x7 stride-2 conv on 3 channels == 4x4 stride-1 conv on 12 channels ----
for dtype in (torch.float32, torch.bfloat16):
    x = x3.to(dtype)
    w = w3.to(dtype)
    b = bias.to(dtype)
    direct = torch.cudnn_convolution_relu(x.contiguous(memory_format=torch.channels_last),
                                          w.contiguous(memory_format=torch.channels_last), b, (2, 2), (3, 3), (1, 1), 1)
    w8 = torch.zeros(64, 3, 8, 8, device="cuda", dtype=dtype)
    w8[:, :, 1:, 1:] = w                                   # tap i' = i + 1, i' = 2a + r
    ws = w8.view(64, 3, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(64, 12, 4, 4).contiguous(memory_format=torch.channels_last)
    # input coordinate 2y + i' - 4 = 2(y + a - 2) + r: z[(c,r,s), Y, X] = in[c, 2Y + r, 2X + s], taps Y = y + a - 2
    z = torch.empty(B, 12, 112 + 3, 112 + 3, device="cuda", dtype=dtype, memory_format=torch.channels_last).zero_()
    zz = x.view(B, 3, 112, 2, 112, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, 12, 112, 112)
    z[:, :, 2:114, 2:114] = zz
    f = lambda: torch.cudnn_convolution_relu(z, ws, b, (1, 1), (0, 0), (1, 1), 1)
    y = f()
    print(f"{dtype} space-to-depth 4x4x12: conv+bias+relu {timed(f):8.1f} us  out {tuple(y.shape)}  "
          f"rel diff vs direct {float((y.float() - direct.float()).norm() / direct.float().norm()):.2e}", flush=True)
    for cpad in (16,):
        zp = torch.empty(B, cpad, 115, 115, device="cuda", dtype=dtype, memory_format=torch.channels_last).zero_()
        zp[:, :12] = z
        wp = torch.zeros(64, cpad, 4, 4, device="cuda", dtype=dtype)
        wp[:, :12] = ws
        wp = wp.contiguous(memory_format=torch.channels_last)
        g = lambda: torch.cudnn_convolution_relu(zp, wp, b, (1, 1), (0, 0), (1, 1), 1)
        y2 = g()
        print(f"{dtype} space-to-depth 4x4x{cpad}: conv+bias+relu {timed(g):8.1f} us  "
              f"rel diff vs direct {float((y2.float() - direct.float()).norm() / direct.float().norm()):.2e}", flush=True)
