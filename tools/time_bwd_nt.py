import sys, torch
sys.path.insert(0, '.')
from heuristique_style_transfer_code_b200 import ops, _lib
g = 32
for (B, C, HW) in [(256, 256, 3136), (256, 512, 784), (256, 1024, 196)]:
    for dt in ("f32", "bf16"):
        x = torch.relu(torch.randn(B, C, HW + (4 if (dt == "bf16" and HW == 196) else 0), device="cuda"))
        if dt == "bf16":
            x = x.bfloat16()
        hw = x.shape[2]
        dd = torch.randn(B, 1, g * g, device="cuda")
        for nt in (64, 128, 192, 256, 0):
            _lib.lib().gh_set_option(b"gram_bwd_nt", nt)
            for _ in range(3):
                ops.gram_pool_bwd(x, g, dd, 0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                ops.gram_pool_bwd(x, g, dd, 0)
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / 10 * 1e3
            ntiles = (hw + (nt or 1) - 1) // (nt or 1) if nt else 0
            print(f"bwd {dt} B={B} C={C} HW={hw} NT={nt:3d} tiles/img={ntiles}: {us:8.1f} us", flush=True)
        del x, dd
