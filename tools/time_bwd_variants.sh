#!/bin/bash
for f in heuristique_style_transfer_code_b200/csrc/variants/libgramhead_*.so; do
  echo "=== $f"
  GRAMHEAD_LIB=$PWD/$f python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from heuristique_style_transfer_code_b200 import ops
g = 32
for (B, C, HW, dt) in [(256, 256, 3136, "f32"), (256, 512, 784, "f32"), (256, 1024, 196, "f32"), (512, 256, 3136, "f32"), (256, 256, 3136, "bf16"), (256, 512, 784, "bf16")]:
    x = torch.relu(torch.randn(B, C, HW, device="cuda"))
    if dt == "bf16":
        x = x.bfloat16()
    dd = torch.randn(B, 1, g * g, device="cuda")
    for _ in range(3):
        ops.gram_pool_bwd(x, g, dd, 0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.gram_pool_bwd(x, g, dd, 0)
    b.record(); torch.cuda.synchronize()
    print(f"bwd {dt} B={B} C={C} HW={HW}: {a.elapsed_time(b)/10*1e3:8.1f} us", flush=True)
    del x, dd
PY
done
