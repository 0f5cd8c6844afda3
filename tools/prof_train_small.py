"""Is the small-shard training step (per-GPU batch 64 = global 512 over 8 ranks) launch-bound or GPU-bound?
    python tools/prof_train_small.py [batch]
Prints the event-timed step, the summed kernel time from torch.profiler (GPU busy), the launch count and the top kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from torchvision import models  # noqa: E402

from heuristique_style_transfer_code_b200 import TruncatedResNet50  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TruncatedResNet50(models.resnet50(weights=None), 7, 4, 32, device=dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
crit = torch.nn.CrossEntropyLoss()
x = torch.randn(B, 3, 224, 224, device=dev)
y = torch.randint(0, 4, (B,), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x), y)
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    step()
e.record()
torch.cuda.synchronize()
print(f"batch {B}: {s.elapsed_time(e) / 10:.2f} ms/step (events, 10 steps)")

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = [k for k in prof.key_averages() if k.device_time_total > 0 and k.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(k.device_time_total for k in ev) / 3 / 1e3
n = sum(k.count for k in ev) / 3
print(f"GPU busy {busy:.2f} ms/step over {n:.0f} kernels+copies per step ({busy * 1e3 / n:.1f} us average)")
for k in sorted(ev, key=lambda k: -k.device_time_total)[:25]:
    print(f"{k.device_time_total / 3 / 1e3:8.3f} ms {k.count / 3:6.0f}x {k.device_time_total / k.count:8.1f} us  {k.key[:110]}")
small = [k for k in ev if k.device_time_total / k.count < 8.0]
print(f"kernels under 8 us: {sum(k.count for k in small) / 3:.0f} per step, {sum(k.device_time_total for k in small) / 3 / 1e3:.2f} ms")
