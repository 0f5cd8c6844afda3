"""Times the encoder and the whole model in each backbone hand-off mode on one GPU.
    python tools/bench_backbone_modes.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torchvision import models  # noqa: E402
from heuristique_style_transfer_code_b200 import TruncatedResNet50, TruncatedResNet50_for_test  # noqa: E402
from heuristique_style_transfer_code_b200.modules import BACKBONE_MODES  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def time_ms(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


torch.manual_seed(0)
model = TruncatedResNet50_for_test(models.resnet50(weights=None), 7, 4, 32, device="cuda").eval()
tmodel = TruncatedResNet50(models.resnet50(weights=None), 7, 4, 32, device="cuda").train()
tmodel.load_state_dict(model.state_dict())
opt = torch.optim.AdamW(tmodel.parameters(), lr=1e-3)
x = torch.randn(B, 3, 224, 224, device="cuda")
y = torch.randint(0, 4, (B,), device="cuda")
ref_logits = None
for mode in BACKBONE_MODES:
    model.set_backbone_mode(mode)
    tmodel.set_backbone_mode(mode)
    with torch.no_grad():
        enc = time_ms(lambda: model._stage_activations(x))
        full = time_ms(lambda: model(x))
        emb, logits = model(x)
    if ref_logits is None:
        ref_logits, ref_emb = logits.float(), emb.float()
    rel = float((logits.float() - ref_logits).norm() / ref_logits.norm())
    rel_e = float((emb.float() - ref_emb).norm() / ref_emb.norm())
    same = bool((logits.argmax(1) == ref_logits.argmax(1)).all())

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(tmodel(x), y)
        loss.backward()
        opt.step()
    tr = time_ms(step, n=5, warm=2)
    _, stages = model._stage_activations(x)
    print(f"mode={mode:20s} B={B}: encoder {enc:7.2f} ms  forward {full:7.2f} ms ({B/full*1e3:8.0f} img/s)  train step {tr:7.2f} ms "
          f"({B/tr*1e3:7.0f} img/s)  logits rel diff vs reference mode {rel:.2e} emb {rel_e:.2e} argmax same={same}  "
          f"stage dtypes/strides: {[(str(s.dtype).split('.')[-1], tuple(s.stride())) for s in stages]}", flush=True)
