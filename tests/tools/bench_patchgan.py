"""Multi-PatchGAN Gram head on one B200: MultiScaleDiscriminator_test (ndf 64, gram_matrix_dim 64,
batch norm, patches 10/70/150) at 224x224, this repo's classes against the fp32 torch port of the reference forward on
the same GPU and weights; per-kernel times and achieved bandwidth of the head kernels.
    python tests/tools/bench_patchgan.py [batch] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from heuristique_style_transfer_code_b200 import ops  # noqa: E402
from heuristique_style_transfer_code_b200.patchgan import MultiScaleDiscriminator_test  # noqa: E402
from oracle.torch_port import patchgan_multiscale_forward  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.manual_seed(0)
m = MultiScaleDiscriminator_test(ndf=64, norm='batch', num_classes=4, gram_matrix_dim=64).cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


with torch.no_grad():
    ours = timed(lambda: m(x), steps)
    port = timed(lambda: patchgan_multiscale_forward(m, x), steps)

    def extractor_only():
        for d in m.scale_discriminators.values():
            y, k = x, 0
            for layer in d.feature_extractor:
                y = layer(y)
                if isinstance(layer, torch.nn.Conv2d):
                    d.projection_layers[k](y)
                    k += 1
    conv = timed(extractor_only, steps)
    ops.PROFILE = []
    for _ in range(steps):
        m(x)
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    e1, o1 = m(x)
    e2, o2 = patchgan_multiscale_forward(m, x)
agg = {}
for name, work, s, e in rec:
    a = agg.setdefault(name, dict(ms=0.0, n=0, bytes=work["bytes"]))
    a["ms"] += s.elapsed_time(e)
    a["n"] += 1
kern = {k: dict(avg_us=round(v["ms"] / v["n"] * 1e3, 1), launches_per_step=v["n"] // steps,
                GBps=round(v["bytes"] / (v["ms"] / v["n"]) / 1e6, 1)) for k, v in agg.items()}
head_ms = sum(v["ms"] for v in agg.values()) / steps
print(json.dumps({"workload": f"MultiScaleDiscriminator_test ndf64 D64 batch {B} 224x224 eval, 3 scales",
                  "ours_ms": round(ours, 3), "torch_port_same_gpu_ms": round(port, 3), "speedup": round(port / ours, 2),
                  "extractor_and_projections_cudnn_ms": round(conv, 3), "head_kernels_ms": round(head_ms, 3),
                  "port_head_ms_estimate": round(port - conv, 3),
                  "images_per_s": round(B / ours * 1e3, 1), "port_images_per_s": round(B / port * 1e3, 1),
                  "emb_rel_diff": float((e1 - e2).norm() / e2.norm()), "out_rel_diff": float((o1 - o2).norm() / o2.norm()),
                  "kernels": kern}))
