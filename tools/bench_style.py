"""Style-transfer iteration: eager loop vs CUDA-graph replay, per-iteration time on one GPU.
    python tools/bench_style.py [layers]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torchvision import models  # noqa: E402
from heuristique_style_transfer_code_b200 import TruncatedResNet50_for_test  # noqa: E402
from heuristique_style_transfer_code_b200.functions import _StyleIteration  # noqa: E402

torch.manual_seed(0)
model = TruncatedResNet50_for_test(models.resnet50(weights=None), 7, 4, 32, device="cuda:0").eval()
out = {}
for layers in ([int(sys.argv[1])] if len(sys.argv) > 1 else [4, 5, 7]):
    encoder = torch.nn.Sequential(*list(model.truncated_encoder.children())[:layers]).to("cuda:0")
    image = torch.randn(1, 3, 224, 224, device="cuda:0")
    with torch.no_grad():
        target = model.gram_matrix(encoder(image))
    res = {"gram": list(target.shape)}
    for use_graph in (False, True):
        it = _StyleIteration(model, encoder, "cuda:0", 0.01, use_graph=use_graph)
        it.start(target, torch.randn(1, 3, 224, 224, device="cuda:0"))
        for _ in range(5):
            it.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 100
        for _ in range(n):
            last = it.step()
        torch.cuda.synchronize()
        res["graph" if use_graph else "eager"] = {"ms_per_iteration": round((time.perf_counter() - t0) / n * 1e3, 3), "loss": last}
    out[f"layers={layers}"] = res
print(json.dumps({"workload": "style_transfer iteration, one 224x224 image (encoder fwd+bwd, dense Gram fwd+bwd, MSE, Adam)", **out}))
