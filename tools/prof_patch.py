"""Single-kernel driver for ncu captures of the Multi-PatchGAN head kernels (batch 256, the six maps of a patch-70
discriminator at 224x224).   python tools/prof_patch.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from heuristique_style_transfer_code_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
maps = [torch.randn(B, 64, s, s, device="cuda") for s in (112, 56, 28, 14, 13, 12)]
a1 = torch.nn.MultiheadAttention(64, 8).cuda()
a2 = torch.nn.MultiheadAttention(64, 8).cuda()
cl = torch.nn.Linear(64, 4).cuda()
fp = torch.nn.Linear(4096, 64).cuda()
for _ in range(4):
    gram, norms = ops.patch_gram(maps)
    feat = ops.gemm_f32(gram.view(-1, 4096), fp.weight.detach().t(), fp.bias.detach()).view(6, B, 64)
    ops.patch_attention(feat, a1, a2, cl)
torch.cuda.synchronize()
print("done", B)
