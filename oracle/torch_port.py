"""fp32 torch-CPU port of the reference's head, op for op  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This is the "reference CPU path" timed beside the B200 numbers (bench.py cpu_baseline / --impl reference): it issues
the same library calls, in the same order and dtype, as the reference does on a CPU device, so MKL/oneDNN do the same
work. It exists because /root/reference is not present on the GPU box. tests/test_oracle_golden.py checks it against
the unmodified reference bit for bit (same torch build => same kernels).

  head()        Models/Models_RESNET50_TRUNCATE_GRAM_with_Attention.py:26-30,51-61  (train class) / :78-82,103-114
  PortModel     the module around it (:13-24, :32-49): stem, per-stage loop, early-out with no stages
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def stage_gram(x: torch.Tensor) -> torch.Tensor:
    n, c, h, w = x.size()
    flat = x.view(n, c, h * w)
    return torch.bmm(flat, flat.transpose(1, 2)).div(h * w)


def head(stage_outputs, g: int, attention: nn.MultiheadAttention, classifier: nn.Linear):
    pooled = [F.adaptive_avg_pool2d(stage_gram(s), (g, g)) for s in stage_outputs]
    tokens = torch.stack(pooled, dim=1).flatten(2).permute(1, 0, 2)     # (L, B, g*g)
    mixed, _ = attention(tokens, tokens, tokens)
    emb = mixed.mean(dim=0)
    emb = emb.view(emb.size(0), -1)
    return emb, classifier(emb)


class PortModel(nn.Module):
    """Same constructor / parameters / forward contract as the reference classes; `return_embeddings` selects the
    `_for_test` return convention."""

    def __init__(self, base_encoder, truncate_after_layer, num_classes, gram_matrix_size, device="cpu",
                 return_embeddings=False):
        super().__init__()
        self.device = device
        self.truncated_encoder = nn.Sequential(*list(base_encoder.children())[:truncate_after_layer]).to(device)
        self.num_classes = num_classes
        self.gram_matrix_size = gram_matrix_size
        self.classifier = nn.Linear(gram_matrix_size ** 2, num_classes).to(device)
        self.attention = nn.MultiheadAttention(embed_dim=gram_matrix_size ** 2, num_heads=1).to(device)
        self.return_embeddings = return_embeddings

    def stage_outputs(self, x):
        x = x.to(self.device)
        for i in range(4):
            x = self.truncated_encoder[i](x)
        outs = []
        for block in self.truncated_encoder[4:]:
            x = block(x)
            outs.append(x)
        return x, outs

    def forward(self, x):
        x, outs = self.stage_outputs(x)
        if not outs:
            return torch.zeros((x.size(0), self.num_classes), requires_grad=True).to(self.device)
        emb, logits = head(outs, self.gram_matrix_size, self.attention, self.classifier)
        return (emb, logits) if self.return_embeddings else logits
