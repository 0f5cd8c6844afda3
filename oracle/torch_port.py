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


def patchgan_forward(model, x: torch.Tensor):
    """fp32 torch restatement of VariablePatchesNLayerDiscriminator_test.forward (Models/Models_Multi_PatchGAN.py:177-258),
    op for op, on the submodules of `model` (any module with the reference's attribute names: the reference class or this
    repo's drop-in), on whatever device they live. Returns (embeddings, output, gram_norms list)."""
    collected, k = [], 0
    for idx, layer in enumerate(model.feature_extractor):
        x = layer(x)
        if torch.isnan(x).any():
            print(f"NaN detected after layer {idx}")
            x = torch.nan_to_num(x, nan=0.0)
        if isinstance(layer, nn.Conv2d):
            proj = model.projection_layers[k](x)
            k += 1
            if torch.isnan(proj).any():
                print(f"NaN detected in projected feature map at layer {idx}")
                proj = torch.nan_to_num(proj, nan=0.0)
            collected.append(F.layer_norm(proj, proj.shape[1:]))
    tokens, norms = [], []
    for i, fm in enumerate(collected):
        pooled = F.adaptive_avg_pool2d(fm, output_size=(4, 4))
        pooled = F.layer_norm(pooled, pooled.shape[1:])
        flat = pooled.view(pooled.size(0), model.gram_matrix_dim, -1)
        gram = torch.bmm(flat, flat.transpose(1, 2)) / (flat.size(-1) + 1e-6)
        norms.append(torch.norm(gram, p='fro', dim=(1, 2)))
        tok = model.feature_projection(gram.view(gram.size(0), -1))
        if torch.isnan(tok).any():
            print(f"NaN detected in projected features at layer {i}, replacing NaNs with zeros.")
            tok = torch.nan_to_num(tok, nan=0.0)
        tokens.append(tok)
    seq = torch.stack(tokens, dim=0)
    seq, _ = model.attention_per_layer(seq, seq, seq)
    seq, _ = model.attention_per_patch(seq, seq, seq)
    emb = seq.mean(dim=0)
    return emb, model.classifier(emb), norms


def patchgan_multiscale_forward(model, x: torch.Tensor):
    """MultiScaleDiscriminator_test.forward (:299-312): every scale sees the same input; plain means over the scales."""
    embs, outs = zip(*[patchgan_forward(d, x)[:2] for d in model.scale_discriminators.values()])
    return torch.stack(embs, 0).mean(0), torch.stack(outs, 0).mean(0)
