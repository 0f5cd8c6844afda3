"""Smallest-shape pass over every kernel family of the library, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tests/tools/sanitizer_smoke.py
    compute-sanitizer --tool racecheck python tests/tools/sanitizer_smoke.py      (one tool per GPU visit)

Each family runs once on a shape that exercises its tile edges (partial K blocks, rows beyond C, x tiles beyond HW) and
is checked against the fp64 oracle, so a clean sanitizer log also means correct results."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from heuristique_style_transfer_code_b200 import _lib, ops  # noqa: E402
from oracle import head_fp64 as O  # noqa: E402


def npf(t):
    return t.detach().float().cpu().numpy()


def main():
    torch.manual_seed(0)
    lib = _lib.lib()
    report = {}
    # Gram forward / backward, both kernel families, NCHW and NHWC, pooled and dense
    for path, flag in (("pair", 1), ("ldg", 0)):
        lib.gh_set_option(b"gram_fwd_pair", flag)
        lib.gh_set_option(b"gram_bwd_pair", flag)
        x = torch.relu(torch.randn(2, 256, 72, device="cuda")).requires_grad_(True)
        d = ops.style_descriptor([x], 32)
        w = torch.randn_like(d)
        (d * w).sum().backward()
        report[f"pool_fwd_{path}"] = O.rel_err(npf(d), O.descriptors([npf(x)], 32))
        report[f"pool_bwd_{path}"] = O.rel_err(npf(x.grad), O.gram_pool_backward(npf(x), 32, npf(w[:, 0])))
        y = torch.relu(torch.randn(1, 64, 8, 9, device="cuda")).requires_grad_(True)
        g = ops.gram_matrix(y)
        dg = torch.randn_like(g)
        (g * dg).sum().backward()
        report[f"dense_fwd_{path}"] = O.rel_err(npf(g), O.gram(npf(y).reshape(1, 64, 72)))
        report[f"dense_bwd_{path}"] = O.rel_err(npf(y.grad).reshape(1, 64, 72),
                                               O.gram_dense_backward(npf(y).reshape(1, 64, 72), npf(dg)))
    lib.gh_set_option(b"gram_fwd_pair", -1)
    lib.gh_set_option(b"gram_bwd_pair", -1)
    z = torch.relu(torch.randn(2, 512, 6, 6, device="cuda")).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    d = ops.style_descriptor([z], 32)
    w = torch.randn_like(d)
    (d * w).sum().backward()
    zf = npf(z).reshape(2, 512, 36)
    report["pool_fwd_nhwc"] = O.rel_err(npf(d), O.descriptors([zf], 32))
    report["pool_bwd_nhwc"] = O.rel_err(npf(z.grad).reshape(2, 512, 36), O.gram_pool_backward(zf, 32, npf(w[:, 0])))
    # attention head, both implementations
    names = ("in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias", "classifier_weight", "classifier_bias")
    for impl, (B, L, g, nc) in (("tma", (5, 3, 8, 4)), ("tma", (3, 2, 16, 3)), ("ldg", (5, 3, 8, 4))):
        ops.ATTN_IMPL = impl
        E = g * g
        desc = torch.randn(B, L, E, device="cuda")
        mha = torch.nn.MultiheadAttention(E, 1).cuda()
        lin = torch.nn.Linear(E, nc).cuda()
        ps = [mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, lin.weight, lin.bias]
        dd = desc.clone().requires_grad_(True)
        emb, logits = ops.attention_head(dd, *ps)
        labels = torch.arange(B, device="cuda") % nc
        torch.nn.functional.cross_entropy(logits, labels).backward()
        params = {k: npf(p) for k, p in zip(names, ps)}
        c = O.attention_forward(npf(desc), *[params[k] for k in names])
        _, dl = O.cross_entropy(c["logits"], labels.cpu().numpy())
        gr = O.attention_backward(c, params, dl, None)
        report[f"attn_fwd_{impl}_{E}"] = O.rel_err(npf(logits), c["logits"])
        report[f"attn_bwd_{impl}_{E}"] = max(O.rel_err(npf(dd.grad), gr["d_desc"]),
                                            O.rel_err(npf(ps[0].grad), gr["in_proj_weight"]))
    ops.ATTN_IMPL = "tma"
    # fused style loss
    t = torch.relu(torch.randn(1, 64, 8, 8, device="cuda"))
    with torch.no_grad():
        target = ops.gram_matrix(torch.relu(torch.randn(1, 64, 8, 8, device="cuda")))
    a = t.clone().requires_grad_(True)
    loss = ops.gram_mse_loss(a, target)
    loss.backward()
    want_loss, want_df = O.style_loss_and_grad(npf(t).reshape(1, 64, 64), npf(target))
    report["style_loss"] = abs(loss.item() - want_loss) / want_loss
    report["style_grad"] = O.rel_err(npf(a.grad).reshape(1, 64, 64), want_df)
    # transpose / max pool / stem staging / PatchGAN head kernels
    u = torch.randn(2, 12, 20, device="cuda")
    report["transpose"] = float((ops.nhwc_to_nchw(torch.randn(2, 16, 3, 5, device="cuda").contiguous(
        memory_format=torch.channels_last)).abs().sum() >= 0))
    mp = torch.randn(2, 8, 9, 9, device="cuda").contiguous(memory_format=torch.channels_last)
    report["maxpool"] = float(torch.equal(ops.maxpool2d_nhwc(mp, 3, 2, 1), torch.nn.functional.max_pool2d(mp, 3, 2, 1)))
    ops.stem_space_to_depth(torch.randn(1, 3, 16, 16, device="cuda"), torch.float32)
    gram, norms = ops.patch_gram([torch.randn(2, 16, 9, 9, device="cuda"), torch.randn(2, 16, 4, 4, device="cuda")])
    report["patch_gram_finite"] = float(torch.isfinite(gram).all() and torch.isfinite(norms).all())
    del u
    torch.cuda.synchronize()
    bad = {k: v for k, v in report.items() if not (v <= 6e-3 or k in ("transpose", "maxpool", "patch_gram_finite"))}
    for k, v in report.items():
        print(f"{k}: {v:.3e}")
    print("device error record:", _lib.last_device_error())
    if bad:
        raise SystemExit(f"out of tolerance: {bad}")
    print("sanitizer smoke: OK")


if __name__ == "__main__":
    main()
