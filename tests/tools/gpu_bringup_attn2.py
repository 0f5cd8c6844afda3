"""GPU bring-up of the TMA-fed attention head (tgemm_pair.cuh / attn_head2.cuh): each case runs in its own process
(a device trap poisons the CUDA context) and prints normwise errors against fp64.
Usage on a B200:  python tests/tools/gpu_bringup_attn2.py [case ...]   (default: all)"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _np(t):
    return t.detach().float().cpu().numpy()


def _rel(a, b):
    import torch
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@case
def split():
    import torch
    from heuristique_style_transfer_code_b200 import ops
    x = torch.randn(1000, 64, device="cuda") * 37.0
    p = ops.split_bf16(x)
    torch.cuda.synchronize()
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    print("split: hi exact", bool((p[0] == hi).all()), "lo exact", bool((p[1] == lo).all()),
          "recon rel", _rel(p[0].float() + p[1].float(), x))


def _gemm_case(M, N, K, a_mn, b_mn, mode, max_split=64):
    import torch
    from heuristique_style_transfer_code_b200 import ops
    torch.manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda")
    Bm = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    ap = ops.split_bf16(A.t().contiguous() if a_mn else A)
    bp = ops.split_bf16(Bm.t().contiguous() if b_mn else Bm)
    ref = A.double() @ Bm.double().t() + bias.double()
    out = ops.gemm_planes(ap, a_mn, bp, b_mn, bias, planes_out=(mode == "planes"), max_split=max_split)
    torch.cuda.synchronize()
    got = out[0].float() + out[1].float() if mode == "planes" else out
    err = _rel(got, ref)
    flag = "" if err < 2e-5 else "   <-- FAIL"
    print(f"gemm M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} {mode} split<={max_split}: rel={err:.3e}{flag}")
    if err >= 2e-5:
        d = (got.double() - ref).abs()
        bm, bn = 32, 32
        rows = [(i, float(d[i:i + bm].max())) for i in range(0, M, bm)]
        cols = [(j, float(d[:, j:j + bn].max())) for j in range(0, N, bn)]
        print("   worst row blocks:", sorted(rows, key=lambda t: -t[1])[:6])
        print("   worst col blocks:", sorted(cols, key=lambda t: -t[1])[:6])
        print("   ref max", float(ref.abs().max()), "got[0,:4]", got[0, :4].tolist(), "ref[0,:4]", ref[0, :4].tolist())


@case
def gemm_kk():
    for (M, N, K) in [(256, 256, 64), (256, 256, 256), (100, 192, 64), (1536, 3072, 1024), (15, 192, 64), (512, 1024, 1024)]:
        _gemm_case(M, N, K, False, False, "f32", 1)
    _gemm_case(768, 3072, 1024, False, False, "f32", 2)
    _gemm_case(512, 1024, 1024, False, False, "f32", 64)


@case
def gemm_planes_out():
    for (M, N, K) in [(256, 256, 64), (99, 64, 128), (512, 1024, 1024)]:
        _gemm_case(M, N, K, False, False, "planes", 1)


@case
def gemm_kmn():
    for (M, N, K) in [(256, 256, 64), (100, 192, 320), (1536, 1024, 3072), (512, 1024, 1024)]:
        _gemm_case(M, N, K, False, True, "f32", 1)
    _gemm_case(1536, 1024, 3072, False, True, "f32", 64)


@case
def gemm_mnmn():
    for (M, N, K) in [(256, 256, 64), (192, 64, 100), (3072, 1024, 1536), (1024, 1024, 512), (64, 64, 15)]:
        _gemm_case(M, N, K, True, True, "f32", 1)
    _gemm_case(3072, 1024, 1536, True, True, "f32", 64)


def _attn_case(B, L, g, nc, time_it=False):
    import torch
    from heuristique_style_transfer_code_b200 import ops
    from oracle import head_fp64 as O
    E = g * g
    torch.manual_seed(0)
    desc = torch.randn(B, L, E, device="cuda") * 2.0
    mha = torch.nn.MultiheadAttention(E, 1).cuda()
    lin = torch.nn.Linear(E, nc).cuda()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.1)
        mha.out_proj.bias.normal_(0, 0.1)
    ps = [mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, lin.weight, lin.bias]
    names = ("in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias", "classifier_weight", "classifier_bias")
    labels = torch.arange(B, device="cuda") % nc
    w = torch.randn(B, E, device="cuda") * 0.01
    res = {}
    for impl in ("tma", "ldg"):
        ops.ATTN_IMPL = impl
        for p in ps:
            p.grad = None
        d = desc.clone().requires_grad_(True)
        emb, logits = ops.attention_head(d, *ps)
        (torch.nn.functional.cross_entropy(logits, labels) + (emb * w).sum()).backward()
        torch.cuda.synchronize()
        res[impl] = dict(emb=emb.detach(), logits=logits.detach(), d_desc=d.grad.detach(),
                         **{k: p.grad.detach().clone() for k, p in zip(names, ps)})
    if B * L * E * E <= 40e9 / 50:
        params = {k: _np(p) for k, p in zip(names, ps)}
        c = O.attention_forward(_np(desc), *[params[k] for k in names])
        _, dl = O.cross_entropy(c["logits"], labels.cpu().numpy())
        gr = O.attention_backward(c, params, dl, _np(w))
        ref = dict(emb=c["emb"], logits=c["logits"], d_desc=gr["d_desc"], **{k: gr[k] for k in names})
        for impl in ("tma", "ldg"):
            errs = {k: O.rel_err(_np(v), ref[k]) for k, v in res[impl].items()}
            bad = [k for k, v in errs.items() if not v <= 2e-5]
            print(f"attn B={B} L={L} E={E} nc={nc} {impl}: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()) +
                  ("   <-- FAIL " + ",".join(bad) if bad else ""))
    else:
        errs = {k: _rel(res["tma"][k], res["ldg"][k]) for k in res["tma"]}
        print(f"attn B={B} L={L} E={E} nc={nc} tma vs ldg: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    if time_it:
        for impl in ("tma", "ldg"):
            ops.ATTN_IMPL = impl
            for _ in range(3):
                d = desc.clone().requires_grad_(True)
                emb, logits = ops.attention_head(d, *ps)
                (logits.sum() + (emb * w).sum()).backward()
            ops.PROFILE = []
            for _ in range(10):
                d = desc.clone().requires_grad_(True)
                torch.cuda._sleep(3_000_000)          # ~1.5 ms of GPU work: the host runs ahead, events see GPU time only
                emb, logits = ops.attention_head(d, *ps)
                torch.cuda._sleep(3_000_000)
                (logits.sum() + (emb * w).sum()).backward()
            torch.cuda.synchronize()
            acc = {}
            for name, work, s, e in ops.PROFILE:
                acc.setdefault(name, []).append(s.elapsed_time(e) * 1e3)
            ops.PROFILE = None
            print(f"  timing {impl}: " + "  ".join(f"{k}: {sorted(v)[len(v) // 2]:.1f} us" for k, v in acc.items()))


@case
def attn_small():
    for cfg in [(5, 3, 8, 4), (4, 1, 16, 3), (7, 4, 8, 10), (1, 3, 32, 4), (33, 3, 32, 4)]:
        _attn_case(*cfg)


@case
def attn_tn():
    """tile-width planning of tgemm_pair: forced 256, forced 128, planned"""
    from heuristique_style_transfer_code_b200 import _lib
    for tn in (256, 128, 0):
        _lib.lib().gh_set_option(b"tgemm_tn", tn)
        print("tgemm_tn =", tn)
        _attn_case(512, 3, 32, 4, time_it=True)
        _attn_case(256, 3, 32, 4, time_it=True)
    _lib.lib().gh_set_option(b"tgemm_tn", 0)


@case
def attn_512():
    _attn_case(512, 3, 32, 4, time_it=True)


@case
def attn_big():
    _attn_case(64, 3, 32, 4, time_it=True)
    _attn_case(256, 3, 32, 4, time_it=True)
    _attn_case(512, 3, 32, 4, time_it=True)


def main():
    names = sys.argv[1:] or list(CASES)
    if len(names) == 1 and os.environ.get("GH_BRINGUP_CHILD") == "1":
        CASES[names[0]]()
        return
    for n in names:
        t0 = time.time()
        env = dict(os.environ, GH_BRINGUP_CHILD="1")
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), n], env=env, capture_output=True, text=True,
                               timeout=240)
            print(f"=== {n} (exit {r.returncode}, {time.time() - t0:.1f}s)")
            print(r.stdout.rstrip())
            if r.returncode != 0:
                print(r.stderr[-3000:])
        except subprocess.TimeoutExpired as e:
            print(f"=== {n} TIMEOUT")
            print((e.stdout or b"").decode()[-2000:] if isinstance(e.stdout, bytes) else (e.stdout or "")[-2000:])
        sys.stdout.flush()


if __name__ == "__main__":
    main()
