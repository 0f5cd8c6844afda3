"""Pooled Gram backward with the generated gradient tile in tensor memory (gh_set_option gram_bwd_ats = 1; the MMAs read
their A operand from TMEM, csrc/gram_bwd_pair.cuh ATS) against the shared-memory form and the fp64 oracle, then timings of
both forms (kernel alone, CUDA events, inputs > L2).
    python tests/tools/check_bwd_ats.py > gpurun_out/bwd_ats.log
The cases run in a child process that is restarted after a failing case: a protocol bug ends in a trap (bounded mbarrier
waits), which poisons the CUDA context. (Under tests/: it evaluates the oracle.)"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = [  # B, C, H, W, g, dtype, channels_last
    (2, 512, 28, 28, 32, "f32", True), (2, 1024, 14, 14, 32, "f32", True), (2, 256, 8, 16, 32, "f32", True),
    (2, 512, 28, 28, 32, "bf16", True), (2, 1024, 14, 14, 32, "bf16", True), (3, 2048, 14, 14, 32, "f32", True),
    (2, 512, 28, 28, 32, "f32", False), (2, 1024, 14, 14, 32, "f32", False), (2, 512, 28, 28, 32, "bf16", False),
    (2, 1024, 10, 20, 32, "bf16", False), (2, 160, 10, 20, 20, "bf16", False), (2, 144, 10, 20, 18, "f32", False),
    (40, 512, 28, 28, 32, "f32", True), (300, 512, 8, 8, 32, "f32", True), (1, 2048, 7, 7, 32, "bf16", True),
    (2, 256, 8, 16, 32, "bf16", True),            # pooling factor 8 < UMMA_K = 16: a k-step holds two table values
    (75, 512, 28, 28, 32, "f32", True),           # 750 units on 74 pairs: uneven unit counts per pair
    (37, 1024, 14, 14, 32, "f32", True),          # 148 units: two per pair
]


def check(i: int) -> int:
    import torch
    from heuristique_style_transfer_code_b200 import _lib, ops
    from oracle import head_fp64 as O
    lib = _lib.lib()
    B, C, H, W, g, dtype, cl = CASES[i]
    torch.manual_seed(0)
    x = torch.relu(torch.randn(B, C, H, W, device="cuda"))
    if dtype == "bf16":
        x = x.bfloat16()
    x = x.contiguous(memory_format=torch.channels_last) if cl else x.reshape(B, C, H * W)
    dd = torch.randn(B, 2, g * g, device="cuda")
    xf = x.float().cpu().numpy().reshape(B, C, H * W)
    ref = O.gram_pool_backward(xf, g, dd[:, 1].cpu().numpy())
    tol = 1e-3 if dtype == "f32" else 6e-3

    def run(ats, nt=0, ch=1):
        assert lib.gh_set_option(b"gram_bwd_ats", ats) == 0 and lib.gh_set_option(b"gram_bwd_nt", nt) == 0
        assert lib.gh_set_option(b"gram_bwd_ch", ch) == 0
        df = ops.gram_pool_bwd(x, g, dd, 1)
        torch.cuda.synchronize()
        lib.gh_set_option(b"gram_bwd_ats", -1)            # the shipped defaults
        lib.gh_set_option(b"gram_bwd_nt", 0)
        lib.gh_set_option(b"gram_bwd_ch", 0)
        return df

    def err(t):
        return O.rel_err(t.float().cpu().numpy().reshape(ref.shape), ref)

    ss = run(0)
    outs = {"tmem-A": run(1), "x2": run(1, 0, 2)}
    if H * W >= 128:                                     # other x-tile widths: other A ring depths
        outs["x2 NT128"] = run(1, 128, 2)
        outs["x2 NT192"] = run(1, 192, 2)
        outs["x1 NT192"] = run(1, 192, 1)
    errs = {k: err(t) for k, t in outs.items()}
    diffs = {k: float((t.float() - ss.float()).norm() / ss.float().norm()) for k, t in outs.items()}
    ok = max(errs.values()) <= tol and max(diffs.values()) <= 1e-6
    print(f"case {CASES[i]}: smem-A err {err(ss):.2e}  " + "  ".join(f"{k} {errs[k]:.2e} (vs smem {diffs[k]:.0e})" for k in outs) +
          f"  bitwise={all(bool(torch.equal(ss, t)) for t in outs.values())}  {'OK' if ok else 'FAIL'}", flush=True)
    return 0 if ok else 1


def timeit(fn, n=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def timings() -> int:
    """Variants interleaved and repeated (boxes drift by several per cent within a run): best of four rounds of 12."""
    import torch
    from heuristique_style_transfer_code_b200 import _lib, ops
    lib = _lib.lib()
    g = 32
    for B in (256, 512):
        for dtype in (torch.float32, torch.bfloat16):
            for C, side in ((256, 56), (512, 28), (1024, 14)):
                x = torch.relu(torch.randn(B, C, side, side, device="cuda")).to(dtype).contiguous(memory_format=torch.channels_last)
                dd = torch.randn(B, 1, g * g, device="cuda")
                variants = [("smem-A", 0, 1, 0), ("tmem-A", 1, 1, 0), ("tmem-A x2", 1, 2, 0)]
                if C == 512:
                    variants += [("smem-A NT224", 0, 1, 224), ("tmem-A x2 NT192", 1, 2, 192), ("tmem-A x2 NT224", 1, 2, 224)]
                best = {}
                for _ in range(4):
                    for name, ats, ch, nt in variants:
                        lib.gh_set_option(b"gram_bwd_ats", ats)
                        lib.gh_set_option(b"gram_bwd_ch", ch)
                        lib.gh_set_option(b"gram_bwd_nt", nt)
                        t = timeit(lambda: ops.gram_pool_bwd(x, g, dd, 0), n=12, warm=2)
                        best[name] = min(best.get(name, 1e9), t)
                lib.gh_set_option(b"gram_bwd_ats", -1)
                lib.gh_set_option(b"gram_bwd_ch", 0)
                lib.gh_set_option(b"gram_bwd_nt", 0)
                print(f"B={B} {str(dtype)[6:]} C={C} HW={side * side}:  " + "  ".join(f"{k} {v:.1f}" for k, v in best.items()), flush=True)
                del x
    return 0


def main() -> int:
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        return timings()
    if len(sys.argv) > 1:
        bad = 0
        for i in range(int(sys.argv[1]), len(CASES)):
            bad += check(i)
            print(f"done {i}", flush=True)
        return 1 if bad else 0
    start, bad = 0, 0
    while start < len(CASES):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), str(start)], capture_output=True, text=True, timeout=900)
        sys.stdout.write("".join(l + "\n" for l in r.stdout.splitlines() if not l.startswith("done ")))
        done = [int(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("done ")]
        bad += sum("FAIL" in l for l in r.stdout.splitlines())
        nxt = (done[-1] + 1) if done else start
        if nxt < len(CASES):                       # the child died inside case nxt
            bad += 1
            print(f"case {CASES[nxt]}: child exit {r.returncode}\n{r.stderr[-1200:]}", flush=True)
            nxt += 1
        start = nxt
    print(f"{len(CASES) - bad}/{len(CASES)} cases OK", flush=True)
    sys.stdout.flush()
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "time"], capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout)
    if r.returncode != 0:
        print(f"timings: exit {r.returncode}\n{r.stderr[-1500:]}", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
