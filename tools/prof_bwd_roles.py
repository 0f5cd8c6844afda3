"""Where each warp role of gram_bwd_pair_kernel spends its cycles (a -DGH_BP_PROFILE build of the library: clock64 around
every mbarrier wait of one thread per role, csrc/gram_bwd_pair.cuh).
    python tools/prof_bwd_roles.py [--build] > gpurun_out/bwd_roles.log
Build the instrumented library first (here, without a GPU):  python tools/prof_bwd_roles.py --build"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "heuristique_style_transfer_code_b200", "csrc")
PROF_LIB = os.path.join(CSRC, "libgramhead_prof.so")


def build() -> None:
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-DGH_BP_PROFILE", "gramhead.cu", "-o", PROF_LIB]
    subprocess.run(cmd, cwd=CSRC, check=True)
    print("built", PROF_LIB)


def main() -> None:
    if "--build" in sys.argv:
        build()
        return
    os.environ["GRAMHEAD_LIB"] = PROF_LIB
    sys.path.insert(0, ROOT)
    import torch
    from heuristique_style_transfer_code_b200 import _lib, ops
    lib = _lib.lib()
    lib.gh_bp_profile_read.restype = ctypes.c_int
    lib.gh_bp_profile_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    buf = (ctypes.c_ulonglong * 16)()
    g, B, reps = 32, 256, 5
    print("per pair and per ring stage (1 or 2 K chunks), in SM cycles (loop = the role's whole main loop / chunks; waits are part of it)")
    for dtype in (torch.float32, torch.bfloat16):
        for C, side in ((256, 56), (512, 28), (1024, 14)):
            x = torch.relu(torch.randn(B, C, side, side, device="cuda")).to(dtype).contiguous(memory_format=torch.channels_last)
            dd = torch.randn(B, 1, g * g, device="cuda")
            for ats, nt, ch in ((0, 0, 1), (1, 0, 1), (1, 0, 2)):
                lib.gh_set_option(b"gram_bwd_ats", ats)
                lib.gh_set_option(b"gram_bwd_nt", nt)
                lib.gh_set_option(b"gram_bwd_ch", ch)
                ops.gram_pool_bwd(x, g, dd, 0)
                assert lib.gh_bp_profile_read(buf) == 0
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    ops.gram_pool_bwd(x, g, dd, 0)
                b.record()
                torch.cuda.synchronize()
                assert lib.gh_bp_profile_read(buf) == 0
                v = [float(t) for t in buf]
                pairs, chunks = max(v[12], 1.0), max(v[4], 1.0)
                per = lambda i: v[i] / chunks                       # noqa: E731
                print(f"{str(dtype)[6:]} C={C} HW={side * side} {'tmem-A' if ats else 'smem-A'}{f' x{ch}' if ch > 1 else ''} NT={nt or 'plan'}: "
                      f"{a.elapsed_time(b) / reps * 1e3:.1f} us  chunks/pair {chunks / pairs:.0f}  "
                      f"issuer loop {per(0):.0f} (tmem_empty {per(1):.0f}, fullA {per(2):.0f}, fullB {per(3):.0f})  "
                      f"epilogue loop {per(5):.0f} (tmem_full {per(6):.0f}, staging {per(7):.0f})  "
                      f"generator loop {per(8):.0f} (emptyA {per(9):.0f})  producer loop {per(10):.0f} (emptyB {per(11):.0f})",
                      flush=True)
            lib.gh_set_option(b"gram_bwd_ats", -1)
            lib.gh_set_option(b"gram_bwd_nt", 0)
            lib.gh_set_option(b"gram_bwd_ch", 0)
            del x


if __name__ == "__main__":
    main()
