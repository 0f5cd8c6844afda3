"""SASS opcode census of csrc/libgramhead.so, per kernel (runs where cuobjdump is: no GPU needed):

    python tools/sass_census.py > profiles/r3_sass_census.txt

Counts the mnemonics that identify the Blackwell paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / .st, UTMALDG / UTMASTG / UTMAREDG = TMA loads / stores / reduce-adds, UBLKCP = bulk copies, HMMA = legacy
mma.sync, BRA.U.ANY = the elect-broadcast retry loop ptxas emits around single-thread instructions guarded by a lane test
(must be 0), LDG / STG / RED / ATOM for the ld.global-fed and atomics paths."""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "heuristique_style_transfer_code_b200", "csrc", "libgramhead.so")
PATTERNS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
            ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTMAREDG", r"\bUTMAREDG"), ("UBLKCP", r"\bUBLKCP"),
            ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("BRA.U.ANY", r"BRA\.U\.ANY"), ("LDG", r"\bLDG"), ("STG", r"\bSTG"),
            ("RED", r"\bRED\b|\bREDG"), ("ATOM", r"\bATOM"), ("STS", r"\bSTS"), ("LDS", r"\bLDS"), ("FFMA", r"\bFFMA")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        kernels[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
        for name, pat in PATTERNS:
            if re.search(pat, line):
                kernels[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    cols = [n for n, _ in PATTERNS]
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a)\n")
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for (name, c), pretty in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", pretty).replace("void ", "").replace("gh::", "")
        print(f"| {short[:70]} | {c['instructions']} | " + " | ".join(str(c[n]) if c[n] else "." for n in cols) + " |")
    total_any = sum(c["BRA.U.ANY"] for c in kernels.values())
    print(f"\nBRA.U.ANY in the whole library: {total_any}")


if __name__ == "__main__":
    main()
