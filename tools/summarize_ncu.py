"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/<tag>_launches.csv, <tag>_*.ncu-rep) into the small
text/JSON summaries kept under profiles/.   usage: python tools/summarize_ncu.py <tag> [<rep-name> ...]"""
from __future__ import annotations

import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def launch_shares(tag):
    path = os.path.join(OUT, f"{tag}_launches.csv")
    if not os.path.isfile(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[h0], rows[h0 + 1:]
    ki, vi, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in data:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki]))
        tot[name] += float(r[vi].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    ours = sum(v for k, v in tot.items() if k.startswith("gh::"))
    lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu` ({tag})",
             f"# launches captured: {sum(cnt.values())}; summed kernel time {total / 1e6:.2f} ms (cold-cache, serialised: compare SHARES)",
             f"# share of this library's kernels (gh::*): {ours / total * 100:.2f} %", "share%      total_us   launches  kernel"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
        lines.append(f"{v / total * 100:6.2f} {v / 1e3:13.1f} {cnt[k]:8d}   {k[:140]}")
    open(os.path.join(PROF, f"{tag}_launch_shares.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


ALGO = {}


def rep_summary(tag, name):
    rep = os.path.join(OUT, f"{tag}_{name}.ncu-rep")
    exported = os.path.join(OUT, f"{tag}_{name}_raw.csv")          # the visit scripts export on the box and drop the .ncu-rep
    if os.path.isfile(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    elif os.path.isfile(exported):
        raw = open(exported).read()
    else:
        return {}
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full --clock-control none --import-source on ({tag}_{name}.ncu-rep), one block per captured launch"]
    traffic = {}
    for r in rows[2:]:
        kname = r[idx["Kernel Name"]]
        lines.append(f"\n== {kname}   [{name}]")
        for k in RAW_KEYS:
            if k in idx:
                lines.append(f"{k:85s} {r[idx[k]]:>18s} {units[idx[k]]}")
        try:
            def val(key):
                v = float(r[idx[key]].replace(",", ""))
                u = units[idx[key]].lower()
                return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            traffic[kname] = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
            if name.startswith("pool_"):                       # tools/prof_one.py pool: (256, 64, 112, 112) channels_last
                dt = "torch.float32" if name.endswith("f32") else "torch.bfloat16"
                key = f"maxpool2d_nhwc[C=64,HW=112x112,{dt}]"
                traffic[key] = traffic[kname]
                ALGO[key] = int(256 * 64 * 112 * 112 * 1.25 * (4 if name.endswith("f32") else 2))
            m = re.match(r"(fwd|bwd)_(\d+)_(\d+)_(f32|bf16)_(nchw|nhwc)", name)     # capture names carry the bench's key
            if m:
                dt = "torch.float32" if m.group(4) == "f32" else "torch.bfloat16"
                tail = ",nhwc" if m.group(5) == "nhwc" else ""
                key = f"gram_pool_{m.group(1)}[C={m.group(2)},HW={m.group(3)},{dt}{tail}]"
                traffic[key] = traffic[kname]
                # algorithmic bytes of the captured launch (tools/prof_one.py: batch 256, g 32), so that bench.py can
                # scale the measured traffic to launches of another batch size (every image is streamed once)
                c, hw, sz = int(m.group(2)), int(m.group(3)), (4 if m.group(4) == "f32" else 2)
                per_image = c * hw * sz + 32 * 32 * 4 + (c * hw * 4 if m.group(1) == "bwd" else 0)
                ALGO[key] = 256 * per_image
                traffic[f"{kname} [C={m.group(2)},HW={m.group(3)},{m.group(5)}]"] = traffic.pop(kname)
        except Exception:
            pass
    open(os.path.join(PROF, f"{tag}_{name}_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
    return traffic


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    launch_shares(tag)
    traffic = {}
    for name in sys.argv[2:]:
        traffic.update(rep_summary(tag, name))
    if traffic:
        # map SASS kernel names to the names bench.py's roofline uses
        mapped = {}
        for k, v in traffic.items():
            m = re.search(r"gram_fwd_kernel<(\d+), (\d+), (\d+)>", k)
            if m:
                kp = int(m.group(2))
                c = kp * 32
                hw = {256: 3136, 512: 784, 1024: 196, 2048: 49}.get(c)
                dt = "torch.float32" if int(m.group(1)) < 2 else "torch.bfloat16"
                mapped[f"gram_pool_fwd[C={c},HW={hw},{dt}]"] = v
            mapped[k] = v
        path = os.path.join(PROF, "ncu_traffic.json")
        old = json.load(open(path)) if os.path.isfile(path) else {}
        old.update(mapped)
        old.setdefault("_algorithmic_bytes_of_captured_launch", {}).update(ALGO)
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)
        print("traffic:", mapped)


if __name__ == "__main__":
    main()
