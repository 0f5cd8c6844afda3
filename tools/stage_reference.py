"""Stages the reference's files for the Gram+attention path under the git-ignored baseline/_ref/ so they travel to the
GPU box with a gpurun snapshot (the box has no /root/reference). Nothing under baseline/_ref is ever committed.
    python tools/stage_reference.py [/root/reference]
Only used by tests/test_gpu_integration.py (reference CLI scripts run UNCHANGED against the drop-in modules)."""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["train_best_RESNET50_Truncate_gram_attention.py", "test_RESNET50_Truncate_gram_attention.py"]


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    dst = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(src):
        raise SystemExit(f"{src} not found")
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copy2(os.path.join(src, f), os.path.join(dst, f))
        print("staged", f)


if __name__ == "__main__":
    main()
