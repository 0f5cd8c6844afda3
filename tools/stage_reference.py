"""Stages the reference's files for the Gram+attention path under the git-ignored baseline/_ref/ so they travel to the
GPU box with a gpurun snapshot (the box has no /root/reference). Nothing under baseline/_ref is ever committed.
    python tools/stage_reference.py [/root/reference]
Used by tests/test_gpu_integration.py (reference CLI scripts run UNCHANGED against the drop-in modules), by
bench.py's reference arm and CPU / same-GPU baselines (the UNMODIFIED reference classes, loaded by file path), and by
the delegated out-of-scope entry points (heuristique_style_transfer_code_b200/_reference.py)."""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["train_best_RESNET50_Truncate_gram_attention.py", "test_RESNET50_Truncate_gram_attention.py",
         os.path.join("Models", "Models_RESNET50_TRUNCATE_GRAM_with_Attention.py"),
         os.path.join("functions", "functions_RESNET50_Truncate_Gram_Attention.py"),
         os.path.join("Models", "Models_Multi_PatchGAN.py"),
         os.path.join("functions", "functions_Multi_PatchGAN.py"),
         "test_Multi_PatchGAN.py"]


def stage(src: str = "/root/reference", verbose: bool = True) -> bool:
    dst = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(src):
        return False
    for f in FILES:
        if not os.path.isfile(os.path.join(src, f)):
            continue
        os.makedirs(os.path.dirname(os.path.join(dst, f)), exist_ok=True)
        shutil.copy2(os.path.join(src, f), os.path.join(dst, f))
        if verbose:
            print("staged", f)
    return True


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    if not stage(src):
        raise SystemExit(f"{src} not found")


if __name__ == "__main__":
    main()
