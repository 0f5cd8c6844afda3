"""Runs one of the reference's CLI scripts UNCHANGED against this repo's drop-in `Models` / `functions` packages:

    python tools/run_ref_script.py /path/to/reference/train_best_RESNET50_Truncate_gram_attention.py --data D --config_path J ...

The script file is executed with runpy from wherever it lives; only sys.path is arranged so that its
`from Models....` / `from functions....` imports resolve to this repository. Offline helpers:
  --gh-seed-hub   pre-seed $TORCH_HOME/hub/checkpoints/resnet50-0676ba61.pth with a random-init state_dict, because the
                  scripts hard-code models.resnet50(weights=IMAGENET1K_V1) (train:69, test:76) and there is no network.
"""
from __future__ import annotations

import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def seed_hub_checkpoint():
    import torch
    from torchvision import models
    hub = os.path.join(torch.hub.get_dir(), "checkpoints")
    os.makedirs(hub, exist_ok=True)
    path = os.path.join(hub, "resnet50-0676ba61.pth")
    if not os.path.isfile(path):
        torch.manual_seed(0)
        torch.save(models.resnet50(weights=None).state_dict(), path)
    return path


def run():
    args = sys.argv[1:]
    if "--gh-seed-hub" in args:
        args.remove("--gh-seed-hub")
        print("seeded", seed_hub_checkpoint())
    if not args:
        raise SystemExit(__doc__)
    script = os.path.abspath(args[0])
    sys.argv = [script] + args[1:]
    sys.path.insert(0, ROOT)
    # import the drop-in packages first: runpy prepends the script's directory (the reference tree, which has its own
    # Models/ and functions/ namespace dirs) to sys.path, but modules already in sys.modules are not looked up again.
    import Models.Models_RESNET50_TRUNCATE_GRAM_with_Attention  # noqa: F401
    import Models.Models_Multi_PatchGAN  # noqa: F401
    import functions.functions_RESNET50_Truncate_Gram_Attention  # noqa: F401
    # modules this repository does not provide (e.g. functions.functions_Multi_PatchGAN, which the Multi-PatchGAN scripts
    # import next to the model classes) still resolve to the reference's files: Models/ and functions/ are namespace
    # packages in both trees, so Python merges the two directories.
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    run()
