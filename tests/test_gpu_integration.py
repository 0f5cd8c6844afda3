"""The reference's own CLI scripts, byte-for-byte unchanged (staged under the git-ignored baseline/_ref/ by
tools/stage_reference.py), driven end to end on a B200 against this repository's drop-in `Models` / `functions`:
train (2 folds, SGD as the script prescribes) -> checkpoint -> test --mode classification -> --mode style_transfer."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "baseline", "_ref")
TRAIN = os.path.join(REF, "train_best_RESNET50_Truncate_gram_attention.py")
TEST = os.path.join(REF, "test_RESNET50_Truncate_gram_attention.py")
CLASSES = ["fog", "rain", "snow", "sun"]


@pytest.fixture(scope="module")
def workspace(tmp_path_factory):
    if not (os.path.isfile(TRAIN) and os.path.isfile(TEST)):
        pytest.skip("reference scripts not staged (run tools/stage_reference.py where /root/reference exists)")
    from PIL import Image
    root = tmp_path_factory.mktemp("ws")
    rng = np.random.default_rng(0)
    for split, n in (("train", 4), ("test", 2)):
        for ci, cls in enumerate(CLASSES):
            d = root / "data" / split / cls
            d.mkdir(parents=True)
            for i in range(n):
                img = (rng.random((80, 96, 3)) * 255).astype(np.uint8)
                img[:, :, ci % 3] //= 2
                Image.fromarray(img).save(d / f"{cls}_{i}.png")
    cfg = dict(hidden_dims=[128], num_layers=1, batch_size=4, lr=1e-3, truncate_layer=7, gram_matrix_size=32)
    (root / "cfg.json").write_text(json.dumps(cfg))
    env = dict(os.environ, TORCH_HOME=str(root / "torch_home"), PYTHONUNBUFFERED="1")
    return root, env


def run(script, args, env, cwd):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_ref_script.py"), "--gh-seed-hub", script] + args
    res = subprocess.run(cmd, env=env, cwd=cwd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + "\n" + res.stderr[-3000:]
    return res.stdout


def test_reference_train_then_test_scripts_run_unchanged(workspace):
    root, env = workspace
    save = root / "saved"
    out = run(TRAIN, ["--data", str(root / "data"), "--config_path", str(root / "cfg.json"), "--epochs", "1",
                      "--k_folds", "2", "--save_dir", str(save)], env, str(root))
    assert "Fold 1, Validation Loss" in out
    ckpt = save / "best_model_fold_1.pth"
    assert ckpt.is_file()
    hp = json.loads((save / "best_hyperparameters_fold_1.json").read_text())
    assert hp["gram_matrix_size"] == 32 and hp["truncate_layer"] == 7
    import torch
    blob = torch.load(str(ckpt), map_location="cpu")
    assert sorted(blob) == ["attention", "classifier", "truncated_encoder"]
    assert blob["attention"]["in_proj_weight"].shape == (3072, 1024)

    res = root / "results"
    run(TEST, ["--data", str(root / "data"), "--model_path", str(ckpt), "--config_path", str(root / "cfg.json"),
               "--mode", "classification", "--save_dir", str(res), "--afficher_params"], env, str(root))
    metrics = json.loads((res / "classification_results.json").read_text())
    assert set(metrics) >= {"precision", "recall", "f1_score"}

    # frozen-encoder training (only the head's parameters get gradients: Gram backward is skipped entirely)
    run(TRAIN, ["--data", str(root / "data"), "--config_path", str(root / "cfg.json"), "--epochs", "1", "--k_folds", "2",
                "--save_dir", str(root / "saved_frozen"), "--freeze_layers"], env, str(root))

    # style-transfer mode drives model.gram_matrix() forward + backward (dense Gram kernels), default --layers 4 -> C = 64
    run(TEST, ["--data", str(root / "data"), "--model_path", str(ckpt), "--config_path", str(root / "cfg.json"),
               "--mode", "style_transfer", "--save_dir", str(res), "--num_iterations", "3", "--num_samples", "2"],
        env, str(root))
    pngs = [p for p in res.rglob("style_transfer_*.png")]
    assert len(pngs) == 2
