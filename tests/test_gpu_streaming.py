"""GPU: camera-mode preprocessing kernel and the CUDA-graph pipeline (streaming.py) against the oracle restatement of
the reference's host-side preprocessing (oracle/pil_resize.py, itself pinned bitwise against Pillow / torchvision)."""
import numpy as np
import pytest
import torch

from oracle import pil_resize as P

pytestmark = pytest.mark.gpu

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


@pytest.fixture(scope="module")
def model():
    from torchvision import models
    from heuristique_style_transfer_code_b200 import TruncatedResNet50_for_test
    torch.manual_seed(0)
    m = TruncatedResNet50_for_test(models.resnet50(weights=None), 7, 4, 32, device="cuda:0")
    return m.eval()


@pytest.mark.parametrize("H,W,resize,crop", [(1080, 1920, (448, 448), None), (1080, 1920, 256, 224), (480, 640, 256, 224),
                                             (300, 200, (224, 224), None), (100, 160, (224, 224), None)])
def test_preprocess_kernel_is_bitwise_the_reference(model, H, W, resize, crop):
    from heuristique_style_transfer_code_b200.streaming import CameraPipeline
    frame = np.random.default_rng(H * 7 + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    pipe = CameraPipeline(model, frame.shape, resize, crop, MEAN, STD, bgr=True, use_graph=False)
    pipe._frame_dev.copy_(torch.from_numpy(frame))
    got = pipe.preprocess_().cpu().numpy()[0]
    ref = P.camera_preprocess(frame, resize, crop, MEAN, STD)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)          # integer resample + IEEE fp32 normalisation: bit exact


def test_pipeline_matches_the_host_path(model):
    from PIL import Image
    from torchvision import transforms
    from heuristique_style_transfer_code_b200.streaming import CameraPipeline
    tf = transforms.Compose([transforms.Resize((448, 448)), transforms.ToTensor(), transforms.Normalize(mean=MEAN, std=STD)])
    rng = np.random.default_rng(3)
    graph = CameraPipeline.from_transform(model, tf, (1080, 1920, 3), use_graph=True)
    eager = CameraPipeline.from_transform(model, tf, (1080, 1920, 3), use_graph=False)
    assert graph is not None and graph._graph is not None
    for _ in range(3):
        frame = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
        with torch.no_grad():
            x = tf(Image.fromarray(np.ascontiguousarray(frame[:, :, ::-1]))).unsqueeze(0).to(model.device)
            _, logits = model(x)
            ref = torch.softmax(logits, dim=1).cpu().numpy()[0]
        pg, pe = graph(frame), eager(frame)
        # same input bits, same kernels; batch-1 Gram launches split K with fp32 atomics, hence not bit exact
        assert np.abs(pg - ref).max() <= 1e-5 and np.abs(pe - ref).max() <= 1e-5
        assert pg.argmax() == ref.argmax() == pe.argmax()
        assert abs(pg.sum() - 1.0) <= 1e-5


def test_run_camera_auto_uses_the_gpu_pipeline(model, tmp_path):
    from torchvision import transforms
    from heuristique_style_transfer_code_b200.functions import run_camera

    class Capture:
        def __init__(self, n):
            self.frames = [np.random.default_rng(i).integers(0, 256, (480, 640, 3), dtype=np.uint8) for i in range(n)]
            self.i = 0

        def isOpened(self):
            return True

        def read(self):
            if self.i >= len(self.frames):
                return False, None
            self.i += 1
            return True, self.frames[self.i - 1].copy()

        def release(self):
            pass

    tf = transforms.Compose([transforms.Resize(256), transforms.CenterCrop(224), transforms.ToTensor(),
                             transforms.Normalize(mean=MEAN, std=STD)])
    names = ["fog", "rain", "snow", "sun"]
    t_gpu = run_camera(model, tf, names, False, str(tmp_path), 0.5, True, capture=Capture(4), display=False, pipeline="auto")
    t_host = run_camera(model, tf, names, False, str(tmp_path), 0.5, False, capture=Capture(4), display=False, pipeline="host")
    assert len(t_gpu) == 4 and len(t_host) == 4 and (tmp_path / "times_camera.json").exists()
