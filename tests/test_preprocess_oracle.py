"""CPU: the camera-preprocessing oracle (oracle/pil_resize.py) against Pillow / torchvision themselves, and the
product-side table builder and transform parser (streaming.py) against the oracle. No GPU needed."""
import numpy as np
import pytest

from oracle import pil_resize as P

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
CASES = [(1080, 1920, (448, 448), None), (1080, 1920, 256, 224), (480, 640, 256, 224), (300, 200, (224, 224), None),
         (100, 160, (224, 224), None), (224, 224, (224, 224), None)]


def _torchvision_reference(frame_bgr, resize, crop):
    from PIL import Image
    from torchvision import transforms
    steps = [transforms.Resize(resize)]
    if crop is not None:
        steps.append(transforms.CenterCrop(crop))
    steps += [transforms.ToTensor(), transforms.Normalize(mean=MEAN, std=STD)]
    rgb = np.ascontiguousarray(frame_bgr[:, :, ::-1])          # cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
    return transforms.Compose(steps)(Image.fromarray(rgb)).numpy(), transforms.Compose(steps)


@pytest.mark.parametrize("H,W,resize,crop", CASES)
def test_oracle_is_bitwise_torchvision(H, W, resize, crop):
    frame = np.random.default_rng(H + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref, _ = _torchvision_reference(frame, resize, crop)
    got = P.camera_preprocess(frame, resize, crop, MEAN, STD)
    assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.parametrize("in_size,out_size", [(1920, 448), (1080, 448), (1920, 455), (100, 224), (224, 224), (7, 3)])
def test_product_tables_match_the_oracle(in_size, out_size):
    from heuristique_style_transfer_code_b200.streaming import resample_tables
    lo, n, kk = P.precompute_coeffs(in_size, out_size)
    lo2, n2, kk2 = resample_tables(in_size, out_size)
    assert np.array_equal(lo, lo2) and np.array_equal(n, n2) and np.array_equal(kk, kk2)
    first, count = out_size // 3, max(1, out_size // 2)
    lo3, n3, kk3 = resample_tables(in_size, out_size, first, count)
    assert np.array_equal(lo3, lo[first:first + count]) and np.array_equal(kk3, kk[first:first + count])
    assert (kk.sum(axis=1) >= (1 << P.PRECISION_BITS) - kk.shape[1]).all()      # rows sum to ~2^22


def test_transform_parser():
    from torchvision import transforms
    from heuristique_style_transfer_code_b200.streaming import parse_transform
    _, tf = _torchvision_reference(np.zeros((8, 8, 3), np.uint8), 256, 224)
    spec = parse_transform(tf)
    assert spec == dict(resize=256, crop=(224, 224), mean=MEAN, std=STD)
    _, tf = _torchvision_reference(np.zeros((8, 8, 3), np.uint8), (448, 448), None)
    assert parse_transform(tf)["resize"] == (448, 448) and parse_transform(tf)["crop"] is None
    assert parse_transform(transforms.Compose([transforms.ToTensor()])) is None
    assert parse_transform(transforms.Compose([transforms.Resize(256, interpolation=transforms.InterpolationMode.NEAREST),
                                               transforms.ToTensor(), transforms.Normalize(MEAN, STD)])) is None
    assert parse_transform(transforms.Compose([transforms.Resize(256), transforms.RandomCrop(224), transforms.ToTensor(),
                                               transforms.Normalize(MEAN, STD)])) is None
