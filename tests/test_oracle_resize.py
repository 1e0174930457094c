"""The numpy restatement of Pillow's bicubic resample (oracle/resize_port.py) against Pillow itself."""

import numpy as np
import pytest
from PIL import Image

from oracle import resize_port as rz

SIZES = [(224, 224, 224), (300, 260, 224), (230, 500, 224), (640, 480, 224), (100, 80, 224), (37, 91, 64), (1000, 333, 224),
         (225, 224, 224), (224, 897, 224), (51, 50, 32)]


@pytest.mark.parametrize("w,h,S", SIZES)
def test_resize_matches_pillow(w, h, S):
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    nw, nh = rz.resized_size(w, h, S)
    want = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BICUBIC))
    assert np.array_equal(rz.resize_bicubic(img, nw, nh), want)


def test_resize_center_crop_matches_torchvision_transform():
    import torchvision.transforms as T

    rng = np.random.default_rng(3)
    tf = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224), T.PILToTensor()])
    for h, w in ((260, 300), (500, 230), (224, 224), (231, 224)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(rz.resize_center_crop(img, 224), tf(Image.fromarray(img)).numpy())


def test_extreme_values_clamp():
    img = np.zeros((64, 64, 3), dtype=np.uint8)
    img[::2] = 255  # bicubic overshoots on a 0/255 grating
    want = np.asarray(Image.fromarray(img).resize((48, 40), Image.BICUBIC))
    assert np.array_equal(rz.resize_bicubic(img, 48, 40), want)
