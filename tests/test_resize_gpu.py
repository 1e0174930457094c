"""GPU bicubic Resize + CenterCrop (slb_resize_bicubic_u8) — byte-identical to Pillow, which is what the reference's
open_clip transform calls per image (foundation_models/clip.py:157-160)."""

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import resize_port as rz

pytestmark = pytest.mark.gpu

SIZES = [(224, 224, 224), (300, 260, 224), (230, 500, 224), (640, 480, 224), (100, 80, 224), (37, 91, 64), (1000, 333, 224),
         (225, 224, 224), (224, 897, 224), (51, 50, 32), (1920, 1080, 224), (500, 375, 256), (7, 9, 32)]


@pytest.fixture(scope="module")
def ops():
    from semanticlens_b200 import ops

    return ops


def pillow_resize_crop(img, S):
    h, w = img.shape[:2]
    nw, nh = rz.resized_size(w, h, S)
    r = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BICUBIC))
    left, top = int(round((nw - S) / 2.0)), int(round((nh - S) / 2.0))
    return r[top:top + S, left:left + S].transpose(2, 0, 1)


@pytest.mark.parametrize("w,h,S", SIZES)
def test_resize_center_crop_is_pillow_exact(ops, w, h, S):
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = ops.resize_center_crop_u8(torch.from_numpy(img).cuda(), S).cpu().numpy()
    assert np.array_equal(got, pillow_resize_crop(img, S))
    assert np.array_equal(got, rz.resize_center_crop(img, S))  # and the numpy oracle


def test_resize_overshoot_clamps_like_pillow(ops):
    img = np.zeros((300, 260, 3), dtype=np.uint8)
    img[::2, 1::2] = 255
    got = ops.resize_center_crop_u8(torch.from_numpy(img).cuda(), 224).cpu().numpy()
    assert np.array_equal(got, pillow_resize_crop(img, 224))


def test_full_window_resize_through_the_c_abi(ops):
    """No crop: the whole resized image, any aspect ratio (up-scaling one axis, shrinking the other)."""
    from semanticlens_b200 import _native as N

    lib = N.load(require_device=True)
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (120, 333, 3), dtype=np.uint8)
    for nw, nh in ((200, 150), (96, 240), (333, 60)):
        src = torch.from_numpy(img).cuda()
        dst = torch.empty((3, nh, nw), dtype=torch.uint8, device="cuda")
        need = lib.slb_resize_workspace_bytes(120, 333, nw, nh, 0, 0, nw, nh)
        ws = torch.empty(need, dtype=torch.uint8, device="cuda")
        N.check(lib.slb_resize_bicubic_u8(src.data_ptr(), 120, 333, nw, nh, 0, 0, nw, nh, dst.data_ptr(), ws.data_ptr(), need,
                                          N.stream_ptr(src.device)), "resize")
        want = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BICUBIC)).transpose(2, 0, 1)
        assert np.array_equal(dst.cpu().numpy(), want)
    assert lib.slb_resize_workspace_bytes(120, 333, 100, 100, 50, 0, 100, 100) == 0  # window outside the resized image
    rc = lib.slb_resize_bicubic_u8(src.data_ptr(), 120, 333, 100, 100, 0, 0, 100, 100, dst.data_ptr(), ws.data_ptr(), 16, None)
    assert rc != 0 and b"workspace" in lib.slb_last_error()


def test_openclip_preprocess_device_resize_matches_reference_transform():
    import torchvision.transforms as T

    from semanticlens_b200.foundation_models import OpenClip

    fm = OpenClip("ViT-B-32", device="cuda", load_weights=False)
    rng = np.random.default_rng(0)
    ims = [Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for h, w in ((224, 224), (260, 300), (500, 230), (97, 131))]
    ims.append(Image.fromarray(rng.integers(0, 256, (240, 250), dtype=np.uint8)))  # mode "L": host path, like the reference
    tf = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224), lambda im: im.convert("RGB"),
                    T.ToTensor(), T.Normalize(fm.cfg.mean, fm.cfg.std)])
    want = torch.stack([tf(im) for im in ims])
    assert torch.equal(fm.preprocess(ims).cpu(), want)


@pytest.mark.parametrize("w,h,S", [(300, 260, 224), (230, 500, 224), (97, 131, 64), (640, 480, 256), (100, 80, 224)])
def test_squash_resize_is_pillow_exact(ops, w, h, S):
    """open_clip's resize_mode="squash" (SigLIP preprocess): Resize((S, S), bicubic), aspect ratio not kept, no crop."""
    rng = np.random.default_rng(w * 3 + h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = ops.resize_center_crop_u8(torch.from_numpy(img).cuda(), S, squash=True).cpu().numpy()
    want = np.asarray(Image.fromarray(img).resize((S, S), Image.BICUBIC)).transpose(2, 0, 1)
    assert np.array_equal(got, want)
    assert np.array_equal(got, rz.resize_bicubic(img, S, S).transpose(2, 0, 1))


def test_siglip_preprocess_squashes_like_the_open_clip_transform():
    import torchvision.transforms as T

    from semanticlens_b200.foundation_models import SigLipV2

    fm = SigLipV2(device="cuda", load_weights=False)
    rng = np.random.default_rng(1)
    ims = [Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for h, w in ((224, 224), (260, 300), (500, 230))]
    ims.append(Image.fromarray(rng.integers(0, 256, (240, 250), dtype=np.uint8)))  # mode "L": host path
    tf = T.Compose([T.Resize((224, 224), interpolation=T.InterpolationMode.BICUBIC), lambda im: im.convert("RGB"),
                    T.ToTensor(), T.Normalize(fm.cfg.mean, fm.cfg.std)])
    assert torch.equal(fm.preprocess(ims).cpu(), torch.stack([tf(im) for im in ims]))
