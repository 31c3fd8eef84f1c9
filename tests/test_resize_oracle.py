"""oracle/pil_resize.py against Pillow itself, bit for bit -- CPU only (Pillow is what the reference calls)."""
import numpy as np
import pytest
from PIL import Image

from oracle import pil_resize as R

SIZES = [(480, 640, 224, 224), (224, 224, 224, 224), (100, 75, 224, 224), (720, 1280, 224, 224), (37, 501, 224, 224),
         (480, 640, 384, 384), (224, 300, 224, 224)]


@pytest.mark.parametrize("H,W,oh,ow", SIZES)
def test_rgb_u8_bit_exact(H, W, oh, ow):
    rng = np.random.Generator(np.random.PCG64(H * 1000 + W))
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[: H // 3] = rng.integers(0, 2, (H // 3, W, 3), dtype=np.uint8) * 255      # hard edges: overshoot must clamp
    ref = np.array(Image.fromarray(img).resize((ow, oh)))
    assert np.array_equal(R.resize_u8(img, ow, oh), ref)


@pytest.mark.parametrize("H,W,oh,ow", SIZES)
def test_depth_f32_bit_exact(H, W, oh, ow):
    rng = np.random.Generator(np.random.PCG64(H * 7 + W))
    d = (rng.random((H, W)) * 0.9).astype(np.float32)
    d[rng.random((H, W)) < 0.05] = 0.0
    ref = np.array(Image.fromarray(d).resize((ow, oh)))
    mine = R.resize_f32(d, ow, oh)
    assert ref.dtype == np.float32 and np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("n_in,n_out", [(640, 224), (480, 224), (224, 224), (75, 224), (1280, 224), (37, 224), (640, 384)])
def test_library_coefficient_tables_bit_exact(n_in, n_out):
    """The C++ table builder of libn1b200 (host-only entry point) vs the oracle: bounds, double weights and 22-bit
    fixed-point weights identical to the last bit."""
    from internnav_b200.preprocess import resize_coeffs
    b, w, f = resize_coeffs(n_in, n_out)
    rb, rw, k = R.coeffs(n_in, n_out)
    assert w.shape[1] == k and np.array_equal(b, rb)
    assert np.array_equal(w.view(np.uint64), rw.view(np.uint64))
    assert np.array_equal(f, R.fixed_point(rw))


def test_unit_scale_division_matches_reference_cast():
    """The reference divides in float64 (`uint8 array / 255.0`) and the model casts to its dtype; the kernel divides in
    fp32.  Same fp32 value for every byte."""
    v = np.arange(256)
    assert np.array_equal((v / 255.0).astype(np.float32), v.astype(np.float32) / np.float32(255.0))
