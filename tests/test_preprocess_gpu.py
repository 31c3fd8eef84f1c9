"""GPU frame preprocessing (row a12) vs Pillow, bit for bit."""
import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W", [(480, 640), (224, 224), (100, 75), (224, 300), (720, 1280)])
def test_resize_matches_pillow_bit_exact(H, W):
    from internnav_b200.preprocess import FramePreprocessor
    pre = FramePreprocessor("cuda:0")
    rng = np.random.Generator(np.random.PCG64(H + W))
    n = 5
    rgb = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    rgb[:, : H // 3] = rng.integers(0, 2, (n, H // 3, W, 3), dtype=np.uint8) * 255
    dep = (rng.random((n, H, W)) * 0.9).astype(np.float32)
    dep[rng.random((n, H, W)) < 0.05] = 0.0
    out_rgb = pre.rgb(rgb).cpu().numpy()
    out_dep = pre.depth(dep).cpu().numpy()
    for i in range(n):
        ref = (np.array(Image.fromarray(rgb[i]).resize((224, 224))) / 255.0).astype(np.float32)
        assert np.array_equal(out_rgb[i], ref), np.abs(out_rgb[i] - ref).max()
        d = np.array(Image.fromarray(dep[i]).resize((224, 224))) * 10.0
        d[d > 5.0] = 5.0
        assert d.dtype == np.float32 and np.array_equal(out_dep[i].view(np.uint32), d.view(np.uint32))


def test_agent_frames_equal_host_path():
    from internnav_b200.agent import s1_frames
    from internnav_b200.preprocess import FramePreprocessor
    pre = FramePreprocessor("cuda:0")
    rng = np.random.Generator(np.random.PCG64(1))
    B = 3
    g_rgb = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(B)]
    c_rgb = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(B)]
    g_d = [rng.random((480, 640, 1)).astype(np.float32) for _ in range(B)]
    c_d = [rng.random((480, 640, 1)).astype(np.float32) for _ in range(B)]
    r, d = pre.s1_frames(g_rgb, g_d, c_rgb, c_d)
    assert r.shape == (B, 2, 224, 224, 3) and d.shape == (B, 2, 224, 224, 1)
    for e in range(B):
        hr, hd = s1_frames(g_rgb[e], g_d[e], c_rgb[e], c_d[e])
        assert torch.equal(r[e].cpu(), hr[0].float()) and torch.equal(d[e].cpu(), hd[0].float())
