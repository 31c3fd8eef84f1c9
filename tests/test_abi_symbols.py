"""The shared library loads without a GPU and exports every function include/n1b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    from internnav_b200 import _lib
    L = _lib.lib()
    with open(os.path.join(ROOT, "include", "n1b200.h")) as fh:
        src = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(n1_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert L.n1_version().decode().startswith("n1b200")


def test_no_gpu_fails_loudly():
    """Compute entry points must not fall back to anything when no sm_100 device is usable."""
    import torch
    from internnav_b200 import _lib
    L = _lib.lib()
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = L.n1_create(ctypes.byref(h), 0)
    assert rc != 0 and b"no CUDA device" in L.n1_last_error() or rc != 0


def test_ddpm_tables_match_oracle():
    import numpy as np
    from internnav_b200 import _lib
    from oracle import ddpm
    L = _lib.lib()
    for K in (20, 50):
        buf = (ctypes.c_float * (K * 5))()
        assert L.n1_ddpm_tables(K, buf) == 0
        mine = np.asarray(list(buf), dtype=np.float32).reshape(K, 5)
        ref = ddpm.DDPMScheduler(num_train_timesteps=K).coef_table()
        assert np.allclose(mine, ref, rtol=2e-5, atol=1e-7), np.abs(mine - ref).max()
