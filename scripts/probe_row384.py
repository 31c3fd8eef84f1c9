"""Developer probe for gemm_row384 (prints progressively; run under `timeout`)."""
import math
import sys

import torch

sys.path.insert(0, ".")
from internnav_b200 import _lib


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


for M, K, use_ln, use_gamma in [(128, 384, False, False), (128, 384, True, False), (300, 384, True, True),
                                (1000, 1536, True, False), (8192, 384, True, False), (65536, 1536, True, False),
                                (65536, 384, False, False)]:
    torch.manual_seed(M + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(384, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(384, device="cuda") * 0.1
    gamma = torch.randn(384, device="cuda") if use_gamma else None
    res = torch.randn(M, 384, device="cuda").bfloat16()
    lw, lb = (1 + 0.1 * torch.randn(384, device="cuda"), 0.1 * torch.randn(384, device="cuda")) if use_ln else (None, None)
    y = a.float() @ w.float().T + bias
    if gamma is not None:
        y = y * gamma
    y = y + res.float()
    print("case", M, K, use_ln, use_gamma, flush=True)
    r = _lib.gemm_row384(a, w, bias, gamma=gamma, residual=res, ln_w=lw, ln_b=lb)
    torch.cuda.synchronize()
    if use_ln:
        out, ln = r
        ref_ln = torch.nn.functional.layer_norm(y.bfloat16().float(), (384,), lw, lb, 1e-5)
        print("   x err %.2e  ln err %.2e" % (rel(out, y), rel(ln, ref_ln)), flush=True)
    else:
        print("   x err %.2e" % rel(r, y), flush=True)
        r2 = res.clone()
        _lib.gemm_row384(a, w, bias, gamma=gamma, residual=r2, out=r2)
        torch.cuda.synchronize()
        print("   in-place err %.2e" % rel(r2, y), flush=True)
    if M == 65536:
        for _ in range(3):
            _lib.gemm_row384(a, w, bias, residual=res, ln_w=lw, ln_b=lb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _lib.gemm_row384(a, w, bias, residual=res, ln_w=lw, ln_b=lb)
        e1.record()
        e1.synchronize()
        print("   time %.1f us" % (e0.elapsed_time(e1) * 100), flush=True)
print("done")
