"""One launch of the LLM gate/up GEMM at the benchmark shape (M = 19456 = 64 x 304 tokens, N = 37888 interleaved
gate/up rows, K = 3584, SwiGLU epilogue) and of down_proj, for `ncu --set full` (dram bytes per launch = roofline.traffic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from internnav_b200 import _lib as L  # noqa: E402

x = torch.randn(19456, 3584, device="cuda").bfloat16()
w = (torch.randn(37888, 3584, device="cuda") / 60).bfloat16()
w2 = (torch.randn(3584, 18944, device="cuda") / 137).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    h = L.gemm(x, w, act=L.ACT_SWIGLU)
    flush.zero_()
    y = L.gemm(h, w2, residual=x)
torch.cuda.synchronize()
print("ok", h.shape, y.shape)
