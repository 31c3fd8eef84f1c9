"""Developer probe: the denoiser's short-K GEMMs and small attention at batch-64 shapes, once each (for ncu)."""
import sys

import torch

sys.path.insert(0, ".")
from internnav_b200 import _lib

R, D = 65536, 384
x = torch.randn(R, D, device="cuda").bfloat16()
for N, K, act in [(1152, 384, 0), (1536, 384, 1), (384, 1536, 0), (384, 384, 0)]:
    a = torch.randn(R, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    for _ in range(2):
        _lib.gemm(a, w, bias=b, act=act)
qkv = torch.randn(R, 3 * D, device="cuda").bfloat16()
for _ in range(2):
    _lib.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], 8, 8, 48, R // 32, 32, 32, causal=True)
ckv = torch.randn(64 * 34, 16 * 2 * D, device="cuda").bfloat16()
q = torch.randn(R, D, device="cuda").bfloat16()
for _ in range(2):
    _lib.attention(q, ckv[:, :D], ckv[:, D:2 * D], 8, 8, 48, R // 32, 32, 34, kv_div=32)
w_ln = torch.ones(D, device="cuda")
for _ in range(2):
    _lib.layernorm(x, w_ln, w_ln)
torch.cuda.synchronize()
print("ok")
