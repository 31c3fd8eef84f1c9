"""One launch of each non-GEMM kernel at the batch-64 dual-system shapes (for `ncu --set full`)."""
import sys
import torch
sys.path.insert(0, ".")
from internnav_b200 import _lib
R, D = 65536, 384
qkv = torch.randn(R, 3 * D, device="cuda").bfloat16()
ckv = torch.randn(64 * 34, 16 * 2 * D, device="cuda").bfloat16()
q = torch.randn(R, D, device="cuda").bfloat16()
x = torch.randn(R, D, device="cuda").bfloat16()
w = torch.ones(D, device="cuda")
T, H = 19456, 3584
xl = torch.randn(T, H, device="cuda").bfloat16()
wl = torch.ones(H, device="cuda")
qkvl = torch.randn(T, 4608, device="cuda").bfloat16()
cu = torch.arange(0, 65 * 304, 304, device="cuda", dtype=torch.int32)
N, Hv = 50176, 1280
qv = torch.randn(N, 3 * Hv, device="cuda").bfloat16()
cuw = torch.arange(0, N + 1, 64, device="cuda", dtype=torch.int32)
cuf = torch.arange(0, N + 1, 784, device="cuda", dtype=torch.int32)
for _ in range(2):
    _lib.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], 8, 8, 48, R // 32, 32, 32, causal=True)
    _lib.attention(q, ckv[:, :D], ckv[:, D:2 * D], 8, 8, 48, R // 32, 32, 34, kv_div=32)
    _lib.layernorm(x, w, w)
    _lib.layernorm(xl, wl, None, rms=True)
    _lib.attention(qkvl[:, :3584], qkvl[:, 3584:4096], qkvl[:, 4096:], 28, 4, 128, 64, 0, 0, cu_q=cu, cu_k=cu, max_seq_q=304, causal=True)
    _lib.attention(qv[:, :Hv], qv[:, Hv:2 * Hv], qv[:, 2 * Hv:], 16, 16, 80, cuw.numel() - 1, 0, 0, cu_q=cuw, cu_k=cuw, max_seq_q=64)
    _lib.attention(qv[:, :Hv], qv[:, Hv:2 * Hv], qv[:, 2 * Hv:], 16, 16, 80, cuf.numel() - 1, 0, 0, cu_q=cuf, cu_k=cuf, max_seq_q=784)
torch.cuda.synchronize()
