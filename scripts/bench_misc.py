"""Developer microbenchmark of the non-GEMM kernels at the batch-64 dual-system shapes."""
import sys
import torch
sys.path.insert(0, ".")
from internnav_b200 import _lib


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters * 1e3


R, D = 65536, 384
qkv = torch.randn(R, 3 * D, device="cuda").bfloat16()
print("denoiser self-attn  T=32   %7.1f us  (x320/step)" % timeit(lambda: _lib.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], 8, 8, 48, R // 32, 32, 32, causal=True)))
ckv = torch.randn(64 * 34, 16 * 2 * D, device="cuda").bfloat16()
q = torch.randn(R, D, device="cuda").bfloat16()
print("denoiser cross-attn 34 keys %7.1f us  (x320)" % timeit(lambda: _lib.attention(q, ckv[:, :D], ckv[:, D:2 * D], 8, 8, 48, R // 32, 32, 34, kv_div=32)))
x = torch.randn(R, D, device="cuda").bfloat16()
w = torch.ones(D, device="cuda")
print("LayerNorm 65536x384         %7.1f us  (x960)" % timeit(lambda: _lib.layernorm(x, w, w)))
T, H = 19456, 3584
xl = torch.randn(T, H, device="cuda").bfloat16()
wl = torch.ones(H, device="cuda")
print("RMSNorm 19456x3584          %7.1f us  (x56)" % timeit(lambda: _lib.layernorm(xl, wl, None, rms=True)))
qkvl = torch.randn(T, 4608, device="cuda").bfloat16()
cu = torch.arange(0, 65 * 304, 304, device="cuda", dtype=torch.int32)
print("LLM attention 64x304 GQA    %7.1f us  (x28)" % timeit(lambda: _lib.attention(qkvl[:, :3584], qkvl[:, 3584:4096], qkvl[:, 4096:], 28, 4, 128, 64, 0, 0, cu_q=cu, cu_k=cu, max_seq_q=304, causal=True)))
N, Hv = 50176, 1280
qv = torch.randn(N, 3 * Hv, device="cuda").bfloat16()
cuw = torch.arange(0, N + 1, 64, device="cuda", dtype=torch.int32)
print("ViT window attention 64-tok %7.1f us  (x28)" % timeit(lambda: _lib.attention(qv[:, :Hv], qv[:, Hv:2 * Hv], qv[:, 2 * Hv:], 16, 16, 80, cuw.numel() - 1, 0, 0, cu_q=cuw, cu_k=cuw, max_seq_q=64)))
cuf = torch.arange(0, N + 1, 784, device="cuda", dtype=torch.int32)
print("ViT full attention 784-tok  %7.1f us  (x4)" % timeit(lambda: _lib.attention(qv[:, :Hv], qv[:, Hv:2 * Hv], qv[:, 2 * Hv:], 16, 16, 80, cuf.numel() - 1, 0, 0, cu_q=cuf, cu_k=cuf, max_seq_q=784)))
xv = torch.randn(N, Hv, device="cuda").bfloat16()
wv = torch.ones(Hv, device="cuda")
print("RMSNorm 50176x1280          %7.1f us  (x65)" % timeit(lambda: _lib.layernorm(xv, wv, None, rms=True)))
