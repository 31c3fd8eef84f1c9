"""Launch the two training-step kernels added in round 2 once each at their largest System-1 shapes, for `ncu --set full`:
  * wgrad_tn_kernel      dW[1536, 384] = dY[98304, 1536]^T X[98304, 384]   (depth-ViT fc1 weight gradient)
  * attn_bwd_mma_kernel  384 sequences x 257 tokens x 6 heads x 64          (depth-ViT attention backward)"""
import sys

import torch

sys.path.insert(0, ".")
from internnav_b200 import _bwd as K, _lib as L  # noqa: E402


def main():
    torch.manual_seed(0)
    M = 98304
    dy = torch.randn(M, 1536, device="cuda").bfloat16()
    x = torch.randn(M, 384, device="cuda").bfloat16()
    for _ in range(3):
        K.wgrad(dy, x)
    B, S, H, hd = 384, 257, 6, 64
    qkv = torch.randn(B * S, 3 * H * hd, device="cuda").bfloat16()
    q, k, v = qkv[:, :384], qkv[:, 384:768], qkv[:, 768:]
    o = L.attention(q, k, v, H, H, hd, B, S, S)
    do = torch.randn(B * S, H * hd, device="cuda").bfloat16()
    for _ in range(3):
        K.attention_bwd(q, k, v, o, do, H, H, hd, B, S, S)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
