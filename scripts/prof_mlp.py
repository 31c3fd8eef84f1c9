import sys
import torch
sys.path.insert(0, ".")
from internnav_b200 import _lib
for M in (2048, 65536):
    x = torch.randn(M, 384, device="cuda").bfloat16()
    w1 = (torch.randn(1536, 384, device="cuda") / 384 ** 0.5).bfloat16()
    w2 = (torch.randn(384, 1536, device="cuda") / 1536 ** 0.5).bfloat16()
    b1, b2 = torch.randn(1536, device="cuda"), torch.randn(384, device="cuda")
    res = torch.randn(M, 384, device="cuda").bfloat16()
    for _ in range(2):
        _lib.fused_mlp(x, w1, b1, w2, b2, residual=res, cluster=1)
torch.cuda.synchronize()
