import sys
import torch
sys.path.insert(0, ".")
from internnav_b200 import _lib
R = 65536
a = torch.randn(R, 384, device="cuda").bfloat16()
w = (torch.randn(384, 384, device="cuda") / 384 ** 0.5).bfloat16()
b = torch.randn(384, device="cuda")
res = torch.randn(R, 384, device="cuda").bfloat16()
for _ in range(2):
    _lib.gemm(a, w, bias=b, residual=res, out=res)
w3 = (torch.randn(1152, 384, device="cuda") / 384 ** 0.5).bfloat16()
b3 = torch.randn(1152, device="cuda")
for _ in range(2):
    _lib.gemm(a, w3, bias=b3)
torch.cuda.synchronize()
