"""Developer probe: one GEMM through the C ABI with explicit synchronisation, for compute-sanitizer runs."""
import math
import sys

import torch

sys.path.insert(0, ".")
from internnav_b200 import _lib

M, N, K = [int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (128, 128, 64))]
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
ref = a.float() @ w.float().T
torch.cuda.synchronize()
out = _lib.gemm(a, w)
try:
    torch.cuda.synchronize()
except Exception as e:
    print("SYNC ERROR:", e)
    sys.exit(1)
err = ((out.float() - ref).norm() / ref.norm()).item()
print("gemm %dx%dx%d rel err %.3e" % (M, N, K, err))
if err > 1e-2:
    d = (out.float() - ref).abs()
    print("max abs", d.max().item(), "bad rows", (d.max(1).values > 0.05).nonzero().flatten()[:20].tolist(),
          "bad cols", (d.max(0).values > 0.05).nonzero().flatten()[:20].tolist())
    print(out[:4, :8].float(), ref[:4, :8])
