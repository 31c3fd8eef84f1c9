"""One-shot GPU check of gemm_skinny.cu: parity on the decode shapes + timing vs the tcgen05 GEMM (M = 64)."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from internnav_b200 import _lib as L  # noqa: E402


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def main():
    torch.manual_seed(0)
    ws = torch.empty(L.lib().n1_op_gemm_skinny_workspace_bytes(), dtype=torch.uint8, device="cuda")
    ok = True
    for M, N, K, act, bias, res in [(64, 3584, 3584, 0, False, True), (64, 4608, 3584, 0, True, False),
                                     (64, 37888, 3584, 3, False, False), (64, 3584, 18944, 0, False, True),
                                     (3, 512, 256, 0, True, True), (17, 1000, 328, 2, True, False),
                                     (64, 152064, 3584, 0, False, False), (1, 40, 72, 1, True, True)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device="cuda") * 0.1 if bias else None
        r = torch.randn(M, N // 2 if act == 3 else N, device="cuda").bfloat16() if res else None
        out = L.gemm_skinny(a, w, bias=b, residual=r, act=act, ws=ws)
        main = L.gemm(a, w, bias=b, residual=r, act=act)
        torch.cuda.synchronize()
        e = rel(out, main)

        def timeit(fn, n=20):
            fn()
            torch.cuda.synchronize()
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(n):
                fn()
            t.record()
            torch.cuda.synchronize()
            return s.elapsed_time(t) / n * 1e3
        if M == 64:
            t_s = timeit(lambda: L.gemm_skinny(a, w, bias=b, residual=r, act=act, ws=ws))
            t_m = timeit(lambda: L.gemm(a, w, bias=b, residual=r, act=act))
            gbs = N * K * 2 / (t_s * 1e-6) / 1e9
            print("M %d N %d K %d act %d: rel vs tcgen05 %.2e | skinny %.1f us (%.0f GB/s of W) vs tcgen05 %.1f us"
                  % (M, N, K, act, e, t_s, gbs, t_m), flush=True)
        else:
            print("M %d N %d K %d act %d: rel vs tcgen05 %.2e" % (M, N, K, act, e), flush=True)
        ok &= e < 4e-3
    print("SKINNY_PARITY", "OK" if ok else "FAIL", flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("elapsed %.1f s" % (time.time() - t0))
