"""NextDiT System 1 at the benchmark batch: 64 environments x 32 trajectories, 10 Euler steps (guidance 1.0 and 3.0).
CUDA events around whole calls; prints one JSON line.  Usage: python scripts/bench_nextdit.py [B] [Ns]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from internnav_b200 import _lib  # noqa: E402
from internnav_b200.manifest import random_nextdit_state_dict  # noqa: E402
from internnav_b200.nextdit import NextDiTSystem1  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    Ns = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    m = NextDiTSystem1().load_state_dict(random_nextdit_state_dict(0))
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(B, 4, 3584, generator=g).bfloat16().cuda()
    img = torch.rand(B, 2, 224, 224, 3, generator=g).cuda()
    x0 = torch.randn(B * Ns, 32, 3, generator=g).bfloat16().cuda()
    out = {"B": B, "Ns": Ns}
    for name, scale in (("guidance_1", 1.0), ("guidance_3", 3.0)):
        for _ in range(2):      # also captures the sampler's CUDA graph (must happen with the profiler off: it records events)
            m.generate_traj(lat, img, guidance_scale=scale, num_sample_trajs=Ns, x_init=x0)
        torch.cuda.synchronize()
        _lib.prof_enable(True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        cond = m.condition_tokens(lat, img)
        torch.cuda.synchronize()
        ev[0].record()
        cond = m.condition_tokens(lat, img)
        ev[1].record()
        m.sample(cond, x0, guidance_scale=scale, num_sample_trajs=Ns)
        ev[2].record()
        m.sample(cond, x0, guidance_scale=scale, num_sample_trajs=Ns, graph=False)
        ev.append(torch.cuda.Event(enable_timing=True))
        ev[3].record()
        torch.cuda.synchronize()
        pr = _lib.prof_read()
        _lib.prof_enable(False)
        out[name] = {"condition_tokens_ms": ev[0].elapsed_time(ev[1]), "sampler_ms": ev[1].elapsed_time(ev[2]), "sampler_eager_ms": ev[2].elapsed_time(ev[3]),
                     "launches": pr["total_launches"], "gemm_tflop": pr["gemm_flops"] / 1e12}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
