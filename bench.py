#!/usr/bin/env python
"""bench.py -- InternVLA-N1 policy-steps/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (BASELINE.json `configs`):
  navdp_denoise  configs[1]: NavDP diffusion denoiser only, 50 DDPM steps, 256 trajectories (8 envs x 32 samples) of
                 horizon 8, bf16.  One "step" = one full 50-step sampling call; one policy step = one environment's
                 32-trajectory sample (8 per call).
Multi-GPU: environments are independent, so every rank runs the same per-GPU workload on its own shard (weak scaling,
no data-path collective).  Timing: CUDA events on the launching stream, per timed step, L2 flushed
between steps, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "navdp_denoise": dict(B=8, Ns=32, T=8, K=50,
                          desc="configs[1]: NavDP denoiser only, 50 denoise steps, 256 trajectories (8 envs x 32) of horizon 8, bf16"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def denoise_flops_per_sample_step(T, D=384, M=34, Ns=32, layers=16):
    # decoder layer = 28 T D^2 + 4 M D^2 / Ns + 4 T^2 D + 4 T M D  (memory K/V once per env)
    return layers * (28 * T * D * D + 4 * M * D * D / Ns + 4 * T * T * D + 4 * T * M * D)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, wl):
    import torch.distributed as dist
    from internnav_b200 import _lib
    from internnav_b200.manifest import random_navdp_state_dict
    from internnav_b200.navdp import NavDP_Policy_DPT_CriticSum_DAT
    from internnav_b200.postprocess import batched_traj_to_actions

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a B200: there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, Ns, T, K = wl["B"], wl["Ns"], wl["T"], wl["K"]

    model = NavDP_Policy_DPT_CriticSum_DAT(memory_size=2, predict_size=32, navdp_version=0.1, device=str(dev))
    model.load_state_dict(random_navdp_state_dict(seed=0))
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    R = B * Ns
    # host-side (pinned) inputs of one step, as the caller of the policy holds them
    h_goal = torch.randn(B, 1, 384, generator=g).bfloat16().pin_memory()
    h_rgbd = torch.randn(B, 32, 384, generator=g).bfloat16().pin_memory()
    h_x0 = torch.randn(R, T, 3, generator=g).pin_memory()
    h_nz = torch.randn(K - 1, R, T, 3, generator=g).pin_memory()
    d_goal, d_rgbd, d_x0, d_nz = (t.to(dev) for t in (h_goal, h_rgbd, h_x0, h_nz))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_resident():
        return model.sample(d_goal, d_rgbd, d_x0, d_nz, num_steps=K)

    def step_e2e():
        goal = h_goal.to(dev, non_blocking=True)
        rgbd = h_rgbd.to(dev, non_blocking=True)
        x0 = h_x0.to(dev, non_blocking=True)
        nz = h_nz.to(dev, non_blocking=True)
        traj = model.sample(goal, rgbd, x0, nz, num_steps=K)
        return batched_traj_to_actions(traj, B)  # D2H of the trajectories + the integer tail

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, use_events=True):
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            if use_events:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                tot += a.elapsed_time(b)
            else:  # includes host work (D2H + numpy tail): wall clock around a synchronised region
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                tot += (time.perf_counter() - t0) * 1e3
        return tot

    for _ in range(max(args.warmup, 3)):
        step_resident()
    _lib.prof_read()
    barrier()
    with ClockSampler(local) as clk:
        ms = timed(step_resident, args.steps)
    barrier()
    launches = _lib.prof_read()
    clocks = clk.summary()
    for _ in range(2):
        step_e2e()
    barrier()
    ms_e2e = timed(step_e2e, args.steps, use_events=False)
    barrier()

    # roofline pass for the dominant kernel (tcgen05 GEMM): per-launch CUDA events, NOT part of the timed runs above
    _lib.prof_read()
    _lib.prof_enable(True)
    step_resident()
    torch.cuda.synchronize()
    prof = _lib.prof_read()
    _lib.prof_enable(False)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    pk = peaks()
    achieved_tf = prof["gemm_flops"] / (prof["gemm_ms"] * 1e-3) / 1e12 if prof["gemm_ms"] > 0 else 0.0
    algo_flops_step = denoise_flops_per_sample_step(T) * R * K
    out = {
        "metric": "InternVLA-N1 policy-steps/sec (batch RGB-D+text->action)", "value": value, "unit": "policy-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "impl": "ours",
        "config": {"workload": args.workload, "description": wl["desc"], "envs_per_gpu": B, "samples_per_env": Ns,
                   "horizon": T, "ddpm_steps": K, "trajectories_per_s": value * Ns, "parallelism": "env-sharded x%d" % world,
                   "weights": "random-init NavDP (98.8M params)", "l2": "flushed (256 MiB memset) between timed steps",
                   "algorithmic_tflop_per_step": algo_flops_step / 1e12,
                   "step_tflops_achieved": algo_flops_step / (ms_per_step * 1e-3) / 1e12},
        "e2e": {"value": e2e_value, "unit": "policy-steps/s",
                "h2d_bytes_per_step": sum(x.numel() * x.element_size() for x in (h_goal, h_rgbd, h_x0, h_nz)),
                "d2h_bytes_per_step": R * T * 3 * 4,
                "api": "NavDP_Policy_DPT_CriticSum_DAT.sample + batched_traj_to_actions, pinned host inputs"},
        "gpu_launches": int(launches["total_launches"]),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "n1::gemm_kernel<BN> (tcgen05)", "achieved": achieved_tf,
                     "peak": pk["tf"], "unit": "TFLOP/s", "frac": achieved_tf / pk["tf"], "traffic": None,
                     "peak_source": pk["src"], "gemm_launches_per_step": int(prof["gemm_launches"]),
                     "gemm_ms_per_step": prof["gemm_ms"], "gemm_share_of_step": prof["gemm_ms"] / ms_per_step},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(wl, budget_s=20.0)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_baseline(wl, budget_s=20.0, threads=None):
    """The reference algorithm (oracle port, fp32 PyTorch eager) on this box's host cores, on a bounded sample of the
    same workload: 1 environment (32 trajectories) for as many denoise steps as fit the budget, scaled linearly to K."""
    from oracle import navdp_oracle as O, weights
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd = weights.make_state_dict(0)
    T, K, Ns = wl["T"], wl["K"], wl["Ns"]
    inp = weights.make_inputs(5, B=1, T=T, Ns=Ns, K=2)
    x, goal, rgbd = inp["x_init"], inp["goal"], inp["rgbd"]
    k = torch.tensor([3])
    with torch.no_grad():
        O.predict_noise(sd, x, k, goal, rgbd)  # warm-up
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s and n < K:
            O.predict_noise(sd, x, k, goal, rgbd)
            n += 1
        dt = time.perf_counter() - t0
    per_env_call = dt / n * K
    return {"value": 1.0 / per_env_call, "unit": "policy-steps/s", "cores": threads, "kind": "port",
            "sample": "%d denoise steps of 1 env x %d trajectories x T=%d (fp32 eager oracle), scaled to %d steps" % (n, Ns, T, K)}


def run_reference(args, wl):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    # a bounded sample per "step": budget split over warmup + steps so the whole run stays within a few minutes
    total = max(args.steps + args.warmup, 1)
    per = max(4.0, min(20.0, 120.0 / total))
    vals = []
    for i in range(total):
        r = cpu_baseline(wl, budget_s=per)
        if i >= args.warmup:
            vals.append(r)
    v = sum(x["value"] for x in vals) / len(vals)
    cb = dict(vals[-1])
    cb["value"] = v
    out = {"metric": "InternVLA-N1 policy-steps/sec (batch RGB-D+text->action)", "value": v, "unit": "policy-steps/s",
           "impl": "reference", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * wl["B"] / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "description": wl["desc"],
                      "note": "reference algorithm on host cores (CPU oracle port; /root/reference is Python and cannot travel)"},
           "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "policy-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="navdp_denoise", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
