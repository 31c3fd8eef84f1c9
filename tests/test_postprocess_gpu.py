"""The device action tail (csrc/postprocess.cu, n1_traj_to_actions) against the reference's own outputs
(tests/golden/traj_to_actions.json = internnav/model/utils/vln_utils.py `traj_to_actions` run on 13 seeded batches) and
against the numpy restatement on random batches.  Integer results: bit-exact."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _golden_cases():
    from test_oracle_s1 import _action_cases
    with open(os.path.join(ROOT, "tests", "golden", "traj_to_actions.json")) as fh:
        return _action_cases(), json.load(fh)


def test_golden_reference_outputs_bit_exact():
    from internnav_b200 import postprocess as P
    cases, gold = _golden_cases()
    # one environment at a time and all 13 as one batch of environments
    for tr, g in zip(cases, gold):
        assert P.batched_traj_to_actions_gpu(tr.cuda(), 1, cap=256)[0] == g
        assert P.s1_action_list(P.batched_traj_to_actions_gpu(tr.cuda(), 1, max_actions=4)[0]) == P.s1_action_list(g)
    batch = torch.cat(cases).cuda()
    assert P.batched_traj_to_actions_gpu(batch, len(cases), cap=256) == gold
    # bf16 trajectories (the dtype the reference hands to traj_to_actions in deployment)
    for tr in cases[:4]:
        tb = tr.bfloat16()
        assert P.batched_traj_to_actions_gpu(tb.cuda(), 1, cap=256)[0] == P.traj_to_actions(tb.clone())


def test_mean_path_and_ids_vs_numpy_on_random_batches():
    from internnav_b200 import postprocess as P
    rng = np.random.Generator(np.random.PCG64(2024))
    n_checked = 0
    for trial in range(6):
        B, Ns, T = [(64, 32, 32), (8, 32, 8), (3, 5, 17), (1, 1, 1), (16, 32, 24), (64, 32, 32)][trial]
        drift = rng.standard_normal((B, 1, 1, 3)).astype(np.float32) * np.float32(0.6)
        a = rng.standard_normal((B, Ns, T, 3)).astype(np.float32) * np.float32(0.25) + drift
        t = torch.from_numpy(a.reshape(B * Ns, T, 3))
        ids, mean = P.batched_traj_to_actions_gpu(t.cuda(), B, cap=256, return_mean=True)
        host = a.reshape(B * Ns, T, 3).copy()
        host[:, :, :2] /= 4.0
        for e in range(B):
            ref_mean = P._mean_trajectory(host[e * Ns:(e + 1) * Ns])
            assert np.array_equal(mean[e].cpu().numpy(), ref_mean), "mean path differs (env %d)" % e   # bit-exact float64
            assert ids[e] == P._discretise(ref_mean), (trial, e)
            n_checked += 1
        short = P.batched_traj_to_actions_gpu(t.cuda(), B, max_actions=4)
        assert [P.s1_action_list(x) for x in short] == [P.s1_action_list(x) for x in ids]
    assert n_checked == 64 + 8 + 3 + 1 + 16 + 64
    # degenerate: all-zero trajectories -> empty list (-> action -1 upstream)
    assert P.batched_traj_to_actions_gpu(torch.zeros(64, 8, 3).cuda(), 2) == [[], []]


def test_product_path_uses_the_kernel():
    """dual_system_step / s1_step_latent hand CUDA trajectories to batched_traj_to_actions: that call must not copy the
    trajectories to the host (the D2H is the ids only)."""
    from internnav_b200 import _lib, postprocess as P
    t = torch.randn(64 * 32, 32, 3, device="cuda")
    _lib.prof_read()
    out = P.batched_traj_to_actions(t, 64, max_actions=4)
    assert _lib.prof_read()["total_launches"] == 1 and len(out) == 64
