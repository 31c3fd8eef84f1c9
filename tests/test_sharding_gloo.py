"""N > 1 host path on CPU: two gloo ranks shard 7 environments, run the integer tail of a policy step on their shards
and gather; the result must equal the single-process result (no data-path collective is involved)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from internnav_b200.postprocess import batched_traj_to_actions, s1_action_list
from internnav_b200.sharding import gather_env_results, shard_indices

N_ENV = 7


def _trajs():
    rng = np.random.Generator(np.random.PCG64(11))
    return [torch.from_numpy(rng.standard_normal((32, 32, 3), dtype=np.float32) * 0.2 +
                             np.array([0.6 * np.cos(i), 0.6 * np.sin(i), 0], dtype=np.float32)) for i in range(N_ENV)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(N_ENV, rank, world)
    tr = _trajs()
    local = torch.cat([tr[i] for i in mine]) if mine else torch.zeros(0, 32, 3)
    acts = [s1_action_list(a) for a in batched_traj_to_actions(local, len(mine))] if mine else []
    full = gather_env_results(mine, acts, N_ENV)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t)
    if rank == 0:
        q.put((full, int(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    assert shard_indices(7, 0, 2) == [0, 2, 4, 6] and shard_indices(7, 1, 2) == [1, 3, 5]
    assert shard_indices(3, 5, 8) == []
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, n = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = [s1_action_list(a) for a in batched_traj_to_actions(torch.cat(_trajs()), N_ENV)]
    assert n == N_ENV and full == ref
