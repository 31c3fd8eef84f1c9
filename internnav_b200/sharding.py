"""Multi-GPU host logic of the policy forward: one process per GPU, environments sharded across ranks, no collective
on the step path; a single gather of per-environment results at the end.

Mirrors the reference's episode striding (`per_scene_eps[rank::world_size]`, internnav/env/habitat_env.py L72) and the
metric gather of DistributedEvaluator (internnav/evaluator/distributed_base.py L98, L121).
"""
import torch.distributed as dist


def shard_indices(n_items, rank, world_size):
    """Indices of the items (environments / episodes) owned by `rank`: rank, rank + W, rank + 2W, ..."""
    return list(range(rank, n_items, world_size))


def gather_env_results(local_indices, local_results, n_items, group=None):
    """Every rank passes the results of its own environments; every rank gets the full, index-ordered list back."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = [None] * n_items
        for i, r in zip(local_indices, local_results):
            out[i] = r
        return out
    world = dist.get_world_size(group)
    bucket = [None] * world
    dist.all_gather_object(bucket, (list(local_indices), list(local_results)), group=group)
    out = [None] * n_items
    for idx, res in bucket:
        for i, r in zip(idx, res):
            out[i] = r
    return out
