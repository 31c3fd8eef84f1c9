"""Gradient buckets of the training step on CPU: layout pinned against torch's own DDP bucket
assignment, arithmetic pinned against DistributedDataParallel on two gloo ranks."""
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from internnav_b200.ddp import GradientBuckets, bucket_assignment, trainable_shapes


def _shapes(seed, n=40):
    rng = np.random.Generator(np.random.PCG64(seed))
    out = OrderedDict()
    for i in range(n):
        shape = (int(rng.integers(1, 900)), int(rng.integers(1, 700))) if rng.random() < 0.6 else (int(rng.integers(1, 5000)),)
        out["p%d" % i] = (shape, torch.float32)
    return out


def test_bucket_layout_matches_torch_ddp():
    for seed, cap in [(1, 1), (2, 2), (3, 25), (4, 100)]:
        shapes = _shapes(seed)
        params = [torch.empty(s, dtype=d) for s, d in shapes.values()]
        idx, _ = dist._compute_bucket_assignment_by_size(params, [1024 * 1024, cap * 1024 * 1024], [False] * len(params))
        names = list(shapes)
        ref = [[names[i] for i in b] for b in reversed(idx)]       # DistributedDataParallel reverses the bucket order
        assert bucket_assignment(shapes, bucket_cap_mb=cap) == ref, (seed, cap)


def test_trainable_set_matches_reference_gradients():
    """The tensors that get a bucket are exactly those the reference's autograd gives a gradient (golden fixture)."""
    from internnav_b200.manifest import navdp_shapes
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "s1_training_reference.npz"))
    mine = trainable_shapes(navdp_shapes(), (1, 4, 3584))
    with_grad = {"model.navdp." + str(n) for n in gold["grad_names"]}
    assert set(mine) - {"model.latent_queries"} >= with_grad
    extra = set(mine) - {"model.latent_queries"} - with_grad
    # tensors of the module that are not touched by this branch at all (critic head etc.) may appear in `mine`;
    # none of them belongs to the detached RGB ViT
    assert not any(".rgb_model." in n for n in mine), "the RGB ViT is frozen in the reference"
    assert all(not n.startswith("model.navdp.decoder.") for n in extra)


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(300, 500)
        self.b = torch.nn.Linear(500, 700)
        self.c = torch.nn.Linear(700, 3)

    def forward(self, x):
        return self.c(torch.relu(self.b(torch.relu(self.a(x)))))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _Net()
    x = torch.randn(8, 300, generator=torch.Generator().manual_seed(100 + rank))
    # local gradients, no communication
    net(x).square().mean().backward()
    local = OrderedDict((n, p.grad.clone()) for n, p in net.named_parameters())
    net.zero_grad()
    # torch DDP on the same module and input
    ddp = torch.nn.parallel.DistributedDataParallel(net, bucket_cap_mb=1)
    ddp(x).square().mean().backward()
    ref = OrderedDict((n, p.grad.clone()) for n, p in net.named_parameters())
    # ours: gradients written into the bucket views, then one all-reduce per bucket
    gb = GradientBuckets(OrderedDict((n, (tuple(p.shape), p.dtype)) for n, p in net.named_parameters()), "cpu",
                         bucket_cap_mb=1)
    for n, g in local.items():
        gb.grads[n].copy_(g)
    gb.all_reduce(async_op=True).wait()
    ok = all(torch.equal(gb.grads[n], ref[n]) for n in ref)
    worst = max(float((gb.grads[n] - ref[n]).abs().max()) for n in ref)
    # the split exchange of the training step: every bucket but the one holding the LAST-arriving gradient first, that one
    # afterwards (DualSystemTrainer.step overlaps the first part with the System-2 backward) -- same result
    gb2 = GradientBuckets(OrderedDict((n, (tuple(p.shape), p.dtype)) for n, p in net.named_parameters()), "cpu",
                          bucket_cap_mb=1)
    for n, g in local.items():
        gb2.grads[n].copy_(g)
    late = gb2.bucket_of("a.weight")
    gb2.all_reduce(async_op=True, skip=(late,))
    gb2.all_reduce(async_op=True, only=(late,), append=True)
    gb2.wait()
    ok = ok and all(torch.equal(gb2.grads[n], ref[n]) for n in ref)
    if rank == 0:
        q.put((ok, worst, len(gb.buffers)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_all_reduce_equals_torch_ddp():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, worst, n_buckets = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_buckets >= 2
    assert ok, "bucketed all-reduce differs from DistributedDataParallel by %g" % worst
