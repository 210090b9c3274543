"""CPU, world_size 2, gloo: the N>1 host logic — shard ranges, flat-bucket gradient all-reduce,
parameter broadcast.  (The rollout itself needs no collective; shard invariance of the kernels is
covered on the GPU by test_gpu_step.py::test_full_size_properties.)"""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "pytorch-rl-enhancedstablebaselines_b200"


def test_shard_range_partition():
    d = importlib.import_module(PKG + ".dist")
    for n, w in ((1_048_576, 8), (65_536, 1), (10, 4), (262_144, 8), (7, 8)):
        shards = [d.shard_range(n, r, w) for r in range(w)]
        assert shards[0][0] == 0 and sum(c for _, c in shards) == n
        for (o0, c0), (o1, _) in zip(shards, shards[1:]):
            assert o1 == o0 + c0  # contiguous, ordered, no gaps
        assert max(c for _, c in shards) - min(c for _, c in shards) <= 1
    with pytest.raises(ValueError):
        d.shard_range(10, 4, 4)


def _worker(rank, world, port, ret):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    d = importlib.import_module(PKG + ".dist")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # same init everywhere; rank 1 is perturbed, then broadcast must repair it
        net = torch.nn.Sequential(torch.nn.Linear(4, 400), torch.nn.ReLU(), torch.nn.Linear(400, 300), torch.nn.ReLU(), torch.nn.Linear(300, 2))
        if rank == 1:
            with torch.no_grad():
                for p in net.parameters():
                    p.add_(1.0)
        d.broadcast_parameters(net.parameters(), src=0)
        ref = torch.nn.Sequential(torch.nn.Linear(4, 400), torch.nn.ReLU(), torch.nn.Linear(400, 300), torch.nn.ReLU(), torch.nn.Linear(300, 2))
        torch.manual_seed(0)
        ref2 = torch.nn.Sequential(torch.nn.Linear(4, 400), torch.nn.ReLU(), torch.nn.Linear(400, 300), torch.nn.ReLU(), torch.nn.Linear(300, 2))
        same = all(torch.equal(a, b) for a, b in zip(net.parameters(), ref2.parameters()))
        # data-parallel gradient: each rank holds half of a batch; averaged grads == full-batch grads
        g = torch.Generator().manual_seed(5)
        x = torch.randn(64, 4, generator=g)
        y = torch.randn(64, 2, generator=g)
        lo, hi = rank * 32, rank * 32 + 32
        torch.nn.functional.mse_loss(net(x[lo:hi]), y[lo:hi]).backward()
        bucket = d.allreduce_gradients(list(net.parameters()), average=True)
        torch.nn.functional.mse_loss(ref2(x), y).backward()
        err = max(float((a.grad - b.grad).abs().max()) for a, b in zip(net.parameters(), ref2.parameters()))
        total = d.global_sum(float(rank + 1))
        # allreduce_flat: the in-place mean of an already flat gradient range (the hook of FusedTD3Update.update)
        flat = torch.arange(8, dtype=torch.float32) * (rank + 1)
        view = flat[2:6]
        d.allreduce_flat(view)
        flat_ok = bool(torch.equal(flat[2:6], torch.arange(2, 6, dtype=torch.float32) * 1.5)) and float(flat[0]) == 0.0 and float(flat[7]) == 7.0 * (rank + 1)
        ret[rank] = (same, err, bucket.numel(), total, flat_ok)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    n_params = 4 * 400 + 400 + 400 * 300 + 300 + 300 * 2 + 2
    for rank in (0, 1):
        same, err, numel, total, flat_ok = ret[rank]
        assert flat_ok, "allreduce_flat must average the given range in place and leave the rest untouched"
        assert same, "broadcast_parameters did not synchronise the ranks"
        assert err < 1e-6, err  # averaged shard gradients == full-batch gradients
        assert numel == n_params == 122_902  # the TD3 actor: one flat bucket
        assert total == 3.0
