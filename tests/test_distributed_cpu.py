"""world_size-2 gloo tests (CPU) of the N>1 host logic: game sharding, counter reduction and the
flat-bucket gradient all-reduce used by the training step."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world_size, port, results):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from deep_mcts_b200 import distributed as dd
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet

    assert dd.world() == (rank, world_size)
    sums, maxima = dd.reduce_counters({"sims": 100.0 * (rank + 1), "evals": 7.0}, {"ms": 10.0 + rank}, "cpu")
    assert sums == {"sims": 300.0, "evals": 14.0} and maxima == {"ms": 11.0}

    torch.manual_seed(rank)  # replicas start different, then get synchronised
    net = ConvolutionalOthelloNet(4, num_residual=1, channels=8, value_head_hidden_units=16)
    dd.broadcast_parameters(net.net, 0)
    ref = [p.detach().clone() for p in net.net.parameters()]
    gathered = [None, None]
    dist.all_gather_object(gathered, [p.sum().item() for p in ref])
    assert gathered[0] == gathered[1]

    torch.manual_seed(100 + rank)
    x = torch.rand(5, 3, 4, 4)
    pi = torch.softmax(torch.rand(5, 17), dim=1)
    z = torch.rand(5, 1) * 2 - 1
    net.grad_sync = dd.allreduce_gradients
    net.train(x, pi, z)
    after = [p.detach().clone() for p in net.net.parameters()]
    sums_after = [None, None]
    dist.all_gather_object(sums_after, [p.double().sum().item() for p in after])
    assert sums_after[0] == sums_after[1], "replicas diverged after a synchronised step"
    assert any(not torch.equal(a, b) for a, b in zip(ref, after))

    # the reduced gradient equals the mean of the per-rank gradients
    g_local = torch.full((3,), float(rank + 1))
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = g_local.clone()
    assert dd.allreduce_gradients([p]) == 3
    assert torch.allclose(p.grad, torch.full((3,), 1.5))

    # GradientBucket: gradients as views into one flat buffer, all-reduced in place (no flatten copies); a
    # training step through it keeps the replicas identical and gives the same update as the copy path
    torch.manual_seed(7)
    net_a = ConvolutionalOthelloNet(4, num_residual=1, channels=8, value_head_hidden_units=16)
    net_b = net_a.copy()
    net_b.load_state_dict(net_a.net.state_dict())
    net_a.grad_sync = dd.allreduce_gradients
    net_b.grad_bucket = dd.GradientBucket(net_b.net.parameters())
    assert net_b.grad_bucket.attached() and net_b.grad_bucket.flat.numel() == sum(q.numel() for q in net_b.net.parameters())
    for step in range(2):      # the second step checks that zero() kept the views alive
        net_a.train(x, pi, z)
        net_b.train(x, pi, z)
        assert net_b.grad_bucket.attached()
    for qa, qb in zip(net_a.net.parameters(), net_b.net.parameters()):
        assert torch.allclose(qa, qb, atol=1e-7), "bucket path differs from the copy path"
    sums_b = [None, None]
    dist.all_gather_object(sums_b, [q.double().sum().item() for q in net_b.net.parameters()])
    assert sums_b[0] == sums_b[1]

    # BatchNorm running statistics drift apart (each rank normalises its own shard); the transfer averages them
    bufs_before = [None, None]
    dist.all_gather_object(bufs_before, [b.double().sum().item() for b in net_b.net.buffers() if b.is_floating_point()])
    assert bufs_before[0] != bufs_before[1]
    dd.allreduce_buffers(net_b.net)
    bufs_after = [None, None]
    dist.all_gather_object(bufs_after, [b.double().sum().item() for b in net_b.net.buffers() if b.is_floating_point()])
    assert bufs_after[0] == bufs_after[1]
    for got, a, b in zip(bufs_after[0], bufs_before[0], bufs_before[1]):
        assert abs(got - (a + b) / 2) < 1e-5
    dd.barrier()
    dist.destroy_process_group()
    results.put(rank)


def test_world_size_two_gloo():
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(results.get(timeout=5) for _ in procs) == [0, 1]


def test_shard_games_covers_everything():
    from deep_mcts_b200.distributed import shard_games
    for total in (0, 1, 7, 4096, 32768 + 3):
        for world in (1, 2, 4, 8):
            parts = [shard_games(total, world, r) for r in range(world)]
            assert sum(parts) == total and max(parts) - min(parts) <= 1
