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
