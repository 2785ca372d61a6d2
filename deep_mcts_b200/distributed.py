"""Multi-GPU plumbing: one process per GPU, games sharded, no collective on the search path.

The reference runs N independent self-play processes and one trainer (reference
deep_mcts/train.py:128-134, 215-227); here every rank owns a pool of games and a replica of the
network, and the only collectives are (a) the gradient all-reduce of the training step and (b)
the reduction of counters / timings for reporting.  Works with any torch.distributed backend
(NCCL on GPUs, gloo in the CPU tests).
"""
from typing import Dict, Iterable, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def barrier() -> None:
    if world()[1] > 1:
        dist.barrier()


def shard_games(total_games: int, world_size: int, rank: int) -> int:
    """Games owned by ``rank`` when ``total_games`` are dealt round-robin (game i -> rank i % world)."""
    return total_games // world_size + (1 if rank < total_games % world_size else 0)


def reduce_counters(sums: Dict[str, float], maxima: Dict[str, float], device) -> Tuple[Dict[str, float], Dict[str, float]]:
    """Sum the work counters and max-reduce the timings over all ranks (whole-job throughput =
    total work / slowest rank's time)."""
    rank, size = world()
    if size == 1:
        return dict(sums), dict(maxima)
    s = torch.tensor([float(v) for v in sums.values()], dtype=torch.float64, device=device)
    m = torch.tensor([float(v) for v in maxima.values()], dtype=torch.float64, device=device)
    if sums:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    if maxima:
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return dict(zip(sums.keys(), s.tolist())), dict(zip(maxima.keys(), m.tolist()))


def _allreduce_mean(flat: torch.Tensor, size: int) -> None:
    """In-place mean over ranks; NCCL averages inside the collective, gloo needs the division."""
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= size


class GradientBucket:
    """All gradients of a module as views into ONE persistent flat buffer, so the data-parallel step is a
    single in-place all-reduce with no flatten / unflatten copies around it (1 579 845 floats =
    6.3 MB for the Othello 8x8 net, latency-bound over NVLink, so one bucket is the right shape).
    Autograd accumulates into an existing ``.grad`` in place, so the views survive ``backward()``;
    clear them with ``zero()`` (``optimizer.zero_grad()`` would drop them)."""

    def __init__(self, parameters: Iterable[torch.nn.Parameter]) -> None:
        self.params = [p for p in parameters if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(total, dtype=first.dtype, device=first.device)
        offset = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[offset:offset + n].view_as(p)
            offset += n

    def zero(self) -> None:
        self.flat.zero_()

    def attached(self) -> bool:
        """False once something replaced a ``.grad`` (e.g. ``zero_grad(set_to_none=True)``)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

    def allreduce(self) -> int:
        rank, size = world()
        if size > 1:
            _allreduce_mean(self.flat, size)
        return self.flat.numel()


def allreduce_buffers(module: torch.nn.Module) -> int:
    """Average the floating-point buffers (BatchNorm running_mean / running_var) over all ranks in one flat
    all-reduce.  Every rank trains on its own shard of the batch, so the running statistics drift apart;
    averaging them at each weight transfer keeps the replicas' self-play networks identical."""
    rank, size = world()
    bufs = [b for b in module.buffers() if b.is_floating_point()]
    if size == 1 or not bufs:
        return sum(b.numel() for b in bufs)
    flat = torch.cat([b.reshape(-1) for b in bufs])
    _allreduce_mean(flat, size)
    offset = 0
    for b in bufs:
        n = b.numel()
        b.copy_(flat[offset:offset + n].view_as(b))
        offset += n
    return offset


def allreduce_gradients(parameters: Iterable[torch.nn.Parameter]) -> int:
    """Average the gradients over all ranks through ONE flat bucket (1.58 M floats = 6.3 MB for the
    Othello 8x8 net: latency-bound on NVLink, so a single bucket is the right shape).  Returns the
    number of elements reduced."""
    rank, size = world()
    grads = [p.grad for p in parameters if p.grad is not None]
    if size == 1 or not grads:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= size
    offset = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[offset:offset + n].view_as(g))
        offset += n
    return offset


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """Make every replica start from rank ``src``'s weights and buffers."""
    rank, size = world()
    if size == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)
