"""Data-parallel plumbing for the 8xB200 box (BASELINE configs[4]).  Device-agnostic, so
the world-size-2 gloo tests exercise exactly this code on CPU.

The replay path shards without communication: actor stream s belongs to shard ``s mod G``; each
rank owns the ring + sum-tree of its streams and samples locally.  The only exchange step is the
learner's gradient all-reduce.  The reference's loss is SUM-reduced over the batch
(``q_loss.mul(weights).sum().backward()``, agent0/deepq/agent.py:154) and its Adam uses
``eps = 1e-2 / batch_size`` (agent.py:102-106), so G ranks at batch B reproduce one learner at
batch G*B when gradients are SUM-all-reduced and ``eps = 1e-2 / (G*B)``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_of_stream(stream, world):
    """Actor stream (env) id -> owning rank."""
    return int(stream) % int(world)


def local_streams(rank, world, num_streams):
    """Stream ids owned by ``rank`` out of ``num_streams`` global actor streams."""
    return [s for s in range(int(num_streams)) if s % int(world) == int(rank)]


def shard_capacity(total_transitions, world):
    """Per-rank ring capacity for a global buffer of ``total_transitions`` (configs[4]: 8 M / G)."""
    return (int(total_transitions) + int(world) - 1) // int(world)


def adam_eps(batch_size, world=1):
    """agent.py:102-106 generalised to G ranks (see module docstring)."""
    return 1e-2 / (int(batch_size) * int(world))


class FlatGradBucket:
    """All parameter gradients as views of ONE flat fp32 buffer: zeroed with one memset and
    all-reduced (SUM) with one collective call -- 6.7-8.8 MB for the deepq nets, far below the
    size where NVLink bandwidth matters, so a single launch-latency-bound call is the cheapest
    schedule."""

    def __init__(self, params, process_group=None):
        self.params = [p for p in params]
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, async_op=False):
        if self.world > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=async_op)
        return None


def global_priority_stats(local_sum, local_top, process_group=None):
    """Sum of priorities and number of sampleable records over all shards (one 2-float SUM
    all-reduce).  Passing ``top=global_top`` and ``sum_offset=global_sum - local_sum`` to
    ``a0_pt_sample`` makes the IS weights those of a single global buffer (trainer.py:91-94)
    up to the per-batch max normalisation, which stays shard-local."""
    t = torch.stack((torch.as_tensor(local_sum, dtype=torch.float32).reshape(()),
                     torch.as_tensor(float(local_top), dtype=torch.float32, device=getattr(local_sum, "device", None))))
    if process_group is not None and dist.get_world_size(process_group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
    return t[0], t[1]
