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


def init_nccl(device_index, high_priority=True):
    """torch.distributed over NCCL for the data-parallel learner, one process per GPU (RANK / WORLD_SIZE /
    MASTER_* from the environment).  high_priority: NCCL's kernels run on a high-priority stream, so the
    all-reduce of the first gradient bucket is scheduled as soon as an SM frees up under the convolution backward
    it overlaps, instead of queueing behind that kernel's remaining CTAs."""
    opts = None
    if high_priority:
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            opts = None
    dev = torch.device("cuda", int(device_index))
    if opts is not None:
        try:
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
            return dist.group.WORLD
        except TypeError:
            pass
    dist.init_process_group("nccl", device_id=dev)
    return dist.group.WORLD


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
            # the view takes the parameter's own (dense) strides, so a channels-last convolution weight gets a
            # channels-last gradient and autograd's layout contract holds; the flat order inside the range follows memory
            seg = self.flat[off:off + p.numel()]
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            p.grad = seg.as_strided(p.size(), p.stride()) if (dense and not p.is_contiguous()) else seg.view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, async_op=False):
        if self.world > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=async_op)
        return None


class OverlappedGradBucket(FlatGradBucket):
    """The same flat buffer, all-reduced in ``n_buckets`` pieces WHILE the backward pass is still running
    ("one flat bucket, overlap with backward").  Parameters keep their registration order in the
    buffer (encoder first, head last); autograd produces gradients in the opposite order, so the buffer is cut
    into contiguous ranges from the END: bucket 0 = the head and the big fully-connected layer (6.9 of the
    7.3 MB of the deepq nets), ready before the convolution backward -- the expensive part -- has started, so
    its NCCL call runs under it; the last bucket (the convolutions, 0.3 MB) is what stays exposed, at launch
    latency.  A post-accumulate hook on every parameter counts its bucket down and issues
    ``all_reduce(async_op=True)`` when the bucket is complete; ``wait()`` joins them before the optimizer step.
    Works under CUDA-graph capture (the collectives are captured on NCCL's stream as parallel branches) and
    with gloo on CPU tensors (tests).  Same sums as the single call, bit for bit per element: every element
    is still reduced exactly once over the same ranks."""

    def __init__(self, params, process_group=None, n_buckets=2, min_bucket_elems=32768, split_frac=0.5):
        super().__init__(params, process_group)
        sizes = [p.numel() for p in self.params]
        total = sum(sizes)
        # contiguous parameter ranges, cut from the end: a bucket closes at the first parameter after which at most
        # split_frac of what was left remains for the earlier layers (deepq nets: right below the 3136x512 layer),
        # and never holds fewer than min_bucket_elems elements
        bounds = [len(self.params)]
        if n_buckets > 1 and total > 2 * min_bucket_elems:
            acc, target = 0, total
            for i in range(len(self.params) - 1, 0, -1):
                acc += sizes[i]
                rest = total - sum(sizes[i:])
                if len(bounds) < n_buckets and acc >= min_bucket_elems and 0 < rest <= split_frac * target:
                    bounds.append(i)
                    target, acc = rest, 0
        bounds.append(0)
        bounds = sorted(set(bounds), reverse=True)
        self.buckets = []                                  # (param_lo, param_hi, elem_lo, elem_hi), latest layers first
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        for hi, lo in zip(bounds[:-1], bounds[1:]):
            self.buckets.append((lo, hi, offs[lo], offs[hi]))
        self._bucket_of = {}
        for b, (lo, hi, _, _) in enumerate(self.buckets):
            for i in range(lo, hi):
                self._bucket_of[id(self.params[i])] = b
        self._left = [hi - lo for lo, hi, _, _ in self.buckets]
        self._handles = []
        self.enabled = True                                # False: no collectives (bench: the step without the exchange)
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)

    def zero_(self):
        self.flat.zero_()
        self._left = [hi - lo for lo, hi, _, _ in self.buckets]
        self._handles = []

    def _hook(self, p):
        b = self._bucket_of[id(p)]
        self._left[b] -= 1
        if self._left[b] == 0 and self.enabled:
            _, _, lo, hi = self.buckets[b]
            self._handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def wait(self):
        for h in self._handles:
            h.wait()
        self._handles = []

    def all_reduce(self, async_op=False):
        """After backward: buckets whose hooks did not fire (a parameter without a gradient this step) are reduced
        now, then everything is joined."""
        if self.world > 1 and self.enabled:
            for b, left in enumerate(self._left):
                if left > 0:
                    _, _, lo, hi = self.buckets[b]
                    self._handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
                    self._left[b] = 0
        self.wait()
        return None


def global_batch_is_weights(prio, local_sum, beta, batch, process_group=None):
    """IS weights normalised over the GLOBAL batch (trainer.py:91-94 for G shards sampled in parallel).

    Rank r draws its B transitions from its own shard with P(i) = p_i / S_r / G (S_r = the shard's priority sum),
    so the weight that corrects towards the uniform law over all N records is (N * P(i))^-beta = c * u_i with
    u_i = (p_i / S_r)^-beta and c = (N / G)^-beta common to every rank.  The reference divides by the maximum over
    the batch; for one learner at batch G*B that maximum runs over all ranks' draws: ONE all-reduce(MAX) of L
    floats per draw.  (With the shard-local maximum -- the default, no collective on the data path -- S_r cancels
    inside a rank's batch but differs between ranks.)  ``prio`` f32 [L*B] priorities of the sampled records
    (K2a's prio_out), ``local_sum`` the shard's root.  Returns w [L*B] with max over the global batch = 1/(1+1e-8/c...)
    i.e. u / (max_global u + 1e-8) -- c cancels except against the 1e-8, exactly as top/S do in the reference."""
    u = (prio / local_sum).pow(-float(beta)).view(-1, int(batch))
    mx = u.max(dim=1)[0].contiguous()
    if process_group is not None and dist.get_world_size(process_group) > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=process_group)
    return (u / (mx.unsqueeze(1) + 1e-8)).view(-1)


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
