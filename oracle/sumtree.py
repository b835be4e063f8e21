"""The fp32 sum-tree the CUDA sampler (K2a/K2b) must match bit-for-bit.  Test infrastructure only.

No reference counterpart: the reference keeps a flat fp32 priority vector and its intended
law is P(i) = priority[i] / sum(priority) via torch.multinomial without replacement
(agent0/deepq/replay.py:39-43, dead code behind a RandomSampler).  The tree is the B200
design's replacement; its arithmetic is *defined* here so that sampled indices are
reproducible:

  * implicit complete binary tree over P = next_pow2(capacity) leaves, node 1 = root,
    children of i are 2i and 2i+1, leaf j lives at P + j;
  * every internal node is fl32(left + right), recomputed from its children (never a delta);
  * stratified draw b of B with uniform u_b:  t = fl32(fl32(fl32(b + u_b) / B) * root);
    descent: go left iff (t < left) or (right <= 0); when going right, t = fl32(t - left).
"""
import numpy as np

f32 = np.float32


def next_pow2(n):
    p = 1
    while p < n:
        p *= 2
    return p


class SumTree:
    def __init__(self, capacity):
        self.capacity = capacity
        self.P = next_pow2(max(capacity, 2))
        self.nodes = np.zeros(2 * self.P, dtype=f32)

    @property
    def root(self):
        return self.nodes[1]

    def leaves(self):
        return self.nodes[self.P:self.P + self.capacity]

    def set(self, idx, values):
        """Set leaves (sequential: the last occurrence of a duplicated index wins), then
        recompute every ancestor from its children, level by level."""
        idx = np.asarray(idx, dtype=np.int64)
        values = np.asarray(values, dtype=f32)
        for i, v in zip(idx, values):
            self.nodes[self.P + i] = v
        level = np.unique((self.P + idx) >> 1)
        while True:
            self.nodes[level] = self.nodes[2 * level] + self.nodes[2 * level + 1]
            if level[0] == 1:
                break
            level = np.unique(level >> 1)

    def descend(self, t):
        node = 1
        t = f32(t)
        while node < self.P:
            left, right = self.nodes[2 * node], self.nodes[2 * node + 1]
            if t < left or right <= 0:
                node = 2 * node
            else:
                t = f32(t - left)
                node = 2 * node + 1
        return node - self.P

    def sample_stratified(self, u):
        u = np.asarray(u, dtype=f32)
        B = u.shape[0]
        root = self.root
        idx = np.empty(B, dtype=np.int64)
        for b in range(B):
            t = f32(f32(f32(b) + u[b]) / f32(B)) * root
            idx[b] = self.descend(f32(t))
        return idx, self.nodes[self.P + idx].copy()


def new_priority(loss, eps, alpha):
    """replay.py:56-58: (loss+eps)^alpha; sqrt when alpha == 0.5 as torch's CPU pow does."""
    x = np.asarray(loss, dtype=f32) + f32(eps)
    return np.sqrt(x) if alpha == 0.5 else np.power(x, f32(alpha)).astype(f32)
