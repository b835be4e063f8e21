"""numpy stand-in for the device side of the replay shard (K1 scatter + K3 gather semantics), used
by the CPU tests of the host index logic.  Test infrastructure only: it executes AppendPlans the
way the CUDA kernels are specified to (include/agent0_b200.h) so that agent0_b200.ring_index can
be exercised without a GPU."""
import numpy as np


class SimDevice:
    def __init__(self, N, NF, F):
        self.N, self.NF, self.F = N, NF, F
        self.frames = np.zeros((NF, F), dtype=np.uint8)
        self.slots = np.zeros((N, 8), dtype=np.int32)
        self.reward = np.zeros(N, dtype=np.float64)
        self.action = np.zeros(N, dtype=np.int64)
        self.done = np.zeros(N, dtype=np.bool_)
        self.link = np.full(N, -1, dtype=np.int32)
        self.leaf = np.zeros(N, dtype=np.float32)

    def execute(self, plan, flat_frames):
        # K2b marks first (evictions, then newly sampleable; last writer wins), then K1
        for p in plan.marks:
            if p >= 0:
                self.leaf[p] = 1.0
            else:
                self.leaf[~p] = 0.0
        if len(plan.new_frame_pos):
            self.frames[plan.new_frame_pos] = flat_frames[plan.new_frame_src]
        for mt in plan.rec_meta:
            pos, link_from, link_to, ad = int(mt[0]), int(mt[1]), int(mt[2]), int(mt[3])
            self.slots[pos] = mt[4:12]
            self.reward[pos] = np.array(mt[12:14], dtype=np.int32).view(np.float64)[0]
            self.action[pos] = ad & 0x7fffffff
            self.done[pos] = bool((ad >> 31) & 1)
            self.link[pos] = link_to
            if link_from >= 0:
                self.link[link_from] = pos

    def gather(self, pos, n_step, gamma):
        """K3 semantics for one record position."""
        p, rs, ds = pos, [], []
        for i in range(n_step):
            rs.append(self.reward[p]); ds.append(self.done[p])
            last = p
            if i + 1 < n_step:
                p = self.link[p]
                assert p >= 0
        r = np.float64(0.0)
        for rt, dt in zip(reversed(rs), reversed(ds)):
            r = r * gamma * (1 - int(dt)) + rt
        frames = np.concatenate((self.frames[self.slots[pos, :4]], self.frames[self.slots[last, 4:]])).reshape(-1)
        return frames, self.action[pos], r, bool(np.any(ds)), int(self.link[last])
