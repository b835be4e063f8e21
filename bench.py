#!/usr/bin/env python
"""bench.py -- replay sample+target throughput of the agent0 deepq hot path on B200.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU pipeline (port)

One "step" = the inner loop of the reference's Trainer.step (agent0/deepq/trainer.py:82-104) with
the CNN excluded (network outputs pre-generated): draw L = learner_steps batches of B transitions
from the prioritized shard (K2a, one launch), gather them (K3, one launch: stack reconstruction +
n-step return), run the fused target/loss kernel once per batch (K4, L launches) and write the new
priorities (K2b, one launch).  Headline workload = BASELINE.json configs[1]: C51 51 atoms, PER,
n_step=3, double+dueling, batch 32, L=20, 1 M-transition ring of synthetic 84x84 uint8 frames.
`value` times that loop as a CUDA graph with everything resident in HBM; `e2e` times the same work
through the public Python API with new transitions arriving from pinned host memory every step
(replay ratio 8 samples per insert, agent0/deepq/config.py) and the losses read back to the host.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_BYTES = 84 * 84
METRIC = "replay_sample_target_transitions_per_sec"
UNIT = "transitions/s"

# BASELINE.json configs, as (algo, per, n_step, double, B, extra)
WORKLOADS = {
    "c51_b32": dict(algo="c51", per=True, n=3, double=True, B=32, desc="configs[1]: C51 51 atoms, PER, n_step=3, double+dueling, batch 32"),
    "dqn_b32_uniform": dict(algo="dqn", per=False, n=1, double=False, B=32, desc="configs[0]: DQN, uniform replay, n_step=1, batch 32"),
    "qr_b512": dict(algo="qr", per=True, n=3, double=True, B=512, desc="configs[2]: QR-DQN 200 quantiles, PER, batch 512"),
    "iqn_b512": dict(algo="iqn", per=True, n=3, double=True, B=512, desc="configs[2]: IQN 64x64 taus, PER, batch 512"),
    "fqf_b512": dict(algo="fqf", per=True, n=3, double=True, B=512, desc="configs[3]: FQF 32 fractions, PER, batch 512"),
    "mdqn_b512": dict(algo="mdqn", per=True, n=3, double=False, B=512, desc="configs[3]: M-DQN, PER, batch 512"),
    "c51_b512": dict(algo="c51", per=True, n=3, double=True, B=512, desc="C51, PER, n_step=3, batch 512"),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="c51_b32", choices=sorted(WORKLOADS))
    p.add_argument("--ring", type=int, default=1_000_000, help="transitions per GPU shard")
    p.add_argument("--total-ring", type=int, default=0, help="if set, shard this many transitions over the GPUs (configs[4]: 8M)")
    p.add_argument("--learner-steps", type=int, default=20)
    p.add_argument("--actions", type=int, default=4)
    p.add_argument("--no-extra", action="store_true", help="skip the secondary workloads and sweeps")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--torch-rng", action="store_true", help="uniforms from torch.Tensor.uniform_ (one more launch per step) instead of the sampler's own Philox")
    p.add_argument("--variant", type=int, default=0, help="K3 variant: 0 TMA ring, 1 LDG/STG, 2 TMA full staging, 3 TMA two CTAs per transition")
    p.add_argument("--cpu-entries", type=int, default=16384, help="deque entries for the CPU baseline sample")
    p.add_argument("--cpu-workers", type=int, default=-1, help="--impl reference DataLoader workers (-1: all cores)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ our arm
def fill_shard(rp, transitions, E, seed, torch):
    from agent0_b200.synth import fill_shard_synthetic
    return fill_shard_synthetic(rp, transitions, E, seed)


def net_outputs(algo, total, A, torch, dev, seed=99):
    """Pre-generated network outputs (the CNN is excluded on both arms): N(0,1)*3 logits."""
    g = torch.Generator(device=dev).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g) * 3.0
    o = {"qsel": rn(total, A)}
    if algo in ("dqn", "mdqn"):
        o.update(online=rn(total, A), tgt_next=rn(total, A), tgt_cur=rn(total, A))
    elif algo == "c51":
        o.update(online=rn(total, A, 51), tgt_next=rn(total, A, 51), atoms=torch.linspace(-10, 10, 51, device=dev))
    elif algo == "qr":
        o.update(online=rn(total, A, 200), tgt_next=rn(total, A, 200))
    elif algo == "iqn":
        o.update(online=rn(total, 64, A), tgt_next=rn(total, 64, A), taus=torch.rand(total, 64, device=dev, generator=g))
    elif algo == "fqf":
        p = torch.softmax(torch.randn(total, 32, device=dev, generator=g), -1)
        taus = torch.cat((torch.zeros(total, 1, device=dev), torch.cumsum(p, -1)), -1)
        o.update(online=rn(total, 32, A), tgt_next=rn(total, 32, A), q_bar=rn(total, 31, A), taus=taus.contiguous(),
                 taus_hat=((taus[:, :-1] + taus[:, 1:]) / 2).contiguous())
    return o


def bytes_per_transition(n):
    """K3 reads (S+n) distinct frames and writes 2S frames per transition."""
    return (4 + min(n, 4) + 8) * F_BYTES


def make_hotpath(rp, wl, L, A, torch, variant=0):
    from agent0_b200.hotloop import ReplayTargetLoop

    class HotPath(ReplayTargetLoop):
        """agent0_b200.hotloop.ReplayTargetLoop (the package's pre-bound form of the Trainer.step inner
        loop) plus what only the benchmark needs: synthetic network outputs, and pools of pre-drawn
        index sets and rotating output buffers so that back-to-back timed gathers neither re-read
        frames nor have their stores absorbed by L2."""

        def __init__(self):
            T = L * wl["B"]
            # timed gathers rotate over enough output buffers (>= ~1 GB) that stores are not absorbed by L2
            self.n_out = min(INNER, max(1, int(1e9 // (T * 8 * F_BYTES))))
            self.frames_pool = torch.empty(self.n_out, T, 8 * F_BYTES, dtype=torch.uint8, device=rp.device)
            super().__init__(rp, wl["algo"], wl["B"], L, A, net_outputs(wl["algo"], T, A, torch, rp.device), n_step=wl["n"],
                             double_q=wl["double"], per=wl["per"], variant=variant, discount=0.99, frames=self.frames_pool[0],
                             rng_seed=RNG_SEED)
            self.wl, self.idx_pool = wl, None

        def draw_pool(self):
            """INNER independent index sets, so that back-to-back timed gathers never re-read frames."""
            self.idx_pool = torch.empty(INNER, self.total, dtype=torch.int64, device=self.dev)
            for i in range(INNER):
                self.u.uniform_(); self.sample()
                self.idx_pool[i].copy_(self.idx)

        def gather(self, count=None, variant=None, pool=None):
            if pool is None:
                return super().gather(count=count, variant=variant)
            return super().gather(self.idx_pool[pool % INNER].data_ptr(), self.frames_pool[pool % self.n_out].data_ptr(),
                                  count, variant)

        loss_k = ReplayTargetLoop.target_loss

        def step_fused_k4(self):
            self.step(fused_k4=True)

    return HotPath()


def HotPath(rp, wl, L, A, torch, variant=0):
    return make_hotpath(rp, wl, L, A, torch, variant)


def time_graphed(hp, steps, warmup, torch, use_graph, barrier, step_fn=None):
    step_fn = step_fn or hp.step
    hp.rp.push_dynamic()
    for _ in range(3):
        step_fn()
    torch.cuda.synchronize()
    runner = step_fn
    if use_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step_fn()
        runner = g.replay
    for _ in range(warmup):
        runner()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        runner()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / 1e3


EMIT = print
INNER = 20
RNG_SEED = 20261017      # the sampler draws its own uniforms (Philox inside K2a); None: torch's uniform_ + a0_pt_sample


def capture_step(hp, torch):
    hp.rp.push_dynamic()
    for _ in range(3):
        hp.step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        hp.step()
    return g


def time_kernel(fn, reps, torch):
    """Average device duration of one launch: INNER back-to-back launches ``fn(i)`` (i selects
    distinct pre-drawn inputs where that matters) captured into one CUDA graph so the host launch
    path is not on the clock; CUDA events on the launching stream around the replays."""
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(INNER):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    rounds = max(1, reps // INNER)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / (rounds * INNER)


def e2e_cabi(hp, steps, warmup, torch, graph=True, depth=2, copy_stream=True, fused=True):
    """The headline end-to-end number: one Trainer.step-shaped pass driven through the C ABI with
    HOST buffers.  Per step: new transitions (replay ratio 8 samples per insert) are ingested from
    page-locked host memory (a0_rb_ingest_steps_dyn: index update, H2D DMA, K2b marks + K1 in one
    launch, which also publishes the shard's top/beta to the device); the L batches are drawn,
    gathered, run through K4 and written back to the tree by replaying ONE CUDA graph of the
    C-ABI launches (``graph=False``: the same launches issued eagerly); the per-sample losses and
    indices reach page-locked host memory -- stored there by the K2b launch that reads them
    (a0_pt_update_report) -- and are read by the host.

    ``fused=False`` is the same loop with the separate pieces: a0_rb_ingest_steps, an
    a0_rb_set_dynamic launch, and two device-to-host copies queued behind the graph.

    depth=1: the host waits for the step's losses before it starts the next step.  depth=2: the
    result buffers are double-buffered -- the host reads step s-1's losses while the device runs
    step s (the reference's own pump is asynchronous too: 2 worker processes and a 3-deep
    prefetch queue, utils.py:59-61).  Every step's H2D input copy and D2H result read, and the
    final drain, are inside the timed region either way.  ``copy_stream``: the ingest's H2D DMA
    runs on the shard's copy stream so it overlaps the previous step's kernels.
    Wall clock around the whole loop."""
    rp, L, B = hp.rp, hp.L, hp.B
    total = hp.total
    new_per_step = max(16, total // 8)
    E = 16
    rng = np.random.RandomState(3)
    host_frames = [torch.randint(0, 256, (new_per_step, F_BYTES), dtype=torch.uint8).pin_memory() for _ in range(depth + 1)]
    streams = np.arange(new_per_step, dtype=np.int64) % E
    ones_new = np.ones(new_per_step, dtype=np.int64)
    actions = rng.randint(0, 4, new_per_step).astype(np.int64)
    zr, zd = np.zeros(new_per_step), np.zeros(new_per_step, dtype=bool)
    if fused:
        rep = hp.bind_report(depth)
        idx_host, loss_host = [r[0] for r in rep], [r[1] for r in rep]
    else:
        loss_host = [torch.empty(total, dtype=torch.float32).pin_memory() for _ in range(depth)]
        idx_host = [torch.empty(total, dtype=torch.int64).pin_memory() for _ in range(depth)]
    done_ev = [torch.cuda.Event() for _ in range(depth)]
    h2d = new_per_step * (F_BYTES + 14 * 4 + 4 + 4)
    d2h = total * (4 + 8)

    g = None
    if graph and fused:
        hp.capture_reporting()
    elif graph:
        g = capture_step(hp, torch)
    else:
        hp.graphs, hp.graph = [], None
    acc = [0.0, 0]

    def consume(slot):
        done_ev[slot].synchronize()
        acc[0] += float(loss_host[slot][0]) + float(loss_host[slot][-1])     # the host reads the result
        acc[1] = max(acc[1], int(idx_host[slot][0]))

    def run(n):
        for s in range(n):
            slot = s % depth
            if s >= depth:
                consume(slot)                       # step s-depth: its buffers are reused below
            rp.append_steps(streams, ones_new, host_frames[s % (depth + 1)], actions, zr, zd, pinned_stable=True,
                            copy_stream=copy_stream, publish_dynamic=fused)
            if fused:
                hp.run(slot=slot, publish=False)
            else:
                rp.push_dynamic()
                if g is not None:
                    g.replay()
                else:
                    hp.step()
                loss_host[slot].copy_(hp.loss, non_blocking=True)
                idx_host[slot].copy_(hp.idx, non_blocking=True)
            done_ev[slot].record()
        for s in range(max(0, n - depth), n):
            consume(s % depth)

    run(warmup)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert all(torch.isfinite(x).all() for x in loss_host) and acc[1] < rp.size and np.isfinite(acc[0])
    if fused:       # what reached the host is what the device holds
        last = (steps - 1) % depth
        assert torch.equal(loss_host[last], hp.loss.cpu()) and torch.equal(idx_host[last], hp.idx.cpu())
    return total * steps / dt, h2d, d2h


def e2e_loop(rp, wl, L, A, steps, warmup, torch):
    """Same work through the public API with host buffers: append new transitions from pinned host
    memory (K1), sample+gather, K4 per batch, update priorities, read losses+indices back."""
    from agent0_b200 import losses as LS
    dev, B, algo = rp.device, wl["B"], wl["algo"]
    total = L * B
    o = net_outputs(algo, total, A, torch, dev)
    new_per_step = max(16, total // 8)            # replay ratio: 8 samples per inserted transition
    E = 16
    rng = np.random.RandomState(3)
    host_frames = torch.randint(0, 256, (new_per_step, F_BYTES), dtype=torch.uint8).pin_memory()
    streams = np.arange(new_per_step, dtype=np.int64) % E
    gam = float(np.float32(0.99 ** wl["n"]))
    h2d = new_per_step * (F_BYTES + 14 * 4 + 4)
    d2h = total * (4 + 8)

    from agent0_b200.replay import split_batches
    os_ = {k: (v.split(B) if v.dim() and v.shape[0] == total else v) for k, v in o.items()}
    use_qsel = wl["double"] or algo in ("iqn", "fqf")
    kw = dict(max_p=rp.max_p_tensor)
    ones_new = np.ones(new_per_step, dtype=np.int64)
    zr, zd = np.zeros(new_per_step), np.zeros(new_per_step, dtype=bool)

    def one():
        # the step's inputs come from pinned host memory; the ingest DMA reads the caller's buffer directly
        rp.append_steps(streams, ones_new, host_frames, rng.randint(0, 4, new_per_step), zr, zd, pinned_stable=True)
        b = rp.sample(B, k_batches=L)
        outs = []
        for k, bk in enumerate(split_batches(b, B)):
            cm = (bk.actions, bk.rewards_f32, bk.terminals_f32, bk.weights, gam)
            qs = os_["qsel"][k] if use_qsel else None
            if algo == "dqn":
                r = LS.dqn_loss(os_["online"][k], os_["tgt_next"][k], *cm, qsel=qs, **kw)
            elif algo == "mdqn":
                r = LS.mdqn_loss(os_["online"][k], os_["tgt_next"][k], os_["tgt_cur"][k], *cm, **kw)
            elif algo == "c51":
                r = LS.c51_loss(os_["online"][k], os_["tgt_next"][k], os_["atoms"], *cm, -10.0, 10.0, qsel=qs, **kw)
            elif algo == "qr":
                r = LS.qr_loss(os_["online"][k], os_["tgt_next"][k], *cm, qsel=qs, **kw)
            elif algo == "iqn":
                r = LS.iqn_loss(os_["online"][k], os_["taus"][k], os_["tgt_next"][k], qs, *cm, **kw)
            else:
                r = LS.fqf_loss(os_["online"][k], os_["taus"][k], os_["taus_hat"][k], os_["tgt_next"][k], os_["q_bar"][k], qs,
                                *cm, **kw)
            outs.append(r.loss)
        loss = torch.cat(outs)
        rp.update_priority(b.indices, loss)
        return loss.cpu(), b.indices.cpu()       # what BaseLearner.train returns (agent.py:163-169)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return total * steps / dt, h2d, d2h


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    from agent0_b200.config import make_config
    from agent0_b200.replay import ReplayDataset
    wl = WORKLOADS[args.workload]
    L, A = args.learner_steps, args.actions
    global RNG_SEED
    if args.torch_rng:
        RNG_SEED = None
    elif RNG_SEED is not None:
        RNG_SEED += rank
    ring = args.total_ring // world if args.total_ring else args.ring
    cfg = make_config(wl["algo"], per=wl["per"], n_step=wl["n"], batch_size=wl["B"], replay_size=ring,
                      double_q=wl["double"], dueling=True, num_envs=16, action_dim=A)
    t_fill = time.perf_counter()
    rp = ReplayDataset(cfg, native_nstep=True, gather_variant=args.variant)
    fill_shard(rp, ring, 16, 1234 + rank, torch)
    t_fill = time.perf_counter() - t_fill
    hp = HotPath(rp, wl, L, A, torch, variant=args.variant)

    with ClockSampler(local) as clk:
        secs = time_graphed(hp, args.steps, args.warmup, torch, not args.no_graph, barrier)
    t = torch.tensor([secs], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t.item())
    total = hp.total
    value = total * args.steps * world / secs

    # ---- roofline of the dominant kernel (K3), same launch shape, fresh indices every launch -------
    hp.draw_pool()
    k3 = time_kernel(lambda i: hp.gather(pool=i), max(40, min(args.steps, 200)), torch)
    bpt = bytes_per_transition(wl["n"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bpt * total / k3 / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k3_traffic.json"))).get(f"{total}")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": ["a0_k3_gather_tma", "a0_k3_gather_ldg", "a0_k3_gather_tma_full", "a0_k3_gather_tma_split"][args.variant],
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback",
                "bytes_per_launch": bpt * total, "launch_us": round(k3 * 1e6, 2), "transitions_per_launch": total}

    # ---- e2e through the public API -----------------------------------------------------------------
    barrier()
    e2e_steps = max(20, args.steps // 2)
    e2e_v, h2d, d2h = e2e_cabi(hp, e2e_steps, 5, torch, graph=not args.no_graph, depth=2, copy_stream=True)
    barrier()
    e2e_sync, _, _ = e2e_cabi(hp, e2e_steps, 5, torch, graph=not args.no_graph, depth=1, copy_stream=False)
    barrier()
    e2e_eager, _, _ = e2e_cabi(hp, e2e_steps, 5, torch, graph=False, depth=2, copy_stream=True)
    barrier()
    e2e_sep, _, _ = e2e_cabi(hp, e2e_steps, 5, torch, graph=not args.no_graph, depth=2, copy_stream=True, fused=False)
    barrier()
    e2e_py, _, _ = e2e_loop(rp, wl, L, A, max(10, args.steps // 4), 3, torch)
    t = torch.tensor([e2e_v, e2e_py, e2e_eager, e2e_sync, e2e_sep], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    e2e_v, e2e_py, e2e_eager, e2e_sync, e2e_sep = (float(x) for x in t.tolist())

    extra = {}
    if world > 1:
        # the learner's only exchange step: one SUM all-reduce of the flat gradient bucket
        # (C51 dueling: 1 814 943 fp32 parameters).  Reported beside the replay+target numbers, not
        # inside them: with the CNN excluded there is no backward pass for it to overlap with.
        bucket = torch.zeros(1_814_943, dtype=torch.float32, device="cuda")
        for _ in range(5):
            dist.all_reduce(bucket)
        torch.cuda.synchronize(); barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            dist.all_reduce(bucket)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50 * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us = float(t.item())
        extra["grad_allreduce"] = {"bytes": bucket.numel() * 4, "us_per_call": round(us, 2),
                                   "bus_GBps": round(2 * (world - 1) / world * bucket.numel() * 4 / (us * 1e-6) / 1e9, 1),
                                   "per_step_us_if_not_overlapped": round(us * L, 1)}
    if not args.no_extra and rank == 0 and world == 1:
        extra = extras(rp, args, torch, peak)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(wl, L, A, args, workers=0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(secs / args.steps * 1e3, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 frames / f32 targets / f64 n-step returns",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "learner_steps_per_step": L, "transitions_per_step_per_gpu": total,
                       "ring_transitions_per_gpu": ring, "frame_ring_GB_per_gpu": round(rp.index.NF * F_BYTES / 1e9, 2),
                       "actions": A, "cuda_graph": not args.no_graph,
                       "uniforms": "torch uniform_ launch" if RNG_SEED is None else "Philox4x32-10 inside K2a (a0_pt_sample_rng)", "sharding": f"{world} independent shards, no data-path collective",
                       "l2_note": "inputs larger than L2: gathers are random reads over the multi-GB frame ring",
                       "fill_seconds": round(t_fill, 1)},
            "clocks": clk.summary(),
            "e2e": {"value": round(e2e_v, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "agent0_b200 public API over the C ABI (ReplayDataset.append_steps + hotloop.ReplayTargetLoop): ingest from pinned host buffers "
                            "(H2D DMA on the shard's copy stream; marks + append + top/beta publication in one launch) + one CUDA-graph "
                            "replay of the C-ABI launches per step; losses and indices stored by the K2b launch into double-buffered "
                            "mapped pinned host memory (a0_pt_update_report), the host reads step s-1's result while step s runs "
                            "(final drain inside the timed region)",
                    "separate_launches_value": round(e2e_sep, 1),
                    "separate_launches_path": "same loop with a0_rb_set_dynamic as its own launch and two device-to-host copies queued behind the graph",
                    "sync_every_step_value": round(e2e_sync, 1),
                    "sync_every_step_path": "same, but the host waits for each step's losses before starting the next (depth 1)",
                    "cabi_eager_value": round(e2e_eager, 1),
                    "python_api_value": round(e2e_py, 1),
                    "python_api_path": "ReplayDataset.append_steps/sample/update_priority + agent0_b200.losses wrappers"},
            "gpu_launches": hp.launches_per_step * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "extra": extra,
        }
        EMIT(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras(rp, args, torch, peak):
    """Secondary numbers: the other BASELINE configs on the same shard, and K3 GB/s by launch size."""
    out = {"workloads": {}, "k3_sweep": []}
    L, A = args.learner_steps, args.actions
    for name in ("c51_b32", "c51_b512", "qr_b512", "iqn_b512", "fqf_b512", "mdqn_b512", "dqn_b32_uniform"):
        wl = WORKLOADS[name]
        if not wl["per"]:
            continue        # the shard was built prioritized; the uniform config is a parity-test case
        hp = HotPath(rp, wl, L, A, torch, variant=args.variant)
        secs = time_graphed(hp, 50, 5, torch, not args.no_graph, lambda: None)
        secs_fused = time_graphed(hp, 50, 5, torch, not args.no_graph, lambda: None, step_fn=hp.step_fused_k4)
        hp.draw_pool()
        k4 = time_kernel(lambda i: hp.loss_k(i % L), 60, torch)
        k3 = time_kernel(lambda i: hp.gather(pool=i), 40, torch)
        k2a = time_kernel(lambda i: hp.sample(), 40, torch)
        k2b = time_kernel(lambda i: hp.update(), 40, torch)
        out["workloads"][name] = {"transitions_per_s": round(hp.total * 50 / secs, 1), "ms_per_step": round(secs / 50 * 1e3, 4),
                                  "transitions_per_s_one_k4_launch_for_all_batches": round(hp.total * 50 / secs_fused, 1),
                                  "k4_us_per_batch": round(k4 * 1e6, 2), "k3_us": round(k3 * 1e6, 2),
                                  "k3_GBps": round(bytes_per_transition(wl["n"]) * hp.total / k3 / 1e9, 1),
                                  "k2a_us": round(k2a * 1e6, 2), "k2b_us": round(k2b * 1e6, 2), "desc": wl["desc"]}
        del hp
    # K3 with the learner's input conversion fused in (a0_rb_gather_f32) against the three-pass
    # alternative it replaces: K3 to u8, then torch's .float() and .div(255) (agent.py:129-135)
    out["k3_f32"] = []
    from agent0_b200 import _lib as LB
    lib = LB.load()
    for name, count in (("c51_b32", 640), ("c51_b512", 10240)):
        wl = WORKLOADS[name]
        hp = HotPath(rp, wl, L, A, torch, variant=args.variant)
        hp.draw_pool()
        nbuf = 4 if count <= 1024 else 1
        obs = torch.empty(nbuf, count, 4 * F_BYTES, dtype=torch.float32, device=hp.dev)
        nxt = torch.empty_like(obs)

        def fused(i):
            LB.check(lib.a0_rb_gather_f32(rp.h, hp.idx_pool[i % INNER].data_ptr(), count, hp.n, 0.99, obs[i % nbuf].data_ptr(),
                                          nxt[i % nbuf].data_ptr(), 1, hp.act.data_ptr(), hp.r64.data_ptr(), hp.r32.data_ptr(),
                                          hp.d8.data_ptr(), hp.d32.data_ptr(), hp.boot.data_ptr(), hp._st()), "a0_rb_gather_f32")

        def three_pass(i):
            hp.gather(pool=i)
            x = hp.frames_pool[i % hp.n_out].view(count, 8, 84, 84).float().div(255.0)
            return x

        t_f = time_kernel(fused, 40, torch)
        t_3 = time_kernel(three_pass, 40, torch)
        by = (4 + min(wl["n"], 4)) * F_BYTES + 8 * F_BYTES * 4
        out["k3_f32"].append({"transitions": count, "fused_us": round(t_f * 1e6, 2), "three_pass_us": round(t_3 * 1e6, 2),
                              "fused_GBps": round(by * count / t_f / 1e9, 1), "frac_of_measured_peak": round(by * count / t_f / 1e9 / peak, 4),
                              "bytes_per_transition": by})
        del hp, obs, nxt
    # the drop-in ingest: ReplayDataset.extend with what the reference actor ships per Trainer.step --
    # 80 steps x 16 envs = 1280 lz4 blocks of concat(st, st_next) (agent.py:78-81, config.py:111-112)
    try:
        from agent0_b200.config import make_config
        from agent0_b200.replay import ReplayDataset
        from agent0_b200.synth import record_stream
        from oracle import cpu_path as CP, reference_replay as OR_
        s_ = record_stream(16, 4 * 80 + 2, seed=77)
        fr_, a_, r_, d_ = OR_.pack_nstep(s_["obs"], s_["action"], s_["reward"], s_["done"], 3, 0.99)
        z = CP.lz4()
        tup = [(z.compress(fr_[i].tobytes()), a_[i], r_[i], d_[i]) for i in range(4 * 1280)]
        rq = ReplayDataset(make_config("c51", per=True, n_step=3, batch_size=32, replay_size=100_000, num_envs=16))
        rq.extend(tup[:1280])
        ms = []
        for i in (1, 2, 3):                      # the first timed call still pays one-off allocations
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rq.extend(tup[i * 1280:(i + 1) * 1280])
            torch.cuda.synchronize()
            ms.append(round((time.perf_counter() - t0) * 1e3, 2))
        out["compat_extend"] = {"transitions": 1280, "ms": min(ms), "ms_all_calls": ms,
                                "note": "lz4 decode + native content de-duplication + staged H2D + K2b marks + K1, one call"}
        del rq
    except Exception as e:          # measurement only
        out["compat_extend"] = {"error": repr(e)}
    wl = dict(WORKLOADS["c51_b512"])
    hp = HotPath(rp, wl, 128, A, torch)           # buffers for up to 65536 transitions
    hp.draw_pool()
    for count in (32, 512, 640, 4096, 10240, 65536):
        for variant in (0, 3, 2, 1):
            dt = time_kernel(lambda i: hp.gather(count=count, variant=variant, pool=i), 40, torch)
            gb = bytes_per_transition(3) * count / dt / 1e9
            out["k3_sweep"].append({"transitions": count, "variant": ["tma_ring4", "ldg", "tma_full8", "tma_split2"][variant], "us": round(dt * 1e6, 2),
                                    "GBps": round(gb, 1), "frac_of_measured_peak": round(gb / peak, 4)})
    return out


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_baseline(wl, L, A, args, workers):
    """The reference's host pipeline (oracle/cpu_path.py) on a bounded sample of the same workload."""
    import torch
    from agent0_b200.synth import record_stream
    from oracle import cpu_path as CP
    cores = os.cpu_count() or 1
    torch.set_num_threads(1 if workers == 0 else cores)
    B, algo = wl["B"], wl["algo"]
    E = 16
    T = max(8, args.cpu_entries // E)
    s = record_stream(E, T, seed=1234)
    rp = CP.CpuReplay(1_000_000, wl["per"])      # 1 M-slot priority vector as in the reference
    n_entries = CP.fill_replay(rp, s, wl["n"])
    fetch, loader = CP.make_fetcher(rp, B, workers)
    o_all = net_outputs(algo, L * B, A, torch, "cpu")
    extra = dict(atoms=o_all.pop("atoms")) if algo == "c51" else {}
    if not (wl["double"] or algo in ("iqn", "fqf")):
        o_all["qsel"] = None

    def outs(it):
        sl = slice(it * B, (it + 1) * B)
        return {k: (v[sl].clone() if v is not None else None) for k, v in o_all.items()}
    gam = 0.99 ** wl["n"]
    step = lambda: CP.trainer_step(rp, fetch, outs, algo, L, gam, extra=extra)
    # calibrate to roughly 10-20 s of CPU work
    t0 = time.perf_counter(); step(); one = time.perf_counter() - t0
    steps = int(max(3, min(400, 12.0 / max(one, 1e-3))))
    n, dt = CP.time_steps(step, steps, 1)
    del loader
    return {"value": round(n / dt, 1), "unit": UNIT, "cores": 1 if workers == 0 else min(cores, workers) + 1,
            "kind": "port", "host_cores_available": cores,
            "sample": f"{steps} Trainer.step loops x {L} batches x {B} on a {n_entries}-entry lz4 deque "
                      f"({'single thread' if workers == 0 else str(workers) + ' DataLoader workers'}; "
                      f"the reference at 1 M entries pays a slower deque[idx])"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    L, A = args.learner_steps, args.actions
    cores = os.cpu_count() or 1
    workers = cores if args.cpu_workers < 0 else args.cpu_workers
    workers = max(1, min(workers, 64))
    import torch
    from agent0_b200.synth import record_stream
    from oracle import cpu_path as CP
    B, algo = wl["B"], wl["algo"]
    E = 16
    s = record_stream(E, max(8, args.cpu_entries // E), seed=1234)
    rp = CP.CpuReplay(1_000_000, wl["per"])
    n_entries = CP.fill_replay(rp, s, wl["n"])
    o_all = net_outputs(algo, L * B, A, torch, "cpu")
    extra = dict(atoms=o_all.pop("atoms")) if algo == "c51" else {}
    if not (wl["double"] or algo in ("iqn", "fqf")):
        o_all["qsel"] = None
    outs = lambda it: {k: (v[it * B:(it + 1) * B].clone() if v is not None else None) for k, v in o_all.items()}
    # The reference ships num_workers=2 and torch's default intra-op threads (= cores).  At batch 32
    # the worker IPC and thread fan-out can cost more than they buy, so the arm calibrates over
    # {in-process, 2 workers (the reference's setting), all cores} x {1, all} intra-op threads on
    # a few steps and times the fastest: the baseline is the best the host path can do here.
    cands = [(w, t) for w in sorted({0, 2, workers}) for t in sorted({1, cores})]
    if args.cpu_workers >= 0:
        cands = [(args.cpu_workers, t) for t in sorted({1, cores})]
    best, tried = None, []
    for w, t in cands:
        torch.set_num_threads(t)
        fetch, loader = CP.make_fetcher(rp, B, w)
        step = lambda: CP.trainer_step(rp, fetch, outs, algo, L, 0.99 ** wl["n"], extra=extra)
        n, dt = CP.time_steps(step, 3, 1)
        tried.append({"workers": w, "intra_op_threads": t, "transitions_per_s": round(n / dt, 1)})
        if best is None or n / dt > best[0]:
            best = (n / dt, w, t)
        del fetch, loader, step
    _, workers, threads = best
    torch.set_num_threads(threads)
    fetch, loader = CP.make_fetcher(rp, B, workers)
    step = lambda: CP.trainer_step(rp, fetch, outs, algo, L, 0.99 ** wl["n"], extra=extra)
    steps = max(1, min(args.steps, 200))
    n, dt = CP.time_steps(step, steps, max(1, min(args.warmup, 5)))
    v = round(n / dt, 1)
    used = max(threads, workers + 1)
    sample = (f"{steps} Trainer.step loops x {L} batches x {B} through the reference's DataLoader pump "
              f"({'in-process, num_workers=0' if workers == 0 else str(workers) + ' worker processes'}) on a {n_entries}-entry "
              f"lz4 deque + torch CPU loss ({threads} intra-op threads); fastest of {tried}")
    cores_used = used
    EMIT(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 frames / f32 targets / f64 n-step returns", "data": "synthetic",
        "config": {"workload": wl["desc"], "learner_steps_per_step": L, "transitions_per_step_per_gpu": L * B},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores_used, "host_cores_available": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))
    del loader


def _json_only_stdout():
    """Everything any library prints to stdout while the benchmark runs (NCCL's version banner, for
    one) goes to stderr; the process's real stdout receives exactly the one JSON line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")
    sys.stdout = sys.stderr

    def emit(line):
        out.write(line + "\n")
        out.flush()
    return emit


if __name__ == "__main__":
    a = parse()
    EMIT = _json_only_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
