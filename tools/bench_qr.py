"""QR-200 K4 alone: the O(N log N) sorted-target kernel against the O(N^2) pair loop, batch 512, A = 4.
CUDA-graph of 20 launches over distinct inputs, CUDA events."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agent0_b200 import _lib, losses as L  # noqa: E402

lib = _lib.load()
B, A, N, K = 512, 4, 200, 20
g = torch.Generator(device="cuda").manual_seed(3)
q = torch.randn(K, B, A, N, device="cuda", generator=g) * 3
tn = torch.randn(K, B, A, N, device="cuda", generator=g) * 3
qs = torch.randn(K, B, A, device="cuda", generator=g)
a = torch.randint(0, A, (B,), device="cuda")
r = torch.randn(B, device="cuda"); d = (torch.rand(B, device="cuda") < 0.1).float(); w = torch.rand(B, device="cuda") + 0.1
gam = float(np.float32(0.99 ** 3))
res = {}
for name, opt in (("sorted_cta_per_sample", 1), ("sorted_warp_per_sample", 2), ("pairwise", 0)):
    lib.a0_set_option(10, opt)
    from agent0_b200.hotloop import ReplayTargetLoop  # noqa: F401  (same C entry point; here through the wrappers)
    outs = [L.qr_loss(q[i], tn[i], a, r, d, w, gam, qsel=qs[i]) for i in range(2)]
    torch.cuda.synchronize()
    loss = torch.empty(B, device="cuda"); grad = torch.empty(B, A, N, device="cuda")
    c = _lib.LossCommon(B=B, A=A, action=a.data_ptr(), reward=r.data_ptr(), done=d.data_ptr(), weight=w.data_ptr(), gamma_n=gam,
                        alpha=0.5, eps=0.01, loss=loss.data_ptr(), prio=None, max_p=None)
    import ctypes as C

    def launch(i):
        _lib.check(lib.a0_loss_quantile(C.byref(c), 0, q[i].data_ptr(), tn[i].data_ptr(), None, qs[i].data_ptr(), N, N, grad.data_ptr(),
                                        None, None, None, None, _lib.stream_ptr()), "a0_loss_quantile")
    launch(0); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(K):
            launch(i)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    res[name] = {"us_per_batch_of_512": round(e0.elapsed_time(e1) * 1e3 / (50 * K), 3), "loss0": float(outs[0].loss[0])}
lib.a0_set_option(10, 1)
print(json.dumps(res))
