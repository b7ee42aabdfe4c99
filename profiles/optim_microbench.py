"""the EMA pass (12 B per parameter) and the fused SGD + EMA pass (28 B per parameter) on the real flat buffer
(1 813 764 parameters: 7.26 MB per array, launch-latency dominated, L2-resident between calls) and on a x64 replicated buffer
(464 MB per array: a pure HBM stream) -- CUDA events over 50 calls through the C ABI:  python profiles/optim_microbench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hpfg_b200 import _lib as L

dev = torch.device("cuda:0")
lib = L.lib()
st = L.stream_ptr(dev)
PEAK = 6553.0


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


for rep in (1, 64):
    n = 1813764 * rep
    p, g, m, e = (torch.randn(n, device=dev) for _ in range(4))
    t_ema = timed(lambda: L.check(lib.hpfg_ema_update(L.ptr(e), L.ptr(p), n, 0.99, st)))
    t_sgd = timed(lambda: L.check(lib.hpfg_sgd_momentum_ema(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(e), n, 0.01, 0.9, 1e-4, 1.0, 0, 0.99, st)))
    for name, t, bpp in (("ema_kernel", t_ema, 12), ("sgd_kernel<EMA>", t_sgd, 28)):
        gbs = n * bpp / t / 1e3
        print("%-16s x%-2d  %9d params  %8.2f us per call  %7.0f GB/s algorithmic = %5.1f %% of %d" % (name, rep, n, t, gbs, 100 * gbs / PEAK, PEAK))
    del p, g, m, e
    torch.cuda.empty_cache()
