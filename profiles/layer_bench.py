"""Per-layer micro-benchmark of the tensor-core convolution kernels (CUDA-event timed, warm, back-to-back launches).
usage: python profiles/layer_bench.py [N]           -> table for every UNet layer shape at batch N (default 32)
       python profiles/layer_bench.py N op cin cout ks res iters   -> one shape (for ncu captures)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from hpfg_b200 import _lib as L  # noqa: E402

LAYERS = [(16, 16, 3, 224), (32, 16, 3, 224), (16, 32, 3, 112), (32, 32, 3, 112), (64, 32, 3, 112), (32, 64, 3, 56),
          (64, 64, 3, 56), (128, 64, 3, 56), (64, 128, 3, 28), (128, 128, 3, 28), (256, 128, 3, 28), (128, 256, 3, 14),
          (256, 256, 3, 14), (256, 128, 1, 14), (128, 64, 1, 28), (64, 32, 1, 56), (32, 16, 1, 112)]


def run(op, n, cin, cout, ks, res, iters=20):
    ms = ctypes.c_float()
    L.check(L.lib().hpfg_conv_tc_bench(op, n, res, res, cin, cout, ks, iters, ctypes.byref(ms), L.stream_ptr(torch.device("cuda:0"))))
    return ms.value


if __name__ == "__main__":
    torch.cuda.init()
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        # bottleneck isolation: python profiles/layer_bench.py sweep N "cin cout ks res" ... ; flags 1 no MMA, 2 no stores, 4 no stats, 8 no TMA
        n = int(sys.argv[2])
        for shape in sys.argv[3:]:
            cin, cout, ks, res = [int(v) for v in shape.split()]
            for op in (0, 1):
                row = []
                for d in [int(v) for v in os.environ.get("SWEEP_FLAGS", "0,1,2,4,8,3,9,11,15").split(",")]:
                    os.environ["HPFG_TC_DBG"] = str(d)
                    row.append("%d:%.1f" % (d, 1e3 * run(op, n, cin, cout, ks, res, 20)))
                os.environ["HPFG_TC_DBG"] = "0"
                print("(%s) op=%d  dbg:us  %s" % (shape, op, "  ".join(row)), flush=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fused":
        # BatchNorm-backward fusions: dgrad plain / two-source loader / GSTAT epilogue / both, wgrad plain / two-source dY
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
        print("batch %d; us per launch" % n)
        print("%-22s %9s %9s %9s %9s %9s %9s" % ("layer (cin,cout,k,res)", "dgrad", "dg+2src", "dg+gstat", "dg+both", "wgrad", "wg+2src"))
        for cin, cout, ks, res in LAYERS:
            ops = (1, 3, 4, 5, 2, 6) if ks == 3 else (1, 1, 4, 4, 2, 2)
            t = [1e3 * run(op, n, cin, cout, ks, res) for op in ops]
            print("%-22s %9.1f %9.1f %9.1f %9.1f %9.1f %9.1f" % ((str((cin, cout, ks, res)),) + tuple(t)), flush=True)
        sys.exit(0)
    if len(sys.argv) > 2:
        n, op, cin, cout, ks, res, iters = [int(v) for v in sys.argv[1:8]]
        print("%.2f us" % (1e3 * run(op, n, cin, cout, ks, res, iters)))
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    print("batch %d; us per launch; HBM-ideal = bf16 in+out once at 6553 GB/s; TC-ideal at 1366 TFLOP/s" % n)
    print("%-22s %9s %9s %9s %10s %9s" % ("layer (cin,cout,k,res)", "fprop", "dgrad", "wgrad", "hbm-ideal", "tc-ideal"))
    for cin, cout, ks, res in LAYERS:
        t = [1e3 * run(op, n, cin, cout, ks, res) for op in (0, 1, 2)]
        px = n * res * res
        hbm = px * (cin + cout) * 2 / 6553e9 * 1e6
        tc = 2.0 * px * cin * cout * ks * ks / 1366.1e12 * 1e6
        print("%-22s %9.1f %9.1f %9.1f %10.1f %9.1f" % (str((cin, cout, ks, res)), t[0], t[1], t[2], hbm, tc))
