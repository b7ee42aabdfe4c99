"""Per-role clock64 trace of CTA 0 for one tensor-core weight-gradient launch: python profiles/trace_wgrad.py N cin cout ks res"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hpfg_b200 import _lib as L
n, cin, cout, ks, res = [int(v) for v in sys.argv[1:6]]
torch.cuda.init()
ms = ctypes.c_float()
st = L.stream_ptr(torch.device("cuda:0"))
L.check(L.lib().hpfg_conv_tc_bench(2, n, res, res, cin, cout, ks, 3, ctypes.byref(ms), st))
os.environ["HPFG_WG_TRACE"] = "1"
os.environ["HPFG_WG_TRACE_DUMP"] = "1"
L.check(L.lib().hpfg_conv_tc_bench(2, n, res, res, cin, cout, ks, 1, ctypes.byref(ms), st))
print("avg us (incl. the dump sync)", ms.value * 1e3)
