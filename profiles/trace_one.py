"""Per-role clock64 trace of CTA 0 for one conv launch: python profiles/trace_one.py N op cin cout ks res [dbg]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hpfg_b200 import _lib as L
n, op, cin, cout, ks, res = [int(v) for v in sys.argv[1:7]]
os.environ["HPFG_TC_DBG"] = sys.argv[7] if len(sys.argv) > 7 else "0"
torch.cuda.init()
ms = ctypes.c_float()
st = L.stream_ptr(torch.device("cuda:0"))
os.environ["HPFG_TC_TRACE"] = "1"
L.check(L.lib().hpfg_conv_tc_bench(op, n, res, res, cin, cout, ks, 3, ctypes.byref(ms), st))
os.environ["HPFG_TC_TRACE_DUMP"] = "1"
L.check(L.lib().hpfg_conv_tc_bench(op, n, res, res, cin, cout, ks, 1, ctypes.byref(ms), st))
print("avg us", ms.value * 1e3)
