import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from hpfg_b200 import _lib as L
from hpfg_b200.losses import ssl_loss_raw
dev = torch.device("cuda:0")
out = torch.randn(32, 4, 224, 224, device=dev)
t_out = torch.randn(32, 4, 224, 224, device=dev)
y = torch.randint(0, 4, (8, 224, 224), device=dev)
for _ in range(3):
    r = ssl_loss_raw(L.LOSS_MT, out, t_out[8:], y, 8, cons_weight=0.01)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    r = ssl_loss_raw(L.LOSS_MT, out, t_out[8:], y, 8, cons_weight=0.01)
e1.record()
torch.cuda.synchronize()
print("loss call avg ms (device):", e0.elapsed_time(e1) / 20)
t0 = time.perf_counter()
for _ in range(20):
    r = ssl_loss_raw(L.LOSS_MT, out, t_out[8:], y, 8, cons_weight=0.01)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("loss call avg ms (host enqueue):", (t1 - t0) / 20 * 1e3)
