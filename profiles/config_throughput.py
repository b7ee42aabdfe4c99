"""Throughput of the fused step drivers on the other BASELINE.json configurations (1 GPU, inputs resident, eager
launches except MT which replays its graph): python profiles/config_throughput.py"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hpfg_b200 as hb

dev = torch.device("cuda:0")


def timed(fn, steps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def batch(n_l, n_u, cin, ncls):
    g = torch.Generator().manual_seed(3)
    x = torch.rand(n_l + n_u, cin, 224, 224, generator=g).to(dev)
    y = torch.randint(0, ncls, (n_l, 224, 224), generator=g).to(dev)
    return x, y


rows = []
for name, cin, ncls, n_l, n_u in [("MT ACDC 1ch/4cls 8+24 (graph replay)", 1, 4, 8, 24), ("MT ISIC 3ch/2cls 12+12 (graph replay)", 3, 2, 12, 12)]:
    torch.manual_seed(0)
    s = hb.UNet(cin, ncls).to(dev)
    t = copy.deepcopy(s)
    st = hb.MeanTeacherStep(s, t)
    st.enable_graph(True)
    x, y = batch(n_l, n_u, cin, ncls)
    ms = timed(lambda: st.step(x, y))
    rows.append((name, n_l + n_u, ms))
    del s, t, st
    torch.cuda.empty_cache()
torch.manual_seed(0)
m1, m2 = hb.UNet(1, 4).to(dev), hb.UNet(1, 4).to(dev)
cps = hb.CPSStep(m1, m2)
x, y = batch(8, 24, 1, 4)
rows.append(("CPS two UNets 1ch/4cls 8+24 (eager)", 32, timed(lambda: cps.step(x, y))))
del m1, m2, cps
torch.cuda.empty_cache()
torch.manual_seed(0)
s = hb.UNet(1, 4).to(dev)
t = hb.UNet(1, 4).to(dev)
ua = hb.UAMTStep(s, t, T=8)
x, y = batch(12, 12, 1, 4)
rows.append(("UAMT T=8 1ch/4cls 12+12 (eager)", 24, timed(lambda: ua.step(x, y), steps=10, warm=3)))
for name, n, ms in rows:
    print("%-42s %7.3f ms/step  %8.0f images/s" % (name, ms, n / ms * 1e3))
