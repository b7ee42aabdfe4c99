"""Throughput of the fused step drivers on the other BASELINE.json configurations and the SURVEY 8f rows (1 GPU, inputs
resident; eager launches and CUDA-graph replay side by side): python profiles/config_throughput.py"""
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
for graph in (False, True):
    tag = "graph replay" if graph else "eager"
    torch.manual_seed(0)
    m1, m2 = hb.UNet(1, 4).to(dev), hb.UNet(1, 4).to(dev)
    cps = hb.CPSStep(m1, m2)
    cps.enable_graph(graph)
    x, y = batch(8, 24, 1, 4)
    rows.append(("CPS two UNets 1ch/4cls 8+24 (%s)" % tag, 32, timed(lambda: cps.step(x, y))))
    del m1, m2, cps
    torch.cuda.empty_cache()
    torch.manual_seed(0)
    s = hb.UNet(1, 4).to(dev)
    t = hb.UNet(1, 4).to(dev)
    ua = hb.UAMTStep(s, t, T=8)
    ua.enable_graph(graph)
    x, y = batch(12, 12, 1, 4)
    rows.append(("UAMT T=8 1ch/4cls 12+12 (%s)" % tag, 24, timed(lambda: ua.step(x, y), steps=10, warm=3)))
    del s, t, ua
    torch.cuda.empty_cache()
    torch.manual_seed(0)
    s = hb.UNet(1, 4).to(dev)
    t = hb.UNet(1, 4).to(dev)
    ict = hb.ICTStep(s, t)
    ict.enable_graph(graph)
    x, y = batch(8, 24, 1, 4)
    lam = torch.rand(12)
    rows.append(("ICT 1ch/4cls 8+24 (%s)" % tag, 32, timed(lambda: ict.step(x, y, lam))))
    del s, t, ict
    torch.cuda.empty_cache()
# inference path (val.py:268-281): a 16-slice volume, eval-mode forward + fused argmax, vs the slice-by-slice loop
torch.manual_seed(0)
m = hb.UNet(1, 4).to(dev)
vol = torch.rand(16, 224, 224, device=dev)
rows.append(("predict_volume 16 slices, one batch", 16, timed(lambda: hb.predict_volume(m, vol))))
rows.append(("predict_volume 16 slices, batch 1 loop", 16, timed(lambda: hb.predict_volume(m, vol, max_batch=1), steps=5, warm=2)))
# loss / argmax kernels alone at the YAML shape (HBM-bound; algorithmic bytes as in DESIGN.md section 3)
lg = torch.randn(32, 4, 224, 224, device=dev)
ms = timed(lambda: hb.argmax_labels(lg, dtype=torch.uint8), steps=50)
print("argmax_labels 32x4x224x224 -> u8: %.2f us, %.0f GB/s algorithmic" % (ms * 1e3, (lg.numel() * 4 + lg.numel() // 4) / ms / 1e6))
t2 = torch.randn(24, 4, 224, 224, device=dev)
yl = torch.randint(0, 4, (8, 224, 224), device=dev)
lamd = torch.rand(12, device=dev)
ms = timed(lambda: hb.ict_loss_raw(lg[:20], t2, lamd, yl, 8, cons_weight=0.1), steps=50)
ict_bytes = (20 * 2 + 24 * 2) * 4 * 50176 * 4 + 20 * 4 * 50176 * 4 + 2 * 8 * 50176 * 8
print("ict_loss 8+12 (2 launches + memset + torch.empty): %.2f us, %.0f GB/s algorithmic (logits read twice)" % (ms * 1e3, ict_bytes / ms / 1e6))
for name, n, ms in rows:
    print("%-42s %7.3f ms/step  %8.0f images/s" % (name, ms, n / ms * 1e3))
