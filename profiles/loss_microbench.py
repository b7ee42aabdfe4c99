"""Device time of the fused loss launches alone (Mean-Teacher mode at the YAML shape 8+24, 4 classes, 224x224), buffers
pre-allocated, 200 back-to-back calls through the C ABI:  HPFG_LOSS_CTAS_PER_SM=<n> python profiles/loss_microbench.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hpfg_b200 import _lib as L

dev = torch.device("cuda:0")
n_l, n_u, C, H, W = 8, 24, 4, 224, 224
g = torch.Generator().manual_seed(1)
s = (2 * torch.randn(n_l + n_u, C, H, W, generator=g)).to(dev)
t = (2 * torch.randn(n_u, C, H, W, generator=g)).to(dev)
y = torch.randint(0, C, (n_l, H, W), generator=g).to(dev)
ds, sc = torch.empty_like(s), torch.empty(8, device=dev)
ws = torch.empty(L.lib().hpfg_ssl_loss_workspace_bytes(L.LOSS_MT, n_l, n_u, C, H, W), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = L.stream_ptr(dev)


def call():
    L.check(L.lib().hpfg_ssl_loss(L.LOSS_MT, L.ptr(s), L.ptr(t), None, 0, L.ptr(y), n_l, n_u, C, H, W, 0.05, 0.0, None, 0.5, 0.5,
                                  L.ptr(ds), None, L.ptr(sc), None, None, L.ptr(ws), L.stream_ptr(dev)))


for _ in range(20):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot, iters = 0.0, 50
for _ in range(iters):           # L2 flushed between calls (the 45 MB of logits would otherwise sit in the 126 MB L2)
    flush.zero_()
    e0.record()
    call()
    e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
us = tot / iters * 1e3
algo = (n_l + n_u) * C * H * W * 4 * 2 + n_u * C * H * W * 4 + n_l * H * W * 8      # every tensor once
# host-independent number: 20 calls captured into one CUDA graph, replayed (L2 warm: the 45 MB of logits stay resident,
# which is also the in-step situation, where the conv kernel has just written them)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(20):
        call()
gr.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    gr.replay()
e1.record()
torch.cuda.synchronize()
us_graph = e0.elapsed_time(e1) / 200 * 1e3
print("graph replay, L2 warm: %.2f us per call (memset + reduce + grad) -> %.0f GB/s algorithmic (%.1f %% of 6553)"
      % (us_graph, algo / us_graph / 1e3, 100 * algo / us_graph / 1e3 / 6553))
print("HPFG_LOSS_CTAS_PER_SM=%s: memset + reduce + grad = %.2f us per call (L2 flushed), %.1f MB algorithmic -> %.0f GB/s (%.1f %% of 6553)"
      % (os.environ.get("HPFG_LOSS_CTAS_PER_SM", "default"), us, algo / 1e6, algo / us / 1e3, 100 * algo / us / 1e3 / 6553), "loss", sc[0].item())
