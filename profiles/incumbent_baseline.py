"""SURVEY 8d "second baseline": the reference's own module graph and step body in plain torch eager on the same B200
(what running the unmodified trainer with cuda: True does), next to this repo's fused step.

The torch forward goes through the SAME nn.Module tree the reference builds (hpfg_b200.UNet registers exactly those
Conv2d / BatchNorm2d / LeakyReLU / Dropout / MaxPool2d / Upsample modules to mirror its state_dict; this script calls
them the way model/unet.py:12-117 wires them), the losses are the torch expressions of utils/loss/medloss.py, the
optimiser is torch.optim.SGD and the EMA the per-parameter loop of utils/utils.py:82-86.  Measurement only.

    python profiles/incumbent_baseline.py [steps]        # fp32, then autocast(bf16) + channels_last
"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import hpfg_b200 as hb

N_L, N_U, C, H, W = 8, 24, 4, 224, 224


def torch_forward(m, x):
    """model/unet.py: Encoder.forward :76-82, UpBlock.forward :53-58, Decoder.forward :101-117 on m's torch submodules."""
    e = m.encoder
    feats = [e.in_conv.conv_conv(x)]
    for i in range(1, 5):
        mp = getattr(e, "down%d" % i).maxpool_conv                 # Sequential(MaxPool2d(2), ConvBlock holder)
        feats.append(mp[1].conv_conv(mp[0](feats[-1])))
    d = m.decoder
    h = feats[4]
    for i in range(1, 5):
        up = getattr(d, "up%d" % i)
        h = up.conv.conv_conv(torch.cat([feats[4 - i], up.up(up.conv1x1(h))], dim=1))
    return d.out_conv(h)


def dice(probs, target, n_classes):
    loss = 0.0
    for i in range(n_classes):                                   # utils/loss/medloss.py:18-26,28-41
        s, t = probs[:, i], (target == i).float()
        loss = loss + (1 - (2 * (s * t).sum() + 1e-5) / ((s * s).sum() + (t * t).sum() + 1e-5))
    return loss / n_classes


def run(dev, steps, warm, amp, h=H, w=W, n_l=N_L, n_u=N_U, in_ch=1, n_cls=C):
    torch.manual_seed(0)
    C = n_cls
    model = hb.UNet(in_ch, n_cls).to(dev)
    ema = copy.deepcopy(model)
    for p in ema.parameters():
        p.requires_grad = False
    model.train(), ema.train()
    if amp:
        model, ema = model.to(memory_format=torch.channels_last), ema.to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n_l + n_u, in_ch, h, w, generator=g).to(dev)
    y = torch.randint(0, C, (n_l, h, w), generator=g).to(dev)
    if amp:
        x = x.contiguous(memory_format=torch.channels_last)
    cuda = dev.type == "cuda"

    def step(it):
        with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=amp):
            out = torch_forward(model, x)
            with torch.no_grad():
                t_out = torch_forward(ema, x)
        out, t_out = out.float(), t_out.float()
        soft, t_soft = torch.softmax(out, 1), torch.softmax(t_out, 1)
        sup = 0.5 * F.cross_entropy(out[:n_l], y, ignore_index=255) + 0.5 * dice(soft[:n_l], y, C)   # medloss.py:54-56
        cons = torch.mean((soft[n_l:] - t_soft[n_l:]) ** 2)                                           # 2017_03...:104
        loss = sup + 0.1 * cons
        opt.zero_grad()
        loss.backward()
        opt.step()
        alpha = min(1 - 1 / (it + 1), 0.99)
        for ep, p in zip(ema.parameters(), model.parameters()):                                       # utils/utils.py:85-86
            ep.data.mul_(alpha).add_(p.data, alpha=1 - alpha)
        return loss

    for it in range(1, warm + 1):
        step(it)
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    import time
    t0 = time.perf_counter()
    for it in range(warm + 1, warm + steps + 1):
        loss = step(it)
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    else:
        ms = (time.perf_counter() - t0) / steps * 1e3
    return ms, float(loss.detach())


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
    small = dev.type == "cpu"
    kw = dict(h=32, w=32, n_l=2, n_u=2) if small else {}
    for amp in (False, True):
        ms, loss = run(dev, steps, 3, amp, **kw)
        n = (kw.get("n_l", N_L) + kw.get("n_u", N_U))
        print("torch eager %-28s %8.3f ms/step  %8.0f images/s  (loss %.4f; cudnn.benchmark %s, conv TF32 %s)"
              % ("autocast bf16 + channels_last" if amp else "fp32", ms, n / ms * 1e3, loss, torch.backends.cudnn.benchmark,
                 torch.backends.cudnn.allow_tf32), flush=True)
