"""Seeded input builders shared by make_golden.py (run against the real reference) and the tests (run
against the oracle / the CUDA path).  torch's CPU generator is deterministic for a given torch version,
so only OUTPUTS are stored in the fixtures; inputs are regenerated from these seeds on both sides."""
import torch

from oracle.unet_ref import ENC_DROPOUT, FT_CHNS, init_unet_state

ENC_PREFIXES = ["encoder.in_conv"] + ["encoder.down%d.maxpool_conv.1" % i for i in range(1, 5)]


def gen(seed):
    return torch.Generator().manual_seed(seed)


def make_state(in_ch, n_cls, seed):
    st = init_unet_state(in_ch, n_cls, generator=gen(seed))
    g = gen(seed + 1)
    for k in st:                      # non-trivial BN affine so gamma/beta gradients are exercised
        if k.endswith(".conv_conv.1.weight") or k.endswith(".conv_conv.5.weight"):
            st[k] = 0.5 + torch.rand(st[k].shape, generator=g)
        elif k.endswith(".conv_conv.1.bias") or k.endswith(".conv_conv.5.bias"):
            st[k] = 0.2 * (torch.rand(st[k].shape, generator=g) - 0.5)
    return st


def make_masks(n, h, w, seed):
    """Keep-masks (bool, NCHW) for the five encoder ConvBlocks' dropout (model/unet.py:161)."""
    g = gen(seed)
    masks = {}
    for lvl, (prefix, p) in enumerate(zip(ENC_PREFIXES, ENC_DROPOUT)):
        shape = (n, FT_CHNS[lvl], h >> lvl, w >> lvl)
        masks[prefix] = torch.rand(shape, generator=g) >= p
    return masks


def make_batch(n_l, n_u, in_ch, n_cls, h, w, seed, ignore_frac=0.0):
    g = gen(seed)
    x_l = torch.rand(n_l, in_ch, h, w, generator=g)
    x_u = torch.rand(n_u, in_ch, h, w, generator=g)
    y = torch.randint(0, n_cls, (n_l, h, w), generator=g)
    if ignore_frac > 0:
        y[torch.rand(y.shape, generator=g) < ignore_frac] = 255
    return x_l, x_u, y


def make_plus_state(in_ch, n_cls, seed):
    """make_state + the 16 projection-neck parameters of UNet_Plus (uniform +-1/sqrt(fan_in), seeded)."""
    from oracle.unet_ref import unet_plus_neck_spec
    st = make_state(in_ch, n_cls, seed)
    g = gen(seed + 2)
    for name, shape in unet_plus_neck_spec(n_cls):
        fan_in = shape[1] if len(shape) > 1 else shape[0]
        b = 1.0 / (fan_in ** 0.5)
        st[name] = (2 * torch.rand(shape, generator=g) - 1) * b
    return st


def make_predict_case(in_ch, n_cls, n, h, w, seed):
    """State + volume for the inference fixture: the output conv is scaled up and the slices are smooth blobs so that a
    random-init network predicts a mix of classes (BN running buffers come from the fixture)."""
    st = make_state(in_ch, n_cls, seed)
    wo = st["decoder.out_conv.weight"]
    st["decoder.out_conv.weight"] = (wo - wo.mean(dim=(1, 2, 3), keepdim=True)) * 40.0
    st["decoder.out_conv.bias"] = torch.zeros_like(st["decoder.out_conv.bias"])
    g = gen(seed + 50)
    coarse = torch.rand(n, 1, 4, 6, generator=g)
    vol = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)[:, 0]
    vol = (vol + 0.05 * torch.rand(n, h, w, generator=g)).contiguous()
    return st, vol


def summarize(t, stride=97, full_below=20000):
    """Compact fingerprint of a tensor: full copy if small, else strided sample + sums (fp64)."""
    t = t.detach().double().flatten()
    d = {"numel": t.numel(), "sum": t.sum().item(), "abs_sum": t.abs().sum().item(),
         "sq_sum": (t * t).sum().item()}
    if t.numel() <= full_below:
        d["full"] = t.float().clone()
    else:
        d["sample"] = t[::stride].float().clone()
        d["stride"] = stride
    return d
