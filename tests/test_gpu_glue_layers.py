"""GPU, layer-isolated: every bf16 "glue" kernel of the backward / forward chains (activated max-pool, bilinear-upsample +
concat, BatchNorm backward, pooling / upsampling adjoints) and the BatchNorm-backward fusions of the tensor-core kernels
(two-source loaders, GSTAT epilogue), each against a plain PyTorch fp32 reference of the same op fed the SAME bf16-rounded
inputs -- the north_star's "per-layer activations and gradients agree within 1e-2 relative error in bf16" bar, for the
kernels the end-to-end bounds of tests/test_gpu_parity.py cannot isolate.  All calls go through the C ABI
(hpfg_glue_debug, hpfg_dgrad_tc_fused_debug, hpfg_wgrad_tc_fused_debug)."""
import copy

import pytest
import torch
import torch.nn.functional as F

import hpfg_b200 as hb
from hpfg_b200 import _lib as L
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-2
SLOPE = 0.01


def _bf(t):
    return t.to(torch.bfloat16).float()


def _nhwc(t):          # NCHW fp32 -> NHWC bf16 device tensor
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(DEV)


def _nchw(t):          # NHWC bf16 -> NCHW fp32 cpu
    return t.float().permute(0, 3, 1, 2).cpu()


def _glue(op, n, h, w, c, a, b=None, cc=None, scale=None, shift=None, mean=None, invstd=None, mask=None, p=0.0, out_shape=None,
          n_f32=0):
    out = torch.empty(out_shape, device=DEV, dtype=torch.bfloat16)
    f32 = torch.zeros(max(n_f32, 1), device=DEV)
    d = lambda t: None if t is None else t.to(DEV).contiguous()
    sc, sh, mu, iv, mk = d(scale), d(shift), d(mean), d(invstd), d(mask)
    L.check(L.lib().hpfg_glue_debug(op, n, h, w, c, L.ptr(a), L.ptr(b), L.ptr(cc), L.ptr(sc), L.ptr(sh), L.ptr(mu), L.ptr(iv), L.ptr(mk),
                                    float(p), L.ptr(out), L.ptr(f32), L.stream_ptr(torch.device(DEV))), "hpfg_glue_debug")
    return out, f32.cpu()


def _affine(c, g):
    return 0.5 + torch.rand(c, generator=g), 0.3 * torch.randn(c, generator=g)


@pytest.mark.parametrize("c,n,h,w", [(16, 2, 32, 48), (64, 3, 24, 40), (128, 2, 16, 16)])
def test_pool_act(c, n, h, w):
    g = torch.Generator().manual_seed(c + h)
    raw = _bf(torch.randn(n, c, h, w, generator=g))
    sc, sh = _affine(c, g)
    got, _ = _glue(0, n, h, w, c, _nhwc(raw), scale=sc, shift=sh, out_shape=(n, h // 2, w // 2, c))
    ref = F.max_pool2d(F.leaky_relu(raw * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1), SLOPE), 2)
    assert rel_l2(_nchw(got), ref) < TOL


@pytest.mark.parametrize("c,n,h,w", [(16, 2, 28, 28), (32, 2, 14, 20), (128, 3, 7, 7)])
def test_upcat(c, n, h, w):
    g = torch.Generator().manual_seed(c + w)
    skip = _bf(torch.randn(n, c, 2 * h, 2 * w, generator=g))
    low = _bf(torch.randn(n, c, h, w, generator=g))
    sc, sh = _affine(c, g)
    got, _ = _glue(1, n, h, w, c, _nhwc(skip), _nhwc(low), scale=sc, shift=sh, out_shape=(n, 2 * h, 2 * w, 2 * c))
    act = F.leaky_relu(skip * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1), SLOPE)
    up = F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=True)
    assert rel_l2(_nchw(got), torch.cat([act, up], 1)) < TOL


def _bn_bwd_ref(dact, raw, sc, sh, mean, invstd, mask, p):
    v = lambda t: t.view(1, -1, 1, 1)
    z = raw * v(sc) + v(sh)
    g = dact * torch.where(z > 0, 1.0, SLOPE)
    if mask is not None:
        g = g * mask.float() / (1.0 - p)
    xhat = (raw - v(mean)) * v(invstd)
    c1, c2 = g.mean(dim=(0, 2, 3)), (g * xhat).mean(dim=(0, 2, 3))
    draw = v(sc) * (g - v(c1) - xhat * v(c2))
    return g, draw, (g * xhat).sum(dim=(0, 2, 3)), g.sum(dim=(0, 2, 3)), c1, c2


@pytest.mark.parametrize("c,n,h,w,drop", [(16, 2, 40, 24, 0.05), (32, 2, 24, 24, 0.0), (256, 4, 14, 14, 0.5), (64, 2, 28, 20, 0.0)])
def test_bn_bwd(c, n, h, w, drop):
    g = torch.Generator().manual_seed(c * 7 + h)
    dact = _bf(torch.randn(n, c, h, w, generator=g))
    raw = _bf(torch.randn(n, c, h, w, generator=g) * 1.5 + 0.2)
    mean, var = raw.mean(dim=(0, 2, 3)), raw.var(dim=(0, 2, 3), unbiased=False)
    invstd = (var + 1e-5).rsqrt()
    gamma, beta = _affine(c, g)
    sc, sh = gamma * invstd, beta - mean * gamma * invstd
    mask = (torch.rand(n, c, h, w, generator=g) >= drop).to(torch.uint8) if drop > 0 else None
    got, f32 = _glue(2, n, h, w, c, _nhwc(dact), _nhwc(raw), scale=sc, shift=sh, mean=mean, invstd=invstd, mask=mask, p=drop,
                     out_shape=(n, h, w, c), n_f32=2 * c)
    _, draw, dgamma, dbeta, _, _ = _bn_bwd_ref(dact, raw, sc, sh, mean, invstd, mask, drop)
    assert rel_l2(_nchw(got), draw) < TOL
    assert rel_l2(f32[:c], dgamma) < 1e-3 and rel_l2(f32[c:2 * c], dbeta) < 1e-3


@pytest.mark.parametrize("c,n,h,w", [(16, 2, 32, 48), (64, 2, 28, 28), (128, 2, 14, 14)])
@pytest.mark.parametrize("fused", [False, True])
def test_skip_pool_bwd(c, n, h, w, fused):
    g = torch.Generator().manual_seed(c + 3 * h + int(fused))
    dcat = _bf(torch.randn(n, 2 * c, h, w, generator=g))
    dpooled = _bf(torch.randn(n, c, h // 2, w // 2, generator=g))
    raw = _bf(torch.randn(n, c, h, w, generator=g))
    mean, var = raw.mean(dim=(0, 2, 3)), raw.var(dim=(0, 2, 3), unbiased=False)
    invstd = (var + 1e-5).rsqrt()
    gamma, beta = _affine(c, g)
    sc, sh = gamma * invstd, beta - mean * gamma * invstd
    act = F.leaky_relu(raw * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1), SLOPE)
    _, idx = F.max_pool2d(act, 2, return_indices=True)
    dact = dcat[:, :c] + F.max_unpool2d(dpooled, idx, 2, output_size=(h, w))
    got, f32 = _glue(5 if fused else 3, n, h, w, c, _nhwc(dcat), _nhwc(dpooled), _nhwc(raw), scale=sc, shift=sh, mean=mean, invstd=invstd,
                     out_shape=(n, h, w, c), n_f32=4 * c)
    if not fused:
        assert rel_l2(_nchw(got), dact) < TOL
        return
    gg, _, dgamma, dbeta, c1, c2 = _bn_bwd_ref(dact, raw, sc, sh, mean, invstd, None, 0.0)
    assert rel_l2(_nchw(got), gg) < TOL
    assert rel_l2(f32[:c], dgamma) < 2e-3 and rel_l2(f32[c:2 * c], dbeta) < 2e-3
    kb = -sc * c2 * invstd
    kd = -sc * c1 - kb * mean
    assert rel_l2(f32[2 * c:3 * c], kb) < 2e-3 and rel_l2(f32[3 * c:4 * c], kd) < 2e-3


@pytest.mark.parametrize("c,n,h,w", [(16, 2, 28, 28), (32, 2, 14, 20), (128, 3, 7, 7)])
def test_up_bwd(c, n, h, w):
    g = torch.Generator().manual_seed(c + 5 * w)
    dcat = _bf(torch.randn(n, 2 * c, 2 * h, 2 * w, generator=g))
    low = torch.zeros(n, c, h, w, requires_grad=True)
    F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=True).backward(dcat[:, c:])
    got, _ = _glue(4, n, h, w, c, _nhwc(dcat), out_shape=(n, h, w, c))
    assert rel_l2(_nchw(got), low.grad) < TOL


# ---------------------------------------------------------------- BatchNorm-backward fusions of the tensor-core kernels
FUSED_SHAPES = [(16, 16), (32, 16), (32, 32), (64, 32), (64, 64), (128, 64), (128, 128), (256, 128), (256, 256), (16, 32), (128, 256)]


def _consts(c, g):
    return 0.5 + torch.rand(c, generator=g), 0.2 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)


@pytest.mark.parametrize("cin,cout", FUSED_SHAPES)
@pytest.mark.parametrize("two,gstat,drop", [(True, False, 0.0), (False, True, 0.3), (True, True, 0.0), (True, True, 0.1)])
def test_dgrad_fused(cin, cout, two, gstat, drop):
    n, h, w = 2, 24, 20                                     # ragged: neither a multiple of the 16 x 8 pixel tile
    g = torch.Generator().manual_seed(cin * 1000 + cout + int(two) * 2 + int(gstat))
    gin = _bf(torch.randn(n, cout, h, w, generator=g))
    raw_in = _bf(torch.randn(n, cout, h, w, generator=g))
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    sc, kb, kd = _consts(cout, g)
    v = lambda t: t.view(1, -1, 1, 1)
    draw = _bf(v(sc) * gin + v(kb) * raw_in + v(kd)) if two else gin            # the kernel rounds the operand to bf16
    din = F.conv_transpose2d(draw, _bf(wt), padding=1)
    raw_out = _bf(torch.randn(n, cin, h, w, generator=g))
    gsc, gsh = _affine(cin, g)
    mask = (torch.rand(n, cin, h, w, generator=g) >= drop).to(torch.uint8) if drop > 0 else None
    ref = din
    if gstat:
        z = raw_out * v(gsc) + v(gsh)
        ref = din * torch.where(z > 0, 1.0, SLOPE)
        if mask is not None:
            ref = ref * mask.float() / (1.0 - drop)
    out = torch.empty((n, h, w, cin), device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * cin, device=DEV)
    d = lambda t: None if t is None else t.to(DEV).contiguous()
    a = dict(sc=d(sc), kb=d(kb), kd=d(kd), w=d(wt), gsc=d(gsc), gsh=d(gsh), mask=d(mask), gin=_nhwc(gin), rin=_nhwc(raw_in), rout=_nhwc(raw_out))
    L.check(L.lib().hpfg_dgrad_tc_fused_debug(n, h, w, cin, cout, 3, L.ptr(a["gin"]), L.ptr(a["rin"]) if two else None, L.ptr(a["sc"]),
                                              L.ptr(a["kb"]), L.ptr(a["kd"]), L.ptr(a["w"]), L.ptr(a["rout"]) if gstat else None,
                                              L.ptr(a["gsc"]), L.ptr(a["gsh"]), L.ptr(a["mask"]), float(drop), L.ptr(out), L.ptr(stats),
                                              L.stream_ptr(torch.device(DEV))), "hpfg_dgrad_tc_fused_debug")
    assert rel_l2(_nchw(out), ref) < TOL
    if gstat:
        s1, s2 = ref.sum(dim=(0, 2, 3)), (ref * raw_out).sum(dim=(0, 2, 3))
        scale = ref.abs().sum(dim=(0, 2, 3))                 # sums of signed values: compare on the scale of the summed magnitudes
        assert ((stats[:cin].cpu() - s1).abs() / scale).max().item() < 2e-3
        assert ((stats[cin:].cpu() - s2).abs() / (ref * raw_out).abs().sum(dim=(0, 2, 3))).max().item() < 2e-3


@pytest.mark.parametrize("cin,cout", FUSED_SHAPES)
def test_wgrad_fused(cin, cout):
    n, h, w = 2, 24, 20
    g = torch.Generator().manual_seed(cin * 31 + cout)
    x = _bf(torch.randn(n, cin, h, w, generator=g))
    gin = _bf(torch.randn(n, cout, h, w, generator=g))
    raw = _bf(torch.randn(n, cout, h, w, generator=g))
    sc, kb, kd = _consts(cout, g)
    xs, xh = _affine(cin, g)
    v = lambda t: t.view(1, -1, 1, 1)
    draw = _bf(v(sc) * gin + v(kb) * raw + v(kd))
    xa = _bf(F.leaky_relu(x * v(xs) + v(xh), SLOPE))           # the loader rounds the activated operand to bf16
    wt = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    bias = torch.zeros(cout, requires_grad=True)
    F.conv2d(xa, wt, bias, padding=1).backward(draw)
    dw = torch.empty(cout, cin, 3, 3, device=DEV)
    db = torch.empty(cout, device=DEV)
    d = lambda t: t.to(DEV).contiguous()
    a = [d(sc), d(kb), d(kd), d(xs), d(xh), _nhwc(x), _nhwc(gin), _nhwc(raw)]
    L.check(L.lib().hpfg_wgrad_tc_fused_debug(n, h, w, cin, cout, 3, L.ptr(a[5]), L.ptr(a[6]), L.ptr(a[7]), L.ptr(a[0]), L.ptr(a[1]),
                                              L.ptr(a[2]), L.ptr(a[3]), L.ptr(a[4]), L.ptr(dw), L.ptr(db),
                                              L.stream_ptr(torch.device(DEV))), "hpfg_wgrad_tc_fused_debug")
    assert rel_l2(dw.cpu(), wt.grad) < 3e-3
    assert rel_l2(db.cpu(), bias.grad) < 3e-3


@pytest.mark.parametrize("in_ch,n_cls,n,h,w", [(1, 4, 3, 48, 80), (3, 2, 2, 64, 64)])
def test_fused_backward_matches_unfused_backward(in_ch, n_cls, n, h, w):
    """Whole network: the backward with BatchNorm backward folded into the dgrad / wgrad kernels against the streaming
    bn_bwd schedule, same weights / masks / batch.  Both are bf16 paths with the same math up to rounding order."""
    from tests.golden.common import make_state, make_masks, make_batch
    st = make_state(in_ch, n_cls, 55)
    x, _, y = make_batch(n, 0, in_ch, n_cls, h, w, 56)
    masks = make_masks(n, h, w, 57)
    grads = {}
    for fused in (False, True):
        m = hb.UNet(in_ch, n_cls, precision="bf16")
        m.load_state_dict(st)
        m = m.to(DEV)
        m.bwd_fusion = fused
        m.set_dropout_masks(masks)
        m.train()
        loss = hb.Med_Sup_Loss(n_cls)(m(x.to(DEV)), y.to(DEV))
        loss.backward()
        grads[fused] = m.last_flat_grad.detach().float().cpu().clone()
        layout = [(nm, o, k) for (nm, _), (o, k, _) in zip(m.named_parameters(), m._layout)]
    worst = ("", 0.0)
    for nm, o, k in layout:
        a, b = grads[True][o:o + k], grads[False][o:o + k]
        if nm.endswith("conv_conv.0.bias") or nm.endswith("conv_conv.4.bias"):
            continue                                          # analytically zero (conv bias ahead of train-mode BatchNorm)
        worst = max(worst, (nm, rel_l2(a, b)), key=lambda t: t[1])
    assert worst[1] < 5e-2, worst
    assert rel_l2(grads[True], grads[False]) < 2e-2
