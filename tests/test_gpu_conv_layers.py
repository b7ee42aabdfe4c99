"""GPU, layer-isolated: each tensor-core convolution (tcgen05 implicit GEMM) against a plain PyTorch fp32
reference of the same op fed the SAME bf16-rounded inputs -- the north_star's "per-layer activations and
gradients agree within 1e-2 relative error in bf16" bar.  Covers every (Cin, Cout, kernel) shape of the UNet,
tile-multiple and ragged spatial sizes, the fused BN+LeakyReLU loader, bias, and the BN statistics epilogue."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from hpfg_b200 import _lib as L
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (cin, cout, ks) of every tensor-core layer in model/unet.py
SHAPES = [(16, 16, 3), (16, 32, 3), (32, 32, 3), (32, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3),
          (128, 256, 3), (256, 256, 3), (256, 128, 3), (128, 64, 3), (64, 32, 3), (32, 16, 3),
          (256, 128, 1), (128, 64, 1), (64, 32, 1), (32, 16, 1)]


def _nhwc_bf16(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _run(op, x_nchw, w, bias=None, scale=None, shift=None, want_stats=False):
    n, _, h, wd = x_nchw.shape
    cout, cin, ks, _ = w.shape
    c_out = cin if op == 1 else cout
    xin = _nhwc_bf16(x_nchw)
    out = torch.empty((n, h, wd, c_out), device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * c_out, device=DEV) if want_stats else None
    L.check(L.lib().hpfg_conv_tc_debug(op, n, h, wd, cin, cout, ks, L.ptr(xin), L.ptr(w.contiguous()), L.ptr(bias),
                                       L.ptr(scale), L.ptr(shift), L.ptr(out), L.ptr(stats),
                                       L.stream_ptr(torch.device(DEV))), "hpfg_conv_tc_debug")
    return out.float().permute(0, 3, 1, 2), stats


@pytest.mark.parametrize("cin,cout,ks", SHAPES)
@pytest.mark.parametrize("n,h,w", [(2, 32, 16), (3, 24, 20)])
def test_fprop_and_dgrad(cin, cout, ks, n, h, w):
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout + ks)
    x = torch.randn(n, cin, h, w, generator=g).to(DEV)
    wt = (torch.randn(cout, cin, ks, ks, generator=g) / (cin * ks * ks) ** 0.5).to(DEV)
    xb = x.to(torch.bfloat16).float()
    wb = wt.to(torch.bfloat16).float()
    # fprop with the BN statistics epilogue
    got, stats = _run(0, x, wt, want_stats=True)
    ref = F.conv2d(xb, wb, padding=ks // 2)
    assert rel_l2(got, ref) < 6e-3, "fprop"
    s1, s2 = ref.sum(dim=(0, 2, 3)), (ref * ref).sum(dim=(0, 2, 3))
    assert torch.allclose(stats[:cout], s1, rtol=1e-3, atol=1e-2 * s2.sqrt().max().item())
    assert torch.allclose(stats[cout:], s2, rtol=1e-3)
    # dgrad: din = conv_transpose(dout, w)
    dy = torch.randn(n, cout, h, w, generator=g).to(DEV)
    got, _ = _run(1, dy, wt)
    ref = F.conv_transpose2d(dy.to(torch.bfloat16).float(), wb, padding=ks // 2)
    assert rel_l2(got, ref) < 6e-3, "dgrad"


@pytest.mark.parametrize("cin,cout,ks", [(16, 16, 3), (64, 32, 3), (256, 128, 1), (128, 128, 3)])
def test_fprop_fused_loader_and_bias(cin, cout, ks):
    n, h, w = 2, 32, 24
    g = torch.Generator(device="cpu").manual_seed(7 + cin)
    x = torch.randn(n, cin, h, w, generator=g).to(DEV)
    wt = (torch.randn(cout, cin, ks, ks, generator=g) / (cin * ks * ks) ** 0.5).to(DEV)
    scale = (0.5 + torch.rand(cin, generator=g)).to(DEV)
    shift = (torch.rand(cin, generator=g) - 0.5).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    got, _ = _run(0, x, wt, bias=bias, scale=scale, shift=shift)
    xb = x.to(torch.bfloat16).float()
    act = F.leaky_relu(xb * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), 0.01).to(torch.bfloat16).float()
    ref = F.conv2d(act, wt.to(torch.bfloat16).float(), bias, padding=ks // 2)
    assert rel_l2(got, ref) < 6e-3


def test_fprop_full_resolution():
    """Benchmark-shape layer (16->16 @224x224, 4 images): catches tile-scheduler / pipeline wrap-around bugs."""
    g = torch.Generator(device="cpu").manual_seed(99)
    x = torch.randn(4, 16, 224, 224, generator=g).to(DEV)
    wt = (torch.randn(16, 16, 3, 3, generator=g) / 12.0).to(DEV)
    got, stats = _run(0, x, wt, want_stats=True)
    ref = F.conv2d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), padding=1)
    assert rel_l2(got, ref) < 6e-3
    assert torch.allclose(stats[16:], (ref * ref).sum(dim=(0, 2, 3)), rtol=1e-3)


@pytest.mark.parametrize("cin,cout,ks", SHAPES)
@pytest.mark.parametrize("n,h,w", [(2, 32, 16), (3, 24, 20)])
def test_wgrad(cin, cout, ks, n, h, w):
    g = torch.Generator(device="cpu").manual_seed(cin * 77 + cout + ks)
    x = torch.randn(n, cin, h, w, generator=g).to(DEV)
    dy = torch.randn(n, cout, h, w, generator=g).to(DEV)
    scale = (0.5 + torch.rand(cin, generator=g)).to(DEV)
    shift = (torch.rand(cin, generator=g) - 0.5).to(DEV)
    dw = torch.empty(cout, cin, ks, ks, device=DEV)
    db = torch.empty(cout, device=DEV)
    xn, dyn = _nhwc_bf16(x), _nhwc_bf16(dy)          # keep alive: ctypes pointers do not hold references
    L.check(L.lib().hpfg_wgrad_tc_debug(n, h, w, cin, cout, ks, L.ptr(xn), L.ptr(dyn), L.ptr(scale),
                                        L.ptr(shift), L.ptr(dw), L.ptr(db), L.stream_ptr(torch.device(DEV))), "wgrad")
    xb, dyb = x.to(torch.bfloat16).float(), dy.to(torch.bfloat16).float()
    act = F.leaky_relu(xb * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), 0.01).to(torch.bfloat16).float()
    act.requires_grad_(False)
    wref = torch.zeros(cout, cin, ks, ks, device=DEV, requires_grad=True)
    bref = torch.zeros(cout, device=DEV, requires_grad=True)
    (F.conv2d(act, wref, bref, padding=ks // 2) * dyb).sum().backward()
    assert rel_l2(dw, wref.grad) < 2e-3, "dW"
    assert rel_l2(db, bref.grad) < 2e-3, "dbias"


def test_wgrad_full_resolution():
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(4, 32, 224, 224, generator=g).to(DEV)
    dy = torch.randn(4, 16, 224, 224, generator=g).to(DEV)
    dw = torch.empty(16, 32, 3, 3, device=DEV)
    db = torch.empty(16, device=DEV)
    xn, dyn = _nhwc_bf16(x), _nhwc_bf16(dy)
    L.check(L.lib().hpfg_wgrad_tc_debug(4, 224, 224, 32, 16, 3, L.ptr(xn), L.ptr(dyn), None, None,
                                        L.ptr(dw), L.ptr(db), L.stream_ptr(torch.device(DEV))), "wgrad")
    wref = torch.zeros(16, 32, 3, 3, device=DEV, requires_grad=True)
    (F.conv2d(x.to(torch.bfloat16).float(), wref, padding=1) * dy.to(torch.bfloat16).float()).sum().backward()
    assert rel_l2(dw, wref.grad) < 2e-3
