"""Stage-by-stage check of csrc/neck.cu through the C ABI against fp64 torch on the same GPU: pooled matrix, hidden layer,
outputs, and every intermediate of hpfg_neck_backward (d_dense rows, hidden gradient, pooled gradient in the scratch buffer).
Prints rel-L2 per stage and the worst rows; run on the GPU box: python tests/probes/neck_stages.py"""
import ctypes
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
import hpfg_b200 as hb          # noqa: E402
from hpfg_b200 import _lib as L  # noqa: E402

DEV = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def worst_rows(a, b, k=4):
    e = (a.double() - b.double()).flatten(1).norm(dim=1) / b.double().flatten(1).norm(dim=1).clamp_min(1e-300)
    v, i = e.topk(min(k, e.numel()))
    return ", ".join("row %d: %.1e" % (int(j), float(x)) for x, j in zip(v, i))


def run(in_dim, hid, out, s, shape, seed=11):
    torch.manual_seed(seed)
    m = hb.projection_conv(in_dim, hid_dim=hid, out_dim=out, s=s).to(DEV)
    ps = [p.detach().contiguous() for p in m._param_list()]
    n, c, h, w = shape
    S, rows = s * s, n * (1 + s * s)
    x = torch.randn(shape, device=DEV)
    pooled = torch.empty(rows, c, device=DEV)
    hidden = torch.empty(rows, hid, device=DEV)
    og, od = torch.empty(n, out, device=DEV), torch.empty(n, out, S, device=DEV)
    pa = (ctypes.c_void_p * 8)(*[p.data_ptr() for p in ps])
    L.check(L.lib().hpfg_neck_forward(L.ptr(x), n, c, h, w, s, hid, out, pa, L.ptr(pooled), L.ptr(hidden), L.ptr(og), L.ptr(od),
                                      L.stream_ptr(DEV)))
    xd = x.double()
    w1, b1, w2, b2, cw1, cb1, cw2, cb2 = [p.double().flatten(1) if p.dim() > 1 else p.double() for p in ps]
    pg = F.adaptive_avg_pool2d(xd, 1).flatten(1)
    pd = F.adaptive_avg_pool2d(xd, s).flatten(2).transpose(1, 2).reshape(n * S, c)
    p_ref = torch.cat([pg, pd])
    pre = torch.cat([pg @ w1.t() + b1, pd @ cw1.t() + cb1])
    h_ref = pre.clamp_min(0)
    g_ref = h_ref[:n] @ w2.t() + b2
    d_ref = (h_ref[n:] @ cw2.t() + cb2).reshape(n, S, out).transpose(1, 2)
    print("case", (in_dim, hid, out, s, shape))
    print("  pooled %.1e  hidden %.1e  out_global %.1e  out_dense %.1e" % (rel(pooled, p_ref), rel(hidden, h_ref), rel(og, g_ref), rel(od, d_ref)))
    dg, dd = torch.randn(n, out, device=DEV), torch.randn(n, out, S, device=DEV)
    grads = [torch.full_like(p, float("nan")) for p in ps]
    dx = torch.full(shape, float("nan"), device=DEV)
    scratch = torch.full((n * S * out + rows * (hid + c),), float("nan"), device=DEV)
    ga = (ctypes.c_void_p * 8)(*[g.data_ptr() for g in grads])
    L.check(L.lib().hpfg_neck_backward(L.ptr(dg), L.ptr(dd), n, c, h, w, s, hid, out, pa, L.ptr(pooled), L.ptr(hidden), ga, L.ptr(dx),
                                       L.ptr(scratch), L.stream_ptr(DEV)))
    torch.cuda.synchronize()
    drows = scratch[:n * S * out].view(n * S, out)
    dhid = scratch[n * S * out:n * S * out + rows * hid].view(rows, hid)
    dpool = scratch[n * S * out + rows * hid:].view(rows, c)
    dO = torch.cat([dg.double(), dd.double().transpose(1, 2).reshape(n * S, out)])
    mask = (hidden > 0).double()                       # the kernel's own mask: isolates the GEMMs from ReLU ties
    dh_ref = torch.cat([dO[:n] @ w2, dO[n:] @ cw2]) * mask
    dp_ref = torch.cat([dh_ref[:n] @ w1, dh_ref[n:] @ cw1])
    refs = [dh_ref[:n].t() @ p_ref[:n], dh_ref[:n].sum(0), dO[:n].t() @ h_ref[:n], dO[:n].sum(0),
            dh_ref[n:].t() @ p_ref[n:], dh_ref[n:].sum(0), dO[n:].t() @ h_ref[n:], dO[n:].sum(0)]
    xg = xd.clone().requires_grad_(True)
    pool_out = torch.cat([F.adaptive_avg_pool2d(xg, 1).flatten(1), F.adaptive_avg_pool2d(xg, s).flatten(2).transpose(1, 2).reshape(n * S, c)])
    (dx_ref,) = torch.autograd.grad((pool_out * dp_ref).sum(), xg, retain_graph=True)
    (dx_from_kernel_dp,) = torch.autograd.grad((pool_out * dpool.double()).sum(), xg)
    print("  d_dense rows %.1e  dhidden %.1e (%s)  dpooled %.1e (%s)" % (rel(drows, dO[n:]), rel(dhid, dh_ref), worst_rows(dhid, dh_ref),
                                                                     rel(dpool, dp_ref), worst_rows(dpool, dp_ref)))
    print("  dx %.1e   dx vs adjoint of the kernel's own dpooled %.1e   nan in dx: %d" % (rel(dx, dx_ref), rel(dx, dx_from_kernel_dp),
                                                                                      int(torch.isnan(dx).sum())))
    names = ["mlp.0.w", "mlp.0.b", "mlp.2.w", "mlp.2.b", "conv.0.w", "conv.0.b", "conv.2.w", "conv.2.b"]
    print("  " + "  ".join("%s %.1e" % (nm, rel(g, r)) for nm, g, r in zip(names, grads, refs)))
    e = (dx.double() - dx_ref).flatten(2).abs().amax(dim=2)
    v, i = e.flatten().topk(5)
    print("  worst dx planes (n,c): " + ", ".join("(%d,%d) %.1e" % (int(j) // c, int(j) % c, float(x)) for x, j in zip(v, i)),
          " typical |dx| %.1e" % dx_ref.abs().mean().item())


if __name__ == "__main__":
    for case in [(256, 2048, 128, 4, (32, 256, 14, 14)), (4, 1024, 128, 4, (32, 4, 224, 224)), (5, 70, 24, 3, (3, 5, 9, 11)),
                 (19, 33, 65, 1, (2, 19, 7, 5))]:
        run(*case)
