"""Drop-in for the reference ``UNet`` (model/unet.py:155-175): an ``nn.Module`` with the same parameter and
buffer names, shapes, registration order and initialisation stream, whose forward/backward run on the
hand-written sm_100a kernels behind the C ABI (``include/hpfg_b200.h``).

Parameters live in ONE flat fp32 buffer (``flat_params``); every ``nn.Parameter`` is a view into it, so
``torch.optim.SGD(model.parameters())``, ``deepcopy(model)``, ``.to(device)``, ``state_dict()`` and the
per-parameter EMA loop of the reference keep working, while the fused SGD/EMA/all-reduce passes operate on
the flat buffer directly."""
import copy
import ctypes

import torch
import torch.nn as nn

from . import _lib as L

FT_CHNS = [16, 32, 64, 128, 256]            # model/unet.py:160
ENC_DROPOUT = [0.05, 0.1, 0.2, 0.3, 0.5]    # model/unet.py:161
ENC_PREFIXES = ["encoder.in_conv"] + ["encoder.down%d.maxpool_conv.1" % i for i in range(1, 5)]


def _conv_block(cin, cout, p):
    # holder modules only: their own forward() is never called, they give the reference's names / init
    blk = nn.Module()
    blk.conv_conv = nn.Sequential(
        nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU(), nn.Dropout(p),
        nn.Conv2d(cout, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU())
    return blk


class _Plan:
    """One C-side plan (workspace + saved activations) for a given batch shape."""

    def __init__(self, n, in_ch, n_cls, h, w, precision, bwd_fusion=None):
        self.handle = ctypes.c_void_p()
        L.check(L.lib().hpfg_unet_plan_create(n, in_ch, n_cls, h, w, precision, ctypes.byref(self.handle)),
                "hpfg_unet_plan_create")
        if bwd_fusion is not None:
            L.check(L.lib().hpfg_unet_plan_set_bwd_fusion(self.handle, int(bool(bwd_fusion))), "hpfg_unet_plan_set_bwd_fusion")
        self.busy = False           # holds activations of a forward whose backward has not run yet

    def __del__(self):
        try:
            if self.handle:
                L.lib().hpfg_unet_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _PlanLease:
    """Marks a plan busy (it holds the activations of a forward whose backward has not run) until backward runs OR the
    autograd node is dropped without one (outputs only inspected, exception, graph freed): no plan is leaked."""

    def __init__(self, plan):
        self.plan = plan
        plan.busy = True

    def release(self):
        if self.plan is not None:
            self.plan.busy = False
            self.plan = None

    __del__ = release


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, *params):
        plan = module._own_machine(module._acquire_plan(x, need_grad=True))
        logits = module._run_forward(plan, x, save=True)
        ctx.lease = _PlanLease(plan)
        ctx.module, ctx.plan = module, plan
        ctx.x = x        # the first layer's weight gradient re-reads the input: keep it alive until backward
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        m, plan = ctx.module, ctx.plan
        grads = torch.empty_like(m.flat_params)        # fresh buffer: param.grad views must not alias a reused one
        L.check(L.lib().hpfg_unet_backward(plan.handle, L.ptr(m.flat_params), L.ptr(dlogits.contiguous().float()),
                                           L.ptr(grads), 0, L.stream_ptr(grads.device)), "hpfg_unet_backward")
        ctx.lease.release()
        m.last_flat_grad = grads
        views = tuple(grads[o:o + n].view(s) for o, n, s in m._layout)
        return (None, None) + views


def _bottleneck(plan, x, num_features=FT_CHNS[-1]):
    feat = torch.empty((x.shape[0], num_features, x.shape[2] // 16, x.shape[3] // 16), device=x.device, dtype=torch.float32)
    L.check(L.lib().hpfg_unet_bottleneck(plan.handle, L.ptr(feat), L.stream_ptr(x.device)), "hpfg_unet_bottleneck")
    return feat


class _UNetPlusFunction(torch.autograd.Function):
    """UNet forward that also returns feature[-1] (model/unet.py:199-201); backward takes both gradients."""

    @staticmethod
    def forward(ctx, x, module, *params):
        plan = module._own_machine(module._acquire_plan(x, need_grad=True))
        logits = module._run_forward(plan, x, save=True)
        feat = _bottleneck(plan, x)
        ctx.lease = _PlanLease(plan)
        ctx.module, ctx.plan = module, plan
        ctx.x = x
        return logits, feat

    @staticmethod
    def backward(ctx, dlogits, dfeat):
        m, plan = ctx.module, ctx.plan
        grads = torch.empty_like(m.flat_params)
        if dlogits is None:
            dlogits = torch.zeros((ctx.x.shape[0], m.num_classes) + tuple(ctx.x.shape[2:]), device=grads.device)
        dfeat = dfeat.contiguous().float() if dfeat is not None else None
        L.check(L.lib().hpfg_unet_backward_ex(plan.handle, L.ptr(m.flat_params), L.ptr(dlogits.contiguous().float()),
                                              L.ptr(dfeat), L.ptr(grads), 0, L.stream_ptr(grads.device)),
                "hpfg_unet_backward_ex")
        ctx.lease.release()
        m.last_flat_grad = grads
        views = tuple(grads[o:o + n].view(s) for o, n, s in m._layout)
        return (None, None) + views


class UNet(nn.Module):
    """``UNet(in_channels=1, num_classes=4)`` -- same signature as model/unet.py:156."""

    def __init__(self, in_channels=1, num_classes=4, precision="bf16"):
        super().__init__()
        self.in_channels, self.num_classes = in_channels, num_classes
        self.precision = precision
        f = FT_CHNS
        enc = nn.Module()
        enc.in_conv = _conv_block(in_channels, f[0], ENC_DROPOUT[0])
        for i in range(1, 5):
            down = nn.Module()
            down.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _conv_block(f[i - 1], f[i], ENC_DROPOUT[i]))
            setattr(enc, "down%d" % i, down)
        dec = nn.Module()
        for i in range(1, 5):
            up = nn.Module()
            up.conv1x1 = nn.Conv2d(f[5 - i], f[4 - i], kernel_size=1)
            up.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            up.conv = _conv_block(2 * f[4 - i], f[4 - i], 0.0)
            setattr(dec, "up%d" % i, up)
        dec.out_conv = nn.Conv2d(f[0], num_classes, kernel_size=3, padding=1)
        self.encoder, self.decoder = enc, dec
        self._plans = {}
        self._dropout_masks = None
        self._no_dropout = False
        self._philox_offset = 0
        self.bwd_fusion = None          # None: library default; True / False: backward schedule of new plans (hpfg_unet_plan_set_bwd_fusion)
        self.last_flat_grad = None
        self._flatten()

    # ------------------------------------------------------------------ flat storage
    def _bn_modules(self):
        return [m for m in self.modules() if isinstance(m, nn.BatchNorm2d)]

    def _core_parameters(self):
        """The 82 encoder/decoder parameters that live in the flat buffer (== parameters() for the plain UNet; UNet_Plus
        adds projection necks, which stay ordinary torch parameters)."""
        return list(self.encoder.parameters()) + list(self.decoder.parameters())

    def _flatten(self):
        params = self._core_parameters()
        assert len(params) == L.NUM_PARAMS
        dev = params[0].device
        with torch.no_grad():
            flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
            layout, off = [], 0
            for p in params:
                n = p.numel()
                p.data = flat[off:off + n].view(p.shape)
                layout.append((off, n, tuple(p.shape)))
                off += n
            bns = self._bn_modules()
            assert len(bns) == L.NUM_BN
            run = torch.cat([torch.cat([b.running_mean.reshape(-1), b.running_var.reshape(-1)]) for b in bns]).float()
            run = run.contiguous().to(dev)
            cnt = torch.stack([b.num_batches_tracked.reshape(()) for b in bns]).to(torch.int64).contiguous().to(dev)
            o = 0
            for i, b in enumerate(bns):
                c = b.num_features
                b.running_mean = run[o:o + c]
                b.running_var = run[o + c:o + 2 * c]
                b.num_batches_tracked = cnt[i]
                o += 2 * c
        self.__dict__["flat_params"] = flat
        self.__dict__["bn_running"] = run
        self.__dict__["bn_counters"] = cnt
        self.__dict__["_layout"] = layout
        # cached object lists: the per-step flatness check must not walk the module tree (it sits in front of the first
        # kernel launch of every step)
        self.__dict__["_flat_params_list"] = params
        self.__dict__["_flat_bn_list"] = bns

    def _is_flat(self):
        flat = self.__dict__.get("flat_params")
        if flat is None:
            return False
        base = flat.data_ptr()
        plist = self.__dict__.get("_flat_params_list")
        if plist is None:
            return False
        for p, (o, n, _) in zip(plist, self._layout):
            if p.data_ptr() != base + 4 * o or p.dtype != torch.float32:
                return False
        bns = self.__dict__["_flat_bn_list"]
        o, rbase = 0, self.bn_running.data_ptr()
        for i, b in enumerate(bns):
            c = b.num_features
            if b.running_mean.data_ptr() != rbase + 4 * o or b.running_var.data_ptr() != rbase + 4 * (o + c):
                return False
            if b.num_batches_tracked.data_ptr() != self.bn_counters.data_ptr() + 8 * i:
                return False
            o += 2 * c
        return True

    def _is_flat_quick(self):
        """First / last parameter and first BatchNorm buffer still alias the flat buffers (3 pointer reads instead of 300)."""
        flat = self.__dict__.get("flat_params")
        plist = self.__dict__.get("_flat_params_list")
        if flat is None or plist is None:
            return False
        base = flat.data_ptr()
        o_last = self._layout[-1][0]
        return (plist[0].data_ptr() == base and plist[-1].data_ptr() == base + 4 * o_last and
                self._flat_bn_list[0].running_mean.data_ptr() == self.bn_running.data_ptr())

    def ensure_flat(self, quick=False):
        """Re-flatten if any parameter / buffer no longer aliases the flat buffers.  quick=True (the step drivers' per-iteration
        call: the full walk over 82 parameters and 18 BatchNorms costs 75 us of host time per model, which an end-to-end loop
        that reads the loss every step cannot hide) checks three sentinel pointers and does the full walk every 64th call."""
        n = self.__dict__["_flat_checks"] = self.__dict__.get("_flat_checks", 0) + 1
        ok = self._is_flat_quick() if (quick and n % 64) else self._is_flat()
        if not ok:
            self._flatten()
        return self.flat_params

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._plans = {}
        self._flatten()
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.ensure_flat()
        return out

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_flat_params_list", "_flat_bn_list"):
                continue                                   # rebuilt by _flatten() below
            new.__dict__[k] = {} if k == "_plans" else (None if k == "last_flat_grad" else copy.deepcopy(v, memo))
        new._flatten()
        return new

    # ------------------------------------------------------------------ test / parity hooks
    def set_dropout_masks(self, masks):
        """masks: None (library Philox stream) or {encoder block prefix: bool/uint8 keep-mask [N,C,H,W]} for all
        five encoder ConvBlocks -- the parity harness feeds the oracle's masks through this."""
        if masks is None:
            self._dropout_masks = None
            return
        dev = self.flat_params.device
        self._dropout_masks = [masks[p].to(device=dev, dtype=torch.uint8).contiguous() for p in ENC_PREFIXES]

    def set_dropout_enabled(self, enabled=True):
        self._no_dropout = not enabled

    def debug_tap(self, name, x_shape):
        plan = self._plans[(tuple(x_shape), self.precision)][-1]
        n, _, h, w = x_shape
        out = torch.empty(n * 256 * h * w, device=self.flat_params.device, dtype=torch.float32)
        L.check(L.lib().hpfg_unet_debug_tap(plan.handle, name.encode(), L.ptr(out), out.numel(),
                                            L.stream_ptr(out.device)), "hpfg_unet_debug_tap")
        return out

    # ------------------------------------------------------------------ forward
    def _acquire_plan(self, x, need_grad):
        key = (tuple(x.shape), self.precision)
        pool = self._plans.setdefault(key, [])
        for pl in pool:
            if not pl.busy:
                return pl
        n, c, h, w = x.shape
        prec = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16}[self.precision]
        pl = _Plan(n, c, self.num_classes, h, w, prec, getattr(self, "bwd_fusion", None))
        pool.append(pl)
        return pl

    @staticmethod
    def _own_machine(plan):
        """A plan last used by a step driver that shares the SMs between two concurrent forwards goes back to the full grid."""
        if getattr(plan, "fwd_ctas", 0):
            L.check(L.lib().hpfg_unet_plan_set_forward_ctas(plan.handle, 0), "hpfg_unet_plan_set_forward_ctas")
            plan.fwd_ctas = 0
        return plan

    def _run_forward(self, plan, x, save, out=None, offset_dev=None):
        dev = x.device
        shape = (x.shape[0], self.num_classes, x.shape[2], x.shape[3])
        logits = out if out is not None else torch.empty(shape, device=dev, dtype=torch.float32)
        assert tuple(logits.shape) == shape and logits.dtype == torch.float32 and logits.is_contiguous()
        masks = None
        if self._dropout_masks is not None:
            masks = (ctypes.c_void_p * L.NUM_DROPOUT)(*[m.data_ptr() for m in self._dropout_masks])
        seed = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()].initial_seed()
        offset = self._philox_offset
        if self.training and offset_dev is None:      # (graph replays advance the offset themselves, once per replay)
            self._philox_offset += 8
        if offset_dev is not None:       # CUDA-graph replays: the Philox offset is read from device memory at run time
            L.check(L.lib().hpfg_unet_forward_dv(plan.handle, L.ptr(self.flat_params), L.ptr(self.bn_running),
                                                 L.ptr(self.bn_counters), L.ptr(x), L.ptr(logits), int(self.training),
                                                 int(self._no_dropout), int(save), seed & (2 ** 64 - 1), L.ptr(offset_dev),
                                                 masks, L.stream_ptr(dev)), "hpfg_unet_forward_dv")
            return logits
        L.check(L.lib().hpfg_unet_forward(plan.handle, L.ptr(self.flat_params), L.ptr(self.bn_running),
                                          L.ptr(self.bn_counters), L.ptr(x), L.ptr(logits), int(self.training),
                                          int(self._no_dropout), int(save), seed & (2 ** 64 - 1), offset, masks,
                                          L.stream_ptr(dev)), "hpfg_unet_forward")
        return logits

    def forward(self, x):
        L.require_cuda(x, "UNet input")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError("expected input [N,%d,H,W], got %s" % (self.in_channels, tuple(x.shape)))
        self.ensure_flat()
        if self.flat_params.device != x.device:
            raise L.HpfgError("model is on %s but the input is on %s" % (self.flat_params.device, x.device))
        x = x.contiguous().float()
        need_grad = torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self._flat_params_list)
        if need_grad:
            return _UNetFunction.apply(x, self, *self._flat_params_list)
        return self._run_forward(self._own_machine(self._acquire_plan(x, False)), x, save=False)


class _NeckFunction(torch.autograd.Function):
    """projection_conv.forward / its adjoint on ``hpfg_neck_forward`` / ``hpfg_neck_backward`` (csrc/neck.cu)."""

    @staticmethod
    def forward(ctx, x, s, *params):
        n, c, h, w = x.shape
        hid, out = params[0].shape[0], params[2].shape[0]
        rows = n * (1 + s * s)
        dev = x.device
        pooled = torch.empty((rows, c), device=dev, dtype=torch.float32)
        hidden = torch.empty((rows, hid), device=dev, dtype=torch.float32)
        out_global = torch.empty((n, out), device=dev, dtype=torch.float32)
        out_dense = torch.empty((n, out, s * s), device=dev, dtype=torch.float32)
        ps = [p.detach().contiguous() for p in params]
        L.check(L.lib().hpfg_neck_forward(L.ptr(x), n, c, h, w, s, hid, out, (ctypes.c_void_p * 8)(*[p.data_ptr() for p in ps]),
                                          L.ptr(pooled), L.ptr(hidden), L.ptr(out_global), L.ptr(out_dense), L.stream_ptr(dev)),
                "hpfg_neck_forward")
        ctx.save_for_backward(pooled, hidden, *ps)
        ctx.geom = (n, c, h, w, s, hid, out)
        return out_global, out_dense

    @staticmethod
    def backward(ctx, d_global, d_dense):
        pooled, hidden, *ps = ctx.saved_tensors
        n, c, h, w, s, hid, out = ctx.geom
        dev = pooled.device
        rows = n * (1 + s * s)
        grads = [torch.empty_like(p) for p in ps]
        dx = torch.empty((n, c, h, w), device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        scratch = torch.empty(n * s * s * out + rows * (hid + c), device=dev, dtype=torch.float32)
        d_global, d_dense = d_global.contiguous().float(), d_dense.contiguous().float()
        L.check(L.lib().hpfg_neck_backward(L.ptr(d_global), L.ptr(d_dense), n, c, h, w, s,
                                           hid, out, (ctypes.c_void_p * 8)(*[p.data_ptr() for p in ps]), L.ptr(pooled), L.ptr(hidden),
                                           (ctypes.c_void_p * 8)(*[g.data_ptr() for g in grads]), L.ptr(dx), L.ptr(scratch),
                                           L.stream_ptr(dev)), "hpfg_neck_backward")
        return (dx, None) + tuple(grads)


class projection_conv(nn.Module):
    """The DenseCL-style neck of model/unet.py:120-152 (same constructor, parameter names and shapes): a global-average-pool +
    fc-relu-fc branch and an adaptive-pool(s x s) + 1x1conv-relu-1x1conv branch, returned as ([N,out_dim], [N,out_dim,s*s]).
    The Linear / Conv2d submodules only hold the parameters; ``forward`` and its gradients run on the library's neck kernels
    (pooling of both branches in one pass, fp32 GEMMs with bias / ReLU epilogues), so there is no CPU path."""

    def __init__(self, in_dim, hid_dim=2048, out_dim=128, s=4):
        super().__init__()
        self.is_s = s
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.mlp = nn.Sequential(nn.Linear(in_dim, hid_dim), nn.ReLU(inplace=True), nn.Linear(hid_dim, out_dim))
        self.mlp_conv = nn.Sequential(nn.Conv2d(in_dim, hid_dim, 1), nn.ReLU(inplace=True), nn.Conv2d(hid_dim, out_dim, 1))
        self.pool = nn.AdaptiveAvgPool2d((s, s)) if self.is_s else None

    def _param_list(self):
        return [self.mlp[0].weight, self.mlp[0].bias, self.mlp[2].weight, self.mlp[2].bias,
                self.mlp_conv[0].weight, self.mlp_conv[0].bias, self.mlp_conv[2].weight, self.mlp_conv[2].bias]

    def forward(self, x):
        L.require_cuda(x, "projection_conv input")
        if x.dim() != 4 or x.shape[1] != self.mlp[0].in_features:
            raise ValueError("expected input [N,%d,H,W], got %s" % (self.mlp[0].in_features, tuple(x.shape)))
        if not self.is_s:
            raise L.HpfgError("projection_conv(s=0) (dense branch without pooling) is not built: UNet_Plus always uses s=4 "
                              "(model/unet.py:192-193)")
        params = self._param_list()
        if any(p.device != x.device or p.dtype != torch.float32 for p in params):
            raise L.HpfgError("projection_conv: parameters must be fp32 on %s" % (x.device,))
        return _NeckFunction.apply(x.contiguous().float(), int(self.is_s), *params)


class UNet_Plus(UNet):
    """``UNet_Plus(in_channels=1, num_classes=4)`` (model/unet.py:178-206): the U-Net plus two projection necks;
    ``forward`` returns ``(logits, high_feature, head_feature)``, ``val`` the logits only.  Parameter names, shapes and
    order match the reference (encoder.*, decoder.*, dense_projection_high.*, dense_projection_head.*)."""

    def __init__(self, in_channels=1, num_classes=4, precision="bf16"):
        super().__init__(in_channels, num_classes, precision)
        self.dense_projection_high = projection_conv(FT_CHNS[-1])
        self.dense_projection_head = projection_conv(num_classes, hid_dim=1024)

    def val(self, x):
        return UNet.forward(self, x)

    def forward(self, x):
        L.require_cuda(x, "UNet_Plus input")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError("expected input [N,%d,H,W], got %s" % (self.in_channels, tuple(x.shape)))
        self.ensure_flat()
        if self.flat_params.device != x.device:
            raise L.HpfgError("model is on %s but the input is on %s" % (self.flat_params.device, x.device))
        x = x.contiguous().float()
        need_grad = torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self._flat_params_list)
        if need_grad:
            output, feature = _UNetPlusFunction.apply(x, self, *self._flat_params_list)
        else:
            plan = self._own_machine(self._acquire_plan(x, False))
            output = self._run_forward(plan, x, save=False)
            feature = _bottleneck(plan, x)
        high_feature = self.dense_projection_high(feature)
        head_feature = self.dense_projection_head(output)
        return output, high_feature, head_feature


def predict_volume(model, slices, max_batch=64, dtype=torch.int64):
    """The network part of val.py:268-281 ``test_single_volume`` for a whole volume at once.

    slices: [S,H,W] (single-channel volumes, as ACDC) or [S,Cin,H,W], already resampled to the test crop size; returns
    ``argmax(softmax(net(slice)), dim=1)`` for every slice as [S,H,W] labels.  The reference loops slice by slice at
    batch 1; in eval mode BatchNorm uses its running statistics and dropout is off, so every slice is independent and
    batching them gives the same labels.  Like the reference (``net.eval()`` at :276) this leaves the model in eval mode.
    Forward = ``hpfg_unet_forward(training=0)``, labels = ``hpfg_argmax_labels`` (one launch per chunk)."""
    from .losses import argmax_labels
    x = slices if slices.dim() == 4 else slices.unsqueeze(1)
    L.require_cuda(x, "slices")
    model.eval()
    out = torch.empty((x.shape[0], x.shape[2], x.shape[3]), device=x.device, dtype=dtype)
    with torch.no_grad():
        for s in range(0, x.shape[0], max_batch):
            out[s:s + max_batch] = argmax_labels(model(x[s:s + max_batch]), dtype=dtype)
    return out
