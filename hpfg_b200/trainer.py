"""Fused training steps for the three north-star trainers, built from the C-ABI calls only:

* ``MeanTeacherStep``  -- 2017_03_NIPS_Mean-Teacher_ACDC.py:89-113
* ``CPSStep``          -- 2021_06_CVPR_CPS_ACDC.py:90-120
* ``UAMTStep``         -- 2019_07_MICCAI_Uncertainty_Aware_ACDC.py:120-170
* ``ICTStep``          -- 2022_02_ISBI_ICT-MedSeg_ACDC.py:96-140

Each ``step()`` enqueues: student forward (activations kept in the plan), teacher/peer forward(s), ONE fused
loss launch pair (value + dlogits), backward into a persistent flat gradient buffer, (data parallel: NCCL
all-reduce of the gradient buckets on a side stream, overlapped with the rest of backward), and ONE fused
SGD-momentum(+EMA) pass over the flat parameter buffers.  Nothing synchronises the host; the returned loss is
a device scalar.  Host-side schedules (Medical_LR, consistency ramp-up, EMA alpha) are python floats exactly as
in the reference (utils/scheduler/medical_lr.py:13-17, utils/utils.py:67-86)."""
import math

import torch

from . import _lib as L
from .losses import ssl_loss_raw, ict_loss_raw, ict_mix_inputs
from .utils import sigmoid_rampup


def medical_lr(cur_itrs, base_lr, max_iterations):
    """Learning rate used at iteration ``cur_itrs`` (1-based) by Medical_LR constructed with last_epoch=-1."""
    return base_lr * (1.0 - (cur_itrs - 2) / max_iterations) ** 0.9


def gradient_buckets(in_channels, num_classes):
    """[(offset, count)] of the flat-gradient buckets in the order backward completes them (tail of the
    parameter list first: out_conv/up4/up3 | up2/up1 | down4 | in_conv..down3)."""
    import ctypes
    offs, cnts = (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)()
    L.check(L.lib().hpfg_unet_bucket_layout(in_channels, num_classes, offs, cnts), "hpfg_unet_bucket_layout")
    return [(int(o), int(c)) for o, c in zip(offs, cnts)]


def allreduce_flat_buckets(flat_grad, buckets, group=None, before_bucket=None, stream=None):
    """Sum-all-reduce ``flat_grad`` bucket by bucket (async), returning the work handles.  ``before_bucket(i)`` is
    called before bucket i is enqueued (the CUDA path makes the comm stream wait for that bucket's event there).
    Works for CPU tensors + gloo (host-logic tests) and CUDA tensors + NCCL alike."""
    import contextlib
    import torch.distributed as dist
    works = []
    for i, (off, cnt) in enumerate(buckets):
        if before_bucket is not None:
            before_bucket(i)
        ctx = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
        with ctx:
            works.append(dist.all_reduce(flat_grad[off:off + cnt], group=group, async_op=True))
    return works


def shard_batch(x_l, x_u, y, rank, world):
    """Data-parallel split that keeps the labeled:unlabeled ratio per rank (both sub-batches are split, because the
    supervised and consistency terms normalise over their own pixels)."""
    n_l, n_u = x_l.shape[0], x_u.shape[0]
    assert n_l % world == 0 and n_u % world == 0, "labeled and unlabeled batch sizes must divide by the world size"
    a, b = n_l // world, n_u // world
    return x_l[rank * a:(rank + 1) * a], x_u[rank * b:(rank + 1) * b], y[rank * a:(rank + 1) * a]


class _StepBase:
    def __init__(self, *, lr=0.01, momentum=0.9, weight_decay=1e-4, total_itrs=30000, consistency=0.1,
                 consistency_rampup=200.0, process_group=None):
        self.base_lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.total_itrs, self.consistency, self.consistency_rampup = total_itrs, consistency, consistency_rampup
        self.cur_itrs = 0
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self._comm = None
        self.last = {}

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    _graph_enabled = False
    _graph_dp = False
    kernels_per_replay = 0
    replayed_kernels = 0          # kernels executed through graph replays (the library's own counter sees host launches only)

    def enable_graph(self, enabled=True, data_parallel=False):
        """From the second iteration on, replay the step as ONE captured CUDA graph (all forwards on their streams, the
        fused loss, backward with its side-stream weight gradients, fused SGD(+EMA)).  The scalars that change per
        iteration -- learning rate, EMA alpha, consistency weight, UAMT threshold, dropout Philox offsets -- live in a
        small device block that is refreshed before every replay (the `_dv` entry points read them at run time).
        data_parallel=True also captures the bucketed NCCL all-reduces of a multi-process run (Mean-Teacher driver only,
        measured on 2/4/8 B200; the other drivers stay eager when data parallel); by default a data-parallel step stays eager."""
        self._graph_enabled = bool(enabled)
        self._graph_dp = bool(data_parallel)      # also capture the NCCL bucket all-reduces (torch >= 2.x captures NCCL work)
        if not enabled:
            self._graph = None
            self._ggraph = None

    _graph_dp_capable = False     # only the Mean-Teacher driver's capture with the NCCL all-reduces inside has been measured

    def _use_graph(self):
        dp_ok = self.world == 1 or (self._graph_dp and self._graph_dp_capable)
        return self._graph_enabled and self.cur_itrs >= 2 and dp_ok

    def _graph_replay(self, inputs, dyn_f, fwd_models, body):
        """Generic capture-once / replay driver (CPS, UAMT, ICT; Mean-Teacher keeps its own copy below).
        inputs: tensors copied into static buffers before each replay; dyn_f: python floats for the device block
        (fp32); fwd_models: one entry per network forward in the eager call order -> one Philox-offset slot each (the
        offsets advance exactly as the eager path advances them); body(static_inputs, dyn_f_dev, dyn_o_dev) enqueues the
        step with the `_dv` entry points and returns its output dict (static tensors)."""
        dev = inputs[0].device
        sig = tuple((tuple(t.shape), t.dtype) for t in inputs) + (len(dyn_f), len(fwd_models))
        if getattr(self, "_gsig", None) != sig:
            self._gin = [torch.empty_like(t) for t in inputs]
            self._gf = torch.zeros(len(dyn_f), device=dev, dtype=torch.float32)
            self._go = torch.zeros(len(fwd_models), device=dev, dtype=torch.int64)
            self._gf_host = torch.zeros(len(dyn_f), dtype=torch.float32).pin_memory()
            self._go_host = torch.zeros(len(fwd_models), dtype=torch.int64).pin_memory()
            self._ggraph, self._gsig = None, sig
        for i, v in enumerate(dyn_f):
            self._gf_host[i] = v
        for i, m in enumerate(fwd_models):
            self._go_host[i] = m._philox_offset
            if m.training:
                m._philox_offset += 8
        self._gf.copy_(self._gf_host, non_blocking=True)
        self._go.copy_(self._go_host, non_blocking=True)
        for s, t in zip(self._gin, inputs):
            s.copy_(t, non_blocking=True)
        if self._ggraph is None:
            for m in fwd_models:
                m.ensure_flat()
            g = torch.cuda.CUDAGraph()
            n0 = L.lib().hpfg_launch_count()
            try:
                with torch.cuda.graph(g):
                    self._gout = body(self._gin, self._gf, self._go)
            except Exception as exc:        # capture not possible here: stay eager on the same device-value path
                import warnings
                warnings.warn("hpfg_b200: CUDA-graph capture of the step failed (%s); falling back to eager launches" % exc)
                self._graph_enabled, self._ggraph = False, None
                torch.cuda.synchronize()
                return body(self._gin, self._gf, self._go)
            self._ggraph = g
            self.kernels_per_replay = int(L.lib().hpfg_launch_count() - n0)   # kernel nodes of the captured step
        self._ggraph.replay()
        self.replayed_kernels += self.kernels_per_replay
        return self._gout

    def _forward_dv(self, model, x, save, out, offset_dev):
        plan = model._acquire_plan(x, need_grad=save)
        return plan, model._run_forward(plan, x, save=save, out=out, offset_dev=offset_dev)

    def _sgd_dv(self, model, grads, buf, dyn_f, ema_model=None):
        """SGD(+EMA) with {lr, ema_alpha, 1-ema_alpha} read from the device block (first_step = 0: replays start at
        iteration 2)."""
        n = model.flat_params.numel()
        st = L.stream_ptr(grads.device)
        if ema_model is None:
            L.check(L.lib().hpfg_sgd_momentum_dv(L.ptr(model.flat_params), L.ptr(grads), L.ptr(buf), n, self.momentum,
                                                 self.weight_decay, 1.0 / self.world, 0, L.ptr(dyn_f), st),
                    "hpfg_sgd_momentum_dv")
        else:
            L.check(L.lib().hpfg_sgd_momentum_ema_dv(L.ptr(model.flat_params), L.ptr(grads), L.ptr(buf),
                                                     L.ptr(ema_model.flat_params), n, self.momentum, self.weight_decay,
                                                     1.0 / self.world, 0, L.ptr(dyn_f), st), "hpfg_sgd_momentum_ema_dv")

    def _dyn_scalars(self, ema_decay=None):
        """[lr, ema_alpha, 1 - ema_alpha (fp32 subtraction, as the by-value entry point), consistency weight]."""
        lr = medical_lr(self.cur_itrs, self.base_lr, self.total_itrs)
        alpha = min(1 - 1 / (self.cur_itrs + 1), ema_decay) if ema_decay is not None else 0.0
        w = self._consistency_weight()
        return [lr, alpha, float(1.0 - torch.tensor(alpha, dtype=torch.float32)), w]

    def _side_stream(self, device):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=device)
        return self._side

    def _consistency_weight(self):
        return self.consistency * sigmoid_rampup(self.cur_itrs // 150, self.consistency_rampup)

    def _forward(self, model, x, save, out=None):
        model.ensure_flat()
        plan = model._acquire_plan(x, need_grad=save)
        logits = model._run_forward(plan, x, save=save, out=out)
        return plan, logits

    def _persistent(self, name, shape, device):
        """Step-persistent fp32 buffer (no per-step allocation: tensors that cross streams would otherwise need
        record_stream, which defeats the caching allocator when the host runs ahead of the device)."""
        buf = self.__dict__.setdefault("_bufs", {}).get(name)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.device != device:
            buf = torch.empty(shape, device=device, dtype=torch.float32)
            self._bufs[name] = buf
        return buf

    def _backward(self, model, plan, dlogits, grads):
        L.check(L.lib().hpfg_unet_backward(plan.handle, L.ptr(model.flat_params), L.ptr(dlogits), L.ptr(grads), 0,
                                           L.stream_ptr(grads.device)), "hpfg_unet_backward")
        if self.world > 1:
            self._allreduce_buckets(plan, grads)

    def _allreduce_buckets(self, plan, grads):
        """Bucketed gradient all-reduce overlapped with the tail of backward: the comm stream waits on the
        per-bucket events the plan recorded, NCCL runs there, the compute stream joins before the SGD pass."""
        import ctypes
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=grads.device)
            self._buckets = gradient_buckets(self.in_channels, self.num_classes)
        lib = L.lib()
        comm = ctypes.c_void_p(self._comm.cuda_stream)
        works = allreduce_flat_buckets(
            grads, self._buckets, self.pg, stream=self._comm,
            before_bucket=lambda b: L.check(lib.hpfg_unet_bucket_wait(plan.handle, b, comm), "hpfg_unet_bucket_wait"))
        for w in works:
            w.wait()            # stream-level wait on the compute stream, not a host sync

    def _sgd(self, model, grads, buf, ema_model=None, ema_alpha=0.0):
        lr = medical_lr(self.cur_itrs, self.base_lr, self.total_itrs)
        n = model.flat_params.numel()
        st = L.stream_ptr(grads.device)
        first = int(self.cur_itrs == 1)
        if ema_model is None:
            L.check(L.lib().hpfg_sgd_momentum(L.ptr(model.flat_params), L.ptr(grads), L.ptr(buf), n, lr, self.momentum,
                                              self.weight_decay, 1.0 / self.world, first, st), "hpfg_sgd_momentum")
        else:
            L.check(L.lib().hpfg_sgd_momentum_ema(L.ptr(model.flat_params), L.ptr(grads), L.ptr(buf),
                                                  L.ptr(ema_model.flat_params), n, lr, self.momentum, self.weight_decay,
                                                  1.0 / self.world, first, ema_alpha, st), "hpfg_sgd_momentum_ema")
        return lr


class MeanTeacherStep(_StepBase):
    _graph_dp_capable = True

    def __init__(self, model, ema_model, *, ema_decay=0.99, **kw):
        super().__init__(**kw)
        self.model, self.ema_model, self.ema_decay = model, ema_model, ema_decay
        self.in_channels, self.num_classes = model.in_channels, model.num_classes
        model.train()
        ema_model.train()                               # the teacher runs in train() mode (2017_03...:70)
        model.ensure_flat()
        ema_model.ensure_flat()
        self.grads = torch.zeros_like(model.flat_params)
        self.mom = torch.zeros_like(model.flat_params)

    def step(self, x, labels):
        """x: [n_l+n_u, C, H, W] fp32 CUDA (labeled slices first); labels: [n_l, H, W] int64 CUDA."""
        self.cur_itrs += 1
        if self._use_graph():
            return self._step_graph(x, labels)
        n_l = labels.shape[0]
        # the teacher forward is independent of the student forward: it runs on a side stream so that each network's
        # small / dependent kernels fill the other's bubbles (the teacher sees the whole batch, 2017_03...:100)
        main = torch.cuda.current_stream(x.device)
        side = main if getattr(self, "serialize", False) else self._side_stream(x.device)   # serialize: profiling only
        side.wait_stream(main)
        shape = (x.shape[0], self.num_classes, x.shape[2], x.shape[3])
        with torch.cuda.stream(side):
            _, t_out = self._forward(self.ema_model, x, False, out=self._persistent("t_out", shape, x.device))
        plan, out = self._forward(self.model, x, True, out=self._persistent("s_out", shape, x.device))
        main.wait_stream(side)
        w = self._consistency_weight()
        r = ssl_loss_raw(L.LOSS_MT, out, t_out[n_l:], labels, n_l, cons_weight=w)
        self._backward(self.model, plan, r["dstudent"], self.grads)
        alpha = min(1 - 1 / (self.cur_itrs + 1), self.ema_decay)      # utils/utils.py:84
        lr = self._sgd(self.model, self.grads, self.mom, self.ema_model, alpha)
        self.last = dict(scalars=r["scalars"], lr=lr, w=w, logits=out, teacher_logits=t_out)
        return r["scalars"][0]

    def _step_graph(self, x, labels):
        dev = x.device
        if getattr(self, "_graph", None) is None or tuple(self._gx.shape) != tuple(x.shape) or tuple(self._gy.shape) != tuple(labels.shape):
            self._gx, self._gy = torch.empty_like(x), torch.empty_like(labels)
            self._dyn_f = torch.zeros(4, device=dev, dtype=torch.float32)        # lr, alpha, 1 - alpha, consistency weight
            self._dyn_o = torch.zeros(2, device=dev, dtype=torch.int64)          # Philox offsets: student, teacher
            self._dyn_f_host = torch.zeros(4, dtype=torch.float32).pin_memory()
            self._dyn_o_host = torch.zeros(2, dtype=torch.int64).pin_memory()
            self._graph = None
        lr = medical_lr(self.cur_itrs, self.base_lr, self.total_itrs)
        alpha = min(1 - 1 / (self.cur_itrs + 1), self.ema_decay)
        w = self._consistency_weight()
        a32 = torch.tensor(alpha, dtype=torch.float32)
        self._dyn_f_host[0], self._dyn_f_host[1], self._dyn_f_host[3] = lr, alpha, w
        self._dyn_f_host[2] = 1.0 - a32                                          # fp32 subtraction, as the by-value entry point
        self._dyn_o_host[0], self._dyn_o_host[1] = self.model._philox_offset, self.ema_model._philox_offset
        self.model._philox_offset += 8
        self.ema_model._philox_offset += 8
        self._dyn_f.copy_(self._dyn_f_host, non_blocking=True)
        self._dyn_o.copy_(self._dyn_o_host, non_blocking=True)
        self._gx.copy_(x, non_blocking=True)
        self._gy.copy_(labels, non_blocking=True)
        if self._graph is None:
            self.model.ensure_flat()
            self.ema_model.ensure_flat()
            g = torch.cuda.CUDAGraph()
            n0 = L.lib().hpfg_launch_count()
            try:
                with torch.cuda.graph(g):
                    self._graph_out = self._step_body_dv(self._gx, self._gy)
            except Exception as exc:        # capture not possible here (e.g. a collective that refuses capture): stay eager
                import warnings
                warnings.warn("hpfg_b200: CUDA-graph capture of the step failed (%s); falling back to eager launches" % exc)
                self._graph_enabled, self._graph = False, None
                torch.cuda.synchronize()
                out = self._step_body_dv(self._gx, self._gy)       # same device-value path, launched eagerly
                self.last = dict(scalars=out["scalars"], lr=lr, w=w, logits=out["logits"], teacher_logits=out["teacher_logits"])
                return out["scalars"][0]
            self._graph = g
            self.kernels_per_replay = int(L.lib().hpfg_launch_count() - n0)   # kernel nodes of the captured step
        self._graph.replay()
        self.replayed_kernels += self.kernels_per_replay
        self.last = dict(scalars=self._graph_out["scalars"], lr=lr, w=w, logits=self._graph_out["logits"],
                         teacher_logits=self._graph_out["teacher_logits"])
        return self._graph_out["scalars"][0]

    def _step_body_dv(self, x, labels):
        """The step with every per-iteration scalar read from the device block (captured once, replayed)."""
        n_l = labels.shape[0]
        dev = x.device
        main, side = torch.cuda.current_stream(dev), self._side_stream(dev)
        shape = (x.shape[0], self.num_classes, x.shape[2], x.shape[3])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            tplan = self.ema_model._acquire_plan(x, need_grad=False)
            t_out = self.ema_model._run_forward(tplan, x, save=False, out=self._persistent("t_out", shape, dev),
                                                offset_dev=self._dyn_o[1:2])
        plan = self.model._acquire_plan(x, need_grad=True)
        out = self.model._run_forward(plan, x, save=True, out=self._persistent("s_out", shape, dev), offset_dev=self._dyn_o[0:1])
        main.wait_stream(side)
        r = ssl_loss_raw(L.LOSS_MT, out, t_out[n_l:], labels, n_l, cons_weight_dev=self._dyn_f[3:4])
        self._backward(self.model, plan, r["dstudent"], self.grads)
        n = self.model.flat_params.numel()
        L.check(L.lib().hpfg_sgd_momentum_ema_dv(L.ptr(self.model.flat_params), L.ptr(self.grads), L.ptr(self.mom),
                                                 L.ptr(self.ema_model.flat_params), n, self.momentum, self.weight_decay,
                                                 1.0 / self.world, 0, L.ptr(self._dyn_f), L.stream_ptr(dev)),
                "hpfg_sgd_momentum_ema_dv")
        return dict(scalars=r["scalars"], logits=out, teacher_logits=t_out)


class CPSStep(_StepBase):
    def __init__(self, model1, model2, **kw):
        super().__init__(**kw)
        self.m1, self.m2 = model1, model2
        self.in_channels, self.num_classes = model1.in_channels, model1.num_classes
        model1.train()
        model2.train()
        model1.ensure_flat()
        model2.ensure_flat()
        self.g1, self.g2 = torch.zeros_like(model1.flat_params), torch.zeros_like(model2.flat_params)
        self.b1, self.b2 = torch.zeros_like(model1.flat_params), torch.zeros_like(model2.flat_params)

    def step(self, x, labels):
        self.cur_itrs += 1
        if self._use_graph():
            dyn = self._dyn_scalars()
            out = self._graph_replay([x, labels], dyn, [self.m1, self.m2], self._body_dv)
            self.last = dict(scalars=out["scalars"], lr=dyn[0], w=dyn[3], logits1=out["logits1"], logits2=out["logits2"])
            return out["scalars"][0]
        n_l = labels.shape[0]
        # the two networks are independent except for the loss: network 2 runs on a side stream (forward, then backward + SGD)
        main = torch.cuda.current_stream(x.device)
        side = main if getattr(self, "serialize", False) else self._side_stream(x.device)
        shape = (x.shape[0], self.num_classes, x.shape[2], x.shape[3])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            p2, o2 = self._forward(self.m2, x, True, out=self._persistent("o2", shape, x.device))
        p1, o1 = self._forward(self.m1, x, True, out=self._persistent("o1", shape, x.device))
        main.wait_stream(side)
        w = self._consistency_weight()
        r = ssl_loss_raw(L.LOSS_CPS, o1, o2, labels, n_l, cons_weight=w, want_pseudo=False)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._backward(self.m2, p2, r["dother"], self.g2)
            self._sgd(self.m2, self.g2, self.b2)
        self._backward(self.m1, p1, r["dstudent"], self.g1)
        lr = self._sgd(self.m1, self.g1, self.b1)
        main.wait_stream(side)
        self.last = dict(scalars=r["scalars"], lr=lr, w=w, logits1=o1, logits2=o2)
        return r["scalars"][0]


    def _body_dv(self, ins, f, o):
        x, labels = ins
        n_l, dev = labels.shape[0], x.device
        main, side = torch.cuda.current_stream(dev), self._side_stream(dev)
        shape = (x.shape[0], self.num_classes, x.shape[2], x.shape[3])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            p2, o2 = self._forward_dv(self.m2, x, True, self._persistent("o2", shape, dev), o[1:2])
        p1, o1 = self._forward_dv(self.m1, x, True, self._persistent("o1", shape, dev), o[0:1])
        main.wait_stream(side)
        r = ssl_loss_raw(L.LOSS_CPS, o1, o2, labels, n_l, cons_weight_dev=f[3:4], want_pseudo=False)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._backward(self.m2, p2, r["dother"], self.g2)
            self._sgd_dv(self.m2, self.g2, self.b2, f)
        self._backward(self.m1, p1, r["dstudent"], self.g1)
        self._sgd_dv(self.m1, self.g1, self.b1, f)
        main.wait_stream(side)
        return dict(scalars=r["scalars"], logits1=o1, logits2=o2)


class UAMTStep(_StepBase):
    def __init__(self, model, ema_model, *, ema_decay=0.99, T=8, **kw):
        super().__init__(**kw)
        self.model, self.ema_model, self.ema_decay, self.T = model, ema_model, ema_decay, T
        self.in_channels, self.num_classes = model.in_channels, model.num_classes
        model.train()
        ema_model.train()
        model.ensure_flat()
        ema_model.ensure_flat()
        self.grads = torch.zeros_like(model.flat_params)
        self.mom = torch.zeros_like(model.flat_params)

    @staticmethod
    def make_noise(like):
        return torch.clamp(torch.randn_like(like) * 0.1, -0.2, 0.2)      # 2019_07...:130,141

    def step(self, x, labels, noise=None, mc_noise=None):
        """noise [n_u,...] / mc_noise [T//2, 2*n_u, ...]: the clamped perturbations; drawn here if None."""
        self.cur_itrs += 1
        n_l = labels.shape[0]
        x_u = x[n_l:]
        n_u = x_u.shape[0]
        if noise is None:
            noise = self.make_noise(x_u)
        if self._use_graph():
            if mc_noise is None:
                mc_noise = torch.clamp(torch.randn((self.T // 2, 2 * n_u) + tuple(x_u.shape[1:]), device=x.device,
                                                   dtype=x.dtype) * 0.1, -0.2, 0.2)
            dyn = self._dyn_scalars(self.ema_decay)
            thr = (0.75 + 0.25 * sigmoid_rampup(self.cur_itrs, self.total_itrs)) * math.log(2)
            out = self._graph_replay([x, labels, noise, mc_noise.contiguous()], dyn + [thr],
                                     [self.model] + [self.ema_model] * (1 + self.T // 2), self._body_dv)
            self.last = dict(scalars=out["scalars"], lr=dyn[0], w=dyn[3], threshold=thr, logits=out["logits"],
                             teacher_logits=out["teacher_logits"], mc_logits=out["mc_logits"])
            return out["scalars"][0]
        xr = x_u.repeat(2, 1, 1, 1)
        noises = [mc_noise[i] if mc_noise is not None else self.make_noise(xr) for i in range(self.T // 2)]
        x_t = (x_u + noise).contiguous()
        x_mc = [(xr + nz).contiguous() for nz in noises]
        ncls, hh, ww = self.num_classes, x.shape[2], x.shape[3]
        mc = self._persistent("mc", (self.T * n_u, ncls, hh, ww), x.device)
        # the five teacher forwards are independent of the student forward: side stream
        main = torch.cuda.current_stream(x.device)
        side = main if getattr(self, "serialize", False) else self._side_stream(x.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _, t_out = self._forward(self.ema_model, x_t, False, out=self._persistent("t_out", (n_u, ncls, hh, ww), x.device))
            for i in range(self.T // 2):
                self._forward(self.ema_model, x_mc[i], False, out=mc[2 * n_u * i:2 * n_u * (i + 1)])
        plan, out = self._forward(self.model, x, True, out=self._persistent("s_out", (x.shape[0], ncls, hh, ww), x.device))
        main.wait_stream(side)
        w = self._consistency_weight()
        thr = (0.75 + 0.25 * sigmoid_rampup(self.cur_itrs, self.total_itrs)) * math.log(2)
        r = ssl_loss_raw(L.LOSS_UAMT, out, t_out, labels, n_l, cons_weight=w, mc_logits=mc, mc_passes=self.T,
                         uamt_threshold=thr)
        self._backward(self.model, plan, r["dstudent"], self.grads)
        alpha = min(1 - 1 / (self.cur_itrs + 1), self.ema_decay)
        lr = self._sgd(self.model, self.grads, self.mom, self.ema_model, alpha)
        self.last = dict(scalars=r["scalars"], lr=lr, w=w, threshold=thr, logits=out, teacher_logits=t_out, mc_logits=mc)
        return r["scalars"][0]


    def _body_dv(self, ins, f, o):
        x, labels, noise, mc_noise = ins
        n_l, dev = labels.shape[0], x.device
        x_u = x[n_l:]
        n_u = x_u.shape[0]
        xr = x_u.repeat(2, 1, 1, 1)
        x_t = (x_u + noise).contiguous()
        x_mc = [(xr + mc_noise[i]).contiguous() for i in range(self.T // 2)]
        ncls, hh, ww = self.num_classes, x.shape[2], x.shape[3]
        mc = self._persistent("mc", (self.T * n_u, ncls, hh, ww), dev)
        main, side = torch.cuda.current_stream(dev), self._side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _, t_out = self._forward_dv(self.ema_model, x_t, False, self._persistent("t_out", (n_u, ncls, hh, ww), dev), o[1:2])
            for i in range(self.T // 2):
                self._forward_dv(self.ema_model, x_mc[i], False, mc[2 * n_u * i:2 * n_u * (i + 1)], o[2 + i:3 + i])
        plan, out = self._forward_dv(self.model, x, True, self._persistent("s_out", (x.shape[0], ncls, hh, ww), dev), o[0:1])
        main.wait_stream(side)
        r = ssl_loss_raw(L.LOSS_UAMT, out, t_out, labels, n_l, mc_logits=mc, mc_passes=self.T, cons_weight_dev=f[3:4],
                         uamt_threshold_dev=f[4:5])
        self._backward(self.model, plan, r["dstudent"], self.grads)
        self._sgd_dv(self.model, self.grads, self.mom, f, self.ema_model)
        return dict(scalars=r["scalars"], logits=out, teacher_logits=t_out, mc_logits=mc)


class ICTStep(_StepBase):
    """Interpolation consistency training (2022_02_ISBI_ICT-MedSeg_ACDC.py:96-140): the student sees the labeled slices
    and a per-sample mix of the two halves of the unlabeled batch; the EMA teacher sees the two un-mixed halves (two
    forwards, as in the reference, so each has its own BatchNorm batch statistics) and the consistency target is the same
    mix of its two softmaxes."""

    def __init__(self, model, ema_model, *, ema_decay=0.99, ict_alpha=0.2, **kw):
        super().__init__(**kw)
        self.model, self.ema_model, self.ema_decay, self.ict_alpha = model, ema_model, ema_decay, ict_alpha
        self.in_channels, self.num_classes = model.in_channels, model.num_classes
        model.train()
        model.ensure_flat()
        ema_model.ensure_flat()
        self.grads = torch.zeros_like(model.flat_params)
        self.mom = torch.zeros_like(model.flat_params)

    def draw_mix_factors(self, n_mixed):
        """np.random.beta(alpha, alpha, size=(n_u//2,1,1,1)) (2022_02...:112-113): host RNG, as in the reference."""
        import numpy as np
        return torch.tensor(np.random.beta(self.ict_alpha, self.ict_alpha, size=(n_mixed,)), dtype=torch.float)

    def step(self, x, labels, mix_factors=None):
        """x: [n_l+n_u, C, H, W] (labeled first, n_u even); mix_factors: [n_u//2] floats (drawn here if None)."""
        self.cur_itrs += 1
        n_l = labels.shape[0]
        n_u = x.shape[0] - n_l
        assert n_u % 2 == 0, "ICT needs an even unlabeled batch"
        n_m = n_u // 2
        dev = x.device
        if mix_factors is None:
            mix_factors = self.draw_mix_factors(n_m)
        lam = mix_factors.reshape(-1).to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
        if self._use_graph():
            dyn = self._dyn_scalars(self.ema_decay)
            out = self._graph_replay([x, labels, lam], dyn, [self.model, self.ema_model, self.ema_model], self._body_dv)
            self.last = dict(scalars=out["scalars"], lr=dyn[0], w=dyn[3], logits=out["logits"],
                             teacher_logits=out["teacher_logits"], mix_factors=lam)
            return out["scalars"][0]
        ux0, ux1 = x[n_l:n_l + n_m], x[n_l + n_m:]
        ncls, hh, ww = self.num_classes, x.shape[2], x.shape[3]
        main = torch.cuda.current_stream(dev)
        side = main if getattr(self, "serialize", False) else self._side_stream(dev)
        side.wait_stream(main)
        t_out = self._persistent("t_out", (n_u, ncls, hh, ww), dev)
        with torch.cuda.stream(side):            # the two teacher forwards are independent of the student forward
            self._forward(self.ema_model, ux0, False, out=t_out[:n_m])
            self._forward(self.ema_model, ux1, False, out=t_out[n_m:])
        x_in = self._persistent("x_in", (n_l + n_m,) + tuple(x.shape[1:]), dev)
        x_in[:n_l].copy_(x[:n_l])
        L.check(L.lib().hpfg_ict_mix(L.ptr(ux0), L.ptr(ux1), L.ptr(lam), n_m, ux0[0].numel(), L.ptr(x_in[n_l:]),
                                     L.stream_ptr(dev)), "hpfg_ict_mix")
        plan, out = self._forward(self.model, x_in, True, out=self._persistent("s_out", (n_l + n_m, ncls, hh, ww), dev))
        main.wait_stream(side)
        w = self._consistency_weight()
        r = ict_loss_raw(out, t_out, lam, labels, n_l, cons_weight=w)
        self._backward(self.model, plan, r["dstudent"], self.grads)
        alpha = min(1 - 1 / (self.cur_itrs + 1), self.ema_decay)
        lr = self._sgd(self.model, self.grads, self.mom, self.ema_model, alpha)
        self.last = dict(scalars=r["scalars"], lr=lr, w=w, logits=out, teacher_logits=t_out, mix_factors=lam)
        return r["scalars"][0]

    def _body_dv(self, ins, f, o):
        x, labels, lam = ins
        n_l, dev = labels.shape[0], x.device
        n_m = (x.shape[0] - n_l) // 2
        ux0, ux1 = x[n_l:n_l + n_m], x[n_l + n_m:]
        ncls, hh, ww = self.num_classes, x.shape[2], x.shape[3]
        main, side = torch.cuda.current_stream(dev), self._side_stream(dev)
        side.wait_stream(main)
        t_out = self._persistent("t_out", (2 * n_m, ncls, hh, ww), dev)
        with torch.cuda.stream(side):
            self._forward_dv(self.ema_model, ux0, False, t_out[:n_m], o[1:2])
            self._forward_dv(self.ema_model, ux1, False, t_out[n_m:], o[2:3])
        x_in = self._persistent("x_in", (n_l + n_m,) + tuple(x.shape[1:]), dev)
        x_in[:n_l].copy_(x[:n_l])
        L.check(L.lib().hpfg_ict_mix(L.ptr(ux0), L.ptr(ux1), L.ptr(lam), n_m, ux0[0].numel(), L.ptr(x_in[n_l:]),
                                     L.stream_ptr(dev)), "hpfg_ict_mix")
        plan, out = self._forward_dv(self.model, x_in, True, self._persistent("s_out", (n_l + n_m, ncls, hh, ww), dev), o[0:1])
        main.wait_stream(side)
        r = ict_loss_raw(out, t_out, lam, labels, n_l, cons_weight_dev=f[3:4])
        self._backward(self.model, plan, r["dstudent"], self.grads)
        self._sgd_dv(self.model, self.grads, self.mom, f, self.ema_model)
        return dict(scalars=r["scalars"], logits=out, teacher_logits=t_out)
