"""Fused training steps for the three north-star trainers, built from the C-ABI calls only:

* ``MeanTeacherStep``  -- 2017_03_NIPS_Mean-Teacher_ACDC.py:89-113
* ``CPSStep``          -- 2021_06_CVPR_CPS_ACDC.py:90-120
* ``UAMTStep``         -- 2019_07_MICCAI_Uncertainty_Aware_ACDC.py:120-170
* ``ICTStep``          -- 2022_02_ISBI_ICT-MedSeg_ACDC.py:96-140
* ``S4CVStep``         -- 2022_08_CVPR_S4CVNet_ACDC.py:108-167
* ``HPFGStep``         -- main.py:128-207, the HPFG iteration itself

Each ``step()`` enqueues: student forward (activations kept in the plan), teacher/peer forward(s), ONE fused
loss launch pair (value + dlogits), backward into a persistent flat gradient buffer, (data parallel: NCCL
all-reduce of the gradient buckets on a side stream, overlapped with the rest of backward), and ONE fused
SGD-momentum(+EMA) pass over the flat parameter buffers.  Nothing synchronises the host; the returned loss is
a device scalar.  Host-side schedules (Medical_LR, consistency ramp-up, EMA alpha) are python floats exactly as
in the reference (utils/scheduler/medical_lr.py:13-17, utils/utils.py:67-86)."""
import math

import torch

from . import _lib as L
from .losses import ssl_loss_raw, ict_loss_raw, ict_mix_inputs, s4cv_loss_raw
from .utils import sigmoid_rampup, linear_rampup


def medical_lr(cur_itrs, base_lr, max_iterations):
    """Learning rate used at iteration ``cur_itrs`` (1-based) by Medical_LR constructed with last_epoch=-1."""
    return base_lr * (1.0 - (cur_itrs - 2) / max_iterations) ** 0.9


def gradient_buckets(in_channels, num_classes):
    """[(offset, count)] of the flat-gradient buckets in the order backward completes them (tail of the
    parameter list first: out_conv/up4/up3 | up2/up1 | down4/down3 | in_conv..down2: the last bucket, whose all-reduce cannot
    overlap backward, is the smallest)."""
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


def allreduce_tensor_list(tensors, group=None):
    """Sum-all-reduce a list of (small) tensors as ONE flat buffer and scatter the sums back in place: the neck gradients of a
    UNet_Plus (16 tensors) travel as one collective instead of 16.  CPU + gloo and CUDA + NCCL alike."""
    import torch.distributed as dist
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, group=group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()


def shard_batch(x_l, x_u, y, rank, world):
    """Data-parallel split that keeps the labeled:unlabeled ratio per rank (both sub-batches are split, because the
    supervised and consistency terms normalise over their own pixels)."""
    n_l, n_u = x_l.shape[0], x_u.shape[0]
    assert n_l % world == 0 and n_u % world == 0, "labeled and unlabeled batch sizes must divide by the world size"
    a, b = n_l // world, n_u // world
    return x_l[rank * a:(rank + 1) * a], x_u[rank * b:(rank + 1) * b], y[rank * a:(rank + 1) * a]


class _StepBase:
    def __init__(self, *, lr=0.01, momentum=0.9, weight_decay=1e-4, total_itrs=30000, consistency=0.1,
                 consistency_rampup=200.0, process_group=None, exact_global=False):
        self.base_lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.total_itrs, self.consistency, self.consistency_rampup = total_itrs, consistency, consistency_rampup
        self.cur_itrs = 0
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self._comm = None
        self.last = {}
        # Data parallel semantics: default = what wrapping the reference in DDP gives (BatchNorm / Dice / CE /
        # consistency means over each rank's shard, gradients averaged).  exact_global=True = the single-GPU reference on the
        # CONCATENATED batch: BatchNorm statistics (forward and backward) and the loss sums are all-reduced over the ranks
        # through the library's hook, parameter gradients are summed.  Eager launches only (one small all-reduce per BatchNorm).
        self.exact_global = bool(exact_global) and self.world > 1
        self._grad_scale = 1.0 if self.exact_global else 1.0 / self.world
        if self.exact_global:
            L.install_allreduce_hook(self.pg)

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
            self._ggraph = None
            self._gsig = None

    _graph_dp_capable = True      # the bucketed NCCL all-reduces are captured with the step (a refused capture falls back to eager launches)

    def _use_graph(self):
        dp_ok = self.world == 1 or (self._graph_dp and self._graph_dp_capable)
        return self._graph_enabled and self.cur_itrs >= 2 and dp_ok and not getattr(self, "exact_global", False)

    class _GlobalSums:
        """Loss sums over the batch of all ranks while a fused loss call is enqueued (hpfg_ssl_loss_set_global_sums)."""

        def __init__(self, on):
            self.on = on

        def __enter__(self):
            if self.on:
                L.check(L.lib().hpfg_ssl_loss_set_global_sums(1), "hpfg_ssl_loss_set_global_sums")

        def __exit__(self, *exc):
            if self.on:
                L.check(L.lib().hpfg_ssl_loss_set_global_sums(0), "hpfg_ssl_loss_set_global_sums")

    def _global_sums(self):
        return self._GlobalSums(getattr(self, "exact_global", False))

    def _global_consistency_value(self, scalars, w):
        """Mean-Teacher / ICT in exact-global mode: the kernel's consistency value is this rank's share of the global mean."""
        if getattr(self, "exact_global", False):
            torch.distributed.all_reduce(scalars[2:3], group=self.pg)
            scalars[0:1].copy_(scalars[1:2] + w * scalars[2:3])

    def _graph_replay(self, inputs, dyn_f, fwd_models, body):
        """Generic capture-once / replay driver (all step drivers).
        inputs: tensors copied into static buffers before each replay; dyn_f: python floats for the device block
        (fp32); fwd_models: one entry per network forward in the eager call order -> one Philox-offset slot each (the
        offsets advance exactly as the eager path advances them); body(static_inputs, dyn_f_dev, dyn_o_dev) enqueues the
        step with the `_dv` entry points and returns its output dict (static tensors).

        The per-iteration scalars travel through a RING of pinned host slots: an async H2D copy reads pinned memory when
        the stream reaches it, not at enqueue time, so a slot is only rewritten after the event recorded behind its copy
        has completed (the host may run up to `_RING` replays ahead of the device without syncing).
        The captured graph bakes in raw pointers (flat parameters, BN buffers, plan workspaces, gradient / momentum
        buffers): they are part of the signature, and a change (`.to()`, `load_state_dict(assign=True)`, ...) drops the
        graph and captures again."""
        dev = inputs[0].device
        models = []
        for m in fwd_models:
            if not any(m is q for q in models):
                models.append(m)
        for m in models:
            m.ensure_flat(quick=True)
        ptrs = tuple((m.flat_params.data_ptr(), m.bn_running.data_ptr(), m.bn_counters.data_ptr(), id(m._plans)) for m in models)
        ptrs += tuple(t.data_ptr() for t in self._state_tensors())
        # (the dropout seed is a by-value launch argument: a re-seeded generator needs a new capture)
        ptrs += (torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()].initial_seed(),)
        sig = tuple((tuple(t.shape), t.dtype) for t in inputs) + (len(dyn_f), len(fwd_models), ptrs)
        if getattr(self, "_gsig", None) != sig:
            self._gin = [torch.empty_like(t) for t in inputs]
            self._gf = torch.zeros(len(dyn_f), device=dev, dtype=torch.float32)
            self._go = torch.zeros(len(fwd_models), device=dev, dtype=torch.int64)
            self._ring = [(torch.zeros(len(dyn_f), dtype=torch.float32).pin_memory(),
                           torch.zeros(len(fwd_models), dtype=torch.int64).pin_memory(), torch.cuda.Event())
                          for _ in range(self._RING)]
            self._ring_used = [False] * self._RING
            self._ring_pos = 0
            self._ggraph, self._gsig = None, sig
        k = self._ring_pos
        self._ring_pos = (k + 1) % self._RING
        f_host, o_host, ev = self._ring[k]
        if self._ring_used[k]:
            ev.synchronize()               # only blocks when the host is a whole ring ahead of the device
        for i, v in enumerate(dyn_f):
            f_host[i] = v
        for i, m in enumerate(fwd_models):
            o_host[i] = m._philox_offset
            if m.training:
                m._philox_offset += 8
        self._gf.copy_(f_host, non_blocking=True)
        self._go.copy_(o_host, non_blocking=True)
        ev.record()
        self._ring_used[k] = True
        for s, t in zip(self._gin, inputs):
            s.copy_(t, non_blocking=True)
        if self._ggraph is None:
            g = torch.cuda.CUDAGraph()
            n0 = L.lib().hpfg_launch_count()
            try:
                with torch.cuda.graph(g):
                    self._gout = body(self._gin, self._gf, self._go)
            except Exception as exc:        # capture not possible here (e.g. a collective that refuses capture): stay eager
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

    _RING = 8

    def _state_tensors(self):
        """Optimiser-side device buffers of this driver (gradient + momentum buffers), in a fixed order."""
        return [getattr(self, n) for n in ("grads", "mom", "g1", "b1", "g2", "b2") if hasattr(self, n)]

    def _models(self):
        return [getattr(self, n) for n in ("model", "ema_model", "m1", "m2") if hasattr(self, n)]

    # ------------------------------------------------------------------ checkpoint / resume (2017_03...:126-133)
    def state_dict(self):
        """{cur_itrs, philox offsets, optimizer}: `optimizer` holds torch.optim.SGD-style per-parameter momentum buffers
        (state[i]['momentum_buffer'] in model.parameters() order) so that a reference checkpoint's optimizer state and
        this driver's flat momentum buffer interchange."""
        out = {"cur_itrs": self.cur_itrs, "philox": [m._philox_offset for m in self._models()], "optimizers": []}
        pairs = [(getattr(self, a, None), getattr(self, b, None)) for a, b in (("model", "mom"), ("m1", "b1"), ("m2", "b2"))]
        for model, buf in pairs:
            if model is None or buf is None:
                continue
            state = {i: {"momentum_buffer": buf[o:o + n].view(shape).clone()} for i, (o, n, shape) in enumerate(model._layout)}
            out["optimizers"].append({"state": state if self.cur_itrs > 0 else {},
                                      "param_groups": [{"lr": medical_lr(self.cur_itrs + 1, self.base_lr, self.total_itrs),
                                                        "momentum": self.momentum, "weight_decay": self.weight_decay,
                                                        "params": list(range(len(model._layout)))}]})
        return out

    def load_state_dict(self, sd):
        self.cur_itrs = int(sd["cur_itrs"])
        for m, off in zip(self._models(), sd.get("philox", [])):
            m._philox_offset = int(off)
        pairs = [(getattr(self, a, None), getattr(self, b, None)) for a, b in (("model", "mom"), ("m1", "b1"), ("m2", "b2"))]
        pairs = [(m, b) for m, b in pairs if m is not None and b is not None]
        for (model, buf), opt in zip(pairs, sd.get("optimizers", [])):
            buf.zero_()
            for i, (o, n, shape) in enumerate(model._layout):
                st = opt["state"].get(i)
                if st is not None and st.get("momentum_buffer") is not None:
                    buf[o:o + n].copy_(st["momentum_buffer"].reshape(-1))

    def _check_buffers(self):
        """The flat gradient / momentum buffers must match the (possibly re-flattened or moved) parameter buffer."""
        for model, names in ((getattr(self, "model", None), ("grads", "mom")), (getattr(self, "m1", None), ("g1", "b1")),
                             (getattr(self, "m2", None), ("g2", "b2"))):
            if model is None:
                continue
            flat = model.flat_params
            for n in names:
                buf = getattr(self, n)
                if buf.device != flat.device or buf.numel() != flat.numel():
                    raise L.HpfgError("step driver buffer '%s' (%s, %d) no longer matches the model's flat parameters (%s, %d): "
                                      "build the step driver after moving the model" % (n, buf.device, buf.numel(), flat.device, flat.numel()))
            if flat.is_cuda and flat.device.index != torch.cuda.current_device():
                raise L.HpfgError("the model lives on %s but the current CUDA device is %d: call torch.cuda.set_device first "
                                  "(the library launches on the current device)" % (flat.device, torch.cuda.current_device()))

    def _broadcast_from_rank0(self):
        """Data parallel: every rank starts from rank 0's parameters / BN buffers (what DDP does at construction)."""
        if self.world <= 1:
            return
        import torch.distributed as dist
        for m in self._models():
            m.ensure_flat()
            for t in (m.flat_params, m.bn_running, m.bn_counters):
                dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)

    # Two forwards enqueued on two streams share the machine: half of the SMs each (hpfg_unet_plan_set_forward_ctas).  Only the
    # drivers whose two streams carry about the same forward work do this (Mean-Teacher, CPS, ICT); with unbalanced streams
    # (UAMT: one student forward next to five teacher forwards; S4CVNet) the longer stream would finish on half a machine.
    SHARED_FORWARD_CTAS = 74
    share_forward = True

    def _set_forward_share(self, plan, shared):
        if getattr(self, "exact_global", False) and not getattr(plan, "sync_bn", False):
            L.check(L.lib().hpfg_unet_plan_set_sync_bn(plan.handle, 1), "hpfg_unet_plan_set_sync_bn")
            plan.sync_bn = True
        want = self.SHARED_FORWARD_CTAS if shared else 0
        if getattr(plan, "fwd_ctas", 0) != want:
            L.check(L.lib().hpfg_unet_plan_set_forward_ctas(plan.handle, want), "hpfg_unet_plan_set_forward_ctas")
            plan.fwd_ctas = want

    def _forward_dv(self, model, x, save, out, offset_dev):
        plan = model._acquire_plan(x, need_grad=save)
        self._set_forward_share(plan, self.share_forward)      # (graph replays always run the forwards on two streams)
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
        model.ensure_flat(quick=True)
        plan = model._acquire_plan(x, need_grad=save)
        self._set_forward_share(plan, self.share_forward and not getattr(self, "serialize", False))
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
                                              self.weight_decay, self._grad_scale, first, st), "hpfg_sgd_momentum")
        else:
            L.check(L.lib().hpfg_sgd_momentum_ema(L.ptr(model.flat_params), L.ptr(grads), L.ptr(buf),
                                                  L.ptr(ema_model.flat_params), n, lr, self.momentum, self.weight_decay,
                                                  self._grad_scale, first, ema_alpha, st), "hpfg_sgd_momentum_ema")
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
        self._check_buffers()
        self._broadcast_from_rank0()

    def step(self, x, labels):
        """x: [n_l+n_u, C, H, W] fp32 CUDA (labeled slices first); labels: [n_l, H, W] int64 CUDA."""
        self.cur_itrs += 1
        self._check_buffers()
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
        with self._global_sums():
            r = ssl_loss_raw(L.LOSS_MT, out, t_out[n_l:], labels, n_l, cons_weight=w)
        self._global_consistency_value(r["scalars"], w)
        self._backward(self.model, plan, r["dstudent"], self.grads)
        alpha = min(1 - 1 / (self.cur_itrs + 1), self.ema_decay)      # utils/utils.py:84
        lr = self._sgd(self.model, self.grads, self.mom, self.ema_model, alpha)
        self.last = dict(scalars=r["scalars"], lr=lr, w=w, logits=out, teacher_logits=t_out)
        return r["scalars"][0]

    def _step_graph(self, x, labels):
        dyn = self._dyn_scalars(self.ema_decay)
        out = self._graph_replay([x, labels], dyn, [self.model, self.ema_model], self._step_body_dv)
        self.last = dict(scalars=out["scalars"], lr=dyn[0], w=dyn[3], logits=out["logits"], teacher_logits=out["teacher_logits"])
        return out["scalars"][0]

    def _step_body_dv(self, ins, dyn_f, dyn_o):
        """The step with every per-iteration scalar read from the device block (captured once, replayed)."""
        x, labels = ins
        n_l = labels.shape[0]
        dev = x.device
        main, side = torch.cuda.current_stream(dev), self._side_stream(dev)
        shape = (x.shape[0], self.num_classes, x.shape[2], x.shape[3])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            tplan = self.ema_model._acquire_plan(x, need_grad=False)
            self._set_forward_share(tplan, True)
            t_out = self.ema_model._run_forward(tplan, x, save=False, out=self._persistent("t_out", shape, dev),
                                                offset_dev=dyn_o[1:2])
        plan = self.model._acquire_plan(x, need_grad=True)
        self._set_forward_share(plan, True)
        out = self.model._run_forward(plan, x, save=True, out=self._persistent("s_out", shape, dev), offset_dev=dyn_o[0:1])
        main.wait_stream(side)
        r = ssl_loss_raw(L.LOSS_MT, out, t_out[n_l:], labels, n_l, cons_weight_dev=dyn_f[3:4])
        self._backward(self.model, plan, r["dstudent"], self.grads)
        n = self.model.flat_params.numel()
        L.check(L.lib().hpfg_sgd_momentum_ema_dv(L.ptr(self.model.flat_params), L.ptr(self.grads), L.ptr(self.mom),
                                                 L.ptr(self.ema_model.flat_params), n, self.momentum, self.weight_decay,
                                                 1.0 / self.world, 0, L.ptr(dyn_f), L.stream_ptr(dev)),
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
        self._check_buffers()
        self._broadcast_from_rank0()

    def step(self, x, labels):
        self.cur_itrs += 1
        self._check_buffers()
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
        with self._global_sums():
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
    share_forward = False

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
        self._check_buffers()
        self._broadcast_from_rank0()

    @staticmethod
    def make_noise(like):
        return torch.clamp(torch.randn_like(like) * 0.1, -0.2, 0.2)      # 2019_07...:130,141

    def step(self, x, labels, noise=None, mc_noise=None, teacher_masks=None):
        """noise [n_u,...] / mc_noise [T//2, 2*n_u, ...]: the clamped perturbations; drawn here if None.
        teacher_masks (parity harness only, eager path): 1 + T//2 dropout keep-mask sets, one per teacher forward in call
        order (the consistency forward on n_u slices, then the T//2 Monte-Carlo forwards on 2*n_u slices) -- every
        stochastic teacher pass draws its own masks, as nn.Dropout does in the reference (2019_07...:134-143)."""
        self.cur_itrs += 1
        self._check_buffers()
        n_l = labels.shape[0]
        x_u = x[n_l:]
        n_u = x_u.shape[0]
        if noise is None:
            noise = self.make_noise(x_u)
        if self._use_graph() and teacher_masks is None:
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
        if teacher_masks is not None:
            assert len(teacher_masks) == 1 + self.T // 2, "one mask set per teacher forward"
            dev_sets = []                  # uploaded on the main stream before the fork; kept alive until the next step
            for ms in teacher_masks:
                self.ema_model.set_dropout_masks(ms)
                dev_sets.append(self.ema_model._dropout_masks)
            self._mask_keepalive = dev_sets
            side.wait_stream(main)
        with torch.cuda.stream(side):
            if teacher_masks is not None:
                self.ema_model._dropout_masks = dev_sets[0]
            _, t_out = self._forward(self.ema_model, x_t, False, out=self._persistent("t_out", (n_u, ncls, hh, ww), x.device))
            for i in range(self.T // 2):
                if teacher_masks is not None:
                    self.ema_model._dropout_masks = dev_sets[1 + i]
                self._forward(self.ema_model, x_mc[i], False, out=mc[2 * n_u * i:2 * n_u * (i + 1)])
        plan, out = self._forward(self.model, x, True, out=self._persistent("s_out", (x.shape[0], ncls, hh, ww), x.device))
        main.wait_stream(side)
        w = self._consistency_weight()
        thr = (0.75 + 0.25 * sigmoid_rampup(self.cur_itrs, self.total_itrs)) * math.log(2)
        with self._global_sums():
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
        self._check_buffers()
        self._broadcast_from_rank0()

    def draw_mix_factors(self, n_mixed):
        """np.random.beta(alpha, alpha, size=(n_u//2,1,1,1)) (2022_02...:112-113): host RNG, as in the reference."""
        import numpy as np
        return torch.tensor(np.random.beta(self.ict_alpha, self.ict_alpha, size=(n_mixed,)), dtype=torch.float)

    def step(self, x, labels, mix_factors=None):
        """x: [n_l+n_u, C, H, W] (labeled first, n_u even); mix_factors: [n_u//2] floats (drawn here if None)."""
        self.cur_itrs += 1
        self._check_buffers()
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
        with self._global_sums():
            r = ict_loss_raw(out, t_out, lam, labels, n_l, cons_weight=w)
        self._global_consistency_value(r["scalars"], w)
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


class S4CVStep(_StepBase):
    """S4CVNet (2022_08_CVPR_S4CVNet_ACDC.py:108-167): two student networks trained with cross pseudo supervision (Dice only,
    weight 7w) plus, from iteration ``mt_start`` on, Mean-Teacher MSE of both students against the EMA teacher of student 2,
    which sees the noise-perturbed unlabeled slices.  w = consistency * linear_rampup(cur_itrs // 150, rampup) (:148-149).
    One fused loss call (``hpfg_s4cv_loss``) produces the value and both logit gradients."""
    share_forward = False

    def __init__(self, model1, model2, ema_model, *, ema_decay=0.99, mt_start=1000, **kw):
        super().__init__(**kw)
        self.m1, self.m2, self.ema_model, self.ema_decay, self.mt_start = model1, model2, ema_model, ema_decay, mt_start
        self.in_channels, self.num_classes = model1.in_channels, model1.num_classes
        model1.train()
        model2.train()
        for m in (model1, model2, ema_model):
            m.ensure_flat()
        self.g1, self.g2 = torch.zeros_like(model1.flat_params), torch.zeros_like(model2.flat_params)
        self.b1, self.b2 = torch.zeros_like(model1.flat_params), torch.zeros_like(model2.flat_params)
        self._check_buffers()
        self._broadcast_from_rank0()

    def _models(self):
        return [self.m1, self.m2, self.ema_model]

    def _consistency_weight(self):
        return self.consistency * linear_rampup(self.cur_itrs // 150, self.consistency_rampup)

    @staticmethod
    def make_noise(like):
        return torch.clamp(torch.randn_like(like) * 0.1, -0.2, 0.2)      # 2022_08...:111

    def step(self, x, labels, noise=None):
        """x: [n_l+n_u, C, H, W] (labeled first); labels: [n_l, H, W]; noise [n_u, ...]: drawn here if None."""
        self.cur_itrs += 1
        self._check_buffers()
        n_l, dev = labels.shape[0], x.device
        x_u = x[n_l:]
        if noise is None:
            noise = self.make_noise(x_u)
        x_t = (x_u + noise).contiguous()
        ncls, hh, ww = self.num_classes, x.shape[2], x.shape[3]
        main = torch.cuda.current_stream(dev)
        side = main if getattr(self, "serialize", False) else self._side_stream(dev)
        shape = (x.shape[0], ncls, hh, ww)
        side.wait_stream(main)
        with torch.cuda.stream(side):               # student 2 and its teacher on the side stream, student 1 on the main stream
            p2, o2 = self._forward(self.m2, x, True, out=self._persistent("o2", shape, dev))
            _, t_out = self._forward(self.ema_model, x_t, False, out=self._persistent("t_out", (x_u.shape[0], ncls, hh, ww), dev))
        p1, o1 = self._forward(self.m1, x, True, out=self._persistent("o1", shape, dev))
        main.wait_stream(side)
        w = self._consistency_weight()
        mt_on = self.cur_itrs >= self.mt_start
        with self._global_sums():
            r = s4cv_loss_raw(o1, o2, t_out if mt_on else None, labels, n_l, cps_weight=7.0 * w, mt_weight=w)
        side.wait_stream(main)
        alpha = min(1 - 1 / (self.cur_itrs + 1), self.ema_decay)
        with torch.cuda.stream(side):
            self._backward(self.m2, p2, r["dother"], self.g2)
            self._sgd(self.m2, self.g2, self.b2, self.ema_model, alpha)      # SGD of student 2 + EMA into its teacher, one pass
        self._backward(self.m1, p1, r["dstudent"], self.g1)
        lr = self._sgd(self.m1, self.g1, self.b1)
        main.wait_stream(side)
        self.last = dict(scalars=r["scalars"], lr=lr, w=w, logits1=o1, logits2=o2, teacher_logits=t_out)
        return r["scalars"][0]


class HPFGStep(_StepBase):
    """The HPFG iteration of main.py:128-207 on three ``UNet_Plus`` networks (model1 sees the CutMix batch, model2 and the EMA
    teacher the plain batch).  U-Net bodies, projection necks and ``Dense_Loss`` run on the library's kernels behind autograd
    Functions; the pixel losses do not go through autograd at all: model2's supervised + Mean-Teacher terms are the MT mode of the
    fused loss kernels, model1's supervised term the SUP mode, its Dice against the CutMix-pasted pseudo-labels ``hpfg_dice_loss``,
    and their logits gradients are fed to ONE ``torch.autograd.backward`` together with the contrastive term.  The optimiser side is
    fused: flat SGD over the 82 U-Net tensors of each student, one launch per neck tensor, the backbone EMA model2 <- model1
    (main.py:68-76) and the teacher EMA (utils/utils.py:82-86) as flat passes.

    step(label_img, target_label, label_img1, target_label1, img_unlabel, cutmix_mask): the tensors main.py:128-150 builds
    (label_img1 / target_label1: the second labeled draw, repeated to the unlabeled batch size here as at :139-140;
    cutmix_mask [n_u,1,H,W] from BoxMaskGenerator)."""

    def __init__(self, model1, model2, ema_model, *, ema_decay=0.99, mt_start=1000, temperature=0.7, **kw):
        super().__init__(**kw)
        from .losses import Dense_Loss
        self.m1, self.m2, self.ema_model, self.ema_decay, self.mt_start = model1, model2, ema_model, ema_decay, mt_start
        self.in_channels, self.num_classes = model1.in_channels, model1.num_classes
        model1.train()
        model2.train()
        for m in (model1, model2, ema_model):
            m.ensure_flat()
        for p in ema_model.parameters():
            p.requires_grad = False
        self.temperature = temperature
        self._dense = None
        self._DenseLoss = Dense_Loss
        self.b1, self.b2 = torch.zeros_like(model1.flat_params), torch.zeros_like(model2.flat_params)
        self._neck_mom = {}                         # id(parameter) -> momentum buffer of the neck tensors (outside the flat buffer)
        self._check_buffers()
        self._broadcast_from_rank0()

    def _models(self):
        return [self.m1, self.m2, self.ema_model]

    def _state_tensors(self):
        return [self.b1, self.b2]

    def _check_buffers(self):
        for model, buf in ((self.m1, self.b1), (self.m2, self.b2)):
            if buf.device != model.flat_params.device or buf.numel() != model.flat_params.numel():
                raise L.HpfgError("HPFGStep: momentum buffer no longer matches the model's flat parameters")

    def _consistency_weight(self):
        return self.consistency * linear_rampup(self.cur_itrs // 150, self.consistency_rampup)

    def _broadcast_from_rank0(self):
        """The flat buffers as in ``_StepBase``, plus the neck tensors (they live outside the flat buffer)."""
        super()._broadcast_from_rank0()
        if self.world <= 1:
            return
        import torch.distributed as dist
        src = dist.get_global_rank(self.pg, 0) if self.pg is not None else 0
        for m in self._models():
            for p in self._neck_params(m):
                dist.broadcast(p.data, src=src, group=self.pg)

    @staticmethod
    def _neck_params(model):
        """The 16 projection-neck tensors of a UNet_Plus: parameters() order after the 82 U-Net tensors of the flat buffer."""
        core = {id(q) for q in model._flat_params_list}
        return [p for p in model.parameters() if id(p) not in core]

    def state_dict(self):
        """As ``_StepBase.state_dict``, with the neck tensors' momentum buffers at their torch.optim.SGD indices (82..97 of
        ``model.parameters()``); tensors that never received a gradient (model1's necks: main.py:148 drops their outputs) have
        no state, as in torch."""
        out = super().state_dict()
        for model, opt in zip((self.m1, self.m2), out["optimizers"]):
            base = len(model._layout)
            necks = self._neck_params(model)
            opt["param_groups"][0]["params"] = list(range(base + len(necks)))
            for j, p in enumerate(necks):
                mom = self._neck_mom.get(id(p))
                if mom is not None and self.cur_itrs > 0:
                    opt["state"][base + j] = {"momentum_buffer": mom.clone()}
        return out

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        self._neck_mom = {}
        for model, opt in zip((self.m1, self.m2), sd.get("optimizers", [])):
            base = len(model._layout)
            for j, p in enumerate(self._neck_params(model)):
                st = opt["state"].get(base + j)
                if st is not None and st.get("momentum_buffer") is not None:
                    self._neck_mom[id(p)] = st["momentum_buffer"].to(device=p.device, dtype=p.dtype).reshape(p.shape).clone()

    def _sgd_model(self, model, buf, lr, first):
        """torch.optim.SGD semantics over all parameters of a UNet_Plus: the flat U-Net buffer in one pass + the neck tensors."""
        st = L.stream_ptr(model.flat_params.device)
        lib = L.lib()
        g = model.last_flat_grad
        L.check(lib.hpfg_sgd_momentum(L.ptr(model.flat_params), L.ptr(g), L.ptr(buf), model.flat_params.numel(), lr, self.momentum,
                                      self.weight_decay, self._grad_scale, first, st), "hpfg_sgd_momentum")
        for p in self._neck_params(model):
            if p.grad is None:
                continue
            mom = self._neck_mom.get(id(p))
            fresh = mom is None                       # torch.optim.SGD: the first gradient a tensor sees initialises its buffer
            if fresh:
                mom = self._neck_mom[id(p)] = torch.zeros_like(p.data)
            L.check(lib.hpfg_sgd_momentum(L.ptr(p.data), L.ptr(p.grad.contiguous()), L.ptr(mom), p.numel(), lr, self.momentum,
                                          self.weight_decay, self._grad_scale, int(fresh), st),
                    "hpfg_sgd_momentum")

    def step(self, label_img, target_label, label_img1, target_label1, img_unlabel, cutmix_mask, lr=None):
        """lr: override of the Medical_LR value of this iteration (resuming with a fresh scheduler; parity harness).
        cutmix_mask: the 0/1 box masks of BoxMaskGenerator (main.py:141-143)."""
        from .utils import update_ema_variables, ema_update_flat
        from .losses import ssl_loss_raw, dice_loss_raw, argmax_labels
        self.cur_itrs += 1
        self._check_buffers()
        m1, m2, ema = self.m1, self.m2, self.ema_model
        label_bs, unlabel_bs = label_img.shape[0], img_unlabel.shape[0]
        rep = int(unlabel_bs // label_bs)
        label_img1 = label_img1.repeat(rep, 1, 1, 1).float()
        target_label1 = target_label1.repeat(rep, 1, 1).long()
        cutmix_mask = cutmix_mask.float()
        batch_un_mix = label_img1 * (1.0 - cutmix_mask) + img_unlabel * cutmix_mask          # main.py:145
        batch_mix = torch.cat([label_img, batch_un_mix], dim=0).float()
        volume_batch = torch.cat([label_img, img_unlabel], dim=0).float()
        if self._dense is None or self._dense.batch_size != label_bs + unlabel_bs:
            self._dense = self._DenseLoss(label_bs + unlabel_bs, volume_batch.device, self.temperature)
        for m in (m1, m2):
            for p in m.parameters():
                p.grad = None
        outputs1, _, _ = m1(batch_mix)
        outputs2, h1, h2 = m2(volume_batch)
        with torch.no_grad():
            ema_output, ema_h1, ema_h2 = ema(volume_batch)
        tl = target_label.long()
        w = self._consistency_weight()
        # model2 (main.py:157-158,183,186): supervised + w * mean((softmax2_u - softmax_ema_u)^2) = the Mean-Teacher mode of the fused
        # loss kernels (value, loss_sup, consistency value and d/d outputs2 in one launch); the MSE term is 0 before mt_start
        r2 = ssl_loss_raw(L.LOSS_MT, outputs2.detach(), ema_output[label_bs:], tl, label_bs,
                          cons_weight=w if self.cur_itrs >= self.mt_start else 0.0)
        # model1 (main.py:154-155,170-175,185): supervised (SUP mode: zero gradient on the unlabeled images) + 7w * Dice of
        # softmax1_u against the teacher's argmax pasted into the second labeled draw's labels by the CutMix mask
        r1 = ssl_loss_raw(L.LOSS_SUP, outputs1.detach(), None, tl, label_bs)
        pseudo = torch.where(cutmix_mask.squeeze(1) > 0.5, argmax_labels(ema_output[label_bs:]), target_label1)      # main.py:171-172
        ps, d_ps = dice_loss_raw(outputs1.detach()[label_bs:], pseudo, softmax=True)
        dlogits1 = r1["dstudent"]
        dlogits1[label_bs:].add_(d_ps, alpha=7.0 * w)
        loss_contrast = self._dense(h1, ema_h1) + self._dense(h2, ema_h2)                     # main.py:166 (autograd: the necks)
        loss_sup = r1["scalars"][1] + r2["scalars"][1]
        loss = r1["scalars"][0] + 7 * w * ps[0] + r2["scalars"][0] + w * loss_contrast
        # one backward pass: the loss kernels' gradients enter at the logits, the contrastive term through the necks of model2
        torch.autograd.backward([outputs1, outputs2, w * loss_contrast], [dlogits1, r2["dstudent"], None])
        if self.world > 1:      # data parallel (main.py itself is single-process): gradients summed here, scaled by 1/world in the SGD pass
            import torch.distributed as dist
            for m in (m1, m2):
                dist.all_reduce(m.last_flat_grad, group=self.pg)
                allreduce_tensor_list([p.grad for p in self._neck_params(m)], group=self.pg)
        lr = medical_lr(self.cur_itrs, self.base_lr, self.total_itrs) if lr is None else lr
        first = int(self.cur_itrs == 1)
        self._sgd_model(m1, self.b1, lr, first)
        self._sgd_model(m2, self.b2, lr, first)
        alpha = min(1 - 1 / (self.cur_itrs + 1), self.ema_decay)
        ema_update_flat(m2.ensure_flat(), m1.ensure_flat(), alpha)          # update_ema_variables_backbone: encoder + decoder = the flat buffer
        update_ema_variables(m2, ema, self.ema_decay, self.cur_itrs)
        self.last = dict(loss=loss.detach(), loss_sup=loss_sup, contrast=loss_contrast.detach(), pseudo=ps[0], cons2=r2["scalars"][2],
                         lr=lr, w=w, outputs1=outputs1.detach(), outputs2=outputs2.detach(), ema_output=ema_output)
        return loss.detach()
