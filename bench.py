#!/usr/bin/env python
"""bench.py -- throughput of the HPFG semi-supervised U-Net training step (Mean-Teacher) on B200.

    python bench.py --gpus N --steps K --warmup W              # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host CPU cores

A "step" is one Mean-Teacher iteration (2017_03_NIPS_Mean-Teacher_ACDC.py:89-113): student forward+backward,
teacher forward, fused SSL loss, SGD and EMA, over one synthetic batch of ACDC-shaped 1x224x224 slices
(config mean_teacher_unet_30k_224x224_ACDC: 8 labeled + 24 unlabeled per GPU, bf16 kernels).  Prints ONE JSON
line (rank 0).  N>1 is launched by torchrun, one rank per GPU, data parallel (weak scaling: 32 images per GPU).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "UNet SSL train images/sec @224x224 (Mean-Teacher step)"
N_L, N_U, IN_CH, N_CLS, H, W = 8, 24, 1, 4, 224, 224
F_FWD = 4516642816.0                 # conv FLOPs per image, forward
TOP_KERNEL_DRAM_BYTES = 58.16e6      # dram__bytes_read.sum + dram__bytes_write.sum of one launch (profiles/r01_ncu_prof_fprop16_final_raw.txt)
LOSS_DRAM_BYTES = 9657856 + 2048 + 48247552 + 169216      # reduce read+write, gradient read+write (profiles/r01_ncu_loss_kernels.csv)
F_IN0 = 2.0 * 9 * H * W * IN_CH * 16
MT_FLOP_PER_IMAGE = 4 * F_FWD - F_IN0    # student fwd+bwd + teacher fwd


def _hbm_roofline(kernel, algo_bytes, cat, pk, note, traffic=None, traffic_source=None):
    """Achieved algorithmic GB/s of one kernel category of the serialized profiling pass (CUDA events in the library)."""
    if cat["ms_per_step"] <= 0:
        return None
    gbs = algo_bytes / (cat["ms_per_step"] * 1e-3) / 1e9
    return {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
            "us_per_step": cat["ms_per_step"] * 1e3, "algorithmic_bytes": algo_bytes, "traffic": traffic,
            "traffic_source": traffic_source, "note": note}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def synthetic_batch(seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(N_L + N_U, IN_CH, H, W, generator=g)
    y = torch.randint(0, N_CLS, (N_L, H, W), generator=g)
    return x, y


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_steps(steps, warmup, n_l, n_u):
    """The reference algorithm (oracle port of the MT step, torch CPU fp32, all host threads) on a bounded sample."""
    import torch
    import oracle
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(1337)
    st = oracle.init_unet_state(IN_CH, N_CLS)
    te = {k: v.clone() for k, v in st.items()}
    opt = oracle.SGDState()
    g = torch.Generator().manual_seed(1337)
    x_l = torch.rand(n_l, IN_CH, H, W, generator=g)
    x_u = torch.rand(n_u, IN_CH, H, W, generator=g)
    y = torch.randint(0, N_CLS, (n_l, H, W), generator=g)
    it = 0
    for _ in range(warmup):
        it += 1
        oracle.mt_step(st, te, opt, x_l, x_u, y, it)
    t0 = time.perf_counter()
    for _ in range(steps):
        it += 1
        oracle.mt_step(st, te, opt, x_l, x_u, y, it)
    dt = time.perf_counter() - t0
    return (n_l + n_u) * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_l, n_u = 2, 6                                   # a quarter of the configured batch per step (bounded sample)
    ips, sec = cpu_reference_steps(args.steps, args.warmup, n_l, n_u)
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "mean_teacher_unet_30k_224x224_ACDC (UNet 1ch/4cls, 8 labeled + 24 unlabeled 224x224 per step)",
                       "note": "reference algorithm on host CPU cores (cuda: False path), oracle port, torch fp32"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "%d+%d images per step (1/4 of the configured 8+24 batch), %d steps" % (n_l, n_u, args.steps)},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import hpfg_b200 as hb
    from hpfg_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()

    torch.manual_seed(1337)                              # identical replicas on every rank
    student = hb.UNet(IN_CH, N_CLS, precision="bf16").to(dev)
    import copy
    teacher = copy.deepcopy(student)
    step = hb.MeanTeacherStep(student, teacher)
    graph_on = not args.eager
    step.enable_graph(graph_on, data_parallel=world > 1)     # whole-step CUDA graph replay (NCCL bucket all-reduces captured too); --eager disables it
    x_cpu, y_cpu = synthetic_batch(1337 + rank)          # rank-distinct data, weak scaling
    x_pin, y_pin = x_cpu.pin_memory(), y_cpu.pin_memory()
    x_dev, y_dev = x_pin.to(dev), y_pin.to(dev)
    loss_pin = torch.empty(1, dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def resident_step():
        return step.step(x_dev, y_dev)

    # End-to-end leg: every step's inputs start in pinned HOST memory and the step's loss is read back to the host and
    # waited for (the reference trainers call loss.item() every iteration).  The copy of step k+1's batch is enqueued
    # on a copy stream while step k computes (what a prefetching data loader does); all K copies and K loss
    # read-backs happen inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(x_dev), torch.empty_like(y_dev)) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0, "primed": False}

    def enqueue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])            # the step that last read this slot has finished
            bufs[slot][0].copy_(x_pin, non_blocking=True)
            bufs[slot][1].copy_(y_pin, non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_step():
        k = state["k"]
        slot = k & 1
        if not state["primed"]:                               # first step of a timed run: its copy is not hidden
            enqueue_copy(slot)
            state["primed"] = True
        main = torch.cuda.current_stream()
        main.wait_event(copied[slot])
        loss = step.step(bufs[slot][0], bufs[slot][1])
        consumed[slot].record(main)
        enqueue_copy(slot ^ 1)                                # prefetch the next step's batch while this step computes
        loss_pin.copy_(loss.reshape(1), non_blocking=True)
        main.synchronize()                                    # the host consumes the loss every step
        state["k"] = k + 1

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        resident_step()
    launches0 = lib.hpfg_launch_count() + step.replayed_kernels
    ms = timed(resident_step, args.steps)
    launches = lib.hpfg_launch_count() + step.replayed_kernels - launches0      # host launches + kernel nodes replayed
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    state["primed"] = False
    ms_e2e = timed(e2e_step, args.steps)
    if sampler:
        sampler.stop_flag = True
    images = (N_L + N_U) * world
    value = images * args.steps / (ms * 1e-3)
    e2e_value = images * args.steps / (ms_e2e * 1e-3)

    # ---- per-category device time (separate, untimed pass): roofline of the dominant kernel family
    prof_steps = 3
    step.enable_graph(False)
    step.serialize = True                                # no side streams: per-category times must not overlap
    resident_step()
    lib.hpfg_profile_begin()
    for _ in range(prof_steps):
        resident_step()
    cat_ms = (ctypes.c_double * 8)()
    cat_calls = (ctypes.c_int64 * 8)()
    lib.hpfg_profile_end(cat_ms, cat_calls)
    step.serialize = False

    def layer_us(op, cin, cout, ks, res, iters=20):
        ms_l = ctypes.c_float()
        L.check(lib.hpfg_conv_tc_bench(op, N_L + N_U, res, res, cin, cout, ks, iters, ctypes.byref(ms_l), L.stream_ptr(dev)), "hpfg_conv_tc_bench")
        return ms_l.value * 1e3
    names = ["conv_tcgen05", "conv_cuda_core", "wgrad_cuda_core", "bn_pool_upsample_glue", "ssl_loss", "sgd_ema", "weight_pack", "wgrad_tcgen05"]
    prof = {names[i]: {"ms_per_step": cat_ms[i] / prof_steps, "calls_per_step": cat_calls[i] / prof_steps} for i in range(8)}
    pk = peaks()
    final_loss = step.last["scalars"][0].item()

    if rank == 0:
        # tensor-core conv family: algorithmic FLOPs = every conv except in_conv.0 / out_conv fwd+dgrad (CUDA cores)
        # and all wgrads that still run on CUDA cores are excluded from the numerator of THIS kernel's roofline.
        n_img = N_L + N_U
        n_params = student.flat_params.numel()
        f_out = 2.0 * 9 * H * W * 16 * N_CLS
        tc_fwd = F_FWD - F_IN0 - f_out                   # per image per forward
        tc_flops = n_img * (2 * tc_fwd + tc_fwd)         # student fwd + teacher fwd + dgrad (same GEMMs transposed)
        tc_ms = prof["conv_tcgen05"]["ms_per_step"]
        achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        # CPU baseline: the FULL configured batch (8+24: BatchNorm and Dice see the real batch), 1 warm-up + 5 timed steps
        # = about 10 s of host work on the 16-core box
        cpu_ips, cpu_sec = cpu_reference_steps(5, 1, N_L, N_U) if world == 1 and not args.no_cpu else (None, None)
        # the single most expensive kernel instance: the 3x3 16->16 conv at 224x224 (4 launches per forward); it is bound by
        # HBM / shared-memory operand bandwidth, not by the tensor pipe: algorithmic bytes = bf16 in + out once
        top_bytes = n_img * H * W * (16 + 16) * 2.0
        top_us = layer_us(0, 16, 16, 3, H) if world == 1 else None
        layers = []
        if world == 1:
            for cin, cout, ks, res in [(16, 16, 3, 224), (32, 16, 3, 224), (32, 32, 3, 112), (64, 64, 3, 56), (128, 128, 3, 28), (256, 256, 3, 14)]:
                px = n_img * res * res
                t3 = [layer_us(op, cin, cout, ks, res) for op in (0, 1, 2)]
                layers.append({"layer": "%dx%d conv %d->%d @%d" % (ks, ks, cin, cout, res), "fprop_us": t3[0], "dgrad_us": t3[1], "wgrad_us": t3[2],
                               "fprop_hbm_frac": px * (cin + cout) * 2.0 / (t3[0] * 1e-6) / 1e9 / pk["hbm"],
                               "fprop_tensor_frac": 2.0 * px * cin * cout * ks * ks / (t3[0] * 1e-6) / 1e12 / pk["tf_burst"]})
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": "mean_teacher_unet_30k_224x224_ACDC (UNet 1ch/4cls, 8 labeled + 24 unlabeled 224x224 per GPU per step)",
                           "global_batch": images, "parallelism": "dp%d" % world, "launch": "cuda-graph replay" if graph_on else "eager (PDL)", "l2": "per-step working set (~2 GB of activations) exceeds the 126 MB L2",
                           "final_loss": final_loss},
                "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": x_pin.numel() * 4 + y_pin.numel() * 8, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches),
                "conv_tensor_fraction_of_step": {"achieved_tflops_whole_step": value / world * MT_FLOP_PER_IMAGE / 1e12,
                                                 "frac_of_peak": value / world * MT_FLOP_PER_IMAGE / 1e12 / pk["tf_sustained"]},
                "roofline": {"kernel": "tc_conv_kernel (tcgen05 implicit-GEMM conv family: fprop + dgrad + 1x1)", "bound": "tensor",
                             "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                             "frac": achieved / pk["tf_sustained"], "traffic": None, "peak_source": pk["src"] + " sustained bf16",
                             "note": "all tensor-core conv launches of a step (serialized profiling pass); FLOPs = SURVEY 8d"},
                "roofline_top_kernel": None if top_us is None else {
                    "kernel": "tc_conv_kernel<3,16,16,RES,MT=4,XF=1> (3x3 16->16 @224, fused BN+LeakyReLU loader, BN-stat epilogue)",
                    "bound": "hbm", "achieved": top_bytes / (top_us * 1e-6) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": top_bytes / (top_us * 1e-6) / 1e9 / pk["hbm"], "us_per_launch": top_us,
                    "traffic": TOP_KERNEL_DRAM_BYTES, "traffic_source": "ncu --set full, profiles/r01_ncu_prof_fprop16_final_raw.txt (dram read+write per launch; the 51 MB output mostly stays in the 126 MB L2)"},
                "roofline_loss": _hbm_roofline(
                    "loss_reduce_kernel<4,MT> + loss_grad_kernel<4> (fused SSL loss value + dlogits, one call per step)",
                    ((N_L + N_U) * N_CLS * H * W * 4.0 * 2 + N_U * N_CLS * H * W * 4.0 + N_L * H * W * 8.0),
                    prof["ssl_loss"], pk, "SURVEY 8d: student logits read + teacher logits read + int64 labels read + dlogits written, "
                    "each once (the labeled logits are read twice: Dice needs batch-wide sums before the gradient); issue-bound "
                    "(accurate expf softmax of student and teacher, ~640 instructions per 4-pixel quad), not HBM-bound: profiles/README.md",
                    traffic=LOSS_DRAM_BYTES, traffic_source="ncu, profiles/r01_ncu_loss_kernels.csv (dram read+write of reduce + gradient, "
                    "L2 flushed before the call: the reduce kernel fetches the labeled 9.7 MB, the gradient kernel all 48.2 MB of inputs; the dlogits stay in L2)"),
                "roofline_sgd_ema": _hbm_roofline(
                    "sgd_kernel<EMA> (SGD momentum + weight decay + EMA teacher, one pass over the flat buffers)",
                    28.0 * n_params, prof["sgd_ema"], pk, "28 B per parameter: read p, g, m, ema; write p, m, ema"),
                "layer_table": layers,
                "kernel_time_per_step": prof,
                "clocks": sampler.summary() if sampler else None}
        if cpu_ips is not None:
            line["cpu_baseline"] = {"value": cpu_ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "the configured 8+24 images per step, 5 timed steps after 1 warm-up (%.1f s per step), torch fp32, all host threads" % cpu_sec}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from the host each step instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
