"""CPU, world_size 2 over gloo: the host-side data-parallel logic (batch sharding that keeps the labeled:unlabeled
ratio, bucketed flat-gradient all-reduce in backward-completion order, 1/world gradient scaling) reproduces what
wrapping the single-process step in DDP would do: identical replicas whose update equals the update from the mean
of the per-rank gradients.  Per-rank compute is the CPU oracle (no GPU here); the all-reduce path is the same
``allreduce_flat_buckets`` / ``gradient_buckets`` / ``shard_batch`` code the CUDA trainer uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import hpfg_b200 as hb
from tests.golden.common import make_state, make_batch, make_masks

IN_CH, N_CLS, H, W = 1, 4, 32, 32
N_L, N_U = 2, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_grads(rank, world):
    st = make_state(IN_CH, N_CLS, 3)
    x_l, x_u, y = make_batch(N_L, N_U, IN_CH, N_CLS, H, W, 4)
    xl, xu, yy = hb.shard_batch(x_l, x_u, y, rank, world)
    n = xl.shape[0] + xu.shape[0]
    masks_all = make_masks(N_L + N_U, H, W, 5)
    idx = list(range(rank * (N_L // world), (rank + 1) * (N_L // world))) + \
        [N_L + i for i in range(rank * (N_U // world), (rank + 1) * (N_U // world))]
    masks = {k: v[idx] for k, v in masks_all.items()}
    names = [nm for nm, _ in oracle.unet_param_spec(IN_CH, N_CLS)]
    leaves = {nm: st[nm].clone().requires_grad_(True) for nm in names}
    view = dict(st)
    view.update(leaves)
    out = oracle.unet_forward(view, torch.cat([xl, xu]), True, masks)
    loss = oracle.med_sup_loss(out[:xl.shape[0]], yy, N_CLS)
    grads = torch.autograd.grad(loss, [leaves[nm] for nm in names])
    assert n == out.shape[0]
    return st, names, torch.cat([g.reshape(-1) for g in grads])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    st, names, flat = _rank_grads(rank, world)
    buckets = hb.gradient_buckets(IN_CH, N_CLS)
    order = []
    works = hb.allreduce_flat_buckets(flat, buckets, before_bucket=order.append)
    for w in works:
        w.wait()
    flat = flat / world                                    # what grad_scale = 1/world does inside the fused SGD kernel
    opt = oracle.SGDState()
    off, grads = 0, {}
    for nm in names:
        n = st[nm].numel()
        grads[nm] = flat[off:off + n].view(st[nm].shape)
        off += n
    oracle.sgd_step(st, grads, opt, 0.01)
    # numpy, not a torch tensor: tensors travel through a file-descriptor side channel that dies with the worker, so a
    # worker that exits before the parent has unpickled its result made the test flaky (FileNotFoundError in q.get)
    q.put((rank, order, torch.cat([st[nm].reshape(-1) for nm in names]).numpy()))
    dist.destroy_process_group()


def test_bucket_layout_covers_parameters_in_completion_order():
    spec = oracle.unet_param_spec(IN_CH, N_CLS)
    sizes = [int(torch.Size(s).numel()) for _, s in spec]
    buckets = hb.gradient_buckets(IN_CH, N_CLS)
    assert len(buckets) == 4 and sum(c for _, c in buckets) == sum(sizes)
    ends = sorted((o, o + c) for o, c in buckets)
    assert ends[0][0] == 0 and all(a[1] == b[0] for a, b in zip(ends, ends[1:]))
    assert [o for o, _ in buckets] == sorted((o for o, _ in buckets), reverse=True)     # tail (out_conv side) first
    names = [n for n, _ in spec]
    start_of = {n: sum(sizes[:i]) for i, n in enumerate(names)}
    assert buckets[0][0] == start_of["decoder.up3.conv1x1.weight"]
    assert buckets[1][0] == start_of["decoder.up1.conv1x1.weight"]
    assert buckets[2][0] == start_of["encoder.down3.maxpool_conv.1.conv_conv.0.weight"]     # (the exposed last bucket stays small)


def test_two_rank_gloo_matches_mean_gradient_update():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    res = [(r, o, torch.from_numpy(a)) for r, o, a in res]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2, 3]                          # buckets enqueued in completion order
    assert torch.equal(res[0][2], res[1][2])                  # replicas stay identical
    # single-process reference: mean of the two per-rank gradients
    st, names, g0 = _rank_grads(0, world)
    _, _, g1 = _rank_grads(1, world)
    flat = (g0 + g1) / world
    off, grads = 0, {}
    for nm in names:
        n = st[nm].numel()
        grads[nm] = flat[off:off + n].view(st[nm].shape)
        off += n
    oracle.sgd_step(st, grads, oracle.SGDState(), 0.01)
    ref = torch.cat([st[nm].reshape(-1) for nm in names])
    assert torch.allclose(res[0][2], ref, rtol=0, atol=1e-7)


def _hpfg_worker(rank, world, port, q):
    """HPFGStep's data-parallel host logic on CPU: replicas built from DIFFERENT seeds are made identical at construction
    (flat U-Net buffers and the neck tensors outside them), and the neck gradients travel as one coalesced sum-all-reduce."""
    import copy
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    torch.manual_seed(100 + rank)
    m1, m2 = hb.UNet_Plus(IN_CH, N_CLS), hb.UNet_Plus(IN_CH, N_CLS)
    step = hb.HPFGStep(m1, m2, copy.deepcopy(m2))
    assert step.world == world and step._grad_scale == 1.0 / world
    digest = torch.cat([p.detach().reshape(-1) for m in (m1, m2, step.ema_model) for p in m.parameters()])
    necks = step._neck_params(m2)
    grads = [torch.full_like(p, float(rank + 1)) * (i + 1) for i, p in enumerate(necks)]
    grads[3] = None                                            # a tensor without gradient is skipped on every rank alike
    hb.allreduce_tensor_list(grads, group=None)
    sums = [float(g.flatten()[0]) if g is not None else None for g in grads]
    q.put((rank, digest.double().sum().item(), digest.abs().double().sum().item(), sums))
    dist.destroy_process_group()


def test_hpfg_step_two_rank_gloo_replicas_and_neck_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_hpfg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:3] == res[1][1:3]                           # all 3 x 98 parameter tensors identical after the broadcast
    want = [3.0 * (i + 1) if i != 3 else None for i in range(16)]        # (1 + 2) * (i + 1)
    assert res[0][3] == want and res[1][3] == want
