"""GPU (>= 2 devices): data-parallel Mean-Teacher step over NCCL against the CPU oracle.  DDP semantics: each rank
runs the step on its shard (per-rank BatchNorm statistics and per-rank loss normalisation), gradients
are averaged by the bucketed all-reduce overlapped with backward, every replica applies the same SGD+EMA update."""
import copy
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

IN_CH, N_CLS, H, W, N_L, N_U = 1, 4, 32, 32, 2, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import hpfg_b200 as hb
    from tests.golden.common import make_state, make_batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    st = make_state(IN_CH, N_CLS, 3)
    student = hb.UNet(IN_CH, N_CLS, precision="fp32")
    student.load_state_dict(st)
    student.to(dev)
    teacher = copy.deepcopy(student)
    student.set_dropout_enabled(False)
    teacher.set_dropout_enabled(False)
    step = hb.MeanTeacherStep(student, teacher)
    assert step.world == world
    losses = []
    for it in (1, 2):
        x_l, x_u, y = make_batch(N_L, N_U, IN_CH, N_CLS, H, W, 10 + it)
        xl, xu, yy = hb.shard_batch(x_l, x_u, y, rank, world)
        losses.append(step.step(torch.cat([xl, xu]).to(dev), yy.to(dev)).item())
    torch.cuda.synchronize()
    q.put((rank, losses, student.flat_params.cpu(), teacher.flat_params.cpu()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_step_matches_oracle_mean_gradient():
    import oracle
    from tests.golden.common import make_state, make_batch
    import hpfg_b200 as hb
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3])      # replicas identical
    # oracle: two independent per-rank steps, gradients averaged, one SGD + EMA update
    names = [n for n, _ in oracle.unet_param_spec(IN_CH, N_CLS)]
    st = make_state(IN_CH, N_CLS, 3)
    students = [{k: v.clone() for k, v in st.items()} for _ in range(world)]
    teachers = [{k: v.clone() for k, v in st.items()} for _ in range(world)]
    opt = oracle.SGDState()
    params = {n: st[n].clone() for n in names}
    ema = {n: st[n].clone() for n in names}
    for it in (1, 2):
        x_l, x_u, y = make_batch(N_L, N_U, IN_CH, N_CLS, H, W, 10 + it)
        grads, loss_r = [], []
        for r in range(world):
            for n in names:
                students[r][n] = params[n].clone()
                teachers[r][n] = ema[n].clone()
            xl, xu, yy = hb.shard_batch(x_l, x_u, y, r, world)
            out = oracle.mt_step(students[r], teachers[r], oracle.SGDState(), xl, xu, yy, it, student_masks={}, teacher_masks={})
            grads.append(out["grads"])
            loss_r.append(out["loss"])
        mean = {n: sum(g[n] for g in grads) / world for n in names}
        oracle.sgd_step(params, mean, opt, oracle.medical_lr(it - 1))
        oracle.update_ema(params, ema, 0.99, it, names)
        for r in range(world):
            assert res[r][1][it - 1] == pytest.approx(loss_r[r], rel=1e-5)
    ref = torch.cat([params[n].reshape(-1) for n in names])
    ref_ema = torch.cat([ema[n].reshape(-1) for n in names])
    assert torch.allclose(res[0][2], ref, rtol=0, atol=2e-6)
    assert torch.allclose(res[0][3], ref_ema, rtol=0, atol=2e-6)


# ---------------------------------------------------------------- exact-global mode (SURVEY 8e caveats 1-3)
def _worker_exact(rank, world, port, q, kind):
    import torch.distributed as dist
    import hpfg_b200 as hb
    from tests.golden.common import make_state, make_batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def net(seed):
        m = hb.UNet(IN_CH, N_CLS, precision="fp32")
        m.load_state_dict(make_state(IN_CH, N_CLS, seed))
        m.to(dev)
        m.set_dropout_enabled(False)
        return m
    a = net(3)
    b = copy.deepcopy(a) if kind == "mt" else net(4)
    b.set_dropout_enabled(False)
    step = hb.MeanTeacherStep(a, b, exact_global=True) if kind == "mt" else hb.CPSStep(a, b, exact_global=True)
    assert step.world == world and step.exact_global
    losses = []
    for it in (1, 2):
        x_l, x_u, y = make_batch(2 * N_L, 2 * N_U, IN_CH, N_CLS, H, W, 20 + it)
        xl, xu, yy = hb.shard_batch(x_l, x_u, y, rank, world)
        losses.append(step.step(torch.cat([xl, xu]).to(dev), yy.to(dev)).item())
    torch.cuda.synchronize()
    sd = a.state_dict()
    q.put((rank, losses, a.flat_params.cpu().numpy(), b.flat_params.cpu().numpy(),
           sd["encoder.down4.maxpool_conv.1.conv_conv.5.running_var"].cpu().numpy(),
           sd["encoder.in_conv.conv_conv.1.running_mean"].cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind", ["mt", "cps"])
def test_two_rank_exact_global_step_equals_single_process_reference(kind):
    """exact_global=True: BatchNorm statistics (forward + backward), Dice / CE / pseudo-label sums and the consistency mean are
    taken over the batch of BOTH ranks and the gradients are summed -- the two-rank step must reproduce the single-process
    oracle step on the concatenated batch (loss, both networks' parameters, BatchNorm running statistics)."""
    import oracle
    from tests.golden.common import make_state, make_batch
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_exact, args=(r, world, port, q, kind)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    names = [n for n, _ in oracle.unet_param_spec(IN_CH, N_CLS)]
    sa = make_state(IN_CH, N_CLS, 3)
    sb = {k: v.clone() for k, v in sa.items()} if kind == "mt" else make_state(IN_CH, N_CLS, 4)
    o1, o2 = oracle.SGDState(), oracle.SGDState()
    ref_losses = []
    for it in (1, 2):
        x_l, x_u, y = make_batch(2 * N_L, 2 * N_U, IN_CH, N_CLS, H, W, 20 + it)
        if kind == "mt":
            r = oracle.mt_step(sa, sb, o1, x_l, x_u, y, it, student_masks={}, teacher_masks={})
        else:
            r = oracle.cps_step(sa, sb, o1, o2, x_l, x_u, y, it, masks1={}, masks2={})
        ref_losses.append(r["loss"])
    for rk in range(world):
        for it in range(2):
            assert res[rk][1][it] == pytest.approx(ref_losses[it], rel=2e-5), (rk, it)
    fa = torch.cat([sa[n].reshape(-1) for n in names]).numpy()
    fb = torch.cat([sb[n].reshape(-1) for n in names]).numpy()
    import numpy as np
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])      # replicas identical
    assert np.allclose(res[0][2], fa, rtol=0, atol=5e-6) and np.allclose(res[0][3], fb, rtol=0, atol=5e-6)
    assert np.allclose(res[0][4], sa["encoder.down4.maxpool_conv.1.conv_conv.5.running_var"].numpy(), rtol=1e-4, atol=1e-6)
    assert np.allclose(res[0][5], sa["encoder.in_conv.conv_conv.1.running_mean"].numpy(), rtol=1e-4, atol=1e-6)
