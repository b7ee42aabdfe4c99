"""CUDA-graph replay of the fused Mean-Teacher step must reproduce the eager step exactly (same kernels, same Philox
offsets, per-iteration scalars read from the device block)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(graph, steps=5):
    import hpfg_b200 as hb
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    student = hb.UNet(1, 4, precision="bf16").to(dev)
    teacher = copy.deepcopy(student)
    step = hb.MeanTeacherStep(student, teacher, total_itrs=100)
    step.enable_graph(graph)
    g = torch.Generator().manual_seed(5)
    losses = []
    for _ in range(steps):
        x = torch.rand(6, 1, 64, 64, generator=g).to(dev)
        y = torch.randint(0, 4, (2, 64, 64), generator=g).to(dev)
        losses.append(step.step(x, y).item())
    torch.cuda.synchronize()
    return losses, {k: v.detach().clone() for k, v in student.state_dict().items()}, {k: v.detach().clone() for k, v in teacher.state_dict().items()}


def test_graph_replay_matches_eager():
    l0, s0, t0 = _run(False)
    l1, s1, t1 = _run(True)
    assert l0 == l1, (l0, l1)
    for k in s0:
        assert torch.equal(s0[k], s1[k]), "student %s differs" % k
        assert torch.equal(t0[k], t1[k]), "teacher %s differs" % k


def _run_other(kind, graph, steps=4):
    """CPS / UAMT / ICT drivers: same seeds, same supplied noise / mix factors, eager vs graph replay."""
    import hpfg_b200 as hb
    dev = torch.device("cuda:0")
    torch.manual_seed(13)
    a = hb.UNet(1, 4, precision="bf16").to(dev)
    b = hb.UNet(1, 4, precision="bf16").to(dev) if kind != "uamt" else copy.deepcopy(a)
    if kind == "cps":
        step = hb.CPSStep(a, b, total_itrs=100)
    elif kind == "uamt":
        step = hb.UAMTStep(a, b, total_itrs=100, T=4)
    else:
        step = hb.ICTStep(a, b, total_itrs=100)
    step.enable_graph(graph)
    g = torch.Generator().manual_seed(6)
    losses = []
    for _ in range(steps):
        x = torch.rand(6, 1, 64, 64, generator=g).to(dev)
        y = torch.randint(0, 4, (2, 64, 64), generator=g).to(dev)
        if kind == "cps":
            loss = step.step(x, y)
        elif kind == "uamt":
            noise = torch.clamp(torch.randn(4, 1, 64, 64, generator=g) * 0.1, -0.2, 0.2).to(dev)
            mc_noise = torch.clamp(torch.randn(2, 8, 1, 64, 64, generator=g) * 0.1, -0.2, 0.2).to(dev)
            loss = step.step(x, y, noise, mc_noise)
        else:
            loss = step.step(x, y, torch.rand(2, generator=g))
        losses.append(loss.item())
    torch.cuda.synchronize()
    assert (step.kernels_per_replay > 0) == bool(graph)
    return losses, {k: v.detach().clone() for k, v in a.state_dict().items()}, {k: v.detach().clone() for k, v in b.state_dict().items()}


@pytest.mark.parametrize("kind", ["cps", "uamt", "ict"])
def test_graph_replay_matches_eager_other_drivers(kind):
    l0, a0, b0 = _run_other(kind, False)
    l1, a1, b1 = _run_other(kind, True)
    assert l0 == l1, (l0, l1)
    for k in a0:
        assert torch.equal(a0[k], a1[k]), "network 1 %s differs" % k
        assert torch.equal(b0[k], b1[k]), "network 2 / teacher %s differs" % k


def _run_async(graph, steps=24):
    """No host synchronisation between steps: the host runs ahead of the device, so the per-iteration scalar block
    (lr, EMA alpha, consistency weight, Philox offsets) must travel through the pinned-slot ring, never through one
    pinned buffer that later steps overwrite before earlier uploads executed."""
    import hpfg_b200 as hb
    dev = torch.device("cuda:0")
    torch.manual_seed(17)
    student = hb.UNet(1, 4, precision="bf16").to(dev)
    teacher = copy.deepcopy(student)
    step = hb.MeanTeacherStep(student, teacher, total_itrs=40, consistency_rampup=0.2)   # lr / alpha / w all move per step
    step.enable_graph(graph)
    g = torch.Generator().manual_seed(7)
    xs = [torch.rand(6, 1, 64, 64, generator=g).to(dev) for _ in range(3)]
    ys = [torch.randint(0, 4, (2, 64, 64), generator=g).to(dev) for _ in range(3)]
    spin = torch.empty(64 << 20, device=dev)
    torch.cuda.synchronize()
    losses = torch.zeros(steps, device=dev)
    for _ in range(6):
        spin.add_(1.0)                              # queue device work first so the enqueue loop below really runs ahead
    for i in range(steps):
        losses[i:i + 1].copy_(step.step(xs[i % 3], ys[i % 3]).reshape(1))
    torch.cuda.synchronize()
    return losses.cpu(), student.flat_params.detach().clone().cpu(), teacher.flat_params.detach().clone().cpu()


def test_graph_replay_matches_eager_without_per_step_sync():
    l0, s0, t0 = _run_async(False)
    l1, s1, t1 = _run_async(True)
    assert torch.equal(l0, l1), (l0, l1)
    assert torch.equal(s0, s1) and torch.equal(t0, t1)


def test_step_driver_state_dict_roundtrip_matches_torch_sgd_layout():
    """state_dict()/load_state_dict() of a step driver (reference checkpoints hold {model, optimizer, cur_itrs},
    2017_03...:126-133): a resumed driver continues bit-identically, and the optimizer entry has torch.optim.SGD's layout."""
    import hpfg_b200 as hb
    dev = torch.device("cuda:0")

    def fresh():
        torch.manual_seed(19)
        s = hb.UNet(1, 4, precision="bf16").to(dev)
        return s, copy.deepcopy(s)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(6, 1, 64, 64, generator=g).to(dev)
    y = torch.randint(0, 4, (2, 64, 64), generator=g).to(dev)
    s, t = fresh()
    step = hb.MeanTeacherStep(s, t)
    for _ in range(2):
        step.step(x, y)
    ck = {"model": copy.deepcopy(s.state_dict()), "ema": copy.deepcopy(t.state_dict()), "step": step.state_dict()}
    ref = [step.step(x, y).item() for _ in range(2)]
    opt = torch.optim.SGD(s.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    opt.load_state_dict({"state": ck["step"]["optimizers"][0]["state"],
                         "param_groups": [dict(opt.state_dict()["param_groups"][0], lr=ck["step"]["optimizers"][0]["param_groups"][0]["lr"])]})
    s2, t2 = fresh()
    s2.load_state_dict(ck["model"])
    t2.load_state_dict(ck["ema"])
    step2 = hb.MeanTeacherStep(s2, t2)
    step2.load_state_dict(ck["step"])
    got = [step2.step(x, y).item() for _ in range(2)]
    assert got == ref, (got, ref)
