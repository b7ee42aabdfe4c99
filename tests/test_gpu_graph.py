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
