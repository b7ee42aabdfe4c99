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
