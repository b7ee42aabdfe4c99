"""CPU, build container only: the oracle against the REAL reference modules loaded by file path, including
torch-RNG dropout (same seed => identical masks) and a full-size Mean-Teacher step that reproduces the
survey's sanity values."""
import copy

import pytest
import torch

import oracle
from oracle.ref_loader import find_reference, load_reference

pytestmark = pytest.mark.skipif(find_reference() is None, reason="reference tree not present on this machine")


def test_names_shapes_and_rng_dropout_forward():
    ref = load_reference()
    torch.manual_seed(1337)
    m = ref.UNet(3, 2)
    assert [(n, tuple(p.shape)) for n, p in m.named_parameters()] == oracle.unet_param_spec(3, 2)
    st = {k: v.clone() for k, v in m.state_dict().items()}
    assert set(st) == {n for n, _ in oracle.unet_param_spec(3, 2)} | {n for n, _, _ in oracle.unet_buffer_spec(3, 2)}
    x = torch.rand(2, 3, 48, 32)
    m.train()
    torch.manual_seed(5)
    a = m(x)
    torch.manual_seed(5)
    b = oracle.unet_forward(st, x, True)
    assert torch.equal(a, b)
    for k, v in m.state_dict().items():
        assert torch.equal(v, st[k]), k


def test_full_size_mt_step_matches_reference_objects():
    """One literal MT step at the benchmark shape (12+12, 1x224x224) with torch-RNG dropout."""
    ref = load_reference()
    torch.manual_seed(1337)
    m = ref.UNet(1, 4)
    ema = copy.deepcopy(m)
    x_l, x_u = torch.rand(12, 1, 224, 224), torch.rand(12, 1, 224, 224)
    y = torch.randint(0, 4, (12, 224, 224))
    student = {k: v.clone() for k, v in m.state_dict().items()}
    teacher = {k: v.clone() for k, v in ema.state_dict().items()}
    opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    sch = ref.Medical_LR(opt, 0.01, 30000)
    med = ref.Med_Sup_Loss(4)
    m.train(), ema.train()
    rng = torch.get_rng_state()
    x = torch.cat([x_l, x_u])
    out = m(x)
    with torch.no_grad():
        eout = ema(x)
    sup = med(out[:12], y)
    cons = torch.mean((torch.softmax(out, 1)[12:] - torch.softmax(eout, 1)[12:]) ** 2)
    w = 0.1 * ref.sigmoid_rampup(1 // 150, 200.0)
    loss = sup + w * cons
    opt.zero_grad()
    loss.backward()
    opt.step()
    sch.step()
    ref.update_ema_variables(m, ema, 0.99, 1)
    # survey sanity values (Appendix B): loss 1.01571846, cons 2.27281800e-3
    assert loss.item() == pytest.approx(1.01571846, abs=5e-5)
    torch.set_rng_state(rng)
    r = oracle.mt_step(student, teacher, oracle.SGDState(), x_l, x_u, y, 1)
    assert r["loss"] == pytest.approx(loss.item(), abs=1e-6)
    assert r["loss_cons"] == pytest.approx(cons.item(), rel=1e-5)
    for k, v in m.state_dict().items():
        assert torch.allclose(student[k].float(), v.float(), atol=1e-6), k
    for k, v in ema.state_dict().items():
        assert torch.allclose(teacher[k].float(), v.float(), atol=1e-6), k


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and list(a.keys()) == list(b.keys()), path
        for k in a:
            _same(a[k], b[k], path + "/" + str(k))
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, path + "[%d]" % i)
    elif torch.is_tensor(a):
        assert torch.is_tensor(b) and a.dtype == b.dtype and a.shape == b.shape, path
        assert torch.allclose(a.double(), b.double(), rtol=1e-5, atol=1e-7), path
    elif isinstance(a, float):
        assert a == pytest.approx(b, rel=1e-5, abs=1e-8), path
    else:
        assert a == b, path


def test_committed_8f_fixtures_are_what_the_reference_produces(tmp_path, monkeypatch):
    """Re-run tests/golden/make_golden.py's SURVEY-8f generators against the REAL reference modules (UNet_Plus, Dense_Loss,
    DiceLoss, Medical_LR, update_ema_variables ...) and compare with the committed fixtures the GPU tests use."""
    import importlib.util
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    monkeypatch.setattr(sys, "argv", ["make_golden.py"])
    spec = importlib.util.spec_from_file_location("_make_golden_check", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    monkeypatch.setattr(mg, "HERE", str(tmp_path))
    mg.golden_f4_losses()
    mg.golden_ict_steps("acdc", 1, 4, 2, 4, 32, 32, 606)
    mg.golden_predict(1, 4, 5, 32, 48, 707)
    mg.golden_unet_plus(1, 4, 3, 64, 64, 808)
    mg.golden_hpfg_step(1, 4, 2, 4, 64, 64, 909)
    for name in ("f4_losses.pt", "ict_steps_acdc.pt", "predict_acdc.pt", "unet_plus_acdc.pt", "hpfg_step_acdc.pt"):
        new = torch.load(os.path.join(str(tmp_path), name), weights_only=False)
        old = torch.load(os.path.join(here, "golden", name), weights_only=False)
        _same(old, new, name)


def test_oracle_necks_and_dense_loss_match_reference_objects():
    """The oracle's projection_conv / dense_loss restatements (what the GPU tests of csrc/neck.cu compare against) give the
    same numbers as the reference modules (model/unet.py:120-152, utils/loss/dense_loss.py) on CPU, values and gradients."""
    import hpfg_b200 as hb
    ref = load_reference()
    from oracle.ref_loader import _load
    import os
    unet = _load("_hpfg_ref_unet_necks", os.path.join(ref.root, "model", "unet.py"))
    torch.manual_seed(3)
    for in_dim, hid, shape in ((256, 2048, (2, 256, 14, 14)), (4, 1024, (2, 4, 64, 64))):
        r = unet.projection_conv(in_dim, hid_dim=hid)
        m = hb.projection_conv(in_dim, hid_dim=hid)
        assert [k for k, _ in m.named_parameters()] == [k for k, _ in r.named_parameters()]
        assert [tuple(p.shape) for p in m._param_list()] == [tuple(p.shape) for p in r.parameters()]   # the kernel's params[8] order
        m.load_state_dict(r.state_dict())
        st = {"neck." + k: v for k, v in r.state_dict().items()}
        f = torch.randn(shape)
        (a1, a2), (b1, b2) = r(f), oracle.projection_conv(st, "neck", f)
        assert torch.equal(a1, b1) and torch.equal(a2, b2)
    for bs in (3, 8):
        x = (torch.randn(bs, 128, requires_grad=True), torch.randn(bs, 128, 16, requires_grad=True))
        y = (torch.randn(bs, 128), torch.randn(bs, 128, 16))
        la = ref.Dense_Loss(batch_size=bs, device=torch.device("cpu"))(x, y)
        lo = oracle.dense_loss(x, y)
        assert la.item() == pytest.approx(lo.item(), rel=1e-6)
        for ga, gb in zip(torch.autograd.grad(la, x), torch.autograd.grad(lo, x)):
            assert torch.allclose(ga, gb, rtol=1e-5, atol=1e-8)
