"""Shared test helpers (golden loading, fingerprint comparison)."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_l2(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def check_summary(t, summ, rtol, atol=0.0, what=""):
    """Compare tensor ``t`` against a summarize() fingerprint (tests/golden/common.py)."""
    t = t.detach().double().flatten().cpu()
    assert t.numel() == summ["numel"], what
    if "full" in summ:
        ref = summ["full"].double()
        got = t
    else:
        ref = summ["sample"].double()
        got = t[::summ["stride"]]
    err = (got - ref).norm().item()
    den = ref.norm().item()
    assert err <= rtol * den + atol, "%s: rel-l2 %.3e (abs %.3e) > %.1e" % (what, err / max(den, 1e-30), err, rtol)
    scale = max(summ["abs_sum"], 1e-30)
    assert abs(t.sum().item() - summ["sum"]) <= 10 * rtol * scale + atol * t.numel(), what + " (sum)"
    assert abs(t.abs().sum().item() - summ["abs_sum"]) <= 10 * rtol * scale + atol * t.numel(), what + " (abs_sum)"
