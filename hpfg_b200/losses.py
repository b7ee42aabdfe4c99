"""Loss classes with the reference's signatures (utils/loss/medloss.py:44-56 ``Med_Sup_Loss``,
utils/loss/diceloss.py:155-191 ``DiceLoss``, :64-81 ``softmax_mse_loss``) plus the fused whole-step SSL losses
of the Mean-Teacher / CPS / UAMT trainers, all running on the fused CUDA loss kernels (csrc/loss.cu).

Each fused call produces the loss value AND d loss / d logits in the same two launches; the autograd
Function just hands the stored gradient back scaled by grad_output."""
import ctypes

import torch
import torch.nn as nn

from . import _lib as L


def _weights_arg(weight, n):
    if weight is None:
        return None
    assert len(weight) == n
    return (ctypes.c_float * n)(*[float(w) for w in weight])


def _workspace(mode, n_l, n_u, c, h, w, device):
    nbytes = L.lib().hpfg_ssl_loss_workspace_bytes(mode, n_l, n_u, c, h, w)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def ssl_loss_raw(mode, student, other, labels, n_l, *, cons_weight=0.0, mc_logits=None, mc_passes=0,
                 uamt_threshold=0.0, class_weights=None, ce_coef=0.5, dice_coef=0.5, want_pseudo=False,
                 out=None, cons_weight_dev=None):
    """Thin wrapper of hpfg_ssl_loss.  Returns dict(scalars[8], dstudent, dother, pseudo1, pseudo2)."""
    L.require_cuda(student, "logits")
    student = student.contiguous().float()
    n, c, h, w = student.shape
    n_u = n - n_l
    dev = student.device
    if labels is not None:
        labels = labels.contiguous().to(torch.int64)
        assert labels.shape == (n_l, h, w), "labels must be [n_l,H,W]"
    if other is not None:
        other = other.contiguous().float()
    dstudent = torch.empty_like(student)
    dother = torch.empty_like(other) if mode == L.LOSS_CPS else None
    scalars = torch.empty(8, device=dev, dtype=torch.float32)
    ws = _workspace(mode, n_l, n_u, c, h, w, dev)
    p1 = p2 = None
    if want_pseudo and mode == L.LOSS_CPS:
        p1 = torch.empty((n_u, h, w), device=dev, dtype=torch.int64)
        p2 = torch.empty((n_u, h, w), device=dev, dtype=torch.int64)
    if cons_weight_dev is not None:      # CUDA-graph replays: the consistency weight is a device scalar
        L.check(L.lib().hpfg_ssl_loss_dv(mode, L.ptr(student), L.ptr(other), L.ptr(mc_logits), mc_passes, L.ptr(labels),
                                         n_l, n_u, c, h, w, L.ptr(cons_weight_dev), float(uamt_threshold),
                                         _weights_arg(class_weights, c), float(ce_coef), float(dice_coef),
                                         L.ptr(dstudent), L.ptr(dother), L.ptr(scalars), L.ptr(p1), L.ptr(p2), L.ptr(ws),
                                         L.stream_ptr(dev)), "hpfg_ssl_loss_dv")
        return dict(scalars=scalars, dstudent=dstudent, dother=dother, pseudo1=p1, pseudo2=p2)
    L.check(L.lib().hpfg_ssl_loss(mode, L.ptr(student), L.ptr(other), L.ptr(mc_logits), mc_passes, L.ptr(labels), n_l,
                                  n_u, c, h, w, float(cons_weight), float(uamt_threshold),
                                  _weights_arg(class_weights, c), float(ce_coef), float(dice_coef), L.ptr(dstudent),
                                  L.ptr(dother), L.ptr(scalars), L.ptr(p1), L.ptr(p2), L.ptr(ws), L.stream_ptr(dev)),
            "hpfg_ssl_loss")
    return dict(scalars=scalars, dstudent=dstudent, dother=dother, pseudo1=p1, pseudo2=p2)


class _SslLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, other, mode, labels, n_l, kw):
        r = ssl_loss_raw(mode, student, other.detach() if other is not None else None, labels, n_l, **kw)
        ctx.save_for_backward(r["dstudent"], r["dother"] if r["dother"] is not None else r["dstudent"].new_empty(0))
        ctx.has_other = r["dother"] is not None
        ctx.extras = r
        return r["scalars"][0].clone()

    @staticmethod
    def backward(ctx, g):
        ds, do = ctx.saved_tensors
        return ds * g, (do * g if ctx.has_other else None), None, None, None, None


class Med_Sup_Loss(nn.Module):
    """0.5*CE(ignore_index=255) + 0.5*Dice(softmax(logits), onehot(y))  (utils/loss/medloss.py:44-56)."""

    def __init__(self, num_classes, ce=0.5, dice=0.5):
        super().__init__()
        self.num_classes, self.ce, self.dice = num_classes, ce, dice

    def forward(self, outputs, target_label):
        assert outputs.shape[1] == self.num_classes
        return _SslLossFn.apply(outputs, None, L.LOSS_SUP, target_label, outputs.shape[0],
                                dict(ce_coef=self.ce, dice_coef=self.dice))


class _DiceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, target, n_classes, weight, softmax):
        L.require_cuda(inputs, "DiceLoss inputs")
        x = inputs.contiguous().float()
        n, c, h, w = x.shape
        t = target.reshape(n, h, w).contiguous().to(torch.int64)
        dx = torch.empty_like(x) if inputs.requires_grad else None
        scalars = torch.empty(1 + c, device=x.device, dtype=torch.float32)
        ws = torch.empty(2048, dtype=torch.uint8, device=x.device)
        L.check(L.lib().hpfg_dice_loss(L.ptr(x), L.ptr(t), n, c, h, w, int(bool(softmax)), _weights_arg(weight, c),
                                       L.ptr(dx), L.ptr(scalars), L.ptr(ws), L.stream_ptr(x.device)), "hpfg_dice_loss")
        if dx is not None:
            ctx.save_for_backward(dx)
        ctx.class_wise = scalars[1:]
        return scalars[0].clone()

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None, None, None, None


class DiceLoss(nn.Module):
    """``DiceLoss(n_classes)(inputs, target, weight=None, softmax=False)`` (utils/loss/diceloss.py:155-191).
    inputs [n,C,H,W] probabilities (or logits with softmax=True); target [n,1,H,W] class ids."""

    def __init__(self, n_classes):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, inputs, target, weight=None, softmax=False):
        onehot_size = (target.shape[0], self.n_classes * target.shape[1]) + tuple(target.shape[2:])
        assert tuple(inputs.size()) == onehot_size, 'predict & target shape do not match'
        return _DiceFn.apply(inputs, target, self.n_classes, weight, softmax)


def softmax_mse_loss(input_logits, target_logits, sigmoid=False):
    """Unreduced (softmax(a)-softmax(b))**2 (utils/loss/diceloss.py:64-81).  Element-wise map kept in torch:
    the fused kernels consume the reduced forms (mean / uncertainty-masked mean) directly."""
    assert input_logits.size() == target_logits.size()
    if sigmoid:
        return (torch.sigmoid(input_logits) - torch.sigmoid(target_logits)) ** 2
    return (torch.softmax(input_logits, dim=1) - torch.softmax(target_logits, dim=1)) ** 2


# ---- fused whole-step losses (what the MT / CPS / UAMT trainers compute inline) ---------------------------
def mean_teacher_loss(student_logits, teacher_logits_u, labels, consistency_weight, ce=0.5, dice=0.5):
    """loss_sup + w*mean((softmax(s_u)-softmax(t_u))^2)  (2017_03_NIPS_Mean-Teacher_ACDC.py:97-106).
    student_logits [n_l+n_u,...]; teacher_logits_u [n_u,...] (the unlabeled slices only)."""
    n_l = labels.shape[0]
    return _SslLossFn.apply(student_logits, teacher_logits_u, L.LOSS_MT, labels, n_l,
                            dict(cons_weight=consistency_weight, ce_coef=ce, dice_coef=dice))


def cps_loss(logits1, logits2, labels, consistency_weight, ce=0.5, dice=0.5):
    """Cross pseudo supervision (2021_06_CVPR_CPS_ACDC.py:99-111); gradients flow to both networks."""
    n_l = labels.shape[0]
    return _SslLossFn.apply(logits1, logits2, L.LOSS_CPS, labels, n_l,
                            dict(cons_weight=consistency_weight, ce_coef=ce, dice_coef=dice))


def uamt_loss(student_logits, teacher_logits_u, mc_logits, labels, consistency_weight, threshold, T=8, ce=0.5,
              dice=0.5):
    """Uncertainty-aware MT (2019_07_MICCAI_Uncertainty_Aware_ACDC.py:145-162)."""
    n_l = labels.shape[0]
    return _SslLossFn.apply(student_logits, teacher_logits_u, L.LOSS_UAMT, labels, n_l,
                            dict(cons_weight=consistency_weight, mc_logits=mc_logits.contiguous().float(), mc_passes=T,
                                 uamt_threshold=threshold, ce_coef=ce, dice_coef=dice))
