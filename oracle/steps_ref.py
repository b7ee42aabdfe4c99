"""fp32 restatement of the trainer step bodies and the host-side schedules around the hot path
(utils/utils.py:67-86, utils/scheduler/medical_lr.py:7-17, utils/__init__.py:14-16,
2017_03_NIPS_Mean-Teacher_ACDC.py:89-113, 2021_06_CVPR_CPS_ACDC.py:90-120,
2019_07_MICCAI_Uncertainty_Aware_ACDC.py:120-170, 2022_02_ISBI_ICT-MedSeg_ACDC.py:96-140)."""
import math

import numpy as np
import torch

from .unet_ref import unet_forward, unet_param_spec
from .losses_ref import (med_sup_loss, mt_consistency, cps_losses, uamt_consistency, dice_loss, ce_loss,
                         ict_losses)


def sigmoid_rampup(current, rampup_length):
    """utils/utils.py:72-79."""
    if rampup_length == 0:
        return 1.0
    current = np.clip(current, 0.0, rampup_length)
    phase = 1.0 - current / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def consistency_weight(cur_itrs, consistency=0.1, consistency_rampup=200.0):
    """get_current_consistency_weight(epoch=cur_itrs // 150) (utils/utils.py:67-69; 2017_03...:105)."""
    return consistency * sigmoid_rampup(cur_itrs // 150, consistency_rampup)


def ema_alpha(global_step, ema_decay=0.99):
    """utils/utils.py:84."""
    return min(1 - 1 / (global_step + 1), ema_decay)


def update_ema(student, teacher, alpha, global_step, names):
    """update_ema_variables (utils/utils.py:82-86): parameters only, BN buffers untouched."""
    a = ema_alpha(global_step, alpha)
    for n in names:
        teacher[n].mul_(a).add_(student[n], alpha=1 - a)


def medical_lr(step_index, base_lr=0.01, max_iterations=30000):
    """Medical_LR (utils/scheduler/medical_lr.py:7-17).  ``step_index`` = number of scheduler.step() calls
    made so far (0 for the first training iteration).  _LRScheduler.__init__ performs one initial step, so
    last_epoch = step_index at the time of iteration ``step_index+1`` and iter_num = last_epoch - 1."""
    iter_num = step_index - 1
    return base_lr * (1.0 - iter_num / max_iterations) ** 0.9


class SGDState:
    """torch.optim.SGD(lr, momentum=0.9, weight_decay=1e-4) state (utils/__init__.py:14-16)."""

    def __init__(self, momentum=0.9, weight_decay=1e-4):
        self.momentum, self.weight_decay, self.buf = momentum, weight_decay, {}


def sgd_step(params, grads, opt, lr):
    """One torch.optim.SGD step: g += wd*p; buf = g (first) or mom*buf + g; p -= lr*buf."""
    for n, g in grads.items():
        if g is None:
            continue
        p = params[n]
        d = g.add(p, alpha=opt.weight_decay) if opt.weight_decay != 0 else g.clone()
        if opt.momentum != 0:
            if n not in opt.buf:
                opt.buf[n] = d.clone()
            else:
                opt.buf[n].mul_(opt.momentum).add_(d)
            d = opt.buf[n]
        p.add_(d, alpha=-lr)


def _names(st):
    cin = st["encoder.in_conv.conv_conv.0.weight"].shape[1]
    ncls = st["decoder.out_conv.weight"].shape[0]
    return [n for n, _ in unet_param_spec(cin, ncls)], ncls


def _grad_forward(st, names, x, dropout_masks):
    leaves = {n: st[n].detach().requires_grad_(True) for n in names}
    view = dict(st)
    view.update(leaves)
    out = unet_forward(view, x, True, dropout_masks)
    for k in st:                                        # carry BN buffer updates back (counters are rebinding-free)
        if k not in leaves:
            st[k] = view[k]
    return out, leaves


def mt_step(student, teacher, opt, x_l, x_u, y, cur_itrs, *, base_lr=0.01, max_iterations=30000,
            ema_decay=0.99, consistency=0.1, consistency_rampup=200.0,
            student_masks=None, teacher_masks=None):
    """One Mean-Teacher iteration (2017_03_NIPS_Mean-Teacher_ACDC.py:89-113).  cur_itrs starts at 1.
    Returns dict(loss, loss_sup, loss_cons, w, lr, logits, teacher_logits, grads)."""
    names, ncls = _names(student)
    lb = x_l.shape[0]
    x = torch.cat([x_l, x_u], dim=0)
    out, leaves = _grad_forward(student, names, x, student_masks)
    with torch.no_grad():
        t_out = unet_forward(teacher, x, True, teacher_masks)         # teacher in train() mode (:70,100)
    loss_sup = med_sup_loss(out[:lb], y, ncls)
    loss_cons = mt_consistency(out[lb:], t_out[lb:])
    w = consistency_weight(cur_itrs, consistency, consistency_rampup)
    loss = loss_sup + w * loss_cons
    grads = dict(zip(names, torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)))
    lr = medical_lr(cur_itrs - 1, base_lr, max_iterations)
    with torch.no_grad():
        sgd_step(student, grads, opt, lr)
        update_ema(student, teacher, ema_decay, cur_itrs, names)
    return dict(loss=float(loss.detach()), loss_sup=float(loss_sup.detach()), loss_cons=float(loss_cons.detach()), w=w, lr=lr,
                logits=out.detach(), teacher_logits=t_out, grads=grads)


def cps_step(m1, m2, opt1, opt2, x_l, x_u, y, cur_itrs, *, base_lr=0.01, max_iterations=30000,
             consistency=0.1, consistency_rampup=200.0, masks1=None, masks2=None):
    """One CPS iteration (2021_06_CVPR_CPS_ACDC.py:90-120)."""
    names, ncls = _names(m1)
    lb = x_l.shape[0]
    x = torch.cat([x_l, x_u], dim=0)
    out1, l1 = _grad_forward(m1, names, x, masks1)
    out2, l2 = _grad_forward(m2, names, x, masks2)
    loss_sup, loss_semi, pl1, pl2 = cps_losses(out1, out2, y, lb, ncls)
    w = consistency_weight(cur_itrs, consistency, consistency_rampup)
    loss = loss_sup + w * loss_semi
    gs = torch.autograd.grad(loss, [l1[n] for n in names] + [l2[n] for n in names], allow_unused=True)
    g1 = dict(zip(names, gs[:len(names)]))
    g2 = dict(zip(names, gs[len(names):]))
    lr = medical_lr(cur_itrs - 1, base_lr, max_iterations)
    with torch.no_grad():
        sgd_step(m1, g1, opt1, lr)
        sgd_step(m2, g2, opt2, lr)
    return dict(loss=float(loss.detach()), loss_sup=float(loss_sup.detach()), loss_semi=float(loss_semi.detach()), w=w, lr=lr,
                logits1=out1.detach(), logits2=out2.detach(), pseudo1=pl1, pseudo2=pl2, grads1=g1, grads2=g2)


def uamt_step(student, teacher, opt, x_l, x_u, y, cur_itrs, noise, mc_noise, *, T=8, base_lr=0.01,
              max_iterations=30000, ema_decay=0.99, consistency=0.1, consistency_rampup=200.0,
              student_masks=None, teacher_masks=None):
    """One UAMT iteration (2019_07_MICCAI_Uncertainty_Aware_ACDC.py:120-170).

    noise [n_u,...] and mc_noise [T//2, 2*n_u, ...] are the already-clamped perturbations
    clamp(randn*0.1, -0.2, 0.2) (:130,141) supplied by the caller so both sides see identical bits.
    teacher_masks: list of T//2+1 dicts (first = the single noisy pass, then the MC passes) or None."""
    names, ncls = _names(student)
    lb, n_u = x_l.shape[0], x_u.shape[0]
    x = torch.cat([x_l, x_u], dim=0)
    out, leaves = _grad_forward(student, names, x, student_masks)
    tm = teacher_masks if teacher_masks is not None else [None] * (T // 2 + 1)
    with torch.no_grad():
        t_out = unet_forward(teacher, x_u + noise, True, tm[0])
        xr = x_u.repeat(2, 1, 1, 1)
        stride = xr.shape[0] // 2
        preds = torch.zeros([stride * T, ncls, x.shape[2], x.shape[3]])
        for i in range(T // 2):
            preds[2 * stride * i:2 * stride * (i + 1)] = unet_forward(teacher, xr + mc_noise[i], True, tm[i + 1])
    loss_ce = ce_loss(out[:lb], y)
    loss_dice = dice_loss(torch.softmax(out, dim=1)[:lb], y.unsqueeze(1), ncls)
    sup = 0.5 * (loss_dice + loss_ce)
    w = consistency_weight(cur_itrs, consistency, consistency_rampup)
    threshold = (0.75 + 0.25 * sigmoid_rampup(cur_itrs, max_iterations)) * math.log(2)
    cons, unc, mask = uamt_consistency(out[lb:], t_out, preds, T, threshold)
    loss = sup + w * cons
    grads = dict(zip(names, torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)))
    lr = medical_lr(cur_itrs - 1, base_lr, max_iterations)
    with torch.no_grad():
        sgd_step(student, grads, opt, lr)
        update_ema(student, teacher, ema_decay, cur_itrs, names)
    return dict(loss=float(loss.detach()), loss_sup=float(sup.detach()), loss_cons=float(cons.detach()), w=w, lr=lr, threshold=threshold,
                logits=out.detach(), teacher_logits=t_out, mc_logits=preds, uncertainty=unc, mask=mask,
                grads=grads)


def ict_step(student, teacher, opt, x_l, x_u, y, cur_itrs, mix_factors, *, base_lr=0.01, max_iterations=30000,
             ema_decay=0.99, consistency=0.1, consistency_rampup=200.0, student_masks=None, teacher_masks=None):
    """One ICT-MedSeg iteration (2022_02_ISBI_ICT-MedSeg_ACDC.py:96-140).  mix_factors [n_u//2] are the
    np.random.beta(ict_alpha, ict_alpha) draws (:112), supplied by the caller so both sides see identical bits.
    teacher_masks: list of two dicts (one per teacher forward) or None.  The teacher is in train() mode (a freshly
    built module that is never switched to eval before the loop, :57,76)."""
    names, ncls = _names(student)
    lb, n_u = x_l.shape[0], x_u.shape[0]
    n_m = n_u // 2
    lam = mix_factors.reshape(n_m, 1, 1, 1).float()
    ux0, ux1 = x_u[:n_m], x_u[n_m:]
    mixed = ux0 * (1.0 - lam) + ux1 * lam                                               # :117
    x = torch.cat([x_l, mixed], dim=0)
    out, leaves = _grad_forward(student, names, x, student_masks)
    tm = teacher_masks if teacher_masks is not None else [None, None]
    with torch.no_grad():
        t0 = unet_forward(teacher, ux0, True, tm[0])                                    # :123
        t1 = unet_forward(teacher, ux1, True, tm[1])                                    # :124
    t_out = torch.cat([t0, t1], dim=0)
    sup, cons = ict_losses(out, t_out, lam, y, lb, ncls)
    w = consistency_weight(cur_itrs, consistency, consistency_rampup)
    loss = sup + w * cons
    grads = dict(zip(names, torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)))
    lr = medical_lr(cur_itrs - 1, base_lr, max_iterations)
    with torch.no_grad():
        sgd_step(student, grads, opt, lr)
        update_ema(student, teacher, ema_decay, cur_itrs, names)
    return dict(loss=float(loss.detach()), loss_sup=float(sup.detach()), loss_cons=float(cons.detach()), w=w, lr=lr,
                logits=out.detach(), teacher_logits=t_out, grads=grads, mixed=mixed)


def linear_rampup(current, rampup_length):
    """utils/utils.py:89-95."""
    if current >= rampup_length:
        return 1.0
    return current / rampup_length


def s4cv_step(m1, m2, teacher, opt1, opt2, x_l, x_u, y, cur_itrs, noise, *, base_lr=0.01, max_iterations=30000, ema_decay=0.99,
              consistency=0.1, consistency_rampup=200.0, mt_start=1000, masks1=None, masks2=None, teacher_masks=None):
    """One S4CVNet iteration (2022_08_CVPR_S4CVNet_ACDC.py:108-167): two students on cat([labeled, unlabeled]), the EMA
    teacher of student 2 on the noise-perturbed unlabeled slices (forward every iteration, :117-133; its output enters the
    loss from iteration `mt_start` on, :143-149), loss of :124-156, two SGD steps, EMA of student 2 into the teacher (:166)."""
    names, ncls = _names(m1)
    lb = x_l.shape[0]
    x = torch.cat([x_l, x_u], dim=0)
    out1, l1 = _grad_forward(m1, names, x, masks1)
    out2, l2 = _grad_forward(m2, names, x, masks2)
    with torch.no_grad():
        t_out = unet_forward(teacher, x_u + noise, True, teacher_masks)
    w = consistency * linear_rampup(cur_itrs // 150, consistency_rampup)
    from .losses_ref import s4cv_losses
    loss, loss_sup, loss_semi, pl1, pl2 = s4cv_losses(out1, out2, t_out if cur_itrs >= mt_start else None, y, lb, ncls, 7 * w, w)
    gs = torch.autograd.grad(loss, [l1[n] for n in names] + [l2[n] for n in names], allow_unused=True)
    g1, g2 = dict(zip(names, gs[:len(names)])), dict(zip(names, gs[len(names):]))
    lr = medical_lr(cur_itrs - 1, base_lr, max_iterations)
    with torch.no_grad():
        sgd_step(m1, g1, opt1, lr)
        sgd_step(m2, g2, opt2, lr)
        update_ema(m2, teacher, ema_decay, cur_itrs, names)
    return dict(loss=float(loss.detach()), loss_sup=float(loss_sup.detach()), loss_semi=float(loss_semi.detach()), w=w, lr=lr,
                logits1=out1.detach(), logits2=out2.detach(), teacher_logits=t_out, grads1=g1, grads2=g2)
