"""fp32 restatement of the SSL losses on the hot path (utils/loss/medloss.py:5-56,
utils/loss/diceloss.py:64-81,155-191 and the inline trainer expressions cited per function).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import torch
import torch.nn.functional as F


def dice_loss(probs, target, n_classes, weight=None, softmax=False):
    """DiceLoss.forward (utils/loss/diceloss.py:178-191 == utils/loss/medloss.py:28-41).

    probs [n,C,H,W]; target [n,1,H,W] (class ids).  Sums run over the whole batch; smooth = 1e-5;
    denominators use squares (diceloss.py:168-176).  ignore_index is NOT honoured (255 matches no class).
    """
    if softmax:
        probs = torch.softmax(probs, dim=1)
    onehot = torch.cat([(target == i) for i in range(n_classes)], dim=1).float()   # _one_hot_encoder :160-166
    if weight is None:
        weight = [1] * n_classes
    assert probs.size() == onehot.size(), 'predict & target shape do not match'
    smooth = 1e-5
    loss = 0.0
    for i in range(n_classes):
        s, t = probs[:, i], onehot[:, i]
        inter = torch.sum(s * t)
        y_sum = torch.sum(t * t)
        z_sum = torch.sum(s * s)
        loss = loss + (1 - (2 * inter + smooth) / (z_sum + y_sum + smooth)) * weight[i]
    return loss / n_classes


def ce_loss(logits, target, ignore_index=255):
    """nn.CrossEntropyLoss(ignore_index=255): mean over non-ignored pixels (medloss.py:49)."""
    return F.cross_entropy(logits, target, ignore_index=ignore_index)


def med_sup_loss(logits, target, n_classes, ce=0.5, dice=0.5):
    """Med_Sup_Loss.forward (utils/loss/medloss.py:54-56)."""
    return ce * ce_loss(logits, target) + dice * dice_loss(torch.softmax(logits, dim=1), target.unsqueeze(1),
                                                           n_classes)


def softmax_mse(input_logits, target_logits):
    """softmax_mse_loss (utils/loss/diceloss.py:64-81; local copy 2019_07...:63-79): elementwise, unreduced."""
    assert input_logits.size() == target_logits.size()
    return (F.softmax(input_logits, dim=1) - F.softmax(target_logits, dim=1)) ** 2


def mt_consistency(student_logits_u, teacher_logits_u):
    """Mean-Teacher consistency (2017_03_NIPS_Mean-Teacher_ACDC.py:97,101,104)."""
    return torch.mean((torch.softmax(student_logits_u, dim=1) - torch.softmax(teacher_logits_u, dim=1)) ** 2)


def cps_losses(out1, out2, target, label_bs, n_classes):
    """CPS supervised + cross pseudo supervision terms (2021_06_CVPR_CPS_ACDC.py:99-110).
    Returns (loss_sup, loss_semi, pseudo1, pseudo2)."""
    soft1 = torch.softmax(out1, dim=1)
    soft2 = torch.softmax(out2, dim=1)
    loss_sup = med_sup_loss(out1[:label_bs], target, n_classes) + med_sup_loss(out2[:label_bs], target, n_classes)
    pl1 = torch.argmax(soft1[label_bs:].detach(), dim=1, keepdim=False)
    pl2 = torch.argmax(soft2[label_bs:].detach(), dim=1, keepdim=False)
    loss_semi = med_sup_loss(out1[label_bs:], pl2, n_classes) + med_sup_loss(out2[label_bs:], pl1, n_classes)
    return loss_sup, loss_semi, pl1, pl2


def uamt_consistency(student_logits_u, teacher_logits_u, teacher_mc_logits, T, threshold):
    """UAMT uncertainty-masked consistency (2019_07_MICCAI_Uncertainty_Aware_ACDC.py:145-160).

    teacher_mc_logits: [T*n_u, C, H, W] logits of the T stochastic teacher passes, in the reference's
    buffer order (pass-major).  Returns (consistency_loss, uncertainty[n_u,1,H,W], mask)."""
    n_u, C, H, W = student_logits_u.shape
    preds = F.softmax(teacher_mc_logits, dim=1).reshape(T, n_u, C, H, W).mean(dim=0)
    uncertainty = -1.0 * torch.sum(preds * torch.log(preds + 1e-6), dim=1, keepdim=True)
    dist = softmax_mse(student_logits_u, teacher_logits_u)
    mask = (uncertainty < threshold).float()
    loss = torch.sum(mask * dist) / (2 * torch.sum(mask) + 1e-16)     # the literal 2 is the reference's
    return loss, uncertainty, mask


def ict_losses(student_logits, teacher_logits_u, mix_factors, target, label_bs, n_classes):
    """ICT-MedSeg supervised + interpolation-consistency terms (2022_02_ISBI_ICT-MedSeg_ACDC.py:121-135).

    student_logits [label_bs + n_m, C, H, W] = model(cat([labeled, mixed])); teacher_logits_u [2*n_m, ...] =
    cat([ema(ux0), ema(ux1)]); mix_factors [n_m,1,1,1].  Returns (supervised_loss, consistency_loss)."""
    n_m = student_logits.shape[0] - label_bs
    soft = torch.softmax(student_logits, dim=1)
    e0 = torch.softmax(teacher_logits_u[:n_m], dim=1)
    e1 = torch.softmax(teacher_logits_u[n_m:], dim=1)
    lam = mix_factors.reshape(n_m, 1, 1, 1).float()
    pred_mixed = e0 * (1.0 - lam) + e1 * lam                                           # :125
    loss_ce = ce_loss(student_logits[:label_bs], target)                               # :127
    loss_dice = dice_loss(soft[:label_bs], target.unsqueeze(1), n_classes)             # :128
    sup = 0.5 * (loss_dice + loss_ce)                                                  # :130
    cons = torch.mean((soft[label_bs:] - pred_mixed) ** 2)                             # :133
    return sup, cons


def s4cv_losses(out1, out2, teacher_logits_u, target, label_bs, n_classes, cps_weight, mt_weight):
    """S4CVNet loss (2022_08_CVPR_S4CVNet_ACDC.py:124-156).  cps_weight = 7*consistency_weight_cps (:148-149);
    teacher_logits_u None = the `cur_itrs < 1000` branch (consistency_loss1/2 = 0.0, :143-145).
    Returns (loss, loss_sup, loss_semi, pseudo1, pseudo2)."""
    soft1, soft2 = torch.softmax(out1, dim=1), torch.softmax(out2, dim=1)
    loss1 = 0.5 * (ce_loss(out1[:label_bs], target) + dice_loss(soft1[:label_bs], target.unsqueeze(1), n_classes))
    loss2 = 0.5 * (ce_loss(out2[:label_bs], target) + dice_loss(soft2[:label_bs], target.unsqueeze(1), n_classes))
    loss_sup = loss1 + loss2
    pl1 = torch.argmax(soft1[label_bs:].detach(), dim=1, keepdim=False)
    pl2 = torch.argmax(soft2[label_bs:].detach(), dim=1, keepdim=False)
    ps1 = dice_loss(soft1[label_bs:], pl2.unsqueeze(1), n_classes)
    ps2 = dice_loss(soft2[label_bs:], pl1.unsqueeze(1), n_classes)
    if teacher_logits_u is None:
        cl1 = cl2 = 0.0
    else:
        ema_soft = torch.softmax(teacher_logits_u, dim=1)
        cl1 = torch.mean((soft1[label_bs:] - ema_soft) ** 2)
        cl2 = torch.mean((soft2[label_bs:] - ema_soft) ** 2)
    loss_semi = (cps_weight * ps1 + mt_weight * cl1) + (cps_weight * ps2 + mt_weight * cl2)
    return loss_sup + loss_semi, loss_sup, loss_semi, pl1, pl2


def dense_contrastive(out_1, out_2, temperature=0.7):
    """Dense_Loss.contrastive_loss (utils/loss/dense_loss.py:18-34) with batch_size = out_1.shape[0]."""
    bs = out_1.shape[0]
    out_1 = F.normalize(out_1, dim=1).flatten(1)
    out_2 = F.normalize(out_2, dim=1).flatten(1)
    out = torch.cat([out_1, out_2], dim=0)
    sim = torch.exp(torch.mm(out, out.t().contiguous()) / temperature)
    mask = (torch.ones_like(sim) - torch.eye(2 * bs)).bool()
    sim = sim.masked_select(mask).view(2 * bs, -1)
    pos = torch.exp(torch.sum(out_1 * out_2, dim=-1) / temperature)
    pos = torch.cat([pos, pos], dim=0)
    return (-torch.log(pos / sim.sum(dim=-1))).mean()


def dense_loss(x, y, temperature=0.7):
    """Dense_Loss.forward (utils/loss/dense_loss.py:36-40): x, y = (global vector, dense map) pairs; y is detached."""
    return 0.5 * (dense_contrastive(x[0], y[0].detach(), temperature) + dense_contrastive(x[1], y[1].detach(), temperature))
