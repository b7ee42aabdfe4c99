"""Generate tests/golden/*.pt by running the REAL reference (loaded by file path from /root/reference or
$HPFG_REF).  Run from the repo root in the build container:  python tests/golden/make_golden.py

Every fixture stores OUTPUTS of the reference's own objects (model/unet.py UNet, utils/loss Med_Sup_Loss /
DiceLoss / softmax_mse_loss, utils/utils.py update_ema_variables, utils/scheduler Medical_LR,
torch.optim.SGD as utils/__init__.py:14-16 builds it) on inputs regenerated from seeds by common.py.
The trainer scripts themselves cannot run here (datasets, tensorboardX, medpy are absent), so the step
fixtures are literal transcriptions of the step bodies calling those reference objects.
"""
import copy
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_loader import load_reference            # noqa: E402
from tests.golden.common import (make_state, make_masks, make_batch, summarize, ENC_PREFIXES, make_predict_case, make_plus_state,
                                 hpfg_main_step, make_hpfg_batch)  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ref = load_reference()
torch.set_num_threads(8)


class Args(dict):
    __getattr__ = dict.__getitem__


def ref_model(st, in_ch, n_cls):
    m = ref.UNet(in_channels=in_ch, num_classes=n_cls)
    m.load_state_dict(st)
    m.train()
    return m


def inject_masks(model, masks):
    """Replace each nn.Dropout's output by inp*mask/(1-p) (what nn.Dropout computes for that mask)."""
    handles = []
    for prefix in ENC_PREFIXES:
        drop = model.get_submodule(prefix + ".conv_conv.3")
        assert isinstance(drop, nn.Dropout)
        if masks is None or prefix not in masks:
            h = drop.register_forward_hook(lambda mod, inp, out: inp[0])
        else:
            mk = masks[prefix]
            h = drop.register_forward_hook(
                lambda mod, inp, out, mk=mk, p=drop.p: inp[0] * mk.to(inp[0].dtype) / (1.0 - p))
        handles.append(h)
    return handles


def golden_unet(tag, in_ch, n_cls, n, h, w, seed, use_masks):
    st = make_state(in_ch, n_cls, seed)
    x_l, x_u, y = make_batch(n, 0, in_ch, n_cls, h, w, seed + 7)
    masks = make_masks(n, h, w, seed + 11) if use_masks else None
    m = ref_model(st, in_ch, n_cls)
    hs = inject_masks(m, masks)
    taps = {}
    for name in ["encoder.in_conv.conv_conv.0", "encoder.in_conv", "encoder.down2", "encoder.down4",
                 "decoder.up1", "decoder.up4"]:
        m.get_submodule(name).register_forward_hook(
            lambda mod, inp, out, name=name: taps.__setitem__(name, out.detach().clone()))
    logits = m(x_l)
    loss = ref.Med_Sup_Loss(n_cls)(logits, y)
    loss.backward()
    out = {"cfg": dict(in_ch=in_ch, n_cls=n_cls, n=n, h=h, w=w, seed=seed, use_masks=use_masks),
           "logits": logits.detach().clone(), "loss": loss.item(),
           "taps": {k: summarize(v) for k, v in taps.items()},
           "grads": {k: summarize(p.grad) for k, p in m.named_parameters()},
           "buffers": {k: v.clone() for k, v in m.state_dict().items() if "running" in k or "tracked" in k}}
    # eval-mode forward with the updated running statistics
    for hnd in hs:
        hnd.remove()
    m.eval()
    with torch.no_grad():
        out["logits_eval"] = m(x_l).clone()
    torch.save(out, os.path.join(HERE, "unet_%s.pt" % tag))
    print("unet_%s: loss %.8f" % (tag, out["loss"]))


def golden_losses():
    g = torch.Generator().manual_seed(4242)
    out = {}
    for tag, (n_l, n_u, C, H, W) in {"c4": (3, 2, 4, 24, 40), "c2": (2, 3, 2, 16, 16)}.items():
        s = (3.0 * torch.randn(n_l + n_u, C, H, W, generator=g)).requires_grad_(True)
        t = 3.0 * torch.randn(n_l + n_u, C, H, W, generator=g)
        y = torch.randint(0, C, (n_l, H, W), generator=g)
        y255 = y.clone()
        y255[torch.rand(y.shape, generator=g) < 0.1] = 255
        rec = {"shape": (n_l, n_u, C, H, W)}
        # Med_Sup_Loss with and without ignored pixels (utils/loss/medloss.py:44-56)
        for nm, yy in (("sup", y), ("sup255", y255)):
            l = ref.Med_Sup_Loss(C)(s[:n_l], yy)
            (gr,) = torch.autograd.grad(l, s)
            rec[nm] = l.item()
            rec[nm + "_grad"] = gr.clone()
        # DiceLoss with class weights and softmax flag (utils/loss/diceloss.py:178-191)
        wts = [0.5 + 0.25 * i for i in range(C)]
        l = ref.DiceLoss(C)(s[:n_l], y.unsqueeze(1), weight=wts, softmax=True)
        (gr,) = torch.autograd.grad(l, s)
        rec["dice_w"], rec["dice_w_grad"], rec["dice_weights"] = l.item(), gr.clone(), wts
        l = ref.DiceLoss(C)(torch.softmax(s, 1)[:n_l], y.unsqueeze(1))
        rec["dice"] = l.item()
        # Mean-Teacher step loss (2017_03...:97-106) with w = 0.37
        w = 0.37
        soft, tsoft = torch.softmax(s, 1), torch.softmax(t, 1)
        sup = ref.Med_Sup_Loss(C)(s[:n_l], y)
        cons = torch.mean((soft[n_l:] - tsoft[n_l:]) ** 2)
        l = sup + w * cons
        (gr,) = torch.autograd.grad(l, s)
        rec["mt"] = dict(w=w, loss=l.item(), sup=sup.item(), cons=cons.item(), grad=gr.clone())
        # CPS (2021_06...:99-111)
        s2 = (3.0 * torch.randn(n_l + n_u, C, H, W, generator=g)).requires_grad_(True)
        med = ref.Med_Sup_Loss(C)
        soft1, soft2 = torch.softmax(s, 1), torch.softmax(s2, 1)
        lsup = med(s[:n_l], y) + med(s2[:n_l], y)
        pl1 = torch.argmax(soft1[n_l:].detach(), dim=1)
        pl2 = torch.argmax(soft2[n_l:].detach(), dim=1)
        lsemi = med(s[n_l:], pl2) + med(s2[n_l:], pl1)
        l = lsup + w * lsemi
        g1, g2 = torch.autograd.grad(l, [s, s2])
        rec["cps"] = dict(w=w, loss=l.item(), sup=lsup.item(), semi=lsemi.item(), pl1=pl1.clone(),
                          pl2=pl2.clone(), grad1=g1.clone(), grad2=g2.clone(), logits2=s2.detach().clone())
        # UAMT (2019_07...:145-160), T = 8
        T = 8
        base = 3.0 * torch.randn(1, n_u, C, H, W, generator=g)      # correlated passes -> a non-trivial mask
        mc = (base + 0.7 * torch.randn(T, n_u, C, H, W, generator=g)).reshape(T * n_u, C, H, W)
        t_u = t[n_l:]
        preds = F.softmax(mc, dim=1).reshape(T, n_u, C, H, W)
        preds = torch.mean(preds, dim=0)
        unc = -1.0 * torch.sum(preds * torch.log(preds + 1e-6), dim=1, keepdim=True)
        loss_ce = nn.CrossEntropyLoss(ignore_index=255)(s[:n_l], y)
        loss_dice = ref.DiceLoss(C)(torch.softmax(s, 1)[:n_l], y.unsqueeze(1))
        sup_u = 0.5 * (loss_dice + loss_ce)
        dist = ref.softmax_mse_loss(s[n_l:], t_u)
        thr = (0.75 + 0.25 * ref.sigmoid_rampup(5000, 30000)) * np.log(2)
        mask = (unc < thr).float()
        cons_u = torch.sum(mask * dist) / (2 * torch.sum(mask) + 1e-16)
        l = sup_u + w * cons_u
        (gr,) = torch.autograd.grad(l, s)
        rec["uamt"] = dict(w=w, T=T, threshold=float(thr), loss=l.item(), sup=sup_u.item(), cons=cons_u.item(),
                           grad=gr.clone(), mc_logits=mc.clone(), uncertainty=unc.clone(),
                           mask_sum=mask.sum().item())
        rec["student"], rec["teacher"], rec["y"], rec["y255"] = s.detach().clone(), t.clone(), y, y255
        out[tag] = rec
    torch.save(out, os.path.join(HERE, "losses.pt"))
    print("losses: sup %.8f mt %.8f cps %.8f uamt %.8f" % (out["c4"]["sup"], out["c4"]["mt"]["loss"],
                                                           out["c4"]["cps"]["loss"], out["c4"]["uamt"]["loss"]))


def golden_schedules():
    out = {"rampup": [(it, ref.get_current_consistency_weight(it // 150, Args(consistency=0.1,
                                                                              consistency_rampup=200.0)))
                      for it in (1, 149, 150, 1500, 15000, 29999, 30000, 45000)],
           "sigmoid": [(c, L, ref.sigmoid_rampup(c, L)) for c, L in ((0, 200.0), (17, 200.0), (250, 200.0),
                                                                      (3, 0))]}
    p = [nn.Parameter(torch.zeros(3))]
    opt = torch.optim.SGD(p, lr=0.01, momentum=0.9, weight_decay=1e-4)
    sch = ref.Medical_LR(optimizer=opt, base_lr=0.01, max_iterations=30000)
    lrs = []
    for _ in range(5):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    out["medical_lr_first5"] = lrs
    # EMA (utils/utils.py:82-86) on two small "models"
    g = torch.Generator().manual_seed(99)
    a, b = nn.Linear(7, 5), nn.Linear(7, 5)
    with torch.no_grad():
        for q in list(a.parameters()) + list(b.parameters()):
            q.copy_(torch.randn(q.shape, generator=g))
    ema_in = {"student": [q.detach().clone() for q in a.parameters()],
              "teacher": [q.detach().clone() for q in b.parameters()]}
    ema_out = {}
    for step in (1, 2, 50, 99, 100, 5000):
        bb = copy.deepcopy(b)
        ref.update_ema_variables(a, bb, 0.99, step)
        ema_out[step] = [q.detach().clone() for q in bb.parameters()]
    out["ema_in"], out["ema_out"] = ema_in, ema_out
    torch.save(out, os.path.join(HERE, "schedules.pt"))
    print("schedules: lrs", lrs)


def golden_mt_steps(tag, in_ch, n_cls, n_l, n_u, h, w, seed, steps=3):
    """Literal Mean-Teacher step (2017_03_NIPS_Mean-Teacher_ACDC.py:55-57,64-70,89-113) on reference objects."""
    st = make_state(in_ch, n_cls, seed)
    model = ref_model(st, in_ch, n_cls)
    ema_model = copy.deepcopy(model)
    for name, p in ema_model.named_parameters():
        p.requires_grad = False
    args = Args(lr=0.01, momentum=0.9, weight_decay=1e-4, total_itrs=30000, ema_decay=0.99, consistency=0.1,
                consistency_rampup=200.0, num_classes=n_cls)
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum,
                                weight_decay=args.weight_decay)
    lr_scheduler = ref.Medical_LR(optimizer=optimizer, base_lr=args.lr, max_iterations=args.total_itrs)
    med_loss = ref.Med_Sup_Loss(args.num_classes)
    model.train()
    ema_model.train()
    recs = []
    cur_itrs = 0
    for it in range(steps):
        cur_itrs += 1
        label_img, unlabel_img, target_label = make_batch(n_l, n_u, in_ch, n_cls, h, w, seed + 100 * cur_itrs)
        hs = inject_masks(model, make_masks(n_l + n_u, h, w, seed + 100 * cur_itrs + 1))
        hs += inject_masks(ema_model, make_masks(n_l + n_u, h, w, seed + 100 * cur_itrs + 2))
        label_bs = label_img.shape[0]
        x = torch.cat([label_img, unlabel_img], dim=0)
        output = model(x)
        output_soft = torch.softmax(output, dim=1)
        with torch.no_grad():
            ema_output = ema_model(x)
            ema_output_soft = torch.softmax(ema_output, dim=1)
        loss_sup = med_loss(output[:label_bs], target_label)
        loss_consistence = torch.mean((output_soft[label_bs:] - ema_output_soft[label_bs:]) ** 2)
        consistency_weight = ref.get_current_consistency_weight(epoch=cur_itrs // 150, args=args)
        loss = loss_sup + consistency_weight * loss_consistence
        optimizer.zero_grad()
        loss.backward()
        lr = optimizer.param_groups[0]["lr"]
        optimizer.step()
        lr_scheduler.step()
        ref.update_ema_variables(model, ema_model, args.ema_decay, cur_itrs)
        for hnd in hs:
            hnd.remove()
        recs.append(dict(loss=loss.item(), sup=loss_sup.item(), cons=loss_consistence.item(),
                         w=consistency_weight, lr=lr,
                         logits=summarize(output), teacher_logits=summarize(ema_output),
                         student_sum=sum(p.double().sum().item() for p in model.parameters()),
                         teacher_sum=sum(p.double().sum().item() for p in ema_model.parameters()),
                         student_out_conv=model.decoder.out_conv.weight.detach().clone(),
                         teacher_out_conv=ema_model.decoder.out_conv.weight.detach().clone(),
                         student_in_conv=model.encoder.in_conv.conv_conv[0].weight.detach().clone(),
                         teacher_rm=ema_model.encoder.in_conv.conv_conv[1].running_mean.clone(),
                         student_rv=model.decoder.up4.conv.conv_conv[5].running_var.clone()))
    out = {"cfg": dict(in_ch=in_ch, n_cls=n_cls, n_l=n_l, n_u=n_u, h=h, w=w, seed=seed, steps=steps),
           "steps": recs}
    torch.save(out, os.path.join(HERE, "mt_steps_%s.pt" % tag))
    print("mt_steps_%s:" % tag, [r["loss"] for r in recs])



def golden_f4_losses():
    """SURVEY 8f.4 loss modes on the reference's own DiceLoss / CrossEntropyLoss objects: the inline expressions of
    2022_02_ISBI_ICT-MedSeg_ACDC.py:119-135 and 2022_08_CVPR_S4CVNet_ACDC.py:124-156, transcribed literally."""
    g = torch.Generator().manual_seed(5151)
    out = {}
    for tag, (n_l, n_m, C, H, W) in {"c4": (3, 2, 4, 24, 40), "c2": (2, 3, 2, 16, 16)}.items():
        rec = {"shape": (n_l, n_m, C, H, W)}
        criterion = nn.CrossEntropyLoss(ignore_index=255)
        dice_loss = ref.DiceLoss(C)
        # ---- ICT: student sees n_l + n_m slices, the teacher the 2*n_m un-mixed ones
        outputs = (3.0 * torch.randn(n_l + n_m, C, H, W, generator=g)).requires_grad_(True)
        ema0 = 3.0 * torch.randn(n_m, C, H, W, generator=g)
        ema1 = 3.0 * torch.randn(n_m, C, H, W, generator=g)
        target_label = torch.randint(0, C, (n_l, H, W), generator=g)
        target_label[torch.rand(target_label.shape, generator=g) < 0.05] = 255
        ict_mix_factors = torch.rand(n_m, 1, 1, 1, generator=g)
        label_bs = n_l
        outputs_soft = torch.softmax(outputs, dim=1)
        ema_output_ux0 = torch.softmax(ema0, dim=1)
        ema_output_ux1 = torch.softmax(ema1, dim=1)
        batch_pred_mixed = ema_output_ux0 * (1.0 - ict_mix_factors) + ema_output_ux1 * ict_mix_factors
        loss_ce = criterion(outputs[:label_bs], target_label)
        loss_dice = dice_loss(outputs_soft[:label_bs], target_label.unsqueeze(1))
        supervised_loss = 0.5 * (loss_dice + loss_ce)
        consistency_weight = 0.37
        consistency_loss = torch.mean((outputs_soft[label_bs:] - batch_pred_mixed) ** 2)
        loss = supervised_loss + consistency_weight * consistency_loss
        (gr,) = torch.autograd.grad(loss, outputs)
        rec["ict"] = dict(w=consistency_weight, loss=loss.item(), sup=supervised_loss.item(), cons=consistency_loss.item(),
                          grad=gr.clone(), student=outputs.detach().clone(), teacher=torch.cat([ema0, ema1]).clone(),
                          mix=ict_mix_factors.reshape(-1).clone(), y=target_label.clone())
        # ---- S4CVNet: two students on n_l + n_u slices, one teacher on the n_u unlabeled ones
        n_u = n_m + 1
        outputs1 = (3.0 * torch.randn(n_l + n_u, C, H, W, generator=g)).requires_grad_(True)
        outputs2 = (3.0 * torch.randn(n_l + n_u, C, H, W, generator=g)).requires_grad_(True)
        ema_output = 3.0 * torch.randn(n_u, C, H, W, generator=g)
        label_batch_size = n_l
        s4 = {}
        for branch, cur_itrs in (("early", 500), ("late", 1500)):
            outputs_soft1 = torch.softmax(outputs1, dim=1)
            outputs_soft2 = torch.softmax(outputs2, dim=1)
            ema_output_soft = torch.softmax(ema_output, dim=1)
            loss1 = 0.5 * (criterion(outputs1[:label_batch_size], target_label) +
                           dice_loss(outputs_soft1[:label_batch_size], target_label.unsqueeze(1)))
            loss2 = 0.5 * (criterion(outputs2[:label_batch_size], target_label) +
                           dice_loss(outputs_soft2[:label_batch_size], target_label.unsqueeze(1)))
            loss_sup = loss1 + loss2
            pseudo_outputs1 = torch.argmax(outputs_soft1[label_batch_size:].detach(), dim=1, keepdim=False)
            pseudo_outputs2 = torch.argmax(outputs_soft2[label_batch_size:].detach(), dim=1, keepdim=False)
            pseudo_supervision1 = dice_loss(outputs_soft1[label_batch_size:], pseudo_outputs2.unsqueeze(1))
            pseudo_supervision2 = dice_loss(outputs_soft2[label_batch_size:], pseudo_outputs1.unsqueeze(1))
            consistency_weight_cps = 0.1 * ref.linear_rampup(cur_itrs // 150, 200.0)
            consistency_weight_mt = 0.1 * ref.linear_rampup(cur_itrs // 150, 200.0)
            if cur_itrs < 1000:
                consistency_loss1 = 0.0
                consistency_loss2 = 0.0
            else:
                consistency_loss1 = torch.mean((outputs_soft1[label_batch_size:] - ema_output_soft) ** 2)
                consistency_loss2 = torch.mean((outputs_soft2[label_batch_size:] - ema_output_soft) ** 2)
            model1_loss = 7 * consistency_weight_cps * pseudo_supervision1 + consistency_weight_mt * consistency_loss1
            model2_loss = 7 * consistency_weight_cps * pseudo_supervision2 + consistency_weight_mt * consistency_loss2
            loss_semi = model1_loss + model2_loss
            loss = loss_sup + loss_semi
            g1, g2 = torch.autograd.grad(loss, [outputs1, outputs2])
            s4[branch] = dict(cur_itrs=cur_itrs, cps_weight=7 * consistency_weight_cps, mt_weight=consistency_weight_mt,
                              loss=loss.item(), sup=loss_sup.item(), semi=loss_semi.item(), grad1=g1.clone(),
                              grad2=g2.clone(), pl1=pseudo_outputs1.clone(), pl2=pseudo_outputs2.clone())
        s4.update(logits1=outputs1.detach().clone(), logits2=outputs2.detach().clone(), teacher=ema_output.clone(),
                  y=target_label.clone(), n_u=n_u)
        rec["s4cv"] = s4
        out[tag] = rec
    torch.save(out, os.path.join(HERE, "f4_losses.pt"))
    print("f4_losses: ict %.8f s4cv early %.8f late %.8f" % (out["c4"]["ict"]["loss"], out["c4"]["s4cv"]["early"]["loss"],
                                                             out["c4"]["s4cv"]["late"]["loss"]))


def golden_ict_steps(tag, in_ch, n_cls, n_l, n_u, h, w, seed, steps=2):
    """Literal ICT-MedSeg step (2022_02_ISBI_ICT-MedSeg_ACDC.py:55-59,66-76,96-140) on reference objects; the Beta draws
    are replaced by seeded uniform mix factors and .cuda() by the CPU device."""
    st = make_state(in_ch, n_cls, seed)
    model = ref_model(st, in_ch, n_cls)
    ema_model = ref_model(make_state(in_ch, n_cls, seed + 7), in_ch, n_cls)      # built separately, as at :57
    for name, param in ema_model.named_parameters():
        param.requires_grad = False
    args = Args(lr=0.01, momentum=0.9, weight_decay=1e-4, total_itrs=30000, ema_decay=0.99, consistency=0.1,
                consistency_rampup=200.0, num_classes=n_cls)
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    lr_scheduler = ref.Medical_LR(optimizer=optimizer, base_lr=args.lr, max_iterations=args.total_itrs)
    criterion = nn.CrossEntropyLoss(ignore_index=255)
    dice_loss = ref.DiceLoss(args.num_classes)
    model.train()
    recs = []
    cur_itrs = 0
    for it in range(steps):
        cur_itrs += 1
        img_labeled, img_unlabeled, target_label = make_batch(n_l, n_u, in_ch, n_cls, h, w, seed + 100 * cur_itrs)
        label_bs = img_labeled.shape[0]
        unlabel_bs = img_unlabeled.shape[0]
        hs = inject_masks(model, make_masks(n_l + unlabel_bs // 2, h, w, seed + 100 * cur_itrs + 1))
        hs += inject_masks(ema_model, make_masks(unlabel_bs // 2, h, w, seed + 100 * cur_itrs + 2))
        ict_mix_factors = torch.rand(unlabel_bs // 2, 1, 1, 1, generator=torch.Generator().manual_seed(seed + 100 * cur_itrs + 3))
        unlabeled_volume_batch_0 = img_unlabeled[0:unlabel_bs // 2, ...]
        unlabeled_volume_batch_1 = img_unlabeled[unlabel_bs // 2:, ...]
        batch_ux_mixed = unlabeled_volume_batch_0 * (1.0 - ict_mix_factors) + unlabeled_volume_batch_1 * ict_mix_factors
        input_volume_batch = torch.cat([img_labeled, batch_ux_mixed], dim=0)
        outputs = model(input_volume_batch)
        outputs_soft = torch.softmax(outputs, dim=1)
        with torch.no_grad():
            ema0 = ema_model(unlabeled_volume_batch_0)
            ema1 = ema_model(unlabeled_volume_batch_1)
            ema_output_ux0 = torch.softmax(ema0, dim=1)
            ema_output_ux1 = torch.softmax(ema1, dim=1)
            batch_pred_mixed = ema_output_ux0 * (1.0 - ict_mix_factors) + ema_output_ux1 * ict_mix_factors
        loss_ce = criterion(outputs[:label_bs], target_label)
        loss_dice = dice_loss(outputs_soft[:label_bs], target_label.unsqueeze(1))
        supervised_loss = 0.5 * (loss_dice + loss_ce)
        consistency_weight = ref.get_current_consistency_weight(epoch=cur_itrs // 150, args=args)
        consistency_loss = torch.mean((outputs_soft[label_bs:] - batch_pred_mixed) ** 2)
        loss = supervised_loss + consistency_weight * consistency_loss
        optimizer.zero_grad()
        loss.backward()
        lr = optimizer.param_groups[0]["lr"]
        optimizer.step()
        lr_scheduler.step()
        ref.update_ema_variables(model, ema_model, args.ema_decay, cur_itrs)
        for hnd in hs:
            hnd.remove()
        recs.append(dict(loss=loss.item(), sup=supervised_loss.item(), cons=consistency_loss.item(), w=consistency_weight,
                         lr=lr, logits=summarize(outputs), teacher_logits=summarize(torch.cat([ema0, ema1])),
                         student_sum=sum(p.double().sum().item() for p in model.parameters()),
                         teacher_sum=sum(p.double().sum().item() for p in ema_model.parameters()),
                         student_out_conv=model.decoder.out_conv.weight.detach().clone(),
                         teacher_out_conv=ema_model.decoder.out_conv.weight.detach().clone(),
                         teacher_rm=ema_model.encoder.in_conv.conv_conv[1].running_mean.clone()))
    out = {"cfg": dict(in_ch=in_ch, n_cls=n_cls, n_l=n_l, n_u=n_u, h=h, w=w, seed=seed, steps=steps), "steps": recs}
    torch.save(out, os.path.join(HERE, "ict_steps_%s.pt" % tag))
    print("ict_steps_%s:" % tag, [r["loss"] for r in recs])


def golden_predict(in_ch, n_cls, n, h, w, seed):
    """val.py:268-281 network part: eval-mode argmax(softmax(net(slice))) slice by slice at batch 1, after a few
    train-mode forwards so that the BatchNorm running statistics are non-trivial."""
    st, vol = make_predict_case(in_ch, n_cls, n, h, w, seed)
    net = ref_model(st, in_ch, n_cls)
    hs = inject_masks(net, None)
    with torch.no_grad():
        for i in range(3):
            net(torch.rand(4, in_ch, h, w, generator=torch.Generator().manual_seed(seed + 10 + i)))
    for hnd in hs:
        hnd.remove()
    net.eval()
    with torch.no_grad():                       # centre the class logits so that the labels are a mix of classes
        net.decoder.out_conv.bias.copy_(-net(vol.unsqueeze(1)).mean(dim=(0, 2, 3)))
    buffers = {k: v.clone() for k, v in net.state_dict().items() if "running_" in k or "num_batches" in k}
    buffers["decoder.out_conv.bias"] = net.decoder.out_conv.bias.detach().clone()
    preds, margins = [], []
    for ind in range(n):
        input = vol[ind].unsqueeze(0).unsqueeze(0).float()
        net.eval()
        with torch.no_grad():
            soft = torch.softmax(net(input), dim=1)
            preds.append(torch.argmax(soft, dim=1).squeeze(0))
            top2 = soft.topk(2, dim=1).values
            margins.append((top2[:, 0] - top2[:, 1]).squeeze(0))
    out = dict(cfg=dict(in_ch=in_ch, n_cls=n_cls, n=n, h=h, w=w, seed=seed), buffers=buffers,
               labels=torch.stack(preds).to(torch.uint8), margin=torch.stack(margins).half())
    torch.save(out, os.path.join(HERE, "predict_acdc.pt"))
    print("predict: label histogram", torch.bincount(out["labels"].flatten().long(), minlength=n_cls).tolist())


def golden_unet_plus(in_ch, n_cls, n, h, w, seed):
    """The reference UNet_Plus + Dense_Loss (model/unet.py:178-206, utils/loss/dense_loss.py) the way main.py:151-170 uses
    them for model2 / ema_model: supervised loss on the logits + contrastive loss between student and teacher necks."""
    st = make_plus_state(in_ch, n_cls, seed)
    model2 = ref.UNet_Plus(in_channels=in_ch, num_classes=n_cls)
    model2.load_state_dict(st)
    model2.train()
    ema_model = ref.UNet_Plus(in_channels=in_ch, num_classes=n_cls)
    ema_model.load_state_dict(make_plus_state(in_ch, n_cls, seed + 9))
    ema_model.train()
    volume_batch, _, target_label = make_batch(n, 0, in_ch, n_cls, h, w, seed + 7)
    hs = inject_masks(model2, make_masks(n, h, w, seed + 11)) + inject_masks(ema_model, make_masks(n, h, w, seed + 12))
    dense_loss = ref.Dense_Loss(batch_size=n, device=torch.device("cpu"))
    criterion = nn.CrossEntropyLoss(ignore_index=255)
    dice_loss = ref.DiceLoss(n_cls)
    outputs2, h1, h2 = model2(volume_batch)
    outputs_soft2 = torch.softmax(outputs2, dim=1)
    with torch.no_grad():
        ema_output, ema_h1, ema_h2 = ema_model(volume_batch)
    loss2 = 0.5 * (criterion(outputs2, target_label) + dice_loss(outputs_soft2, target_label.unsqueeze(1)))
    loss_constrivate = dense_loss(h1, ema_h1) + dense_loss(h2, ema_h2)
    weight = 0.3
    loss = loss2 + weight * loss_constrivate
    loss.backward()
    for hnd in hs:
        hnd.remove()
    model2.eval()
    with torch.no_grad():
        val = model2.val(volume_batch)
    out = dict(cfg=dict(in_ch=in_ch, n_cls=n_cls, n=n, h=h, w=w, seed=seed, weight=weight), loss=loss.item(), sup=loss2.item(),
               contrast=loss_constrivate.item(), logits=summarize(outputs2), val_logits=summarize(val),
               h1=[h1[0].detach().clone(), h1[1].detach().clone()], h2=[h2[0].detach().clone(), h2[1].detach().clone()],
               ema_h1=[ema_h1[0].clone(), ema_h1[1].clone()], ema_h2=[ema_h2[0].clone(), ema_h2[1].clone()],
               grads={k: summarize(p.grad) for k, p in model2.named_parameters()})
    torch.save(out, os.path.join(HERE, "unet_plus_acdc.pt"))
    print("unet_plus: loss %.8f (sup %.8f, contrast %.8f)" % (out["loss"], out["sup"], out["contrast"]))


def golden_hpfg_step(in_ch, n_cls, n_l, n_u, h, w, seed, cur_itrs=1500):
    """One iteration of main.py (HPFG, :128-207) on the reference's UNet_Plus x3, Dense_Loss, DiceLoss, SGD, Medical_LR."""
    args = Args(num_classes=n_cls, batch_size=n_l, unlabel_batch_size=n_u, consistency=0.1, consistency_rampup=200.0,
                ema_decay=0.99)
    model1 = ref.UNet_Plus(in_channels=in_ch, num_classes=n_cls)
    model1.load_state_dict(make_plus_state(in_ch, n_cls, seed))
    model2 = ref.UNet_Plus(in_channels=in_ch, num_classes=n_cls)
    model2.load_state_dict(make_plus_state(in_ch, n_cls, seed + 20))
    ema_model = copy.deepcopy(model2)
    for name, param in ema_model.named_parameters():
        param.requires_grad = False
    optimizer1 = torch.optim.SGD(model1.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0005)
    optimizer2 = torch.optim.SGD(model2.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0005)
    sch1 = ref.Medical_LR(optimizer=optimizer1, base_lr=0.01, max_iterations=30000)
    sch2 = ref.Medical_LR(optimizer=optimizer2, base_lr=0.01, max_iterations=30000)
    model1.train(), model2.train()
    hs = inject_masks(model1, make_masks(n_l + n_u, h, w, seed + 31)) + inject_masks(model2, make_masks(n_l + n_u, h, w, seed + 32))
    hs += inject_masks(ema_model, make_masks(n_l + n_u, h, w, seed + 33))
    r = hpfg_main_step(ref, model1, model2, ema_model, optimizer1, optimizer2, sch1, sch2,
                       make_hpfg_batch(n_l, n_u, in_ch, n_cls, h, w, seed + 40), cur_itrs, args, torch.device("cpu"))
    for hnd in hs:
        hnd.remove()
    out = dict(cfg=dict(in_ch=in_ch, n_cls=n_cls, n_l=n_l, n_u=n_u, h=h, w=w, seed=seed, cur_itrs=cur_itrs),
               scalars={k: v for k, v in r.items() if isinstance(v, float)},
               outputs1=summarize(r["outputs1"]), outputs2=summarize(r["outputs2"]), ema_output=summarize(r["ema_output"]),
               after={nm: {k: summarize(v, full_below=512) for k, v in m.state_dict().items() if "tracked" not in k}
                      for nm, m in (("model1", model1), ("model2", model2), ("ema_model", ema_model))})
    torch.save(out, os.path.join(HERE, "hpfg_step_acdc.pt"))
    print("hpfg_step:", out["scalars"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "f4":      # only the SURVEY 8f.3 / 8f.4 fixtures (leaves the others untouched)
        golden_f4_losses()
        golden_ict_steps("acdc", 1, 4, 2, 4, 32, 32, 606)
        golden_predict(1, 4, 5, 32, 48, 707)
        golden_unet_plus(1, 4, 3, 64, 64, 808)
        golden_hpfg_step(1, 4, 2, 4, 64, 64, 909)
        sys.exit(0)
    golden_unet("acdc_masks", 1, 4, 2, 32, 48, 101, True)
    golden_unet("acdc_nodrop", 1, 4, 3, 32, 32, 202, False)
    golden_unet("isic_masks", 3, 2, 2, 48, 32, 303, True)
    golden_losses()
    golden_schedules()
    golden_mt_steps("acdc", 1, 4, 2, 2, 32, 32, 404)
    golden_mt_steps("isic", 3, 2, 1, 3, 32, 32, 505, steps=2)
    golden_f4_losses()
    golden_ict_steps("acdc", 1, 4, 2, 4, 32, 32, 606)
    golden_predict(1, 4, 5, 32, 48, 707)
    golden_unet_plus(1, 4, 3, 64, 64, 808)
    golden_hpfg_step(1, 4, 2, 4, 64, 64, 909)
