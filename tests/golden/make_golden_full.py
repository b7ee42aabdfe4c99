"""Generate tests/golden/full_*.pt: the REAL reference (loaded by file path, see oracle/ref_loader.py) run on CPU at
the BASELINE.json shapes -- 224x224, config 1 (Mean-Teacher 12+12), config 2 (Mean-Teacher 8+24, two iterations),
config 3 (CPS 8+24), config 4 (UAMT 12+12, T=8, a separate dropout mask set for every teacher forward) and config 5
(ISIC-shape 3ch / 2 classes 12+12).  Only fingerprints are stored (strided samples + fp64 sums, a few hundred KB per
config); inputs, weights, dropout masks and noise are regenerated from seeds by tests/golden/common.py on both sides.

    python tests/golden/make_golden_full.py            # all configs (~15 min on 8 cores, ~25 GB of RAM: fp32 pass + fp64 pass)
    python tests/golden/make_golden_full.py mt_cfg1    # one config

The step bodies are literal transcriptions of 2017_03_NIPS_Mean-Teacher_ACDC.py:89-113, 2021_06_CVPR_CPS_ACDC.py:90-120
and 2019_07_MICCAI_Uncertainty_Aware_ACDC.py:120-170 on the reference's own UNet / Med_Sup_Loss / DiceLoss /
softmax_mse_loss / update_ema_variables / Medical_LR objects (the scripts themselves need datasets, tensorboardX and a
GPU; `.cuda()` at 2019_07...:138 is replaced by the CPU device and torch.randn_like by the seeded noise of common.py)."""
import copy
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_loader import load_reference            # noqa: E402
from tests.golden.common import (make_state, make_masks, make_batch, summarize, ENC_PREFIXES, make_uamt_noise, make_uamt_teacher_state,
                                 FULL_CONFIGS)  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ref = load_reference()
torch.set_num_threads(os.cpu_count())


class Args(dict):
    __getattr__ = dict.__getitem__


ARGS = dict(lr=0.01, momentum=0.9, weight_decay=1e-4, total_itrs=30000, ema_decay=0.99, consistency=0.1,
            consistency_rampup=200.0)


DTYPE = torch.float32        # the second pass of every config runs in fp64: the reference's own fp32 rounding noise is measured


def ref_model(st, in_ch, n_cls):
    m = ref.UNet(in_channels=in_ch, num_classes=n_cls)
    m.load_state_dict(st)
    m = m.to(DTYPE)
    m.train()
    return m


def batch_of(*a, **k):
    x_l, x_u, y = make_batch(*a, **k)
    return x_l.to(DTYPE), x_u.to(DTYPE), y


class MaskFeed:
    """Forward hooks on the five encoder nn.Dropout modules that replace their output by inp*mask/(1-p) for the mask set
    currently in ``self.masks`` (what nn.Dropout computes for that draw); the set is swapped between forwards."""

    def __init__(self, model):
        self.masks = None
        for prefix in ENC_PREFIXES:
            drop = model.get_submodule(prefix + ".conv_conv.3")
            assert isinstance(drop, nn.Dropout)
            drop.register_forward_hook(lambda mod, inp, out, prefix=prefix, p=drop.p:
                                       inp[0] * self.masks[prefix].to(inp[0].dtype) / (1.0 - p))


def conv_taps(model, store):
    """Every convolution's raw output (bias included), fingerprinted: the per-layer activations of north_star."""
    for name, mod in model.named_modules():
        if isinstance(mod, nn.Conv2d):
            mod.register_forward_hook(lambda m, i, o, name=name: store.__setitem__(name, summarize(o, stride=4999)))


def grads_of(model):
    return {k: summarize(p.grad, stride=211) for k, p in model.named_parameters()}


def after_of(model):
    sd = model.state_dict()
    return dict(param_sum=sum(p.double().sum().item() for p in model.parameters()),
                out_conv=sd["decoder.out_conv.weight"].clone(), in_conv=sd["encoder.in_conv.conv_conv.0.weight"].clone(),
                rm_first=sd["encoder.in_conv.conv_conv.1.running_mean"].clone(),
                rv_last=sd["decoder.up4.conv.conv_conv.5.running_var"].clone(),
                rv_deep=sd["encoder.down4.maxpool_conv.1.conv_conv.5.running_var"].clone())


def labels_fp(logits):
    """argmax(softmax(logits)) as the trainers compute it, fingerprinted: class histogram + strided uint8 sample + the
    top-2 softmax margin at the sampled pixels (so a bf16 run can report agreement away from near-ties)."""
    soft = torch.softmax(logits, dim=1)
    lab = torch.argmax(soft, dim=1)
    top2 = soft.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]).flatten()
    return dict(hist=torch.bincount(lab.flatten(), minlength=logits.shape[1]), stride=101,
                sample=lab.flatten()[::101].to(torch.uint8).clone(), margin=margin[::101].half().clone())


def mean_teacher(c):
    in_ch, n_cls, n_l, n_u, h, w, seed = c["in_ch"], c["n_cls"], c["n_l"], c["n_u"], c["h"], c["w"], c["seed"]
    args = Args(num_classes=n_cls, **ARGS)
    model = ref_model(make_state(in_ch, n_cls, seed), in_ch, n_cls)
    ema_model = copy.deepcopy(model)
    for name, p in ema_model.named_parameters():
        p.requires_grad = False
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    lr_scheduler = ref.Medical_LR(optimizer=optimizer, base_lr=args.lr, max_iterations=args.total_itrs)
    med_loss = ref.Med_Sup_Loss(args.num_classes)
    model.train()
    ema_model.train()
    fs, ft = MaskFeed(model), MaskFeed(ema_model)
    taps = {}
    conv_taps(model, taps)
    recs, cur_itrs = [], 0
    for it in range(c["steps"]):
        cur_itrs += 1
        label_img, unlabel_img, target_label = batch_of(n_l, n_u, in_ch, n_cls, h, w, seed + 100 * cur_itrs)
        fs.masks = make_masks(n_l + n_u, h, w, seed + 100 * cur_itrs + 1)
        ft.masks = make_masks(n_l + n_u, h, w, seed + 100 * cur_itrs + 2)
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
        rec = dict(loss=loss.item(), sup=loss_sup.item(), cons=loss_consistence.item(), w=consistency_weight, lr=lr,
                   logits=summarize(output, stride=1009), teacher_logits=summarize(ema_output, stride=1009),
                   labels=labels_fp(output.detach()), grads=grads_of(model), taps=dict(taps) if it == 0 else None)
        optimizer.step()
        lr_scheduler.step()
        ref.update_ema_variables(model, ema_model, args.ema_decay, cur_itrs)
        rec["student_after"], rec["teacher_after"] = after_of(model), after_of(ema_model)
        recs.append(rec)
        print("  it %d loss %.8f sup %.8f cons %.6e" % (cur_itrs, rec["loss"], rec["sup"], rec["cons"]), flush=True)
    return recs


def cps(c):
    in_ch, n_cls, n_l, n_u, h, w, seed = c["in_ch"], c["n_cls"], c["n_l"], c["n_u"], c["h"], c["w"], c["seed"]
    args = Args(num_classes=n_cls, **ARGS)
    model1 = ref_model(make_state(in_ch, n_cls, seed), in_ch, n_cls)
    model2 = ref_model(make_state(in_ch, n_cls, seed + 7), in_ch, n_cls)
    optimizer1 = torch.optim.SGD(model1.parameters(), lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    optimizer2 = torch.optim.SGD(model2.parameters(), lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    lr_scheduler1 = ref.Medical_LR(optimizer=optimizer1, base_lr=args.lr, max_iterations=args.total_itrs)
    lr_scheduler2 = ref.Medical_LR(optimizer=optimizer2, base_lr=args.lr, max_iterations=args.total_itrs)
    med_loss = ref.Med_Sup_Loss(args.num_classes)
    model1.train()
    model2.train()
    f1, f2 = MaskFeed(model1), MaskFeed(model2)
    recs, cur_itrs = [], 0
    for it in range(c["steps"]):
        cur_itrs += 1
        label_img, unlabel_img, target_label = batch_of(n_l, n_u, in_ch, n_cls, h, w, seed + 100 * cur_itrs)
        f1.masks = make_masks(n_l + n_u, h, w, seed + 100 * cur_itrs + 1)
        f2.masks = make_masks(n_l + n_u, h, w, seed + 100 * cur_itrs + 2)
        label_bs = label_img.shape[0]
        x = torch.cat([label_img, unlabel_img], dim=0)
        output1 = model1(x)
        output2 = model2(x)
        output_soft1 = torch.softmax(output1, dim=1)
        output_soft2 = torch.softmax(output2, dim=1)
        loss_sup1 = med_loss(output1[:label_bs], target_label)
        loss_sup2 = med_loss(output2[:label_bs], target_label)
        loss_sup = loss_sup1 + loss_sup2
        pseudo_label1 = torch.argmax(output_soft1[label_bs:].detach(), dim=1, keepdim=False)
        pseudo_label2 = torch.argmax(output_soft2[label_bs:].detach(), dim=1, keepdim=False)
        loss_semi = med_loss(output1[label_bs:], pseudo_label2) + med_loss(output2[label_bs:], pseudo_label1)
        consistency_weight = ref.get_current_consistency_weight(epoch=cur_itrs // 150, args=args)
        loss = loss_sup + consistency_weight * loss_semi
        optimizer1.zero_grad()
        optimizer2.zero_grad()
        loss.backward()
        lr = optimizer1.param_groups[0]["lr"]
        rec = dict(loss=loss.item(), sup=loss_sup.item(), semi=loss_semi.item(), w=consistency_weight, lr=lr,
                   logits1=summarize(output1, stride=1009), logits2=summarize(output2, stride=1009),
                   labels1=labels_fp(output1.detach()), labels2=labels_fp(output2.detach()),
                   grads1=grads_of(model1), grads2=grads_of(model2))
        optimizer1.step()
        optimizer2.step()
        lr_scheduler1.step()
        lr_scheduler2.step()
        rec["m1_after"], rec["m2_after"] = after_of(model1), after_of(model2)
        recs.append(rec)
        print("  it %d loss %.8f sup %.8f semi %.8f" % (cur_itrs, rec["loss"], rec["sup"], rec["semi"]), flush=True)
    return recs


def uamt(c):
    in_ch, n_cls, n_l, n_u, h, w, seed = c["in_ch"], c["n_cls"], c["n_l"], c["n_u"], c["h"], c["w"], c["seed"]
    args = Args(num_classes=n_cls, **ARGS)
    model = ref_model(make_state(in_ch, n_cls, seed), in_ch, n_cls)
    ema_model = ref_model(make_uamt_teacher_state(in_ch, n_cls, seed + 7, c["teacher_gain"]), in_ch, n_cls)   # a separately built network (2019_07...:55)
    for name, param in ema_model.named_parameters():
        param.requires_grad = False
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    lr_scheduler = ref.Medical_LR(optimizer=optimizer, base_lr=args.lr, max_iterations=args.total_itrs)
    criterion = nn.CrossEntropyLoss(ignore_index=255)
    dice_loss = ref.DiceLoss(args.num_classes)
    model.train()                                           # the teacher is a fresh module: train mode as well (:94)
    fs, ft = MaskFeed(model), MaskFeed(ema_model)
    recs, cur_itrs = [], 0
    for it in range(c["steps"]):
        cur_itrs += 1
        img_labeled, unlabeled_volume_batch, target_label = batch_of(n_l, n_u, in_ch, n_cls, h, w, seed + 100 * cur_itrs)
        noise0, mc_noise = [t.to(DTYPE) for t in make_uamt_noise(n_u, in_ch, h, w, c["T"], seed + 100 * cur_itrs + 3)]
        fs.masks = make_masks(n_l + n_u, h, w, seed + 100 * cur_itrs + 1)
        label_bs = img_labeled.shape[0]
        volume_batch = torch.cat([img_labeled, unlabeled_volume_batch], dim=0)
        outputs = model(volume_batch)
        outputs_soft = torch.softmax(outputs, dim=1)
        with torch.no_grad():
            noise = noise0
            ema_inputs = unlabeled_volume_batch + noise
            ft.masks = make_masks(n_u, h, w, seed + 100 * cur_itrs + 10)
            ema_output = ema_model(ema_inputs)
        T = c["T"]
        _, _, w_, h_ = unlabeled_volume_batch.shape
        volume_batch_r = unlabeled_volume_batch.repeat(2, 1, 1, 1)
        stride = volume_batch_r.shape[0] // 2
        preds = torch.zeros([stride * T, args.num_classes, w_, h_], dtype=DTYPE)
        for i in range(T // 2):
            ema_inputs = volume_batch_r + mc_noise[i]
            ft.masks = make_masks(2 * n_u, h, w, seed + 100 * cur_itrs + 11 + i)
            with torch.no_grad():
                preds[2 * stride * i:2 * stride * (i + 1)] = ema_model(ema_inputs)
        mc_fp = summarize(preds, stride=4999)
        preds = F.softmax(preds, dim=1)
        preds = preds.reshape(T, stride, args.num_classes, w_, h_)
        preds = torch.mean(preds, dim=0)
        uncertainty = -1.0 * torch.sum(preds * torch.log(preds + 1e-6), dim=1, keepdim=True)
        loss_ce = criterion(outputs[:label_bs], target_label)
        loss_dice = dice_loss(outputs_soft[:label_bs], target_label.unsqueeze(1))
        supervised_loss = 0.5 * (loss_dice + loss_ce)
        consistency_weight = ref.get_current_consistency_weight(epoch=cur_itrs // 150, args=args)
        consistency_dist = ref.softmax_mse_loss(outputs[label_bs:], ema_output)
        threshold = (0.75 + 0.25 * ref.sigmoid_rampup(cur_itrs, args.total_itrs)) * np.log(2)
        mask = (uncertainty < threshold).float()
        consistency_loss = torch.sum(mask * consistency_dist) / (2 * torch.sum(mask) + 1e-16)
        loss = supervised_loss + consistency_weight * consistency_loss
        optimizer.zero_grad()
        loss.backward()
        lr = optimizer.param_groups[0]["lr"]
        rec = dict(loss=loss.item(), sup=supervised_loss.item(), cons=consistency_loss.item(), w=consistency_weight, lr=lr,
                   threshold=float(threshold), mask_sum=mask.sum().item(), uncertainty=summarize(uncertainty, stride=1009),
                   logits=summarize(outputs, stride=1009), teacher_logits=summarize(ema_output, stride=1009), mc_logits=mc_fp,
                   grads=grads_of(model))
        optimizer.step()
        lr_scheduler.step()
        ref.update_ema_variables(model, ema_model, args.ema_decay, cur_itrs)
        rec["student_after"], rec["teacher_after"] = after_of(model), after_of(ema_model)
        recs.append(rec)
        print("  it %d loss %.8f sup %.8f cons %.6e mask %.0f" % (cur_itrs, rec["loss"], rec["sup"], rec["cons"], rec["mask_sum"]),
              flush=True)
    return recs


RUNNERS = {"mt": mean_teacher, "cps": cps, "uamt": uamt}

if __name__ == "__main__":
    want = sys.argv[1:] or list(FULL_CONFIGS)
    for tag in want:
        c = FULL_CONFIGS[tag]
        print("full_%s: %s" % (tag, c), flush=True)
        DTYPE = torch.float32
        steps = RUNNERS[c["kind"]](c)
        # fp64 pass of the same iterations: the gradients / loss the fp32 reference is itself an approximation of.  At these
        # sizes the reference's fp32 gradients differ from the fp64 ones by 1e-3 .. 7e-3 (rel-L2 per parameter: sums over 1.2 M
        # pixels behind 18 train-mode BatchNorms cancel heavily), so an fp32 implementation is judged against the fp64 values
        # with the reference's own fp32 error as the yardstick (tests/test_gpu_full_size.py).
        DTYPE = torch.float64
        print("  fp64 pass", flush=True)
        steps64 = RUNNERS[c["kind"]](c)
        for r32, r64 in zip(steps, steps64):
            for k in list(r64):
                if k.startswith("grads") or k == "loss":
                    r32[k + "64"] = r64[k]
        out = {"cfg": c, "steps": steps, "torch": torch.__version__}
        torch.save(out, os.path.join(HERE, "full_%s.pt" % tag))
        print("  -> %.0f KB" % (os.path.getsize(os.path.join(HERE, "full_%s.pt" % tag)) / 1024), flush=True)
