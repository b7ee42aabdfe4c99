"""Seeded input builders shared by make_golden.py (run against the real reference) and the tests (run
against the oracle / the CUDA path).  torch's CPU generator is deterministic for a given torch version,
so only OUTPUTS are stored in the fixtures; inputs are regenerated from these seeds on both sides."""
import torch

from oracle.unet_ref import ENC_DROPOUT, FT_CHNS, init_unet_state

ENC_PREFIXES = ["encoder.in_conv"] + ["encoder.down%d.maxpool_conv.1" % i for i in range(1, 5)]


def gen(seed):
    return torch.Generator().manual_seed(seed)


def make_state(in_ch, n_cls, seed):
    st = init_unet_state(in_ch, n_cls, generator=gen(seed))
    g = gen(seed + 1)
    for k in st:                      # non-trivial BN affine so gamma/beta gradients are exercised
        if k.endswith(".conv_conv.1.weight") or k.endswith(".conv_conv.5.weight"):
            st[k] = 0.5 + torch.rand(st[k].shape, generator=g)
        elif k.endswith(".conv_conv.1.bias") or k.endswith(".conv_conv.5.bias"):
            st[k] = 0.2 * (torch.rand(st[k].shape, generator=g) - 0.5)
    return st


def make_masks(n, h, w, seed):
    """Keep-masks (bool, NCHW) for the five encoder ConvBlocks' dropout (model/unet.py:161)."""
    g = gen(seed)
    masks = {}
    for lvl, (prefix, p) in enumerate(zip(ENC_PREFIXES, ENC_DROPOUT)):
        shape = (n, FT_CHNS[lvl], h >> lvl, w >> lvl)
        masks[prefix] = torch.rand(shape, generator=g) >= p
    return masks


def make_batch(n_l, n_u, in_ch, n_cls, h, w, seed, ignore_frac=0.0):
    g = gen(seed)
    x_l = torch.rand(n_l, in_ch, h, w, generator=g)
    x_u = torch.rand(n_u, in_ch, h, w, generator=g)
    y = torch.randint(0, n_cls, (n_l, h, w), generator=g)
    if ignore_frac > 0:
        y[torch.rand(y.shape, generator=g) < ignore_frac] = 255
    return x_l, x_u, y


def make_plus_state(in_ch, n_cls, seed):
    """make_state + the 16 projection-neck parameters of UNet_Plus (uniform +-1/sqrt(fan_in), seeded)."""
    from oracle.unet_ref import unet_plus_neck_spec
    st = make_state(in_ch, n_cls, seed)
    g = gen(seed + 2)
    for name, shape in unet_plus_neck_spec(n_cls):
        fan_in = shape[1] if len(shape) > 1 else shape[0]
        b = 1.0 / (fan_in ** 0.5)
        st[name] = (2 * torch.rand(shape, generator=g) - 1) * b
    return st


def make_predict_case(in_ch, n_cls, n, h, w, seed):
    """State + volume for the inference fixture: the output conv is scaled up and the slices are smooth blobs so that a
    random-init network predicts a mix of classes (BN running buffers come from the fixture)."""
    st = make_state(in_ch, n_cls, seed)
    wo = st["decoder.out_conv.weight"]
    st["decoder.out_conv.weight"] = (wo - wo.mean(dim=(1, 2, 3), keepdim=True)) * 40.0
    st["decoder.out_conv.bias"] = torch.zeros_like(st["decoder.out_conv.bias"])
    g = gen(seed + 50)
    coarse = torch.rand(n, 1, 4, 6, generator=g)
    vol = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)[:, 0]
    vol = (vol + 0.05 * torch.rand(n, h, w, generator=g)).contiguous()
    return st, vol


def summarize(t, stride=97, full_below=20000):
    """Compact fingerprint of a tensor: full copy if small, else strided sample + sums (fp64)."""
    t = t.detach().double().flatten()
    d = {"numel": t.numel(), "sum": t.sum().item(), "abs_sum": t.abs().sum().item(),
         "sq_sum": (t * t).sum().item()}
    if t.numel() <= full_below:
        d["full"] = t.float().clone()
    else:
        d["sample"] = t[::stride].float().clone()
        d["stride"] = stride
    return d


def make_cutmix_masks(n, h, w, seed):
    """Seeded stand-in for BoxMaskGenerator.generate_params (main.py:92-97,140-142): [n,1,h,w] float masks that are 1
    outside one random box and 0 inside (invert=True convention)."""
    g = gen(seed)
    m = torch.ones(n, 1, h, w)
    for i in range(n):
        y0, x0 = int(torch.randint(0, h // 2, (1,), generator=g)), int(torch.randint(0, w // 2, (1,), generator=g))
        hh, ww = int(torch.randint(h // 4, h // 2, (1,), generator=g)), int(torch.randint(w // 4, w // 2, (1,), generator=g))
        m[i, :, y0:y0 + hh, x0:x0 + ww] = 0.0
    return m


def hpfg_main_step(ns, model1, model2, ema_model, optimizer1, optimizer2, lr_scheduler1, lr_scheduler2, batch, cur_itrs,
                   args, device):
    """main.py:128-207 (the HPFG iteration) transcribed once and run with EITHER the reference's objects (make_golden.py,
    CPU) OR hpfg_b200's drop-ins (tests, GPU): ``ns`` supplies DiceLoss / Dense_Loss / update_ema_variables /
    linear_rampup, everything else is the caller code of the reference, unchanged."""
    import torch.nn as nn
    label_img, target_label, label_img1, target_label1, img_unlabel, cutmix_mask = batch
    criterion = nn.CrossEntropyLoss(ignore_index=255)
    dice_loss = ns.DiceLoss(args.num_classes)
    dense_loss = ns.Dense_Loss(args.batch_size + args.unlabel_batch_size, device)
    label_bs = label_img.shape[0]
    unlabel_bs = img_unlabel.shape[0]
    label_img = label_img.to(device).float()
    img_unlabel = img_unlabel.to(device).float()
    label_img1 = label_img1.repeat(int(unlabel_bs // label_bs), 1, 1, 1).to(device).float()
    target_label1 = target_label1.repeat(int(unlabel_bs // label_bs), 1, 1).to(device).long()
    target_label = target_label.to(device).long()
    cutmix_mask = cutmix_mask.clone().float().to(device)
    batch_un_mix = label_img1 * (1.0 - cutmix_mask) + img_unlabel * cutmix_mask
    batch_mix = torch.cat([label_img, batch_un_mix], dim=0).to(device).float()
    outputs1, _, _ = model1(batch_mix)
    outputs_soft1 = torch.softmax(outputs1, dim=1)
    volume_batch = torch.cat([label_img, img_unlabel], dim=0).to(device).float()
    outputs2, h1, h2 = model2(volume_batch)
    outputs_soft2 = torch.softmax(outputs2, dim=1)
    with torch.no_grad():
        ema_output, ema_h1, ema_h2 = ema_model(volume_batch)
        ema_output_soft = torch.softmax(ema_output.detach(), dim=1)
    loss1 = 0.5 * (criterion(outputs1[:label_bs], target_label) + dice_loss(outputs_soft1[:label_bs],
                                                                            target_label.unsqueeze(1)))
    loss2 = 0.5 * (criterion(outputs2[:label_bs], target_label) + dice_loss(outputs_soft2[:label_bs],
                                                                            target_label.unsqueeze(1)))
    loss_sup = loss1 + loss2
    loss_constrivate = dense_loss(h1, ema_h1) + dense_loss(h2, ema_h2)
    cutmix_mask = cutmix_mask.squeeze(1)
    pseudo_outputs1 = torch.argmax(ema_output_soft[label_bs:], dim=1, keepdim=False)
    pseudo_outputs1 = target_label1 * (1.0 - cutmix_mask) + pseudo_outputs1 * cutmix_mask
    pseudo_supervision1 = dice_loss(outputs_soft1[label_bs:], pseudo_outputs1.unsqueeze(1))
    consistency_weight_cps = args.consistency * ns.linear_rampup(cur_itrs // 150, args.consistency_rampup)
    consistency_weight_mt = args.consistency * ns.linear_rampup(cur_itrs // 150, args.consistency_rampup)
    if cur_itrs < 1000:
        consistency_loss1 = 0.0
        consistency_loss2 = 0.0
    else:
        consistency_loss1 = 0.0
        consistency_loss2 = torch.mean((outputs_soft2[label_bs:] - ema_output_soft[label_bs:]) ** 2)
    model1_loss = 7 * consistency_weight_cps * pseudo_supervision1 + consistency_weight_mt * consistency_loss1
    model2_loss = consistency_weight_mt * consistency_loss2 + consistency_weight_mt * loss_constrivate
    loss_semi = model1_loss + model2_loss
    loss = loss_sup + loss_semi
    optimizer1.zero_grad()
    optimizer2.zero_grad()
    loss.backward()
    optimizer1.step()
    optimizer2.step()
    # update_ema_variables_backbone(model1, model2, ...) -- main.py:68-76: model2's backbone tracks model1's
    alpha = min(1 - 1 / (cur_itrs + 1), args.ema_decay)
    for ema_param, param in zip(model2.encoder.parameters(), model1.encoder.parameters()):
        ema_param.data.mul_(alpha).add_(param.data, alpha=1 - alpha)
    for ema_param, param in zip(model2.decoder.parameters(), model1.decoder.parameters()):
        ema_param.data.mul_(alpha).add_(param.data, alpha=1 - alpha)
    ns.update_ema_variables(model2, ema_model, args.ema_decay, cur_itrs)
    lr_scheduler1.step()
    lr_scheduler2.step()
    return dict(loss=loss.item(), loss_sup=loss_sup.item(), loss_semi=loss_semi.item(),
                contrast=loss_constrivate.item(), pseudo=pseudo_supervision1.item(),
                cons2=float(consistency_loss2.detach()) if torch.is_tensor(consistency_loss2) else 0.0, outputs1=outputs1.detach(), outputs2=outputs2.detach(),
                ema_output=ema_output.detach())


def make_hpfg_batch(n_l, n_u, in_ch, n_cls, h, w, seed):
    x_l, x_u, y = make_batch(n_l, n_u, in_ch, n_cls, h, w, seed)
    x_l1, _, y1 = make_batch(n_l, 0, in_ch, n_cls, h, w, seed + 1)
    return x_l, y, x_l1, y1, x_u, make_cutmix_masks(n_u, h, w, seed + 2)


# ---- BASELINE.json shapes (tests/golden/make_golden_full.py -> full_<tag>.pt; tests/test_gpu_full_size.py)
FULL_CONFIGS = {
    "mt_cfg1": dict(kind="mt", in_ch=1, n_cls=4, n_l=12, n_u=12, h=224, w=224, seed=1100, steps=1),
    "mt_cfg2": dict(kind="mt", in_ch=1, n_cls=4, n_l=8, n_u=24, h=224, w=224, seed=1200, steps=2),
    "cps": dict(kind="cps", in_ch=1, n_cls=4, n_l=8, n_u=24, h=224, w=224, seed=1300, steps=1),
    "uamt": dict(kind="uamt", in_ch=1, n_cls=4, n_l=12, n_u=12, h=224, w=224, seed=1400, steps=1, T=8, teacher_gain=40.0),
    "mt_isic": dict(kind="mt", in_ch=3, n_cls=2, n_l=12, n_u=12, h=224, w=224, seed=1500, steps=1),
}


def make_uamt_noise(n_u, in_ch, h, w, T, seed):
    """The clamped Gaussian perturbations of 2019_07_MICCAI_Uncertainty_Aware_ACDC.py:130,141, seeded: one for the
    teacher's consistency forward [n_u,...] and T//2 for the Monte-Carlo forwards over the batch repeated twice."""
    g = gen(seed)
    noise = torch.clamp(torch.randn(n_u, in_ch, h, w, generator=g) * 0.1, -0.2, 0.2)
    mc = torch.clamp(torch.randn(T // 2, 2 * n_u, in_ch, h, w, generator=g) * 0.1, -0.2, 0.2)
    return noise, mc


def make_uamt_teacher_state(in_ch, n_cls, seed, gain):
    """Teacher weights for the UAMT fixture: a random-init 4-class network is maximally uncertain everywhere (entropy ~ ln 4 >
    the 0.75 ln 2 threshold of 2019_07...:158), which would leave the masked consistency term identically zero; the
    output conv is centred and scaled by ``gain`` so that part of the pixels falls under the threshold."""
    st = make_state(in_ch, n_cls, seed)
    wo = st["decoder.out_conv.weight"]
    st["decoder.out_conv.weight"] = (wo - wo.mean(dim=(1, 2, 3), keepdim=True)) * gain
    return st
