"""Functional fp32 restatement of the reference UNet (model/unet.py:12-117,155-175).

State is a flat ``dict[str, Tensor]`` with exactly the reference ``state_dict`` keys, so a real reference
``UNet(...).state_dict()`` can be fed in unchanged.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

FT_CHNS = [16, 32, 64, 128, 256]            # model/unet.py:160
ENC_DROPOUT = [0.05, 0.1, 0.2, 0.3, 0.5]    # model/unet.py:161 (decoder blocks use 0.0, :94-97)
BN_EPS = 1e-5                               # nn.BatchNorm2d default
BN_MOMENTUM = 0.1
LEAKY_SLOPE = 0.01                          # nn.LeakyReLU() default (model/unet.py:20,24)


def _convblock_names(prefix):
    # nn.Sequential indices of model/unet.py:17-25: 0 conv, 1 bn, 2 lrelu, 3 dropout, 4 conv, 5 bn, 6 lrelu
    return [prefix + ".conv_conv.0", prefix + ".conv_conv.1", prefix + ".conv_conv.4", prefix + ".conv_conv.5"]


def _blocks(in_channels):
    """[(block_prefix, cin, cout, dropout_p)] in registration order (model/unet.py:70-74,94-97)."""
    f = FT_CHNS
    enc = [("encoder.in_conv", in_channels, f[0], ENC_DROPOUT[0])]
    for i in range(1, 5):
        enc.append(("encoder.down%d.maxpool_conv.1" % i, f[i - 1], f[i], ENC_DROPOUT[i]))
    dec = []
    for i in range(1, 5):
        c1, c2 = f[5 - i], f[4 - i]           # UpBlock(in_channels1, in_channels2, out)  :94-97
        dec.append(("decoder.up%d" % i, c1, c2))
    return enc, dec


def unet_param_spec(in_channels=1, num_classes=4):
    """Ordered (name, shape) of the 82 parameters, == list(reference UNet.named_parameters())."""
    enc, dec = _blocks(in_channels)
    spec = []

    def convblock(prefix, cin, cout):
        c0, b0, c1, b1 = _convblock_names(prefix)
        spec.extend([(c0 + ".weight", (cout, cin, 3, 3)), (c0 + ".bias", (cout,)),
                     (b0 + ".weight", (cout,)), (b0 + ".bias", (cout,)),
                     (c1 + ".weight", (cout, cout, 3, 3)), (c1 + ".bias", (cout,)),
                     (b1 + ".weight", (cout,)), (b1 + ".bias", (cout,))])

    for prefix, cin, cout, _ in enc:
        convblock(prefix, cin, cout)
    for prefix, c1, c2 in dec:
        spec.extend([(prefix + ".conv1x1.weight", (c2, c1, 1, 1)), (prefix + ".conv1x1.bias", (c2,))])
        convblock(prefix + ".conv", 2 * c2, c2)
    spec.extend([("decoder.out_conv.weight", (num_classes, FT_CHNS[0], 3, 3)),
                 ("decoder.out_conv.bias", (num_classes,))])
    return spec


def unet_buffer_spec(in_channels=1, num_classes=4):
    """Ordered (name, shape, dtype) of the 54 BatchNorm buffers."""
    enc, dec = _blocks(in_channels)
    spec = []
    for prefix, cout in [(p, co) for p, _, co, _ in enc] + [(p + ".conv", c2) for p, _, c2 in dec]:
        _, b0, _, b1 = _convblock_names(prefix)
        for b in (b0, b1):
            spec.extend([(b + ".running_mean", (cout,), torch.float32),
                         (b + ".running_var", (cout,), torch.float32),
                         (b + ".num_batches_tracked", (), torch.int64)])
    return spec


def init_unet_state(in_channels=1, num_classes=4, generator=None):
    """Random state with the reference's *distributions* (kaiming_uniform(a=sqrt 5) conv weights,
    U(-1/sqrt(fan_in), +) biases, BN gamma=1 beta=0).  NOT the same RNG stream as constructing the
    reference module; tests that need identical init build the real module or use golden fixtures."""
    st = OrderedDict()
    for name, shape in unet_param_spec(in_channels, num_classes):
        if len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / fan_in ** 0.5
            st[name] = (torch.rand(shape, generator=generator) * 2 - 1) * bound
            last_bound = bound
        elif ".conv_conv.1." in name or ".conv_conv.5." in name:
            st[name] = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
        else:
            st[name] = (torch.rand(shape, generator=generator) * 2 - 1) * last_bound
    for name, shape, dt in unet_buffer_spec(in_channels, num_classes):
        st[name] = torch.ones(shape, dtype=dt) if name.endswith("running_var") else torch.zeros(shape, dtype=dt)
    return st


def _bn(st, name, x, training):
    # nn.BatchNorm2d forward: batch statistics in training (biased var for normalisation, unbiased for the
    # running estimate), running statistics in eval; num_batches_tracked += 1 per training forward.
    if training:
        st[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, st[name + ".running_mean"], st[name + ".running_var"], st[name + ".weight"],
                        st[name + ".bias"], training, BN_MOMENTUM, BN_EPS)


def _conv_block(st, prefix, x, p, training, masks, taps):
    """ConvBlock.forward (model/unet.py:12-28)."""
    c0, b0, c1, b1 = _convblock_names(prefix)
    x = F.conv2d(x, st[c0 + ".weight"], st[c0 + ".bias"], padding=1)
    taps[c0] = x
    x = F.leaky_relu(_bn(st, b0, x, training), LEAKY_SLOPE)
    if masks is not None and prefix in masks:          # explicit keep-mask (1 = keep), scale 1/(1-p)
        x = x * masks[prefix] / (1.0 - p)
    else:
        if training and p > 0.0 and masks is None:
            ones = torch.ones_like(x)
            kept = F.dropout(ones, p, True)            # consumes torch's RNG exactly like nn.Dropout
            taps[prefix + ".dropout_keep"] = (kept != 0)
            x = x * kept
    taps[prefix + ".mid"] = x
    x = F.conv2d(x, st[c1 + ".weight"], st[c1 + ".bias"], padding=1)
    taps[c1] = x
    x = F.leaky_relu(_bn(st, b1, x, training), LEAKY_SLOPE)
    taps[prefix] = x
    return x


def unet_forward(st, x, training=True, dropout_masks=None, return_taps=False):
    """UNet.forward (model/unet.py:172-175 -> Encoder :76-82 -> Decoder :101-117).

    dropout_masks: None -> draw with torch's global RNG as nn.Dropout would (p=0 blocks draw nothing);
                   dict {block_prefix: bool/float keep mask [N,C,H,W]} -> use these (blocks absent from
                   the dict get no dropout).
    BN running buffers in ``st`` are updated in place when training (as the reference module does).
    """
    in_channels = st["encoder.in_conv.conv_conv.0.weight"].shape[1]
    enc, dec = _blocks(in_channels)
    taps = OrderedDict()
    feats = []
    h = x
    for i, (prefix, _, _, p) in enumerate(enc):
        if i > 0:
            h = F.max_pool2d(h, 2)                     # DownBlock (model/unet.py:31-42)
        h = _conv_block(st, prefix, h, p, training, dropout_masks, taps)
        feats.append(h)
    taps["feature4"] = feats[4]                        # feature[-1]: what UNet_Plus feeds its high projection neck (:201)
    h = feats[4]
    for i, (prefix, _, _) in enumerate(dec):           # UpBlock.forward (model/unet.py:53-58)
        skip = feats[3 - i]
        h = F.conv2d(h, st[prefix + ".conv1x1.weight"], st[prefix + ".conv1x1.bias"])
        taps[prefix + ".conv1x1"] = h
        h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=True)
        h = torch.cat([skip, h], dim=1)
        taps[prefix + ".cat"] = h
        h = _conv_block(st, prefix + ".conv", h, 0.0, training, dropout_masks, taps)
    out = F.conv2d(h, st["decoder.out_conv.weight"], st["decoder.out_conv.bias"], padding=1)   # :99,116
    return (out, taps) if return_taps else out


# ---- UNet_Plus (model/unet.py:120-152,178-206) -----------------------------------------------------------------------
NECKS = (("dense_projection_high", FT_CHNS[-1], 2048), ("dense_projection_head", None, 1024))   # (name, in_dim, hid_dim)


def unet_plus_neck_spec(num_classes, out_dim=128):
    """[(name, shape)] of the 16 projection-neck parameters in registration order (after the 82 U-Net parameters)."""
    spec = []
    for name, in_dim, hid in NECKS:
        in_dim = num_classes if in_dim is None else in_dim
        spec += [(name + ".mlp.0.weight", (hid, in_dim)), (name + ".mlp.0.bias", (hid,)),
                 (name + ".mlp.2.weight", (out_dim, hid)), (name + ".mlp.2.bias", (out_dim,)),
                 (name + ".mlp_conv.0.weight", (hid, in_dim, 1, 1)), (name + ".mlp_conv.0.bias", (hid,)),
                 (name + ".mlp_conv.2.weight", (out_dim, hid, 1, 1)), (name + ".mlp_conv.2.bias", (out_dim,))]
    return spec


def projection_conv(st, prefix, x, s=4):
    """projection_conv.forward (model/unet.py:140-152): (global vector [N,128], dense map [N,128,s*s])."""
    x1 = F.adaptive_avg_pool2d(x, (1, 1)).reshape(x.size(0), -1)
    x1 = F.linear(F.relu(F.linear(x1, st[prefix + ".mlp.0.weight"], st[prefix + ".mlp.0.bias"])),
                  st[prefix + ".mlp.2.weight"], st[prefix + ".mlp.2.bias"])
    xp = F.adaptive_avg_pool2d(x, (s, s)) if s else x
    x2 = F.conv2d(F.relu(F.conv2d(xp, st[prefix + ".mlp_conv.0.weight"], st[prefix + ".mlp_conv.0.bias"])),
                  st[prefix + ".mlp_conv.2.weight"], st[prefix + ".mlp_conv.2.bias"])
    return x1, x2.view(x2.size(0), x2.size(1), -1)


def unet_plus_forward(st, x, training=True, dropout_masks=None):
    """UNet_Plus.forward (model/unet.py:198-204): (logits, high_feature, head_feature)."""
    out, taps = unet_forward(st, x, training, dropout_masks, return_taps=True)
    return out, projection_conv(st, "dense_projection_high", taps["feature4"]), projection_conv(st, "dense_projection_head", out)
