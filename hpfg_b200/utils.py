"""Host-side helpers with the reference's names (utils/utils.py:67-95)."""
import numpy as np
import torch

from . import _lib as L
from .unet import UNet


def sigmoid_rampup(current, rampup_length):
    """Exponential rampup (utils/utils.py:72-79)."""
    if rampup_length == 0:
        return 1.0
    current = np.clip(current, 0.0, rampup_length)
    phase = 1.0 - current / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def linear_rampup(current, rampup_length):
    """utils/utils.py:89-95."""
    assert current >= 0 and rampup_length >= 0
    if current >= rampup_length:
        return 1.0
    return current / rampup_length


def get_current_consistency_weight(epoch, args):
    """utils/utils.py:67-69."""
    return args.consistency * sigmoid_rampup(epoch, args.consistency_rampup)


def ema_update_flat(ema_flat, param_flat, alpha):
    """One 128-bit vectorised pass over the flat buffers: ema <- alpha*ema + (1-alpha)*param."""
    L.require_cuda(ema_flat, "ema buffer")
    assert ema_flat.dtype == torch.float32 and param_flat.dtype == torch.float32
    assert ema_flat.is_contiguous() and param_flat.is_contiguous() and ema_flat.numel() == param_flat.numel()
    L.check(L.lib().hpfg_ema_update(L.ptr(ema_flat), L.ptr(param_flat), ema_flat.numel(), float(alpha),
                                    L.stream_ptr(ema_flat.device)), "hpfg_ema_update")


def update_ema_variables(model, ema_model, alpha, global_step):
    """``update_ema_variables(model, ema_model, alpha, global_step)`` (utils/utils.py:82-86): parameters only,
    BN buffers untouched.  Two hpfg_b200 UNets -> a single pass over their flat parameter buffers; any other
    pair of CUDA modules -> one launch per parameter tensor."""
    alpha = min(1 - 1 / (global_step + 1), alpha)
    pairs = zip(ema_model.parameters(), model.parameters())
    if isinstance(model, UNet) and isinstance(ema_model, UNet):
        ema_update_flat(ema_model.ensure_flat(), model.ensure_flat(), alpha)
        # UNet_Plus: the projection-neck parameters live outside the flat buffer (they follow the 82 U-Net parameters)
        core = {id(q) for q in ema_model._flat_params_list}
        pairs = [(e, p) for e, p in zip(ema_model.parameters(), model.parameters()) if id(e) not in core]
    for ema_param, param in pairs:
        e, p = ema_param.data, param.data
        if not (e.is_contiguous() and p.is_contiguous() and e.data_ptr() % 16 == 0 and p.data_ptr() % 16 == 0):
            raise L.HpfgError("update_ema_variables: parameters must be contiguous, 16-byte aligned CUDA fp32 tensors")
        ema_update_flat(e.view(-1), p.view(-1), alpha)
