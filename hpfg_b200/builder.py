"""``build_model(args)`` -- the reference's model factory (model/builder.py:14-62) for the path this
library accelerates.  ``args.model == 'unet'`` (model/builder.py:29-30) and ``'unet_plus'`` (:31-32) are served; every
other name raises ``NotImplementedError`` exactly like the reference's final ``else`` branch (:59-60)."""
from .unet import UNet, UNet_Plus


def build_model(args):
    if args.model == 'unet':
        precision = getattr(args, "precision", None) or (args.get("precision", "bf16") if hasattr(args, "get") else "bf16")
        model = UNet(in_channels=args.in_channels, num_classes=args.num_classes, precision=precision)
    elif args.model == 'unet_plus':
        precision = getattr(args, "precision", None) or (args.get("precision", "bf16") if hasattr(args, "get") else "bf16")
        model = UNet_Plus(in_channels=args.in_channels, num_classes=args.num_classes, precision=precision)
    else:
        raise NotImplementedError
    return model
