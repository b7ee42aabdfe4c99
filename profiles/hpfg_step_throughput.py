"""Throughput of main.py's HPFG iteration (HPFGStep: three UNet_Plus networks, CutMix batch, necks + Dense_Loss) at the ACDC
shape, 8 labeled + 24 unlabeled 224x224 per step, bf16 U-Nets, 1 GPU, inputs resident:  python profiles/hpfg_step_throughput.py"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hpfg_b200 as hb
from hpfg_b200 import _lib as L

dev = torch.device("cuda:0")
n_l, n_u, H, W = 8, 24, 224, 224
torch.manual_seed(0)
m1, m2 = hb.UNet_Plus(1, 4).to(dev), hb.UNet_Plus(1, 4).to(dev)
ema = copy.deepcopy(m2)
Step = hb.HPFGStep
if os.environ.get("HPFG_STEP_PREV"):      # A/B: `git show <rev>:hpfg_b200/trainer.py > hpfg_b200/_trainer_prev.py` first
    from hpfg_b200._trainer_prev import HPFGStep as Step
step = Step(m1, m2, ema, weight_decay=0.0005)
step.cur_itrs = 1500                      # Mean-Teacher term on (main.py:178-183)
g = torch.Generator().manual_seed(3)
label_img, img_unlabel = torch.rand(n_l, 1, H, W, generator=g).to(dev), torch.rand(n_u, 1, H, W, generator=g).to(dev)
label_img1 = torch.rand(n_l, 1, H, W, generator=g).to(dev)
y, y1 = torch.randint(0, 4, (n_l, H, W), generator=g).to(dev), torch.randint(0, 4, (n_l, H, W), generator=g).to(dev)
mask = torch.zeros(n_u, 1, H, W)
mask[:, :, 40:150, 60:180] = 1.0
mask = mask.to(dev)


def one():
    return step.step(label_img, y, label_img1, y1, img_unlabel, mask)


for _ in range(5):
    one()
torch.cuda.synchronize()
k0 = L.lib().hpfg_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 20
e0.record()
for _ in range(steps):
    loss = one()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("HPFGStep 8+24 @224 bf16: %.3f ms/step  %.0f images/s  (%d library launches per step, loss %.4f, contrast %.4f)"
      % (ms, (n_l + n_u) / ms * 1e3, (L.lib().hpfg_launch_count() - k0) // steps, loss.item(), step.last["contrast"].item()))
# the necks + Dense_Loss alone (forward + backward of model2's two necks and both contrastive terms)
feat = torch.randn(32, 256, 14, 14, device=dev, requires_grad=True)
logits = torch.randn(32, 4, H, W, device=dev, requires_grad=True)
with torch.no_grad():
    t_hi, t_hd = ema.dense_projection_high(feat), ema.dense_projection_head(logits)
dense = hb.Dense_Loss(32, dev)


def necks():
    l = dense(m2.dense_projection_high(feat), t_hi) + dense(m2.dense_projection_head(logits), t_hd)
    l.backward()


for _ in range(3):
    necks()
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    necks()
e1.record()
torch.cuda.synchronize()
print("model2 necks + 2 x Dense_Loss, forward + backward: %.3f ms" % (e0.elapsed_time(e1) / 20))
