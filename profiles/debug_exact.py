import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import hpfg_b200 as hb
from hpfg_b200 import _lib as L
from tests.golden.common import make_state, make_batch
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
# 1. does the hook all-reduce in place?
L.install_allreduce_hook(None)
t = torch.full((8,), float(rank + 1), device=dev, dtype=torch.float64)
cb = L._hook_keepalive["cb"]
cb(None, t.data_ptr(), 8, 1, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("rank", rank, "hook result", t.tolist()[:2], flush=True)
import oracle
import ctypes
calls = []
def verbose_install(group=None):
    world = dist.get_world_size(group)
    def _hook(ctx, ptr, count, is_double, stream):
        d = torch.cuda.current_device()
        t = torch.as_tensor(L._DevBuf(ptr, count, is_double), device=torch.device("cuda", d))
        es = torch.cuda.ExternalStream(stream or 0, device=d)
        if len(calls) < 0:
            es.synchronize(); before = t[:2].tolist()
        if os.environ.get("HSYNC") == "2": torch.cuda.synchronize()
        with torch.cuda.stream(es):
            dist.all_reduce(t, group=group)
        if os.environ.get("HSYNC"): torch.cuda.synchronize()
        if len(calls) < 0:
            es.synchronize(); print("rank", rank, "hook call", len(calls), "ptr %x" % ptr, "count", count, "dbl", is_double, "stream", stream, "before", before, "after", t[:2].tolist(), "t.ptr %x" % t.data_ptr(), flush=True)
        calls.append(count)
    cb = L.ALLREDUCE_FN(_hook)
    L._hook_keepalive["cb"] = cb
    L.check(L.lib().hpfg_set_allreduce_hook(ctypes.cast(cb, ctypes.c_void_p), None, world))
    return world
L.install_allreduce_hook = verbose_install
IN_CH, N_CLS, H, W, N_L, N_U = 1, 4, 32, 32, 2, 2
m = hb.UNet(IN_CH, N_CLS, precision="fp32"); m.load_state_dict(make_state(IN_CH, N_CLS, 3)); m.to(dev); m.set_dropout_enabled(False)
t_ = copy.deepcopy(m); t_.set_dropout_enabled(False)
step = hb.MeanTeacherStep(m, t_, exact_global=True)
step.serialize = bool(int(os.environ.get("SER", "0")))
x_l, x_u, y = make_batch(2 * N_L, 2 * N_U, IN_CH, N_CLS, H, W, 21)
xl, xu, yy = hb.shard_batch(x_l, x_u, y, rank, world)
loss = step.step(torch.cat([xl, xu]).to(dev), yy.to(dev))
sc = step.last["scalars"].cpu().tolist()
sa = make_state(IN_CH, N_CLS, 3); sb = {k: v.clone() for k, v in sa.items()}
r = oracle.mt_step(sa, sb, oracle.SGDState(), x_l, x_u, y, 1, student_masks={}, teacher_masks={})
sa2 = make_state(IN_CH, N_CLS, 3); sb2 = {k: v.clone() for k, v in sa2.items()}
rl = oracle.mt_step(sa2, sb2, oracle.SGDState(), xl, xu, yy, 1, student_masks={}, teacher_masks={})
print("rank", rank, "ours", sc[:6], "| oracle global loss %.7f sup %.7f cons %.3e | oracle local loss %.7f sup %.7f" % (r["loss"], r["loss_sup"], r["loss_cons"], rl["loss"], rl["loss_sup"]), flush=True)
lg = step.last["logits"].cpu()
idx = list(range(rank * N_L, (rank + 1) * N_L)) + [2 * N_L + i for i in range(rank * N_U, (rank + 1) * N_U)]
print("rank", rank, "logits vs global oracle", float((lg - r["logits"][idx]).norm() / r["logits"][idx].norm()), "vs local oracle", float((lg - rl["logits"]).norm() / rl["logits"].norm()), flush=True)
ost = make_state(IN_CH, N_CLS, 3)
_, taps = oracle.unet_forward(ost, torch.cat([x_l, x_u]), True, {}, return_taps=True)
for k in [kk for kk in taps if kk.endswith("conv_conv.0") or kk.endswith("conv_conv.4")]:
    tt = taps[k][idx]
    got = m.debug_tap(k, (4, 1, H, W))[:tt.numel()].view(tt.shape).cpu() + ost[k + ".bias"].view(1, -1, 1, 1)
    print("rank", rank, "tap", k, "rel err vs global oracle", float((got - tt).norm() / tt.norm()), flush=True)
rm = m.state_dict()["encoder.in_conv.conv_conv.1.running_mean"].cpu()
print("rank", rank, "running_mean[:4] ours", rm[:4].tolist(), "global oracle", sa["encoder.in_conv.conv_conv.1.running_mean"][:4].tolist(), "local oracle", sa2["encoder.in_conv.conv_conv.1.running_mean"][:4].tolist(), flush=True)
print("rank", rank, "hook calls", len(calls), calls[:8], flush=True)
dist.barrier(); dist.destroy_process_group()
