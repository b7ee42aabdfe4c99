"""Does a high-priority main stream (data-gradient chain) over the plan's priority-0 weight-gradient stream change the step?
Eager launches and graph replay, 1 GPU: python profiles/ab_priority.py"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hpfg_b200 as hb

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1337)
x = torch.rand(32, 1, 224, 224, generator=g).to(dev)
y = torch.randint(0, 4, (8, 224, 224), generator=g).to(dev)


def run(prio, graph):
    torch.manual_seed(1337)
    s = hb.UNet(1, 4, precision="bf16").to(dev)
    t = copy.deepcopy(s)
    step = hb.MeanTeacherStep(s, t)
    step.enable_graph(graph)
    st = torch.cuda.Stream(device=dev, priority=prio) if prio is not None else torch.cuda.current_stream(dev)
    if prio is not None:
        step._side = torch.cuda.Stream(device=dev, priority=prio)      # the teacher's forward stream at the same priority
    with torch.cuda.stream(st):
        for _ in range(6):
            step.step(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            step.step(x, y)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 30


for graph in (False, True):
    for prio in (None, -1, None, -1):
        print("graph=%s main-stream priority %s: %.3f ms/step" % (graph, prio, run(prio, graph)), flush=True)
