"""Host-side cost of enqueuing one Mean-Teacher step (python + ctypes + CUDA launches), measured with an empty GPU
queue: python profiles/cpu_enqueue.py"""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hpfg_b200 as hb

dev = torch.device("cuda:0")
torch.manual_seed(0)
student = hb.UNet(1, 4, precision="bf16").to(dev)
teacher = copy.deepcopy(student)
step = hb.MeanTeacherStep(student, teacher)
step.enable_graph(bool(int(os.environ.get("GRAPH", "1"))))
x = torch.rand(32, 1, 224, 224, device=dev)
y = torch.randint(0, 4, (8, 224, 224), device=dev)
for _ in range(5):
    step.step(x, y)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step.step(x, y)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t0))
print("enqueue ms (median): %.3f   enqueue+drain ms: %.3f" % (sorted(t[0] for t in ts)[5] * 1e3, sorted(t[1] for t in ts)[5] * 1e3))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step.step(x, y)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
