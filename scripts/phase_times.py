"""dev: how much of a steady-state 512x512 / T=8 frame each phase costs ON THE CRITICAL PATH (graph-replayed frames, CUDA
events): the default schedule, memorize inline (OTVM_OVERLAP=0 in a second process), and frames that do not memorize at
all (no Encoder_M / KV_M pass: the bank keeps its 8 frames).   python scripts/phase_times.py [precision]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from otvm_b200.fixtures import make_frame
torch.set_grad_enabled(False)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x2"
m = bench.build(prec)
fr = [tuple(t.cuda() for t in make_frame(0, i, 512, 512)) for i in range(4)]
kw = dict(last_frame=False, memorize=True, max_memory_num=8)
m(*fr[0], first_frame=True, **kw)
for i in range(1, 12): m(*fr[i % 4], first_frame=False, **kw)

def timed(n, **k):
    for i in range(6): m(*fr[i % 4], first_frame=False, **k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): m(*fr[i % 4], first_frame=False, **k)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t_mem = timed(100, **kw)
t_nomem = timed(100, last_frame=False, memorize=False, max_memory_num=8)
print(f"overlap={os.environ.get('OTVM_OVERLAP', '1')} {prec}: memorize=True {t_mem:.3f} ms/frame, memorize=False {t_nomem:.3f} ms/frame, "
      f"memorize pass costs {1e3 * (t_mem - t_nomem):.0f} us on the critical path")
