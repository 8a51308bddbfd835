"""dev tool: per-conv-shape device time of one steady-state 512x512 frame (CUDA events, eager launches)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from otvm_b200 import ops
from otvm_b200.fixtures import make_frame
torch.set_grad_enabled(False)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
m = bench.build(prec)
kw = dict(last_frame=False, memorize=True, max_memory_num=8)
fr = [tuple(t.cuda() for t in make_frame(0, i, 512, 512)) for i in range(4)]
m(*fr[0], first_frame=True, **kw)
for i in range(1, 10): m(*fr[i % 4], first_frame=False, **kw)
ops.PROFILE_SHAPES = True; ops.PROFILER = ops.Profiler()
n = 3
for i in range(n):
    torch.cuda._sleep(50_000_000)
    m(*fr[i % 4], first_frame=False, **kw)
    torch.cuda.synchronize()
s = ops.PROFILER.summary(); ops.PROFILER = None
tot = sum(v["ms"] for v in s.values())
print(f"total {tot/n:.3f} ms/frame")
for k, v in sorted(s.items(), key=lambda kv: -kv[1]["ms"])[:60]:
    tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["flops"] else 0
    print(f"{v['ms']/n*1e3:9.1f} us/frame  x{v['calls']//n:3d}  {tf:7.1f} TF/s  {v['bytes']/max(v['ms'],1e-9)/1e6:8.1f} GB/s  {k}")
