"""dev/profiling tool: warm up, then run N steady-state 512x512 frames eagerly between cudaProfilerStart/Stop
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/one_frame.py"""
import os, sys, torch
os.environ["OTVM_CUDA_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from otvm_b200.fixtures import make_frame
torch.set_grad_enabled(False)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
size = int(sys.argv[3]) if len(sys.argv) > 3 else 512
m = bench.build(prec)
kw = dict(last_frame=False, memorize=True, max_memory_num=8)
fr = [tuple(t.cuda() for t in make_frame(0, i, size, size)) for i in range(4)]
m(*fr[0], first_frame=True, **kw)
for i in range(1, 9): m(*fr[i % 4], first_frame=False, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(n): m(*fr[i % 4], first_frame=False, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
