"""dev: which GroupNorm convolutions of a 512x512 frame are fused / not fused"""
import os, sys, torch
os.environ["OTVM_CUDA_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from otvm_b200 import ops
from otvm_b200.fixtures import make_frame
torch.set_grad_enabled(False)
m = bench.build("bf16")
kw = dict(last_frame=False, memorize=True, max_memory_num=8)
fr = [tuple(t.cuda() for t in make_frame(0, i, 512, 512)) for i in range(2)]
m(*fr[0], first_frame=True, **kw)
orig = ops.conv2d
log = []
def spy(x, w, b, out, **k):
    r = orig(x, w, b, out, **k)
    if k.get("gn_fuse") is not None or k.get("gn_stats") is not None:
        log.append((bool(r is True), tuple(x.shape[1:]), tuple(w.shape), k.get("stride", 1), k.get("dil", 1), k.get("res") is not None, k.get("gn_fuse") is not None))
    return r
ops.conv2d = spy
m(*fr[1], first_frame=False, **kw)
torch.cuda.synchronize()
for l in log: print(l)
print("fused", sum(1 for l in log if l[0]), "of", len(log))
