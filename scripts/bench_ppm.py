import sys, torch
sys.path.insert(0, '/root/repo')
from otvm_b200 import ops
from otvm_b200.split import SplitArena
ar = SplitArena(2, 1 << 29, "cuda")
x = ar.alloc((1, 64, 64, 3072)); ar.write(x, torch.randn(1, 64, 64, 3072, device="cuda"))
conv5 = x[..., :2048]
pooled = ar.alloc((50, 2048)); rows = torch.empty(64 * 12 * 2048, device="cuda")
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for flush in (True, False):
    ts = []
    for i in range(12):
        if flush: big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.ppm_pool(conv5, pooled, rows); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("ppm_pool (rows + cells)", "cold L2" if flush else "warm L2", f"{sorted(ts)[len(ts)//2]:.1f} us")
