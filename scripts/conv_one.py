"""dev/profiling tool: a few conv shapes of the 512x512 frame in an element format (bf16 | bf16x2 | bf16x3), each launched
once after warm-up (`ncu --set full -k regex:conv_tc`), or (argument `gn`) the GroupNorm-apply kernel on the frame's big
tensors.   python scripts/conv_one.py [bf16x2] [gn]"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops
from otvm_b200.split import SplitArena, split_planes
fmt = next((a for a in sys.argv[1:] if a.startswith("bf16")), "bf16x2")
planes = {"bf16": 1, "bf16x2": 2, "bf16x3": 3}[fmt]
ar = SplitArena(planes, 1 << 30, "cuda")
def act(shape, rand=True):
    t = ar.alloc(shape)
    if rand: ar.write(t, torch.randn(shape, device="cuda"))
    return t
if "gn" in sys.argv[1:]:
    for C, H, W in [(64, 512, 512), (256, 128, 128), (2048, 64, 64), (256, 64, 64)]:
        x = act((1, H, W, C)); st = torch.zeros(64, dtype=torch.float64, device="cuda")
        g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
        ops.gn_stats(x, st)
        for _ in range(2): ops.gn_apply(x, st, g, b, x, act=ops.ACT_RELU)
    torch.cuda.synchronize(); sys.exit(0)
# (Cin, Cout, k, dil, H, W): the heaviest layer, a full-resolution refine conv (persistent kernel), two latency-bound small
# grids, a 1x1 with one K iteration, the PPM-fed decoder conv, a wide 1x1
shapes = [(256, 256, 3, 1, 128, 128), (64, 64, 3, 1, 512, 512), (256, 256, 3, 1, 32, 32), (64, 256, 1, 1, 128, 128),
          (3072, 256, 3, 1, 64, 64), (512, 2048, 1, 1, 64, 64)]
ws = torch.empty(16 << 20, device="cuda")
for Cin, Cout, k, d, H, W in shapes:
    x = act((1, H, W, Cin))
    w = torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)
    w = split_planes(w, planes) if planes > 1 else w.bfloat16()
    out = act((1, H, W, Cout), rand=False); b = torch.zeros(Cout, device="cuda")
    for _ in range(2): ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d, workspace=ws)
torch.cuda.synchronize()
