"""dev/profiling tool: a few bf16 conv shapes of the 512x512 frame, each launched once after warm-up
(`ncu --set full -k regex:conv_tc_kernel`), or (argument `gn`) the GroupNorm-apply kernel on the frame's big tensors"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops
if len(sys.argv) > 1 and sys.argv[1] == "gn":
    for C, H, W in [(64, 512, 512), (256, 128, 128), (2048, 64, 64), (256, 64, 64)]:
        x = torch.randn(1, H, W, C, device="cuda").bfloat16(); st = torch.zeros(64, dtype=torch.float64, device="cuda")
        g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
        ops.gn_stats(x, st)
        for _ in range(2): ops.gn_apply(x, st, g, b, x, act=ops.ACT_RELU)
    torch.cuda.synchronize(); sys.exit(0)
# (Cin, Cout, k, dil, H, W): the heaviest layer, a full-resolution refine conv, two latency-bound small grids,
# a 1x1 with one K iteration, the PPM-fed decoder conv
shapes = [(256, 256, 3, 1, 128, 128), (64, 64, 3, 1, 512, 512), (256, 256, 3, 1, 32, 32), (64, 256, 1, 1, 128, 128),
          (3072, 256, 3, 1, 64, 64), (512, 2048, 1, 1, 64, 64)]
torch.cuda.profiler.stop()
for Cin, Cout, k, d, H, W in shapes:
    x = torch.randn(1, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)).bfloat16()
    out = torch.empty(1, H, W, Cout, device="cuda", dtype=torch.bfloat16); b = torch.zeros(Cout, device="cuda")
    ws = torch.empty(16 << 20, device="cuda")
    for _ in range(2): ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d, workspace=ws)
torch.cuda.synchronize()
