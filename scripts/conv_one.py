"""dev/profiling tool: run a few bf16 conv shapes (for `ncu --set full -k regex:conv_tc`)"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops
shapes = [(64, 256, 1, 1, 128, 128), (256, 256, 3, 1, 128, 128), (256, 1024, 1, 1, 32, 32), (64, 64, 3, 1, 512, 512)]
for Cin, Cout, k, d, H, W in shapes:
    x = torch.randn(1, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)).bfloat16()
    out = torch.empty(1, H, W, Cout, device="cuda", dtype=torch.bfloat16); b = torch.zeros(Cout, device="cuda")
    for _ in range(2): ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d)
torch.cuda.synchronize()
