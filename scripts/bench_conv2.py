"""dev: warm device time of conv shapes, halo vs per-tap, measured by replaying a CUDA graph of 20 launches
(no host launch overhead in the number)"""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load(); lib.otvm_debug_set_conv_halo.argtypes = [ctypes.c_int]
shapes = [(64, 64, 3, 1, 512, 512), (96, 64, 3, 1, 512, 512), (96, 32, 3, 1, 512, 512), (32, 16, 3, 1, 512, 512),
          (256, 256, 3, 1, 128, 128), (512, 256, 3, 1, 128, 128), (320, 64, 3, 1, 256, 256), (3072, 256, 3, 1, 64, 64),
          (512, 512, 3, 4, 64, 64), (256, 256, 3, 2, 64, 64), (256, 256, 3, 1, 64, 64), (128, 128, 3, 1, 64, 64),
          (256, 256, 3, 1, 32, 32), (1024, 512, 3, 1, 32, 32), (64, 64, 3, 1, 128, 128),
          (64, 256, 1, 1, 128, 128), (1024, 2048, 1, 1, 64, 64), (256, 1024, 1, 1, 32, 32)]
for Cin, Cout, k, d, H, W in shapes:
    x = torch.randn(1, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)).bfloat16()
    out = torch.empty(1, H, W, Cout, device="cuda", dtype=torch.bfloat16); b = torch.zeros(Cout, device="cuda")
    ws = torch.empty(16 << 20, device="cuda")
    res = []
    for halo in ((1, 0) if k == 3 else (1,)):
        lib.otvm_debug_set_conv_halo(halo)
        for _ in range(2): ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d, workspace=ws)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d, workspace=ws)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        res.append(f"{'halo' if halo else 'tap '}: {us:7.1f} us {2.0 * H * W * Cout * Cin * k * k / us / 1e6:7.1f} TF/s")
    lib.otvm_debug_set_conv_halo(1)
    print(f"Cin={Cin:5d} Cout={Cout:5d} k={k} d={d} {H}x{W}: " + "   ".join(res), flush=True)
