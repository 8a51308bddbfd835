"""dev: warm device time of conv shapes, halo vs per-tap and ring budgets, measured by replaying a CUDA graph of 20
launches (no host launch overhead in the number).  usage: bench_conv2.py [budget_kb ...]"""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load(); lib.otvm_debug_set_conv_halo.argtypes = [ctypes.c_int]; lib.otvm_debug_set_conv_budget_kb.argtypes = [ctypes.c_int]
budgets = [int(v) for v in sys.argv[1:]] or [0]
shapes = [(64, 64, 3, 1, 512, 512), (96, 64, 3, 1, 512, 512), (96, 32, 3, 1, 512, 512), (32, 16, 3, 1, 512, 512), (64, 32, 3, 1, 512, 512),
          (256, 256, 3, 1, 128, 128), (512, 256, 3, 1, 128, 128), (320, 64, 3, 1, 256, 256), (3072, 256, 3, 1, 64, 64),
          (512, 512, 3, 4, 64, 64), (256, 256, 3, 2, 64, 64), (256, 256, 3, 1, 64, 64), (128, 128, 3, 1, 64, 64),
          (256, 256, 3, 1, 32, 32), (1024, 512, 3, 1, 32, 32), (64, 64, 3, 1, 128, 128),
          (64, 256, 1, 1, 128, 128), (1024, 2048, 1, 1, 64, 64), (256, 1024, 1, 1, 32, 32), (16, 64, 7, 1, 512, 512)]
for Cin, Cout, k, d, H, W in shapes:
    stride = 2 if k == 7 else 1
    x = torch.randn(1, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)).bfloat16()
    out = torch.empty(1, H // stride, W // stride, Cout, device="cuda", dtype=torch.bfloat16); b = torch.zeros(Cout, device="cuda")
    ws = torch.empty(16 << 20, device="cuda")
    res = []
    for halo in ((-1, 1, 0) if k == 3 else (-1,)):
        for bud in budgets:
            lib.otvm_debug_set_conv_halo(halo); lib.otvm_debug_set_conv_budget_kb(bud)
            kw = dict(pad=d * (k // 2), dil=d, workspace=ws, stride=stride)
            for _ in range(2): ops.conv2d(x, w, b, out, **kw)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20): ops.conv2d(x, w, b, out, **kw)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            res.append(f"{'A' if halo < 0 else 'H' if halo else 'T'}{bud}:{us:6.1f}")
    lib.otvm_debug_set_conv_halo(-1); lib.otvm_debug_set_conv_budget_kb(0)
    best = min(float(r.split(':')[1]) for r in res)
    print(f"Cin={Cin:5d} Cout={Cout:5d} k={k} d={d} {H}x{W}: " + " ".join(res) + f"  best {2.0 * (H // stride) * (W // stride) * Cout * Cin * k * k / best / 1e6:6.0f} TF/s", flush=True)
