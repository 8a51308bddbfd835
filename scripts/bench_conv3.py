"""dev: bench_conv2.py for split operands (bf16x2): warm device time of 3x3 shapes with the patch ("halo") mode auto /
forced on / off, 20 launches replayed from a CUDA graph.   python scripts/bench_conv3.py"""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
from otvm_b200.split import SplitArena, split_planes
lib = _lib.load(); lib.otvm_debug_set_conv_halo.argtypes = [ctypes.c_int]
ar = SplitArena(2, 1 << 30, "cuda")
shapes = [(320, 64, 3, 1, 256, 256), (512, 256, 3, 1, 128, 128), (256, 256, 3, 1, 128, 128), (256, 256, 3, 2, 64, 64),
          (512, 512, 3, 4, 64, 64), (128, 128, 3, 1, 64, 64), (256, 256, 3, 1, 32, 32), (64, 64, 3, 1, 128, 128), (256, 3, 3, 1, 128, 128)]
ws = torch.empty(16 << 20, device="cuda")
for Cin, Cout, k, d, H, W in shapes:
    x = ar.alloc((1, H, W, Cin)); ar.write(x, torch.randn(1, H, W, Cin, device="cuda"))
    w = split_planes(torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k), 2)
    out = ar.alloc((1, H, W, max(Cout, 8)))[..., :Cout]; b = torch.zeros(Cout, device="cuda")
    res = []
    for halo in (-1, 1, 0):
        lib.otvm_debug_set_conv_halo(halo)
        kw = dict(pad=d, dil=d, workspace=ws, act=ops.ACT_RELU)
        try:
            for _ in range(2): ops.conv2d(x, w, b, out, **kw)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20): ops.conv2d(x, w, b, out, **kw)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            res.append(f"{'auto' if halo < 0 else 'halo' if halo else 'taps'} {e0.elapsed_time(e1) / 20 * 1e3:6.1f} us")
        except Exception as e:
            res.append(f"{'auto' if halo < 0 else 'halo' if halo else 'taps'} failed ({type(e).__name__})")
    lib.otvm_debug_set_conv_halo(-1)
    print(f"Cin={Cin:4d} Cout={Cout:4d} k={k} d={d} {H}x{W}: " + "  ".join(res), flush=True)
