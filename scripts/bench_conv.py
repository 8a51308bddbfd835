"""micro-benchmark of the conv kernels on the layer shapes of one 512x512 frame (dev tool)"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops
DEV = "cuda"
shapes = [(3072, 256, 3, 1, 64, 64), (512, 512, 3, 4, 64, 64), (1024, 2048, 1, 1, 64, 64), (256, 256, 3, 1, 128, 128),
          (64, 64, 3, 1, 512, 512), (96, 64, 3, 1, 512, 512), (96, 32, 3, 1, 512, 512), (32, 16, 3, 1, 512, 512),
          (1024, 512, 3, 1, 32, 32), (1024, 128, 3, 1, 32, 32), (256, 1024, 1, 1, 32, 32), (512, 256, 3, 1, 128, 128),
          (320, 64, 3, 1, 256, 256), (64, 256, 1, 1, 128, 128)]
for dt in (torch.bfloat16, torch.float32):
    for Cin, Cout, k, d, H, W in shapes:
        x = torch.randn(1, H, W, Cin, device=DEV).to(dt); w = (torch.randn(Cout, k, k, Cin, device=DEV) / math.sqrt(Cin * k * k)).to(dt)
        out = torch.empty(1, H, W, Cout, device=DEV, dtype=dt); b = torch.zeros(Cout, device=DEV)
        pad = d * (k // 2)
        for _ in range(3): ops.conv2d(x, w, b, out, pad=pad, dil=d)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20 if dt == torch.bfloat16 else 5
        e0.record()
        for _ in range(n): ops.conv2d(x, w, b, out, pad=pad, dil=d)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * H * W * Cout * Cin * k * k
        print(f"{str(dt)[6:]:9s} Cin={Cin:5d} Cout={Cout:5d} k={k} d={d} {H}x{W}: {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
