"""dev: per-CTA phase timestamps (clock64) of the tcgen05 conv kernel for a few shapes"""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load()
lib.otvm_debug_set_conv_timestamps.argtypes = [ctypes.c_void_p]
names = {0: "start", 1: "setup done (after pdl wait)", 2: "first TMA issued", 3: "first stage full (mma)", 4: "all MMA issued",
         5: "accum ready (epi)", 8: "first tmem ld done", 9: "chunk loop done", 6: "epilogue done", 7: "cta end"}
for Cin, Cout, k, d, H, W in [(64, 256, 1, 1, 128, 128), (256, 64, 1, 1, 128, 128), (256, 256, 3, 1, 32, 32), (1024, 256, 1, 1, 32, 32), (256, 256, 3, 1, 128, 128)]:
    x = torch.randn(1, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)).bfloat16()
    out = torch.empty(1, H, W, Cout, device="cuda", dtype=torch.bfloat16); b = torch.zeros(Cout, device="cuda")
    for _ in range(2): ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d)
    dbg = torch.zeros(4096, 64, dtype=torch.int64, device="cuda")
    lib.otvm_debug_set_conv_timestamps(ctypes.c_void_p(dbg.data_ptr()))
    ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d)
    torch.cuda.synchronize()
    lib.otvm_debug_set_conv_timestamps(None)
    t = dbg[dbg[:, 0] > 0].cpu()
    rel = (t - t[:, :1]).float()
    print(f"Cin={Cin} Cout={Cout} k={k} {H}x{W}: ctas={len(t)} kernel span={int(t[:, 7].max() - t[:, 0].min())} cyc, CTA start spread={int(t[:, 0].max() - t[:, 0].min())}")
    print("   " + " | ".join(f"{n}={rel[:, i].median():.0f}" for i, n in names.items() if i and (rel[:, i] > 0).any()), flush=True)
