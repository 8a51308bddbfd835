"""dev: per-iteration stamps of the one-tile conv kernel's K loop: when the TMA producer issued iteration i and when
the MMA warp saw stage i full (CTA 0), to tell which role paces the loop"""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load()
lib.otvm_debug_set_conv_timestamps.argtypes = [ctypes.c_void_p]
for Cin, Cout, k, d, H, W in [(1024, 256, 1, 1, 32, 32), (256, 256, 3, 1, 32, 32), (256, 256, 3, 1, 64, 64), (256, 256, 3, 1, 128, 128)]:
    x = torch.randn(1, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)).bfloat16()
    out = torch.empty(1, H, W, Cout, device="cuda", dtype=torch.bfloat16); b = torch.zeros(Cout, device="cuda")
    for _ in range(2): ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d)
    dbg = torch.zeros(4096, 64, dtype=torch.int64, device="cuda")
    lib.otvm_debug_set_conv_timestamps(ctypes.c_void_p(dbg.data_ptr()))
    ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d)
    torch.cuda.synchronize()
    lib.otvm_debug_set_conv_timestamps(None)
    t = dbg[dbg[:, 0] > 0].cpu()
    rel = (t - t[:, :1])
    print(f"Cin={Cin} Cout={Cout} k={k} {H}x{W}: ctas={len(t)}  all-MMA-issued={int(rel[0, 4])} accum-ready={int(rel[0, 5])} epi-done={int(rel[0, 6])}")
    nk = k * k * (Cin // 64)
    per = (rel[:, 4] - rel[:, 3]).float().median() / max(1, nk - 1)
    print(f"   K iterations={nk}  first stage full={int(rel[0, 3])}  MMA-loop period={per:.0f} cycles/iteration (256 = tensor-bound for a 128x128x64 stage)")
