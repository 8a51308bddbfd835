import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load(); lib.otvm_debug_set_read_timestamps.argtypes = [ctypes.c_void_p]
for hw_side, T in [(32, 8), (32, 16)]:
    HW = hw_side * hw_side; M = T * HW
    k = torch.randn(M, 128, device="cuda").bfloat16(); v = torch.randn(512, M, device="cuda").bfloat16()
    q = (torch.randn(1, hw_side, hw_side, 128, device="cuda") * 1.5).bfloat16()
    out = torch.zeros(1, hw_side, hw_side, 1024, device="cuda", dtype=torch.bfloat16)
    ws = torch.zeros(ops.memory_read_workspace(M, HW, 128, 512, torch.bfloat16) // 4 + 1, device="cuda")
    for _ in range(2): ops.memory_read(k, v, M, q, out[..., :512], M, ws)
    dbg = torch.zeros(1024, 64, dtype=torch.int64, device="cuda")
    lib.otvm_debug_set_read_timestamps(ctypes.c_void_p(dbg.data_ptr()))
    ops.memory_read(k, v, M, q, out[..., :512], M, ws); torch.cuda.synchronize()
    lib.otvm_debug_set_read_timestamps(None)
    t = dbg[dbg[:, 0] > 0].cpu(); rel = (t - t[:, :1])
    print(f"T={T}: ctas={len(t)} start spread={(t[:,0]-t[:,0].min()).max().item()} setup={rel[:,1].median().item()} o_done={rel[:,2].median().item()} end={rel[:,3].median().item()}")
    c = 5  # some CTA
    print("   s_full seen :", [int(x) for x in rel[c, 8:20]])
    print("   p written   :", [int(x) for x in rel[c, 20:32]])
    print("   S issue     :", [int(x) for x in rel[c, 32:44]])
    print("   PV issue    :", [int(x) for x in rel[c, 44:56]])
    print("   softmax J=4 : s_full", int(rel[c, 12]), "tmem_ld", int(rel[c, 56]), "row max agreed", int(rel[c, 57]), "exps done", int(rel[c, 58]), "p_empty passed", int(rel[c, 59]), "published", int(rel[c, 24]))
