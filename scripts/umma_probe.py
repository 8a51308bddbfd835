"""dev: hardware probe for row-shifted SWIZZLE_128B UMMA operands (see csrc/umma_probe.cu)"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import _lib
lib = _lib.load()
lib.otvm_debug_umma_shift_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
torch.manual_seed(0)
A = torch.randn(136, 64, device="cuda").bfloat16(); W = torch.randn(64, 64, device="cuda").bfloat16()
for policy in (0, 1):
    out = torch.zeros(9, 128, 64, device="cuda")
    rc = lib.otvm_debug_umma_shift_probe(A.data_ptr(), W.data_ptr(), out.data_ptr(), policy, None)
    torch.cuda.synchronize()
    errs = []
    for s in range(9):
        want = A[s:s + 128].float() @ W.float().t()
        errs.append(float((out[s] - want).abs().max() / want.abs().max()))
    print(f"policy={policy} (base_offset {'=(addr>>7)&7' if policy else '=0'}) rc={rc} rel_err per shift:", " ".join(f"{e:.1e}" for e in errs), flush=True)
