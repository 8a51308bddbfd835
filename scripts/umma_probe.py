"""dev: hardware probe for row-shifted SWIZZLE_128B UMMA operands (scripts/umma_probe.cu).  The probe is NOT part of
the product library: it is compiled here into its own shared object, linked against the library's object files
(tensor-map helper) that `python -m otvm_b200.build` leaves in otvm_b200/build/."""
import ctypes as C, glob, os, subprocess, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from otvm_b200.build import build, NVCC, FLAGS
build()
so = "/tmp/libotvm_probe.so"
subprocess.check_call([NVCC, *FLAGS, "-shared", "-I", os.path.join(ROOT, "otvm_b200", "csrc"), "-o", so,
                       os.path.join(ROOT, "scripts", "umma_probe.cu"), *glob.glob(os.path.join(ROOT, "otvm_b200", "build", "*.o")),
                       "-lcuda"])
lib = C.CDLL(so)
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
