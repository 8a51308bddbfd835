"""dev: is a graph replay bound by the host (cudaGraphLaunch cost) or by the device?  For a 40-conv dependent chain and
for the steady-state 512x512 frame graph: host time of replay() (no sync), device time of ONE isolated replay
(events, device idle before), and the back-to-back period."""
import math, os, sys, time, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load()


def measure(name, replay, n_kernels, reps=20):
    for _ in range(3): replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iso = []
    for _ in range(5):
        torch.cuda.synchronize(); e0.record(); replay(); e1.record(); torch.cuda.synchronize()
        iso.append(e0.elapsed_time(e1) * 1e3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps): replay()
    t1 = time.perf_counter()
    e1.record(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    dev = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name}: kernels={n_kernels}  host replay() {(t1 - t0) * 1e6 / reps:8.1f} us   isolated device {sorted(iso)[2]:8.1f} us   "
          f"back-to-back period {dev:8.1f} us  ({dev / n_kernels:.2f} us/kernel; wall {(t2 - t0) * 1e6 / reps:.1f} us)", flush=True)


def conv_chain(ci, k, H, n=40):
    x = [torch.randn(1, H, H, ci, device="cuda").bfloat16() for _ in range(2)]
    w = (torch.randn(ci, k, k, ci, device="cuda") / math.sqrt(ci * k * k)).bfloat16(); b = torch.zeros(ci, device="cuda")
    work = torch.empty(16 << 20, dtype=torch.float32, device="cuda")
    def body():
        for i in range(n): ops.conv2d(x[i % 2], w, b, x[1 - i % 2], pad=k // 2, workspace=work)
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): body()
    return g


if __name__ == "__main__":
    for pdl in (1, 0):
        lib.otvm_set_pdl(pdl)
        g = conv_chain(256, 1, 32)
        measure(f"chain 256->256 k1 32^2 pdl={pdl}", g.replay, 40)
    lib.otvm_set_pdl(1)
    g = conv_chain(64, 3, 128)
    measure("chain 64->64 k3 128^2 pdl=1", g.replay, 40)
    # the frame graph
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import bench
    torch.set_grad_enabled(False)
    model = bench.build("bf16")
    from otvm_b200.fixtures import make_frame
    dev = [tuple(t.cuda() for t in make_frame(0, i, 512, 512)) for i in range(8)]
    kw = dict(last_frame=False, memorize=True, max_memory_num=8)
    model(*dev[0], first_frame=True, **kw)
    for i in range(1, 24): model(*dev[i % 8], first_frame=False, **kw)
    torch.cuda.synchronize()
    eng = model.engine
    key, g = next(iter(eng.graphs.items()))
    measure("frame graph (replay only)", g.replay, eng.graph_launches[key])
    it = [0]
    def full():
        it[0] += 1
        model(*dev[it[0] % 8], first_frame=False, **kw)
    measure("frame via EvalModel.forward (device inputs)", full, eng.graph_launches[key])
