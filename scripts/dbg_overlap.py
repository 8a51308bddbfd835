import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from frames_util import build_model
from util import rel_err
from otvm_b200 import _lib, ops
from otvm_b200.fixtures import make_frame
def run(overlap, pdl, fuse, graphs="1", n=12):
    os.environ["OTVM_OVERLAP"] = overlap; os.environ["OTVM_CUDA_GRAPHS"] = graphs
    _lib.load().otvm_set_pdl(int(pdl)); ops.GN_FUSE = fuse
    model, _ = build_model("tempered", "bf16")
    res = []
    for i in range(n):
        a, fg, bg = make_frame(0, i, 128, 160)
        out = model(a.cuda(), fg.cuda(), bg.cuda(), first_frame=(i == 0), last_frame=(i == n - 1), memorize=(i % 3 != 2), max_memory_num=4)
        res.append((out[3].clone().cpu(), out[1].clone().cpu()))
    return res
def cmp(name, A, B):
    print(name, " ".join(f"{rel_err(a1, a0):.1e}/{rel_err(t1, t0):.1e}" for (a0, t0), (a1, t1) in zip(A, B)), flush=True)
base = run("0", 0, False, graphs="0")
cmp("same cfg rerun (eager, nofuse)      ", base, run("0", 0, False, graphs="0"))
cmp("graphs                               ", base, run("0", 0, False))
cmp("overlap                              ", base, run("1", 0, False))
cmp("pdl                                  ", base, run("0", 1, False))
cmp("fuse (eager)                         ", base, run("0", 0, True, graphs="0"))
f = run("0", 0, True, graphs="0")
cmp("fuse rerun vs fuse                   ", f, run("0", 0, True, graphs="0"))
cmp("fuse+graphs vs fuse                  ", f, run("0", 0, True))
cmp("fuse+pdl vs fuse                     ", f, run("0", 1, True))
cmp("fuse+overlap vs fuse                 ", f, run("1", 0, True))
cmp("fuse+overlap+pdl vs fuse             ", f, run("1", 1, True))
