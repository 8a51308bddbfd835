"""worker of tests/test_host_logic.py::test_clip_sharding_two_ranks_gloo (launched under torch.distributed.run)"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.environ["OTVM_ROOT"])
from otvm_b200.fixtures import make_frame  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
clips = [c for c in range(6) if c % world == rank]
sig = torch.tensor([float(make_frame(c, 0, 32, 32)[1].sum()) for c in clips], dtype=torch.float64)
ms = torch.tensor([10.0 + rank])                     # per-rank elapsed time -> job time is the max over ranks
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
out = [torch.zeros_like(sig) for _ in range(world)]
dist.all_gather(out, sig)
if rank == 0:
    allsig = torch.cat(out)
    assert float(ms) == 10.0 + world - 1
    assert len(set(allsig.tolist())) == 6, "each clip processed exactly once across ranks"
    print("OK", world)
dist.destroy_process_group()
