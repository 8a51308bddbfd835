"""SURVEY.md section 8(f) rank 2 / BASELINE configs[4] (scoped to the operator this repo owns): one training step of the
STM read block -- KeyValue projections, the FUSED Memory.read with its recompute backward, decoder head -- in bf16
autocast under DistributedDataParallel, the way train.py:349-375 steps the stage-4 model.  Gradients are all-reduced by
NCCL (gloo on CPU, where the composite read of the tests stands in for the CUDA kernels).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 scripts/train_step_ddp.py
    python scripts/train_step_ddp.py                     # single process

Prints one JSON line (rank 0): step time (CUDA events, max over ranks), all-reduce bytes per step, the loss trajectory,
and the check that every rank holds identical parameters after the steps (what DDP guarantees iff the all-reduce ran)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from otvm_b200 import train  # noqa: E402


def composite_read(m_in, m_out, q_in, q_out):
    """Memory.forward as the reference writes it (STM.py:144-163), for CPU runs of the DDP plumbing only"""
    import math
    B, De, T, h, w = m_in.shape
    mi = m_in.reshape(B, De, T * h * w).transpose(1, 2)
    p = torch.softmax(torch.bmm(mi, q_in.reshape(B, De, h * w)) / math.sqrt(De), dim=1)
    mem = torch.bmm(m_out.reshape(B, -1, T * h * w), p).view(B, -1, h, w)
    return torch.cat([mem, q_out], dim=1)


def main():
    world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cuda = torch.cuda.is_available() and os.environ.get("OTVM_TRAIN_CPU", "0") != "1"
    dev = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl" if cuda else "gloo", **({"device_id": dev} if cuda else {}))
    steps, warmup = int(os.environ.get("STEPS", 8)), int(os.environ.get("WARMUP", 2))
    h = w = int(os.environ.get("FEAT", 20))                   # 320x320 crops (config.py:27) at 1/16 resolution
    T = int(os.environ.get("T_MEM", 2))                       # memory frames of a 3-frame sample (config.py:31-32)
    torch.manual_seed(111)
    full = os.environ.get("MODEL", "read_block") == "stage4"   # the whole stage-4 model (BASELINE configs[4])
    if full:
        from otvm_b200 import train_stage4
        from otvm_b200.fixtures import make_state_dict, make_train_sample
        model = train_stage4.Stage4Model(read_fn=None if cuda else composite_read)
        model.load_state_dict(make_state_dict("tempered"))      # strict: the reference's 785 stage-4 keys
        model = model.to(dev)
        size, S = int(os.environ.get("SIZE", 320)), int(os.environ.get("FRAMES", 3))    # config.py:27,31-32
    else:
        model = train.STMReadBlock(read_fn=None if cuda else composite_read).to(dev)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local] if cuda else None) if world > 1 else model
    opt = torch.optim.RAdam(ddp.parameters(), lr=1e-4)        # train.py:262 uses RAdam
    losses = []

    def one(i):
        if full:
            sample = [t.to(dev) for t in make_train_sample(1000 * i + rank, S, size, size)]
            return train_stage4.step(ddp, opt, sample, torch.bfloat16 if cuda else None)[0]
        batch = train.synthetic_batch(1, T, h, w, seed=1000 * i + rank, device=dev)       # a different sample per rank
        return train.train_step(ddp, opt, batch, torch.bfloat16 if cuda else None)

    for i in range(warmup):
        one(i)
    if cuda:
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    for i in range(steps):
        losses.append(one(warmup + i))
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    else:
        ms = (time.perf_counter() - t0) * 1e3 / steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # identical parameters on every rank <=> the gradient all-reduce ran every step
    sig = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    if world > 1:
        dist.all_gather(sigs, sig)
    else:
        sigs = [sig]
    if rank == 0:
        what = (f"stage-4 joint step (train.py:349-375): {S} frames of {size}x{size} per rank, FBA + STM networks, four losses, "
                "fused Memory.read fwd+bwd on the path" if full else
                "STM read block (KeyValue x2, fused Memory.read fwd+bwd, convFM, pred), one sample per rank")
        print(json.dumps({"what": what,
                          "n_ranks": world, "backend": ("nccl" if cuda else "gloo") if world > 1 else None,
                          "autocast": "bf16" if cuda else None,
                          "feature_map": [size // 16, size // 16] if full else [h, w], "memory_frames": S - 1 if full else T,
                          "ms_per_step": round(float(t), 3), "steps": steps,
                          "allreduce_bytes_per_step": train.allreduce_bytes(model) if world > 1 else 0,
                          "parameters": sum(p.numel() for p in model.parameters()),
                          "loss_first_last": [round(float(losses[0]), 5), round(float(losses[-1]), 5)],
                          "ranks_in_sync": bool(all(float(s) == float(sigs[0]) for s in sigs)),
                          "fused_read": cuda}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
