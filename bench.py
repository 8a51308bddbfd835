#!/usr/bin/env python
"""Headline benchmark: frames/s of the OTVM per-frame inference loop (segment -> alpha -> memorize) on
synthetic 512x512 clips with a T=8 memory bank (BASELINE.json configs[1]), one clip per GPU.

    python bench.py --gpus 1 --steps 200 --warmup 4           # B200 arm (this repo), default precision bf16x2
    python bench.py --precision strict                        # three-plane mode (<= 1e-3 of the reference)
    python bench.py --impl reference --steps 8 --warmup 1     # CPU arm: oracle port of the reference
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).
  value      frames/s with inputs resident in HBM (CUDA events around the timed region, max over ranks)
  e2e        the same loop through EvalModel.forward with PINNED HOST inputs (H2D of a/fg/bg through the engine's copy
             stream and D2H of the alpha inside the timed region)
  dtype      arithmetic of the path: "bf16x2" = every value stored as two bf16 planes, three tcgen05 products per K
             step, fp32 accumulation (otvm_b200/split.py); `parity` next to it = max-norm errors of one frame of THIS
             workload against the oracle, computed in the cpu_baseline leg
  roofline   the dominant kernel family (tcgen05 convolutions): ALGORITHMIC FLOPs per frame (686.7
             GFLOP at 512^2, unpadded channels) / the family's time inside the replayed frame / the measured sustained
             bf16 peak.  The family's time = its share of the serialised device time of one frame (CUDA events around
             every C-ABI call of an eager frame, measured live here; the ncu launch list under profiles/ gives the same
             share) x the graph-replayed ms_per_step, so it can never exceed the frame.
  roofline_memory_read / _T16   the fused Memory.read inside the frame (T=8) and alone at the north-star point
             (512^2, T=16), FLOPs against the burst tensor peak, HBM GB/s beside it
  cfg3, precision_modes   BASELINE configs[2] (1024^2 / T=16) and the other precision modes, same protocol, fewer steps
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H = W = 512
T_MEM = 8
RADIUS = 12
METRIC = "frames/sec at 512x512, T=8 memory"
CONV_GFLOP_512 = 686.74          # every convolution of one steady-state frame at 512^2
PLANE_PRODUCTS = {"bf16": 1, "bf16x2": 3, "bf16x3": 6, "fast": 3, "strict": 6, "fp32": 1}


def workload(size, T):
    cfg = {512: 1, 1024: 2}.get(size)
    tag = f" (BASELINE configs[{cfg}])" if cfg is not None and T == (8 if size == 512 else 16) else ""
    return (f"{size}x{size} synthetic clip, T={T} memory frames, full eval.py trimap->alpha per-frame loop{tag}; "
            "one independent clip per GPU, no collective")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True).start()
            time.sleep(0.25)                          # first sample lands before the timed region starts
            self.rows.clear()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for l in self.rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def build(precision, radius=RADIUS):
    import otvm_b200
    from otvm_b200.fixtures import make_state_dict
    cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(STAGE=4))
    mt = otvm_b200.get_model_trimap(cfg, "Test", radius)
    ma = otvm_b200.get_model_alpha(cfg, mt, "Test", radius)
    ma.load_state_dict(make_state_dict("tempered"))
    return ma.cuda().eval().set_precision(precision)


class Clip:
    """one model + one synthetic clip at (size, T): bank filled, graphs captured, ready to time steady-state frames"""

    def __init__(self, precision, size, T, rank, warmup):
        from otvm_b200.fixtures import make_frame
        self.size, self.T = size, T
        self.model = build(precision)
        self.n_src = 8                                     # distinct synthetic frames cycled through
        self.host = [tuple(t.pin_memory() for t in make_frame(rank, i, size, size)) for i in range(self.n_src)]
        self.dev = [tuple(t.cuda() for t in f) for f in self.host]
        self.alpha_host = torch.empty(1, 1, 1, size, size).pin_memory()
        self.kw = dict(last_frame=False, memorize=True, max_memory_num=T)
        # fill the bank: frame 0 + T-1 memorize frames (models/alpha/model.py:472-493)
        self.model(*self.dev[0], first_frame=True, **self.kw)
        for i in range(1, T):
            self.step(i)
        eng = self.model.engine
        assert eng.bank(eng.plan(size, size)).T == T
        for i in range(max(warmup, T) + 1):                # >= T frames so every bank-slot graph is captured before timing
            self.step(i)

    def step(self, i):
        return self.model(*self.dev[i % self.n_src], first_frame=False, **self.kw)

    def e2e_step(self, i):
        # the call a user makes (eval.py:170-175): HOST tensors in, alpha read back to the host
        out = self.model(*self.host[i % self.n_src], first_frame=False, **self.kw)
        self.alpha_host.copy_(out[3], non_blocking=True)


def run_b200(args):
    from otvm_b200 import ops
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug lines go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.set_grad_enabled(False)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, reduce=True):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier() if reduce else torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier() if reduce else torch.cuda.synchronize()
        ms_local = e0.elapsed_time(e1)
        if dist is None or not reduce:
            return ms_local, [ms_local]
        t = torch.tensor([ms_local], device="cuda")
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = [float(x) for x in allr]
        return max(per_rank), per_rank                      # the job is as slow as its slowest rank

    clip = Clip(args.precision, H, T_MEM, rank, args.warmup)
    model = clip.model
    clocks = ClockSampler(local)
    clocks.start()
    l0 = ops.launch_count() + model.engine.replayed_launches
    ms_dev, ms_ranks = timed(clip.step, args.steps)
    launches = ops.launch_count() + model.engine.replayed_launches - l0
    clk = clocks.stop()
    if dist is not None:                                   # whole-job count, like `value`
        lt = torch.tensor([float(launches)], device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    for i in range(3):
        clip.e2e_step(i)
    ms_e2e, e2e_ranks = timed(clip.e2e_step, args.steps)

    base = {"metric": METRIC, "value": round(world * args.steps / (ms_dev * 1e-3), 3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_dev / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": model.precision,
            "data": "synthetic"}
    bytes_in = (1 + 3 + 3) * H * W * 4
    e2e = {"value": round(world * args.steps / (ms_e2e * 1e-3), 3), "unit": "frames/s", "h2d_bytes_per_step": bytes_in,
           "d2h_bytes_per_step": H * W * 4, "ms_per_step": round(ms_e2e / args.steps, 4),
           "api": "EvalModel.forward(a, fg, bg, ...) with pinned host tensors; alpha copied back to pinned host memory"}
    config = {"workload": workload(H, T_MEM), "frame": [H, W], "memory_frames": T_MEM,
              "precision": f"{model.precision}: {PLANE_PRODUCTS[model.precision]} tcgen05 plane product(s) per K step",
              "weights": "random-init (fixtures 'tempered', seed 111)",
              "l2": "per-frame working set (activations + 150-300 MB of weights) exceeds the 126 MB L2; 8 distinct "
                    "frames are cycled, no explicit flush"}
    if world > 1:
        base["ms_per_step_per_rank"] = [round(m / args.steps, 4) for m in ms_ranks]
        e2e["ms_per_step_per_rank"] = [round(m / args.steps, 4) for m in e2e_ranks]
    if args.no_profile:
        if rank == 0:
            base.update(e2e=e2e, gpu_launches=int(launches), clocks=clk, config=config)
            print(json.dumps(base), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- instrumented pass: serialised device time per kernel family (CUDA events around every C-ABI call of an eager
    # frame).  Each instrumented frame is queued behind a device-side sleep so the GPU never waits for the host between
    # launches (an idle GPU would stamp the start event early and charge the host-side launch latency to the kernel).
    prof_frames = 4
    ops.PROFILER = ops.Profiler()
    for i in range(prof_frames):
        torch.cuda._sleep(100_000_000)
        clip.step(i)
        torch.cuda.synchronize()
    fam = ops.PROFILER.summary()
    ops.PROFILER = None
    pk = peaks()
    total_ms = sum(v["ms"] for v in fam.values())
    ms_step = ms_dev / args.steps
    scale = (H * W) / (512.0 * 512.0)

    def roof(name, tensor_peak=None):
        v = fam[name]
        share = v["ms"] / total_ms
        in_frame_ms = share * ms_step                      # this family's time inside the replayed frame
        if tensor_peak is not None:
            gflop = CONV_GFLOP_512 * scale if name == "conv_tcgen05" else v["flops"] / prof_frames / 1e9
            ach = gflop / in_frame_ms                      # GFLOP / ms = TFLOP/s
            r = {"bound": "tensor", "achieved": round(ach, 2), "peak": tensor_peak, "unit": "TFLOP/s",
                 "algorithmic_gflop_per_frame": round(gflop, 2)}
        else:
            ach = v["bytes"] / prof_frames / (in_frame_ms * 1e-3) / 1e9
            r = {"bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s"}
        r.update(frac=round(r["achieved"] / r["peak"], 4), traffic=None, kernel=name,
                 launches_per_frame=v["calls"] // prof_frames, ms_per_frame=round(in_frame_ms, 4),
                 share_of_serialised_frame=round(share, 4), serialised_ms_per_frame=round(v["ms"] / prof_frames, 4),
                 algorithmic_mb_per_frame=round(v["bytes"] / prof_frames / 1e6, 2), peak_source=pk["src"])
        return r

    rl = roof("conv_tcgen05", pk["tf_sust"])
    rl["timing"] = ("share of the serialised per-call CUDA-event time of an eager frame x graph-replayed ms_per_step "
                    "(profiles/: ncu launch list of the same frame gives the same share)")
    rl["plane_products_per_k_step"] = PLANE_PRODUCTS[model.precision]
    rl["executed_mma_frac_of_peak"] = round(rl["frac"] * PLANE_PRODUCTS[model.precision], 4)
    rl_read = roof("memory_read", pk["tf_burst"])
    rl_read["hbm_gbs"] = round(fam["memory_read"]["bytes"] / prof_frames / (rl_read["ms_per_frame"] * 1e-3) / 1e9, 1)
    rl_read["hbm_frac"] = round(rl_read["hbm_gbs"] / pk["hbm"], 4)
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):                      # DRAM bytes per frame per family from the committed ncu pass
        tr = json.load(open(tp))
        if tr.get("precision") == model.precision and H == 512:
            for r_ in (rl, rl_read):
                t = tr.get("per_frame", {}).get(r_["kernel"])
                if t:
                    r_["traffic"] = t["dram_bytes"]
                    r_["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per frame over this family's "
                                          f"launches ({tr.get('source', 'profiles/ncu_traffic.json')})")

    line = dict(base)
    line.update(config=config, e2e=e2e, gpu_launches=int(launches), clocks=clk, roofline=rl, roofline_memory_read=rl_read,
                kernel_families_serialised_ms_per_frame={k: round(v["ms"] / prof_frames, 4) for k, v in sorted(fam.items())})

    if world == 1 and not args.quick:
        line["roofline_memory_read_T16"] = read_microbench(model.precision, pk)
        del clip
        torch.cuda.empty_cache()
        # the other precision modes and BASELINE configs[2], same protocol with fewer steps
        modes = {}
        for prec in ("bf16x3", "bf16"):
            if prec == model.precision:
                continue
            c = Clip(prec, H, T_MEM, rank, 3)
            ms, _ = timed(c.step, 48, reduce=False)
            for i in range(3):
                c.e2e_step(i)                              # (staging buffers / copy stream are created on first use)
            ms2, _ = timed(c.e2e_step, 48, reduce=False)
            modes[prec] = {"value": round(48 / (ms * 1e-3), 2), "e2e": round(48 / (ms2 * 1e-3), 2), "unit": "frames/s",
                           "note": {"bf16x3": "strict mode: three planes, <= 1e-3 of the reference (tests/test_gpu_frames.py)",
                                    "bf16": "plain bf16 storage: misses the reference by up to 0.7 on alpha, NOT a parity mode"}[prec]}
            del c
            torch.cuda.empty_cache()
        line["precision_modes"] = modes
        if H == 512:
            c = Clip(model.precision, 1024, 16, rank, 3)
            ms, _ = timed(c.step, 24, reduce=False)
            for i in range(3):
                c.e2e_step(i)
            ms2, _ = timed(c.e2e_step, 24, reduce=False)
            line["cfg3"] = {"workload": workload(1024, 16), "value": round(24 / (ms * 1e-3), 2), "unit": "frames/s",
                            "ms_per_step": round(ms / 24, 3), "e2e": round(24 / (ms2 * 1e-3), 2), "steps": 24,
                            "dtype": model.precision,
                            "conv_tflops_algorithmic": round(CONV_GFLOP_512 * 4 / (ms / 24), 1)}
            del c
            torch.cuda.empty_cache()
        line["frame_io"] = frame_io_bench(model.precision)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], line["parity"] = run_reference_sample(3, 1, parity_precision=model.precision)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def read_microbench(precision, pk):
    """the fused Memory.read alone at the north-star point (512x512 frame -> HW = 1024 queries, T = 16 -> 16384 memory
    locations), operands in the precision mode's element format, L2 flushed between launches; algorithmic FLOPs / bytes of
    SURVEY.md section 8(d)"""
    from otvm_b200 import ops
    from otvm_b200.engine import PRECISION_MODES
    from otvm_b200.split import SplitArena
    planes = max(1, PRECISION_MODES[precision][1])
    HW, T, De, Do = 1024, 16, 128, 512
    M = T * HW
    ar = SplitArena(planes, 64 << 20, "cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    keys, vals = ar.alloc((M, De)), ar.alloc((Do, M))
    q, out = ar.alloc((1, 32, 32, De)), ar.alloc((1, 32, 32, 2 * Do))
    for t in (keys, vals, q):
        ar.write(t, torch.randn(t.shape, device="cuda", generator=g))
    ws = torch.zeros(ops.memory_read_workspace(M, HW, De, Do) // 4 + 1, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    simt = planes == 3                                     # strict mode reads through the fp32 FFMA kernel (engine.segment)
    n, tot = 20, 0.0
    for i in range(n + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.memory_read(keys, vals, M, q, out[..., :Do], M, ws, force_simt=simt)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            tot += e0.elapsed_time(e1)
    ms = tot / n
    gflop = 2.0 * M * HW * (De + Do) / 1e9
    es = 2 * planes
    nbytes = (De + Do) * M * es + De * HW * es + Do * HW * es
    tf = gflop / ms
    return {"bound": "tensor", "achieved": round(tf, 1), "peak": pk["tf_burst"], "unit": "TFLOP/s",
            "frac": round(tf / pk["tf_burst"], 4), "traffic": None, "kernel": "memory_read (kernel + split combine)",
            "us_per_launch": round(ms * 1e3, 2), "algorithmic_gflop": round(gflop, 2),
            "algorithmic_mb": round(nbytes / 1e6, 2), "hbm_gbs": round(nbytes / (ms * 1e-3) / 1e9, 1),
            "hbm_frac": round(nbytes / (ms * 1e-3) / 1e9 / pk["hbm"], 4),
            "workload": "512x512 frame, T=16: HW=1024 queries x THW=16384 memory locations, L2 flushed between launches",
            "plane_products_per_k_step": 1 if planes == 1 else 3, "ffma": simt, "peak_source": pk["src"]}


def frame_io_bench(precision, n=96):
    """the same clip from PNG files to PNG mattes, (a) with otvm_b200.frame_io (decoder /
    writer threads, 8-bit transfers, device-side unpack / conversion) and (b) the way eval.py does it (decode, fp32
    conversion and upload, synchronous fp32 read-back, host conversion and imwrite on the main thread,
    dataset.py:857-920 + eval.py:195-217).  Same model, same frames; frames/s of the whole loop."""
    import shutil
    import tempfile
    try:
        import cv2
        import numpy as np
    except Exception as e:                      # no OpenCV on this box: nothing to measure
        return {"unavailable": repr(e)}
    from otvm_b200.frame_io import run_sequence
    d = tempfile.mkdtemp(prefix="otvm_io_")
    try:
        r = np.random.RandomState(0)
        yy, xx = np.mgrid[0:H, 0:W]
        base_f = r.randint(0, 256, (H, W, 4)).astype(np.uint8); base_b = r.randint(0, 256, (H, W, 3)).astype(np.uint8)
        fgs, bgs = [], []
        for i in range(n):
            f = np.roll(base_f, 3 * i, axis=1)
            f[..., 3] = np.clip(255 - (np.hypot(yy - H / 2, xx - W / 2 - i) - H / 4) * 12, 0, 255).astype(np.uint8)
            fp, bp = os.path.join(d, f"fg_{i:04d}.png"), os.path.join(d, f"bg_{i:04d}.png")
            cv2.imwrite(fp, f); cv2.imwrite(bp, np.roll(base_b, 2 * i, axis=0))
            fgs.append(fp); bgs.append(bp)
        model = build(precision)
        kw = dict(max_memory_num=T_MEM, memory_skip_frame=1)
        run_sequence(model, fgs[:12], bgs[:12], os.path.join(d, "warm"), **kw)       # graphs / plans
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_sequence(model, fgs, bgs, os.path.join(d, "out"), **kw)
        torch.cuda.synchronize()
        t_pipe = time.perf_counter() - t0
        os.makedirs(os.path.join(d, "ref"), exist_ok=True)
        t0 = time.perf_counter()
        for i in range(n):                      # eval.py's own sequence of host steps around the same model
            _f = cv2.imread(fgs[i], cv2.IMREAD_UNCHANGED)
            fg = np.float32(_f[..., :-1]); a = np.float32(_f[..., -1:]) / 255.
            bg = np.float32(cv2.imread(bgs[i], cv2.IMREAD_COLOR))
            t = lambda x: torch.from_numpy(x).permute(2, 0, 1).unsqueeze(0).unsqueeze(0).float()
            torch.cuda.synchronize()
            out = model(t(a).cuda(), t(fg).cuda(), t(bg).cuda(), first_frame=(i == 0), last_frame=(i == n - 1),
                        memorize=False, max_memory_num=T_MEM)
            torch.cuda.synchronize()
            img = (out[3] * 255).byte().cpu().squeeze(0).squeeze(0).squeeze(0).numpy()
            cv2.imwrite(os.path.join(d, "ref", f"fg_{i:04d}.png"), img)
        t_ref = time.perf_counter() - t0
        same = all(np.abs(cv2.imread(os.path.join(d, "out", f"fg_{i:04d}.png"), cv2.IMREAD_UNCHANGED).astype(int) -
                          cv2.imread(os.path.join(d, "ref", f"fg_{i:04d}.png"), cv2.IMREAD_UNCHANGED).astype(int)).max() <= 1
                   for i in range(0, n, 7))
        return {"value": round(n / t_pipe, 2), "unit": "frames/s", "frames": n, "frame": [H, W],
                "what": "PNG files -> decode -> model -> 8-bit matte -> PNG files, wall clock of the whole clip",
                "pipelined": "otvm_b200.frame_io: 2 decoder + 2 writer threads, 7 B/px upload, 1 B/px read-back",
                "eval_py_style_host_loop": round(n / t_ref, 2), "outputs_match": bool(same)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_reference_sample(steps, warmup, parity_precision=None):
    """The reference's algorithm on the host cores: the oracle port (oracle/otvm_oracle.py, plain PyTorch CPU
    fp32 ops, validated bit-for-bit against the unmodified reference in tests/test_oracle_golden.py).
    Bounded sample: the bank is filled to T by repeating frame 0's memory (work per frame does not depend on the bank's
    content), then `steps` steady-state frames are timed.  With ``parity_precision`` the oracle also serves as the
    CHECKER of the engine: the first timed frame is run by both on the same bank and compared."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import otvm_oracle as O
    from otvm_b200.fixtures import make_frame, make_state_dict
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    parity = None
    with torch.no_grad():
        om = O.OracleEvalModel(make_state_dict("tempered"), dilate_kernel=RADIUS)
        a, fg, bg = make_frame(0, 0, H, W)
        om(a, fg, bg, first_frame=True, last_frame=False, memorize=True, max_memory_num=T_MEM)
        om.memories = {k: v.repeat(1, 1, 1, T_MEM, 1, 1) for k, v in om.memories.items()}
        kw = dict(first_frame=False, last_frame=False, memorize=True, max_memory_num=T_MEM)
        for i in range(warmup):
            om(*make_frame(0, 1 + i, H, W), **kw)
        frames = [make_frame(0, 1 + warmup + i, H, W) for i in range(steps)]
        bank0 = {k: v.clone() for k, v in om.memories.items()}
        t0 = time.perf_counter()
        outs = [om(*f, **kw) for f in frames[:1]]
        seg0 = om.trace["seg_logit"].clone()
        for f in frames[1:]:
            om(*f, **kw)
        dt = time.perf_counter() - t0
        if parity_precision is not None and torch.cuda.is_available():
            parity = check_parity(parity_precision, frames[0], bank0, outs[0], seg0, kw)
    eager = None
    if parity_precision is not None and torch.cuda.is_available():
        eager = torch_eager_same_gpu(O, kw)
    cpu = {"value": round(steps / dt, 4), "unit": "frames/s", "cores": cores, "kind": "port",
           "torch_eager_same_gpu": eager,
           "sample": f"{steps} steady-state {H}x{W} frames at T={T_MEM} (after {warmup} warm-up), oracle port of the "
                     f"reference on {cores} host threads, torch {torch.__version__} CPU fp32",
           "ms_per_step": round(dt / steps * 1e3, 1)}
    return cpu, parity


def torch_eager_same_gpu(O, kw, steps=12, warmup=3):
    """SURVEY.md section 2.3's bar: the reference's OWN GPU path on this very B200 -- the same PyTorch modules in eager
    mode through cuDNN / cuBLAS (strict fp32, and with TF32 allowed), host distance transform and all
    (utils/utils.py:12-23).  It is the oracle port with its tensors on the device; reported
    for context next to the CPU number, measured with the same steady-state protocol (T-frame bank, wall clock with a
    synchronise per frame as eval.py:195-197 does)."""
    from otvm_b200.fixtures import make_frame, make_state_dict
    out = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            om = O.OracleEvalModel(make_state_dict("tempered"), dilate_kernel=RADIUS, device="cuda")
            fr = [tuple(t.cuda() for t in make_frame(0, i, H, W)) for i in range(4)]
            om(*fr[0], first_frame=True, last_frame=False, memorize=True, max_memory_num=T_MEM)
            om.memories = {k: v.repeat(1, 1, 1, T_MEM, 1, 1) for k, v in om.memories.items()}
            for i in range(warmup):
                om(*fr[i % 4], **kw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(steps):
                om(*fr[i % 4], **kw)
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out["tf32" if tf32 else "fp32"] = round(steps / dt, 2)
        del om
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.cuda.empty_cache()
    return {"unit": "frames/s", "frames_per_s": out, "steps": steps,
            "what": "oracle port (plain PyTorch eager, cuDNN / cuBLAS, host EDT) with its tensors on this GPU; "
                    f"{H}x{W}, T={T_MEM}; 'tf32' = torch.backends.*.allow_tf32 on"}


def check_parity(precision, frame, bank, ref, ref_seg, kw):
    """one steady-state frame of the benchmarked workload on the engine vs the oracle, same T-frame bank: scale-relative
    max errors (max|got - want| / max|want|) of the propagated trimap logits, the trimap and the alpha matte"""
    from otvm_b200.fixtures import make_frame
    model = build(precision)
    a0, fg0, bg0 = make_frame(0, 0, H, W)
    model(a0.cuda(), fg0.cuda(), bg0.cuda(), first_frame=True, last_frame=False, memorize=True, max_memory_num=T_MEM)
    eng = model.engine
    pl = eng.plan(H, W)
    b = eng.bank(pl)
    eng.flush(pl)
    key, val = bank["key"][0, 0], bank["val"][0, 0]
    T = key.shape[1]
    b.store(key.permute(1, 2, 3, 0).reshape(T * b.hw, -1).cuda(), val.reshape(val.shape[0], T * b.hw).cuda())
    b.order = list(range(T))
    out = model(*(t.cuda() for t in frame), **kw)
    torch.cuda.synchronize()

    def err(got, want):
        return float((got.double().cpu() - want.double()).abs().max() / max(float(want.abs().max()), 1e-12))
    seg = pl.bufs["seg_logits"][0, pl.pad_top:pl.pad_top + H, pl.pad_left:pl.pad_left + W, :3].permute(2, 0, 1)
    # pixels whose propagated-trimap class differs (near-ties of the oracle's own logits: the class is an argmax)
    flips = int((seg.float().cpu().argmax(0) != ref_seg[0].argmax(0)).sum())
    return {"precision": precision, "frame": f"steady-state {H}x{W} frame, T={T} bank shared with the oracle",
            "propagated_class_flips": flips,
            "metric": "max|got-want| / max|want| against the CPU oracle (bit-identical to the reference on tests/golden)",
            "seg_logit": err(seg, ref_seg[0]), "trimap": err(out[1], ref[1]), "alpha": err(out[3], ref[3]),
            "tolerance": {"bf16x2": 1e-2, "fast": 1e-2, "bf16x3": 1e-3, "strict": 1e-3, "fp32": 1e-3}.get(precision)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 16)); warmup = max(1, min(args.warmup, 2))
    cpu, _ = run_reference_sample(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "frames/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": steps, "warmup": warmup,
            "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload(H, T_MEM), "frame": [H, W], "memory_frames": T_MEM,
                       "host": "the reference's per-frame loop on the host CPU cores (oracle port, all threads); each step "
                               "is one steady-state frame, the run is a bounded sample of the clip"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16x2", choices=["bf16x2", "bf16x3", "fast", "strict", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="dev: skip the instrumented per-family pass")
    ap.add_argument("--quick", action="store_true", help="skip the cfg3 / precision-mode / T=16 read sub-records")
    ap.add_argument("--size", type=int, default=512, help="frame height = width (BASELINE configs[2]: 1024)")
    ap.add_argument("--memory", type=int, default=8, help="memory-bank frames T (BASELINE configs[2]: 16)")
    args = ap.parse_args()
    global H, W, T_MEM, METRIC
    H = W = args.size
    T_MEM = args.memory
    METRIC = f"frames/sec at {H}x{W}, T={T_MEM} memory"
    args.precision = {"fast": "bf16x2", "strict": "bf16x3"}.get(args.precision, args.precision)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
