"""CPU-only checks of the host side: state-dict contract, weight packing, memory-bank policy, C ABI exports,
clip sharding across ranks (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from util import ROOT


def test_state_dict_contract():
    """785 keys with the reference's names/shapes; strict load like eval.py:79"""
    import types
    import otvm_b200
    from otvm_b200.fixtures import make_state_dict
    from otvm_b200.spec import state_spec
    spec = state_spec()
    assert len(spec) == 785
    assert spec["NET.encoder.conv1.weight"].shape == (64, 11, 7, 7)
    assert spec["NET.refine.conv1.0.weight"].shape == (64, 73, 3, 3)
    assert spec["NET.refine.pred.4.weight"].shape == (10, 16, 1, 1)
    assert spec["trimap.model.Encoder_M.conv1_h.weight"].shape == (64, 16, 7, 7)
    assert spec["trimap.model.KV_M_r4.Value.weight"].shape == (512, 1024, 3, 3)
    assert sum(int(torch.tensor(e.shape).prod()) if e.shape else 1 for e in spec.values()) == 73949006
    cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(STAGE=4))
    mt = otvm_b200.get_model_trimap(cfg, "Test", 12)
    ma = otvm_b200.get_model_alpha(cfg, mt, "Test", 12)
    sd = make_state_dict("tempered")
    assert set(ma.state_dict()) == set(spec) == set(sd)
    ma.load_state_dict(sd)                      # strict
    with pytest.raises(RuntimeError):           # no CPU fallback: forward needs a CUDA device
        if not torch.cuda.is_available():
            ma.engine
        else:
            raise RuntimeError("gpu present")


def test_fixture_weights_are_stable():
    """name-keyed RandomState draws: the same bits here, in the golden generator and on the GPU box"""
    from otvm_b200.fixtures import make_frame, make_state_dict
    sd = make_state_dict("tempered")
    w = sd["trimap.model.KV_Q_r4.Key.weight"]
    assert abs(float(w.double().sum()) - float(make_state_dict("tempered")["trimap.model.KV_Q_r4.Key.weight"].double().sum())) == 0
    a, fg, bg = make_frame(0, 3, 64, 96)
    assert a.shape == (1, 1, 1, 64, 96) and fg.shape == (1, 1, 3, 64, 96)
    assert 0 <= float(a.min()) and float(a.max()) <= 1 and 0 <= float(fg.min()) and float(fg.max()) < 255
    assert ((a > 0) & (a < 1)).any() and (a == 0).any() and (a == 1).any()


def test_weight_packing_matches_torch_semantics():
    """BN folding, weight standardisation and the fused 22-channel Encoder_M stem against plain torch ops"""
    import torch.nn.functional as F
    from otvm_b200.engine import PackedWeights, _ws
    from otvm_b200.fixtures import make_state_dict
    sd = make_state_dict("default")
    pw = PackedWeights(sd, torch.float32, "cpu")
    g = torch.Generator().manual_seed(0)
    # folded BN == conv -> batch_norm(eval)
    p = "trimap.model.Encoder_Q.res3.0"
    x = torch.randn(1, 256, 8, 8, generator=g)
    want = F.batch_norm(F.conv2d(x, sd[p + ".downsample.0.weight"], None, 2), sd[p + ".downsample.1.running_mean"],
                        sd[p + ".downsample.1.running_var"], sd[p + ".downsample.1.weight"],
                        sd[p + ".downsample.1.bias"], False, 0.0, 1e-5)
    w, b = pw.conv[p + ".downsample"]
    got = F.conv2d(x, w.permute(0, 3, 1, 2), b, 2)
    assert float((got - want).abs().max()) < 1e-4
    # weight standardisation folded once == layers_WS.Conv2d.forward
    w, b = pw.conv["NET.encoder.layer1.0.conv2"]
    assert float((w.permute(0, 3, 1, 2) - _ws(sd["NET.encoder.layer1.0.conv2.weight"])).abs().max()) == 0
    # five Encoder_M stems summed (STM.py:63,67) == one conv over the concatenated 22 channels (+2 zero pads)
    e = "trimap.model.Encoder_M"
    f, m, o, a, h = (torch.randn(1, c, 16, 16, generator=g) for c in (3, 1, 1, 1, 16))
    conv = lambda n, t: F.conv2d(t, sd[f"{e}.{n}.weight"], None, 2, 3)
    pre = conv("conv1", f) + conv("conv1_m", m) + conv("conv1_o", o) + conv("conv1_a", a) + conv("conv1_h", h)
    want = F.batch_norm(pre, sd[e + ".bn1.running_mean"], sd[e + ".bn1.running_var"], sd[e + ".bn1.weight"],
                        sd[e + ".bn1.bias"], False, 0.0, 1e-5)
    w, b = pw.conv[e + ".stem"]
    xin = torch.cat([f, m, o, a, h, torch.zeros(1, w.shape[3] - 22, 16, 16)], 1)
    got = F.conv2d(xin, w.permute(0, 3, 1, 2), b, 2, 3)
    assert float((got - want).abs().max()) < 1e-4


@pytest.mark.parametrize("planes", [1, 2])
def test_encoder_pair_weights_and_heads(planes):
    """the grouped Encoder_Q / Encoder_M launches read ONE pair tensor per layer whose halves ARE the per-encoder entries
    (views, no copy, same values as a stand-alone packing), and the 1x1 heads keep their fp32 weights"""
    from otvm_b200.engine import PackedWeights
    from otvm_b200.fixtures import make_state_dict
    from otvm_b200.split import split_planes
    sd = make_state_dict("default")
    pw = PackedWeights(sd, torch.bfloat16, "cpu", planes=planes)
    q, m = "trimap.model.Encoder_Q", "trimap.model.Encoder_M"
    assert len(pw.pair) == 3 * 3 + 4 * 3 + 6 * 3 + 3            # 13 bottlenecks x 3 convolutions + 3 downsample branches
    for sfx, (wp, bp) in pw.pair.items():
        (wq, bq), (wm, bm) = pw.conv[q + sfx], pw.conv[m + sfx]
        cout = wq.shape[-4]
        assert wp.shape[-4] == 2 * cout and bp.numel() == 2 * cout
        half = (lambda t, i: t[:, i * cout:(i + 1) * cout]) if planes > 1 else (lambda t, i: t[i * cout:(i + 1) * cout])
        assert torch.equal(half(wp, 0), wq) and torch.equal(half(wp, 1), wm)
        assert torch.equal(bp[:cout], bq) and torch.equal(bp[cout:], bm)
        assert wq.untyped_storage().data_ptr() == wp.untyped_storage().data_ptr()      # a view of the pair tensor
    # values: the same BN-folded weights a single-plane fp32 packing gives, rounded / split
    ref = PackedWeights(sd, torch.float32, "cpu")
    name = q + ".res3.1.conv2"
    w32 = ref.conv[name][0]
    want = split_planes(w32, planes) if planes > 1 else w32.to(torch.bfloat16)
    assert torch.equal(pw.conv[name][0], want)
    for name, cout in (("NET.decoder.conv_up4.4", 7), ("NET.refine.pred.4", 10)):
        w, b = pw.head[name]
        assert w.dtype == torch.float32 and tuple(w.shape) == (cout, 16) and torch.equal(w, sd[name + ".weight"].view(cout, 16))
        assert torch.equal(b, sd[name + ".bias"])


def _reference_policy(events, max_n):
    """models/alpha/model.py:472-493 on a list of frame ids"""
    mem = None
    for i, (first, memorize) in enumerate(events):
        new = [i]
        if max_n == 0:
            if first:
                mem = new
        elif max_n == 1:
            mem = new
        else:
            if first:
                mem = new
            elif memorize:
                mem = mem + new
            else:
                mem = (mem + new) if len(mem) == 1 else (mem[:-1] + new)
            if len(mem) > max_n:
                mem = mem[:1] + mem[2:]
    return mem


@pytest.mark.parametrize("max_n", [0, 1, 2, 3, 5, 8])
def test_memory_bank_policy_matches_reference(max_n):
    """in-place slot replacement holds exactly the frames the reference's torch.cat policy would hold"""
    import random
    from otvm_b200.engine import MemoryBank
    rng = random.Random(max_n)
    bank = MemoryBank(hw=4, cap=16, dtype=torch.float32, device="cpu")
    content = {}
    events = [(True, True)] + [(False, rng.random() < 0.4) for _ in range(60)]
    for i, (first, memorize) in enumerate(events):
        if first:
            bank.reset()
        slot, order = bank.next_slot(first, memorize, max_n)
        if slot is not None:
            content[slot] = i
            bank.order = order
        want = _reference_policy(events[:i + 1], max_n)
        assert [content[s] for s in bank.order] == want, (i, bank.order, want)
        assert len(set(bank.order)) == len(bank.order) and all(s < 16 for s in bank.order)


def test_c_abi_exports_every_declared_symbol():
    """the shared library loads without a GPU and exports exactly what include/otvm_b200.h declares"""
    from otvm_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "otvm_b200.h")).read()
    declared = set(re.findall(r"OTVM_API\s+[\w\s\*]+?\b(otvm_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name)
    dbg = open(os.path.join(ROOT, "include", "otvm_b200_debug.h")).read()
    for name in set(re.findall(r"\b(otvm_debug_\w+)\s*\(", dbg)):          # the diagnostic hooks are exported too
        assert hasattr(lib, name), name
    assert lib.otvm_version() == 5
    assert lib.otvm_strerror(-3).decode().startswith("unsupported")
    assert ctypes.sizeof(_lib.ConvParams) == 208 and ctypes.sizeof(_lib.ReadParams) == 112   # sizeof() of the C structs


def test_clip_sharding_two_ranks_gloo():
    """bench.py's multi-GPU layout: clip c -> rank c mod N, no data-path collective; only the timing max-reduce"""
    env = dict(os.environ, OTVM_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "_gloo_worker.py")],
                       env=env, capture_output=True, text=True, timeout=240)
    assert "OK 2" in r.stdout, r.stdout + r.stderr


def test_trimap_wrapper_runs_alone_and_has_no_cpu_fallback():
    """FullModel_eval (models/trimap/model.py:173) holds exactly the `trimap.*` keys, loads them strictly, packs the
    STM weights without the alpha network, and refuses to compute on a CPU device (no fallback path exists)"""
    import types
    import otvm_b200
    from otvm_b200.engine import FramePlan, PackedWeights
    from otvm_b200.fixtures import make_state_dict
    sd = make_state_dict("tempered")
    tri_sd = {k[len("trimap."):]: v for k, v in sd.items() if k.startswith("trimap.")}
    cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(STAGE=4))
    mt = otvm_b200.get_model_trimap(cfg, "Test", 12)
    assert set(mt.state_dict()) == set(tri_sd)
    mt.load_state_dict(tri_sd)                                   # strict
    with pytest.raises(RuntimeError, match="CUDA device only"):
        mt(None, torch.zeros(1, 3, 32, 32), None, segment=True, memories={"key": None, "val": None})
    with pytest.raises(NotImplementedError):
        mt(torch.zeros(1, 1, 1, 32, 32), torch.zeros(1, 1, 3, 32, 32), torch.zeros(1, 1, 3, 32, 32))
    w = PackedWeights({"trimap." + k: v for k, v in tri_sd.items()}, torch.bfloat16, "cpu", fba=False)
    assert not w.norm and "trimap.model.KV_M_r4.Key" in w.conv and not any(k.startswith("NET.") for k in w.conv)
    assert w.conv["trimap.model.Encoder_M.stem"][0].shape == (64, 7, 7, 32)          # 22 live channels padded to 32
    # STM.memorize / STM.segment pad to 16 (STM.py:204,241), the eval frame to 32 (models/alpha/model.py:408-410)
    p16, p32 = FramePlan(88, 120, torch.float32, "cpu", multiple=16), FramePlan(88, 120, torch.float32, "cpu")
    assert (p16.Hp, p16.Wp, p16.pad_top, p16.pad_left) == (96, 128, 4, 4)
    assert (p32.Hp, p32.Wp) == (96, 128)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm): one JSON line on stdout with the
    contract's keys, timed on the oracle port; under torchrun only rank 0 prints"""
    import json
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count()))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--size", "128", "--memory", "2"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == os.cpu_count()
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True, text=True,
                        timeout=600, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout (build container only)")
def test_dropin_runs_the_eval_py_sequence(tmp_path):
    """eval.py:15,41-94 with otvm_b200/dropin first on PYTHONPATH: the REFERENCE's own helpers.py (`from helpers import *`)
    builds the models through its lazy `import models.trimap.model` / `import models.alpha.model`, which resolve to the
    B200 classes; strict load_state_dict, DataParallel wrap, model.eval(), format_time -- everything eval.py's main()
    touches before and after the clip loop.  (A forward needs a GPU: it must raise here, not fall back.)"""
    script = r'''
import os, sys, types
_popen = os.popen
class _Fake:
    def read(self): return "24 80"
os.popen = lambda cmd, *a, **k: _Fake() if str(cmd).startswith("stty") else _popen(cmd, *a, **k)   # helpers.py:211 needs a TTY
import torch
from torch import nn
from helpers import *                                    # eval.py:15 -- the reference's helpers
import helpers, models.alpha.model as ma_mod, models.trimap.model as mt_mod
assert helpers.__file__.startswith("/root/reference"), helpers.__file__
assert "otvm_b200/dropin" in ma_mod.__file__ and "otvm_b200/dropin" in mt_mod.__file__
cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(STAGE=4), SYSTEM=types.SimpleNamespace(RANDOM_SEED=111))
MODEL = get_model_name(cfg)                              # eval.py:46
assert MODEL == "s4_OTVM"
model_trimap = get_model_trimap(cfg, mode='Test', dilate_kernel=12)          # eval.py:74
model = get_model_alpha(cfg, model_trimap, mode='Test', dilate_kernel=12)    # eval.py:75
import otvm_b200.models as M
assert isinstance(model, M.EvalModel) and isinstance(model_trimap, M.FullModel_eval)
from otvm_b200.fixtures import make_state_dict, make_frame
model.load_state_dict(make_state_dict("tempered"))       # eval.py:79 (strict)
model = nn.DataParallel(model)                           # eval.py:80 (no device here: DataParallel is a pass-through)
model.eval()                                             # eval.py:118
assert model.module.memories == {"key": None, "val": None}
a, fg, bg = make_frame(0, 0, 64, 64)
try:
    model(a, fg, bg, tri=None, tri_gt=None, first_frame=True, last_frame=False, memorize=False, max_memory_num=5, large_input=False)
    raise SystemExit("forward ran without a CUDA device: there must be no CPU fallback")
except RuntimeError as e:
    assert "CUDA" in str(e), e
print("done | Total time: {}".format(format_time(12.5)))  # eval.py:94
'''
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "otvm_b200", "dropin"), ROOT, "/root/reference"]),
               CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "done | Total time:" in r.stdout


def test_read_block_ddp_step_two_ranks_gloo():
    """the DDP plumbing of scripts/train_step_ddp.py on CPU (gloo, world_size 2) with the composite read standing in for
    the CUDA kernels: gradients are all-reduced (ranks stay in sync), bytes per step = fp32 size of the parameters"""
    import json
    env = dict(os.environ, OTVM_TRAIN_CPU="1", STEPS="2", WARMUP="1", FEAT="6", T_MEM="2", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "scripts", "train_step_ddp.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, r.stdout + r.stderr
    d = json.loads(line[-1])
    assert d["n_ranks"] == 2 and d["backend"] == "gloo" and d["ranks_in_sync"] is True
    assert d["allreduce_bytes_per_step"] == 4 * d["parameters"] and d["parameters"] > 14_000_000


def _stage4_against_golden(device, read_fn, tol, tol_grad=None):
    """one stage-4 training step (forward, four losses, backward) of otvm_b200.train_stage4 against the REFERENCE's own
    step on the same weights / sample (tests/golden/train_step_s4.npz, oracle/make_golden_train.py)"""
    import numpy as np
    import torch
    from otvm_b200 import train_stage4 as S4
    from otvm_b200.fixtures import make_state_dict, make_train_sample
    g = np.load(os.path.join(ROOT, "tests", "golden", "train_step_s4.npz"))
    H, W, S = [int(x) for x in g["meta"]]
    m = S4.Stage4Model(read_fn=read_fn)
    m.load_state_dict(make_state_dict("tempered"))                              # strict, 785 keys
    m = m.to(device)
    out = m(*[t.to(device) for t in make_train_sample(0, S, H, W)[:3]], ignore_region=None,
            tri=make_train_sample(0, S, H, W)[3].to(device))
    losses = [o.mean() for o in out[:4]]
    sum(losses).backward()
    for got, want in zip(losses, g["losses"]):
        assert abs(float(got.detach()) - want) <= tol * max(1.0, abs(want)), (float(got.detach()), want)
    named = dict(m.named_parameters())
    gsq = sum(float(v.grad.double().pow(2).sum()) for v in named.values() if v.grad is not None)
    assert abs(gsq ** 0.5 - float(g["grad_norm"])) <= tol * float(g["grad_norm"])
    tol_grad = tol_grad or tol
    for k in g.files:
        if k.startswith("g:"):
            gr = named[k[2:]].grad.flatten().cpu()
            gr = gr[:: max(1, gr.numel() // 256)][:256].numpy()
            assert np.abs(gr - g[k]).max() <= tol_grad * max(np.abs(g[k]).max(), 1e-3), k
    assert np.abs(out[4][0, :, 0].detach().cpu().numpy()[:, ::2, ::2] - g["alphas"]).max() < tol
    assert np.abs(out[5][0].detach().cpu().numpy()[:, :, ::4, ::4] - g["preds_trimap"]).max() < tol


def test_stage4_step_matches_reference_golden_cpu():
    """host logic of the stage-4 step (graph, losses, parameter names) with the composite read standing in for the CUDA
    kernels; the GPU test of the same name in test_gpu_train.py runs it with the fused read"""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from train_step_ddp import composite_read
    _stage4_against_golden("cpu", composite_read, 2e-3)


def test_stage4_ddp_step_two_ranks_gloo():
    """BASELINE configs[4] plumbing on CPU: the whole stage-4 model under DDP (gloo, world_size 2), 295.5 MB of fp32
    gradients all-reduced per step (SURVEY 8(d) cfg 5), ranks in sync afterwards"""
    import json
    env = dict(os.environ, OTVM_TRAIN_CPU="1", MODEL="stage4", SIZE="64", STEPS="1", WARMUP="0", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29619", os.path.join(ROOT, "scripts", "train_step_ddp.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, r.stdout + r.stderr
    d = json.loads(line[-1])
    assert d["n_ranks"] == 2 and d["ranks_in_sync"] is True
    assert d["parameters"] == 73_887_684 and d["allreduce_bytes_per_step"] == 295_550_736
