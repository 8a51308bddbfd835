"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (``/root/reference``) on CPU.

TEST INFRASTRUCTURE.  Runs only in the build container (the GPU box has no ``/root/reference``).  The three
import-time shims are the ones listed in SURVEY.md §8(c); no reference source is edited or copied.

    python oracle/make_golden.py            # writes tests/golden/

Weights and frames come from ``otvm_b200/fixtures.py`` (name-keyed, seed-stable), loaded STRICTLY into the
reference model (``eval.py:77-79``), which also proves ``otvm_b200/spec.py`` names all 785 keys correctly.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OTVM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def import_reference():
    """Import the reference's ``helpers`` with the SURVEY §8(c) shims."""
    _popen = os.popen

    class _Fake:
        def read(self):
            return "24 80"

    os.popen = lambda cmd, *a, **k: _Fake() if str(cmd).startswith("stty") else _popen(cmd, *a, **k)
    import torchvision
    _r50 = torchvision.models.resnet50
    torchvision.models.resnet50 = lambda pretrained=False, **k: _r50(weights=None)
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: "cpu"
    sys.path.insert(0, REF)
    import helpers  # noqa: the reference's helpers.py
    return helpers


def build_reference(helpers, state_dict, dilate_kernel=12):
    cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(STAGE=4))
    mt = helpers.get_model_trimap(cfg, "Test", dilate_kernel=dilate_kernel)
    ma = helpers.get_model_alpha(cfg, mt, "Test", dilate_kernel=dilate_kernel)
    ma.load_state_dict(state_dict)          # strict, like eval.py:79
    return ma.eval()


def run_clip(ma, H, W, n_frames, max_mem, clip=0, keep=None, stride=1):
    """Drive the model like eval.py:170-175 and capture the tensors the parity tests compare."""
    from otvm_b200.fixtures import make_frame
    cap = {}
    hooks = [
        ma.trimap.model.Decoder.register_forward_hook(lambda m, i, o: cap.__setitem__("seg_logit", o)),
        ma.NET.register_forward_hook(lambda m, i, o: cap.__setitem__("net", o)),
        ma.trimap.model.Memory.register_forward_hook(lambda m, i, o: cap.__setitem__("m4", o)),
    ]
    out = {}
    for i in range(n_frames):
        cap.clear()
        a, fg, bg = make_frame(clip, i, H, W)
        r = ma(a, fg, bg, tri=None, tri_gt=None, first_frame=(i == 0), last_frame=False,
               memorize=True, max_memory_num=max_mem)
        if keep is not None and i not in keep:
            continue
        s = stride
        out[f"f{i}_alpha"] = r[3][0, 0, 0].numpy()[::s, ::s].copy()
        out[f"f{i}_trimap"] = r[1][0, 0].numpy()[:, ::s, ::s].copy()
        net = cap["net"]
        out[f"f{i}_dec_alpha"] = net[0][0, 0].numpy()[::s, ::s].copy()
        out[f"f{i}_hid"] = net[1][0].numpy()[:, ::4 * s, ::4 * s].copy()
        out[f"f{i}_refine_fb"] = net[2][0, 1:].numpy()[:, ::4 * s, ::4 * s].copy()
        out[f"f{i}_refine_trimap"] = net[3][0].numpy()[:, ::2 * s, ::2 * s].copy()
        if "seg_logit" in cap:
            out[f"f{i}_seg_logit"] = cap["seg_logit"][0].numpy()[:, ::2 * s, ::2 * s].copy()
            out[f"f{i}_m4_mem"] = cap["m4"][0, :512].numpy()[::8].copy()
        out[f"f{i}_bank_T"] = np.asarray(ma.memories["key"].shape[3])
        out[f"f{i}_key_last"] = ma.memories["key"][0, 0, :, -1].numpy()[::4].copy()
        out[f"f{i}_val_last"] = ma.memories["val"][0, 0, :, -1].numpy()[::16].copy()
    for h in hooks:
        h.remove()
    return out


def user_trimap_cases(ma, H=128, W=128):
    """two 2-frame clips whose first frame is seeded by ``tri=`` / ``tri_gt=`` (models/alpha/model.py:395-401)"""
    from otvm_b200.fixtures import make_frame, user_trimap
    out = {}
    for kind in ("tri", "tri_gt"):
        for i in range(2):
            a, fg, bg = make_frame(3, i, H, W)
            kw = {kind: user_trimap(kind, H, W)} if i == 0 else {}
            r = ma(a, fg, bg, first_frame=(i == 0), last_frame=False, memorize=True, max_memory_num=8, **kw)
            out[f"{kind}_f{i}_alpha"] = r[3][0, 0, 0].numpy().copy()
            out[f"{kind}_f{i}_trimap"] = r[1][0, 0].numpy().copy()
            out[f"{kind}_f{i}_tri_gt"] = r[2][0, 0].numpy().copy()
    return out


def memory_read_cases(helpers):
    """Known-answer vectors for ``Memory.forward`` (STM.py:144-163) straight from the reference class."""
    from models.trimap.STM import Memory
    mem = Memory()
    out = {}
    cases = [  # (T, h, w, key/query scale)
        (1, 4, 4, 1.0), (3, 6, 5, 1.0), (2, 8, 8, 6.0), (5, 7, 9, 0.3), (8, 16, 16, 2.0),
    ]
    for ci, (T, h, w, sc) in enumerate(cases):
        r = np.random.RandomState(1000 + ci)
        t = lambda *s: torch.from_numpy(r.standard_normal(s).astype(np.float32))
        m_in, m_out = t(1, 128, T, h, w) * sc, t(1, 512, T, h, w)
        q_in, q_out = t(1, 128, h, w) * sc, t(1, 512, h, w)
        y = mem(m_in, m_out, q_in, q_out)
        out[f"c{ci}_shape"] = np.asarray([T, h, w])
        out[f"c{ci}_scale"] = np.asarray(sc, np.float32)
        out[f"c{ci}_out"] = y[0].numpy()
    return out


def stm_standalone_inputs(H=88, W=120, seed=5):
    """the inputs of tests/test_gpu_frames.py::test_trimap_wrapper_standalone (re-drawn there from the same seed)"""
    g = torch.Generator().manual_seed(seed)
    frames = [torch.rand(1, 3, H, W, generator=g) for _ in range(3)]
    tri = torch.nn.functional.one_hot(torch.randint(0, 3, (1, H, W), generator=g), 3).permute(0, 3, 1, 2).float()
    alpha = torch.rand(1, 1, H, W, generator=g)
    hid = torch.randn(1, 16, H, W, generator=g) * 0.5
    return frames, tri, alpha, hid


def stm_standalone_cases(helpers, state_dict):
    """``FullModel_eval.forward(memorize=True)`` / ``(segment=True)`` of the reference called directly
    (models/trimap/model.py:247-264) on a size that exercises the pad-16 of STM.memorize / STM.segment."""
    cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(STAGE=4))
    mt = helpers.get_model_trimap(cfg, "Test", dilate_kernel=12)
    mt.load_state_dict({k[len("trimap."):]: v for k, v in state_dict.items() if k.startswith("trimap.")})   # strict
    mt.eval()
    frames, tri, alpha, hid = stm_standalone_inputs()
    out, keys, vals = {}, [], []
    for i, f in enumerate(frames[:2]):
        mem = mt(alpha, None, f, tri=tri, memorize=True, hid=hid)
        out[f"m{i}_key"] = mem["key"][0, 0, :, 0].numpy().copy()
        out[f"m{i}_val"] = mem["val"][0, 0, :, 0].numpy().copy()
        keys.append(mem["key"]); vals.append(mem["val"])
    bank = {"key": torch.cat(keys, dim=3), "val": torch.cat(vals, dim=3)}
    out["seg_logit"] = mt(None, frames[2], None, segment=True, memories=bank)[0].numpy().copy()
    return out


def main():
    from otvm_b200.fixtures import make_state_dict
    torch.set_grad_enabled(False)
    torch.set_num_threads(os.cpu_count())
    helpers = import_reference()
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)

    only = sys.argv[1:]
    if not only or "memory_read" in only:
        np.savez_compressed(os.path.join(gdir, "memory_read.npz"), **memory_read_cases(helpers))
        print("memory_read.npz")
    if not only or "stm_standalone" in only:
        np.savez_compressed(os.path.join(gdir, "stm_standalone.npz"), **stm_standalone_cases(helpers, make_state_dict("tempered")))
        print("stm_standalone.npz")

    if not only or "user_trimap_128" in only:
        ma = build_reference(helpers, make_state_dict("tempered"))
        np.savez_compressed(os.path.join(gdir, "user_trimap_128.npz"), **user_trimap_cases(ma))
        print("user_trimap_128.npz")

    jobs = [  # name, kind, H, W, frames, max_mem, keep, stride
        ("clip_tempered_256", "tempered", 256, 256, 3, 8, None, 1),
        ("clip_default_128", "default", 128, 128, 3, 8, None, 1),
        ("clip_tempered_120x152", "tempered", 120, 152, 3, 2, None, 1),     # pad-to-32 path + eviction (T<=2)
        ("clip_tempered_512_T8", "tempered", 512, 512, 10, 8, (8, 9), 4),   # BASELINE configs[1], strided sample
    ]
    for name, kind, H, W, n, mm, keep, stride in jobs:
        if only and name not in only:
            continue
        ma = build_reference(helpers, make_state_dict(kind))
        out = run_clip(ma, H, W, n, mm, keep=keep, stride=stride)
        out["meta"] = np.asarray([H, W, n, mm, stride])
        np.savez_compressed(os.path.join(gdir, name + ".npz"), **out)
        a = out[f"f{n - 1}_alpha"]
        print(name, "alpha range", float(a.min()), float(a.max()), "mean", float(a.mean()))


if __name__ == "__main__":
    main()
