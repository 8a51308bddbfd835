"""The oracle (oracle/otvm_oracle.py) against the reference's own outputs (tests/golden, made by
oracle/make_golden.py from the unmodified reference).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from util import ROOT, golden, rel_err

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import otvm_oracle as O  # noqa: E402
from otvm_b200.fixtures import make_frame, make_state_dict  # noqa: E402

TOL = 2e-4   # fp32 CPU vs fp32 CPU, same library kernels, different association order only


def test_memory_read_known_answers():
    g = golden("memory_read")
    ci = 0
    while f"c{ci}_out" in g:
        T, h, w = (int(v) for v in g[f"c{ci}_shape"])
        sc = float(g[f"c{ci}_scale"])
        r = np.random.RandomState(1000 + ci)
        t = lambda *s: torch.from_numpy(r.standard_normal(s).astype(np.float32))
        m_in, m_out = t(1, 128, T, h, w) * sc, t(1, 512, T, h, w)
        q_in, q_out = t(1, 128, h, w) * sc, t(1, 512, h, w)
        y = O.memory_read(m_in, m_out, q_in, q_out)
        assert rel_err(y[0], g[f"c{ci}_out"]) < 1e-5, ci
        ci += 1
    assert ci == 5


def test_edt_matches_cv2_and_bruteforce():
    r = np.random.RandomState(3)
    for shape, p in (((37, 53), 0.02), ((64, 64), 0.3), ((5, 90), 0.5), ((48, 48), 0.001)):
        m = r.uniform(size=shape) > p          # nonzero = not a seed
        m[r.randint(shape[0]), r.randint(shape[1])] = False
        d2 = O.edt_sq(m)
        yy, xx = np.indices(shape)
        zy, zx = np.nonzero(~m)
        brute = ((yy[..., None] - zy) ** 2 + (xx[..., None] - zx) ** 2).min(-1)
        assert np.array_equal(d2, brute)
        try:
            import cv2
        except ImportError:
            continue
        d = cv2.distanceTransform(m.astype(np.uint8) * 255, cv2.DIST_L2, 0)
        # utils/utils.py:21 — cv2 returns sqrt(exact integer squared distance); its float sqrt is within 1 ulp
        assert np.array_equal(np.rint(d.astype(np.float64) ** 2).astype(np.int64), d2)
        assert np.allclose(d, np.sqrt(d2.astype(np.float32)), rtol=2.5e-7, atol=0)


def _check_clip(name, kind, frames=None):
    g = golden(name)
    H, W, n, mm, s = (int(v) for v in g["meta"])
    model = O.OracleEvalModel(make_state_dict(kind), dilate_kernel=12)
    for i in range(n):
        a, fg, bg = make_frame(0, i, H, W)
        out = model(a, fg, bg, first_frame=(i == 0), last_frame=False, memorize=True, max_memory_num=mm)
        if f"f{i}_alpha" not in g:
            continue
        tr = model.trace
        assert int(g[f"f{i}_bank_T"]) == model.memories["key"].shape[3]
        assert rel_err(out[3][0, 0, 0][::s, ::s], g[f"f{i}_alpha"]) < TOL
        assert rel_err(out[1][0, 0][:, ::s, ::s], g[f"f{i}_trimap"]) < TOL
        assert rel_err(tr["output"][0, 0][::s, ::s], g[f"f{i}_dec_alpha"]) < TOL
        assert rel_err(tr["hid"][0][:, ::4 * s, ::4 * s], g[f"f{i}_hid"]) < TOL
        assert rel_err(tr["refine_output"][0, 1:][:, ::4 * s, ::4 * s], g[f"f{i}_refine_fb"]) < TOL
        assert rel_err(tr["refine_trimap"][0][:, ::2 * s, ::2 * s], g[f"f{i}_refine_trimap"]) < TOL
        if f"f{i}_seg_logit" in g:
            # the golden holds the Decoder output before STM.segment's crop; sizes here need no pad-16 crop
            Hp, Wp = g[f"f{i}_seg_logit"].shape[1:]
            got = tr["seg_logit"][0][:, ::2 * s, ::2 * s]
            if got.shape[1:] == (Hp, Wp):
                assert rel_err(got, g[f"f{i}_seg_logit"]) < TOL
            assert rel_err(tr["m4"][0, :512][::8], g[f"f{i}_m4_mem"]) < TOL
        assert rel_err(model.memories["key"][0, 0, :, -1][::4], g[f"f{i}_key_last"]) < TOL
        assert rel_err(model.memories["val"][0, 0, :, -1][::16], g[f"f{i}_val_last"]) < TOL


def test_clip_tempered_256():
    _check_clip("clip_tempered_256", "tempered")


def test_clip_default_128():
    _check_clip("clip_default_128", "default")


def test_clip_padded_120x152_with_eviction():
    _check_clip("clip_tempered_120x152", "tempered")


@pytest.mark.skipif(os.environ.get("OTVM_SLOW", "0") != "1", reason="set OTVM_SLOW=1 (10 frames at 512x512 on CPU)")
def test_clip_tempered_512_T8():
    _check_clip("clip_tempered_512_T8", "tempered")


def test_user_trimap_first_frame():
    """frame 0 seeded by ``tri=`` (BGR trimap image) or ``tri_gt=`` (one-hot), models/alpha/model.py:395-401: the oracle
    against the unmodified reference's outputs (tests/golden/user_trimap_128.npz), both frames of each clip"""
    from otvm_b200.fixtures import make_frame, user_trimap
    g = golden("user_trimap_128")
    sd = make_state_dict("tempered")
    for kind in ("tri", "tri_gt"):
        model = O.OracleEvalModel(sd, dilate_kernel=12)
        for i in range(2):
            a, fg, bg = make_frame(3, i, 128, 128)
            kw = {kind: user_trimap(kind, 128, 128)} if i == 0 else {}
            out = model(a, fg, bg, first_frame=(i == 0), last_frame=False, memorize=True, max_memory_num=8, **kw)
            assert rel_err(out[3][0, 0, 0], g[f"{kind}_f{i}_alpha"]) < TOL, (kind, i)
            assert rel_err(out[1][0, 0], g[f"{kind}_f{i}_trimap"]) < TOL, (kind, i)
            assert rel_err(out[2][0, 0], g[f"{kind}_f{i}_tri_gt"]) < TOL, (kind, i)


def _stm_standalone_inputs(H=88, W=120, seed=5):
    """same draw as oracle/make_golden.py::stm_standalone_inputs and tests/test_gpu_frames.py"""
    g = torch.Generator().manual_seed(seed)
    frames = [torch.rand(1, 3, H, W, generator=g) for _ in range(3)]
    tri = torch.nn.functional.one_hot(torch.randint(0, 3, (1, H, W), generator=g), 3).permute(0, 3, 1, 2).float()
    alpha = torch.rand(1, 1, H, W, generator=g)
    hid = torch.randn(1, 16, H, W, generator=g) * 0.5
    return frames, tri, alpha, hid


def test_stm_memorize_segment_match_reference_wrapper():
    """oracle STM.memorize / STM.segment == the reference's FullModel_eval.forward(memorize=) / (segment=) called
    directly (models/trimap/model.py:247-264) on an 88x120 frame (pad-16 path); golden from the unmodified reference"""
    g = golden("stm_standalone")
    sd = make_state_dict("tempered")
    frames, tri, alpha, hid = _stm_standalone_inputs()
    keys, vals = [], []
    with torch.no_grad():
        for i, f in enumerate(frames[:2]):
            k4, v4 = O.stm_memorize(sd, f, torch.cat([tri, alpha, hid], dim=1))
            assert rel_err(k4[0, :, 0], g[f"m{i}_key"]) < TOL and rel_err(v4[0, :, 0], g[f"m{i}_val"]) < TOL
            keys.append(k4); vals.append(v4)
        logit, _ = O.stm_segment(sd, frames[2], torch.cat(keys, dim=2), torch.cat(vals, dim=2))
    assert logit.shape == (1, 3, 88, 120)
    assert rel_err(logit[0], g["seg_logit"]) < TOL
