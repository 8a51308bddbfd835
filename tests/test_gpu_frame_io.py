"""Frame I/O around the loop (otvm_b200/frame_io.py) against the reference's host-side
recipe (dataset.py:857-920 decode, eval.py:209-217 write-back), restated here with numpy / cv2: bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _write_clip(d, n, H, W, seed=0, with_alpha=True):
    r = np.random.RandomState(seed)
    fgs, bgs = [], []
    for i in range(n):
        fg = r.randint(0, 256, (H, W, 4 if with_alpha else 3)).astype(np.uint8)
        if with_alpha:
            yy, xx = np.mgrid[0:H, 0:W]
            fg[..., 3] = np.clip(255 - (np.hypot(yy - H / 2, xx - W / 2 - 2 * i) - H / 4) * 12, 0, 255).astype(np.uint8)
        bg = r.randint(0, 256, (H, W, 3)).astype(np.uint8)
        fp, bp = os.path.join(d, f"fg_{i:04d}.png"), os.path.join(d, f"bg_{i:04d}.png")
        cv2.imwrite(fp, fg); cv2.imwrite(bp, bg)
        fgs.append(fp); bgs.append(bp)
    return fgs, bgs


def _reference_decode(fgp, bgp):
    """dataset.py:857-905 (EvalDataset.get_data with a = tri_gt = None), on the host"""
    _f = cv2.imread(fgp, cv2.IMREAD_UNCHANGED)
    fg = np.float32(_f[..., :-1]); a = np.float32(_f[..., -1:]) / 255.
    bg = np.float32(cv2.imread(bgp, cv2.IMREAD_COLOR))
    t = lambda x: torch.from_numpy(x).permute(2, 0, 1).unsqueeze(0).float()
    return t(fg), t(bg), t(a)


@pytest.mark.parametrize("H,W", [(64, 96), (250, 333)])
def test_frame_source_is_bit_identical_to_the_host_decode(tmp_path, H, W):
    from otvm_b200.frame_io import FrameSource
    fgs, bgs = _write_clip(str(tmp_path), 7, H, W)
    src = FrameSource(fgs, bgs, depth=3, workers=2)
    got = [(a.clone(), fg.clone(), bg.clone(), name) for a, fg, bg, name in src]
    torch.cuda.synchronize()
    assert len(got) == 7
    for i, (a, fg, bg, name) in enumerate(got):
        rfg, rbg, ra = _reference_decode(fgs[i], bgs[i])
        assert name == f"fg_{i:04d}.jpg"
        assert a.shape == (1, 1, 1, H, W) and fg.shape == (1, 1, 3, H, W)
        assert torch.equal(a[0].cpu(), ra) and torch.equal(fg[0].cpu(), rfg) and torch.equal(bg[0].cpu(), rbg)


def test_alpha_writer_is_bit_identical_to_the_host_conversion(tmp_path):
    from otvm_b200.frame_io import AlphaWriter
    g = torch.Generator(device="cuda").manual_seed(1)
    w = AlphaWriter(depth=3, workers=2, keep=True)
    alphas = []
    for i in range(9):
        a = torch.rand(1, 1, 1, 120, 200, device="cuda", generator=g)
        a[..., :5, :] = 0.0; a[..., -5:, :] = 1.0; a[..., 60, 100] = (i + 0.5) / 255.0
        alphas.append(a.clone())
        w.write(a, str(tmp_path / f"a{i}.png"))
        a.zero_()                                   # the caller may reuse its buffer right away (stream-ordered)
    w.close()
    for i, a in enumerate(alphas):
        want = (a * 255).byte().cpu().squeeze(0).squeeze(0).squeeze(0).numpy()          # eval.py:209
        assert np.array_equal(w.kept[str(tmp_path / f"a{i}.png")], want)
        assert np.array_equal(cv2.imread(str(tmp_path / f"a{i}.png"), cv2.IMREAD_UNCHANGED), want)


def test_run_sequence_matches_the_plain_loop(tmp_path):
    """the eval.py:162-217 loop of one clip with the I/O pipeline == the same frames decoded on the host and pushed
    through EvalModel.forward one by one with a synchronous read-back"""
    from frames_util import build_model
    from otvm_b200.frame_io import run_sequence
    H, W, n = 96, 128, 6
    os.makedirs(tmp_path / "in")
    fgs, bgs = _write_clip(str(tmp_path / "in"), n, H, W, seed=3)
    model, _ = build_model("tempered", "bf16x2")
    assert run_sequence(model, fgs, bgs, str(tmp_path / "out"), max_memory_num=5, memory_skip_frame=1) == n
    model2, _ = build_model("tempered", "bf16x2")
    for i in range(n):
        rfg, rbg, ra = _reference_decode(fgs[i], bgs[i])
        out = model2(ra.unsqueeze(0).cuda(), rfg.unsqueeze(0).cuda(), rbg.unsqueeze(0).cuda(), first_frame=(i == 0),
                     last_frame=(i == n - 1), memorize=False, max_memory_num=5)
        want = (out[3] * 255).byte().cpu().squeeze(0).squeeze(0).squeeze(0).numpy()
        got = cv2.imread(str(tmp_path / "out" / f"fg_{i:04d}.png"), cv2.IMREAD_UNCHANGED)
        assert got is not None and got.shape == (H, W)
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 1, i     # GroupNorm atomics order: last-bit noise only
