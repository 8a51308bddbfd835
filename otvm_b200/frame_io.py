"""Frame I/O around the per-frame loop.

What the reference does per frame, all on the main thread (``num_workers=0``, ``eval.py:136-140``):
``cv2.imread`` of the foreground (BGRA PNG, alpha in the 4th channel) and the background, ``np.float32`` / ``/ 255`` /
HWC->CHW on the host (``dataset.py:857-920``), a 28 B/pixel fp32 upload, the model, ``(alphas * 255).byte().cpu()``
(a synchronous fp32 read-back + host conversion, ``eval.py:209``) and ``cv2.imwrite`` of the PNG (``eval.py:217``).
Once the frame itself takes a few milliseconds on a B200 that host work IS the loop.  Here:

* :class:`FrameSource` -- reader threads decode frames ahead of the loop (``cv2.imread`` releases the GIL) into PINNED
  8-bit buffers; the upload is the 7 B/pixel the decoder produced, on a copy stream, and ``otvm_unpack_frame_u8``
  turns it into the exact fp32 tensors ``EvalModel.forward`` takes (bit-identical to the host decode).
* :class:`AlphaWriter` -- ``otvm_alpha_to_u8`` converts the matte on the device (bit-identical to
  ``(alphas * 255).byte()``), the 1 B/pixel result goes to a pinned ring asynchronously, and a writer thread encodes
  the PNGs while the GPU runs the next frames.
* :func:`run_sequence` -- the ``eval.py:162-217`` loop of one clip on top of the two.

PyTorch only owns memory and streams here; decoding / encoding stay with OpenCV on host threads (nvJPEG is not in this
image, and the VideoMatting108 foregrounds are PNGs).
"""
from __future__ import annotations

import os
import queue
import threading
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops


def _cv2():
    import cv2
    return cv2


class FrameSource:
    """Iterates ``(a, fg, bg, name)`` device tensors shaped like an ``eval.py`` batch ([1,1,C,H,W] fp32, BGR 0..255).

    ``fg_paths[i]`` is a BGRA (or BGR) image, ``bg_paths[i]`` a BGR image (``None``: the foreground is its own
    background, ``dataset.py:893-894``).  ``depth`` frames are decoded ahead by ``workers`` threads."""

    def __init__(self, fg_paths: Sequence[str], bg_paths: Optional[Sequence[str]] = None, device="cuda", depth: int = 4,
                 workers: int = 2):
        self.fg_paths, self.bg_paths = list(fg_paths), list(bg_paths) if bg_paths is not None else None
        self.device = torch.device(device)
        self.depth = max(2, depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots: List[dict] = []                   # pinned staging + device buffers, allocated on the first frame
        self._jobs: "queue.Queue" = queue.Queue()
        self._done = [threading.Event() for _ in self.fg_paths]
        self._decoded: List[Optional[tuple]] = [None] * len(self.fg_paths)
        self._next_job = 0
        self._lock = threading.Lock()
        self._threads = [threading.Thread(target=self._worker, daemon=True) for _ in range(max(1, workers))]
        for t in self._threads:
            t.start()
        for _ in range(min(self.depth, len(self.fg_paths))):
            self._submit()

    def __len__(self):
        return len(self.fg_paths)

    # -- host side ------------------------------------------------------------------------------------------
    def _submit(self):
        with self._lock:
            if self._next_job < len(self.fg_paths):
                self._jobs.put(self._next_job)
                self._next_job += 1

    def _worker(self):
        cv2 = _cv2()
        while True:
            i = self._jobs.get()
            if i is None:
                return
            try:
                f = cv2.imread(self.fg_paths[i], cv2.IMREAD_UNCHANGED)            # dataset.py:860
                if f is None:
                    raise FileNotFoundError(self.fg_paths[i])
                if f.ndim == 2:
                    f = np.stack([f] * 3, axis=-1)
                if self.bg_paths is None:
                    b = np.ascontiguousarray(f[..., :3])
                else:
                    bp = self.bg_paths[i]
                    if not os.path.exists(bp):
                        bp = os.path.splitext(bp)[0] + ".png"                     # dataset.py:897-898
                    b = cv2.imread(bp, cv2.IMREAD_COLOR)
                    if b is None:
                        raise FileNotFoundError(bp)
                self._decoded[i] = (np.ascontiguousarray(f), np.ascontiguousarray(b))
            except Exception as e:                                                 # surfaced by __iter__
                self._decoded[i] = e
            self._done[i].set()

    def close(self):
        for _ in self._threads:
            self._jobs.put(None)

    # -- device side ----------------------------------------------------------------------------------------
    def _slot(self, k, H, W, c):
        while len(self._slots) <= k:
            self._slots.append({})
        s = self._slots[k]
        if s.get("shape") != (H, W, c):
            s.update(shape=(H, W, c),
                     fg_h=torch.empty(H, W, c, dtype=torch.uint8).pin_memory(),
                     bg_h=torch.empty(H, W, 3, dtype=torch.uint8).pin_memory(),
                     fg_d=torch.empty(H, W, c, dtype=torch.uint8, device=self.device),
                     bg_d=torch.empty(H, W, 3, dtype=torch.uint8, device=self.device),
                     a=torch.empty(1, 1, 1, H, W, device=self.device),
                     fg=torch.empty(1, 1, 3, H, W, device=self.device),
                     bg=torch.empty(1, 1, 3, H, W, device=self.device), free=None, h2d=None)
        return s

    def _upload(self, i):
        """decoded frame i -> device tensors of slot i % depth (on the copy stream); returns the slot + a ready event"""
        self._done[i].wait()
        d = self._decoded[i]
        if isinstance(d, Exception):
            raise d
        f, b = d
        self._decoded[i] = None
        H, W, c = f.shape
        if b.shape[:2] != (H, W):
            raise ValueError(f"foreground {f.shape} and background {b.shape} sizes differ: {self.fg_paths[i]}")
        s = self._slot(i % self.depth, H, W, c)
        if s["h2d"] is not None:
            s["h2d"].synchronize()                          # the previous upload from this pinned staging has left the host
        with torch.cuda.stream(self.stream):
            if s["free"] is not None:
                self.stream.wait_event(s["free"])           # the consumer of this slot's previous frame has finished
            s["fg_h"].numpy()[...] = f                      # (pinned staging: the H2D below is truly asynchronous)
            s["bg_h"].numpy()[...] = b
            s["fg_d"].copy_(s["fg_h"], non_blocking=True)
            s["bg_d"].copy_(s["bg_h"], non_blocking=True)
            s["h2d"] = torch.cuda.Event()
            s["h2d"].record(self.stream)
            ops.unpack_frame_u8(s["fg_d"], s["bg_d"], s["a"], s["fg"], s["bg"])
            ready = torch.cuda.Event()
            ready.record(self.stream)
        self._submit()                                      # keep the decoders `depth` frames ahead
        return s, ready

    def __iter__(self):
        n = len(self.fg_paths)
        pending = [self._upload(i) for i in range(min(2, n))]          # upload runs one frame ahead of the consumer
        for i in range(n):
            s, ready = pending.pop(0)
            torch.cuda.current_stream().wait_event(ready)
            name = os.path.splitext(os.path.basename(self.fg_paths[i]))[0] + ".jpg"      # dataset.py:907
            yield s["a"], s["fg"], s["bg"], name
            s["free"] = torch.cuda.Event()
            s["free"].record(torch.cuda.current_stream())   # everything the consumer queued on these tensors
            if i + 2 < n:
                pending.append(self._upload(i + 2))
        self.close()


class AlphaWriter:
    """``write(alpha, path)``: queue the 8-bit matte of a frame for writing; returns immediately.

    The conversion and the D2H copy are asynchronous on a side stream into a ring of pinned buffers; a writer thread
    waits for the copy's event and encodes the PNG (``cv2.imwrite``, ``eval.py:217``).  ``close()`` drains."""

    def __init__(self, device="cuda", depth: int = 8, workers: int = 2, keep: bool = False):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self._ring: List[dict] = [{} for _ in range(depth)]
        self._i = 0
        self._q: "queue.Queue" = queue.Queue()
        self.keep = keep                                  # tests: keep the u8 arrays instead of (only) writing files
        self.kept: dict = {}
        self._threads = [threading.Thread(target=self._worker, daemon=True) for _ in range(max(1, workers))]
        for t in self._threads:
            t.start()

    def _worker(self):
        cv2 = _cv2()
        while True:
            job = self._q.get()
            if job is None:
                return
            slot, path, ev = job
            ev.synchronize()
            img = slot["host"].numpy()
            if self.keep:
                self.kept[path] = img.copy()
            if path is not None:
                os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
                cv2.imwrite(path, img)
            slot["busy"].set()

    def write(self, alpha: torch.Tensor, path: Optional[str]):
        H, W = alpha.shape[-2:]
        slot = self._ring[self._i % self.depth]
        self._i += 1
        if slot.get("shape") != (H, W):
            slot.update(shape=(H, W), dev=torch.empty(H, W, dtype=torch.uint8, device=self.device),
                        host=torch.empty(H, W, dtype=torch.uint8).pin_memory(), busy=threading.Event())
            slot["busy"].set()
        slot["busy"].wait()                               # the writer thread has finished with this slot's previous frame
        slot["busy"].clear()
        produced = torch.cuda.Event()
        produced.record(torch.cuda.current_stream())      # the frame that produced `alpha`
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(produced)
            ops.alpha_to_u8(alpha if alpha.is_contiguous() else alpha.contiguous(), slot["dev"])
            converted = torch.cuda.Event()
            converted.record(self.stream)
            slot["host"].copy_(slot["dev"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        # the model reuses its output buffer for the next frame: that frame must wait until the conversion has read it
        torch.cuda.current_stream().wait_event(converted)
        self._q.put((slot, path, ev))

    def close(self):
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()


def run_sequence(model, fg_paths, bg_paths, out_dir, *, max_memory_num=5, memory_skip_frame=1, tri_first=None,
                 tri_gt_first=None, writer: Optional[AlphaWriter] = None, source: Optional[FrameSource] = None):
    """One clip through ``model`` the way ``eval.py:162-217`` drives it (first / last frame flags, ``memorize`` every
    ``memory_skip_frame`` frames when that is > 2, ``eval.py:180-188``), with the frame I/O above.  Returns the number
    of frames; the mattes are in ``out_dir/<name>.png``."""
    own_w, own_s = writer is None, source is None
    writer = writer or AlphaWriter(model.IMG_MEAN.device)
    source = source or FrameSource(fg_paths, bg_paths, model.IMG_MEAN.device)
    n = len(source)
    for i, (a, fg, bg, name) in enumerate(source):
        first, last = i == 0, i == n - 1
        memorize = (i % memory_skip_frame) == 0 if memory_skip_frame > 2 else False
        out = model(a, fg, bg, tri=tri_first if first else None, tri_gt=tri_gt_first if first else None,
                    first_frame=first, last_frame=last, memorize=memorize, max_memory_num=max_memory_num)
        writer.write(out[3], os.path.join(out_dir, os.path.splitext(name)[0] + ".png") if out_dir else None)
    if own_w:
        writer.close()
    if own_s:
        source.close()
    return n
