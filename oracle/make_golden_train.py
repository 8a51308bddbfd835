"""Golden vectors of the reference's STAGE-4 TRAINING step (forward, four losses, backward) -> tests/golden/train_step_s4.npz.

TEST INFRASTRUCTURE, build container only (imports the UNMODIFIED /root/reference with the shims of make_golden.py).
The reference objects are the ones train.py builds (helpers.get_model_trimap / get_model_alpha with mode='Train',
train.py:93-97), put in the mode train.py:311-316 puts them in (model.train(), BatchNorm of the trimap network in eval),
called like train.py:357 and differentiated like train.py:358-372.

    python oracle/make_golden_train.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden import import_reference  # noqa: E402

GRAD_KEYS = ["NET.encoder.conv1.weight", "NET.encoder.layer3.2.conv2.weight", "NET.decoder.conv_up1.0.weight",
             "NET.refine.pred.4.weight", "NET.refine.layer1.bn1.weight", "trimap.model.Encoder_M.conv1_h.weight",
             "trimap.model.Encoder_Q.res3.1.conv2.weight", "trimap.model.KV_M_r4.Key.weight",
             "trimap.model.KV_Q_r4.Value.bias", "trimap.model.Decoder.pred.weight"]


def main():
    from otvm_b200.fixtures import make_state_dict, make_train_sample
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    helpers = import_reference()
    from torch import nn
    cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(STAGE=4))
    mt = helpers.get_model_trimap(cfg, "Train", dilate_kernel=None)
    ma = helpers.get_model_alpha(cfg, mt, "Train", dilate_kernel=None)
    ma.load_state_dict(make_state_dict("tempered"))                    # strict
    ma.train()                                                         # train.py:311
    for m in ma.trimap.modules():                                      # train.py:312-316
        if isinstance(m, nn.BatchNorm2d):
            m.eval()
    H = W = 64
    a, fg, bg, tri = make_train_sample(0, 3, H, W)
    out = ma(a, fg, bg, ignore_region=None, tri=tri)                   # train.py:357
    losses = [out[i].mean() for i in range(4)]                         # train.py:358-367
    loss = sum(losses)
    ma.zero_grad()
    loss.backward()
    res = {"meta": np.asarray([H, W, 3]), "losses": np.asarray([float(l) for l in losses], np.float64),
           "alphas": out[6][0, :, 0].detach().numpy()[:, ::2, ::2].copy(),
           "preds_trimap": out[11][0].detach().numpy()[:, :, ::4, ::4].copy()}
    named = dict(ma.named_parameters())
    gsq = 0.0
    for k, v in named.items():
        if v.grad is not None:
            gsq += float(v.grad.double().pow(2).sum())
    res["grad_norm"] = np.asarray(gsq ** 0.5)
    for k in GRAD_KEYS:
        g = named[k].grad.detach().flatten()
        res["g:" + k] = g[:: max(1, g.numel() // 256)][:256].numpy().copy()
        res["gn:" + k] = np.asarray(float(g.double().norm()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "train_step_s4.npz"), **res)
    print("losses", res["losses"], "grad_norm", float(res["grad_norm"]))


if __name__ == "__main__":
    main()
