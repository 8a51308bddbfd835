"""Factories with the reference's names and signatures (helpers.py:323-362), so that
``helpers.get_model_trimap(cfg, 'Test', dilate_kernel)`` / ``helpers.get_model_alpha(cfg, model_trimap, 'Test',
dilate_kernel)`` in ``eval.py:74-75`` return the B200 implementation."""
from __future__ import annotations

from .models import EvalModel, FullModel_eval


def get_model_name(cfg):
    return {1: "s1_OTVM_alpha", 2: "s2_OTVM_alpha", 3: "s3_OTVM", 4: "s4_OTVM"}[cfg.TRAIN.STAGE]


def get_model_trimap(cfg, mode="Test", dilate_kernel=None):
    if mode != "Test":
        raise NotImplementedError("otvm_b200 covers the inference path (mode='Test') only")
    return FullModel_eval(eps=0, stage=cfg.TRAIN.STAGE, dilate_kernel=dilate_kernel, hdim=16)


def get_model_alpha(cfg, model_trimap, mode="Test", dilate_kernel=None):
    if mode != "Test":
        raise NotImplementedError("otvm_b200 covers the inference path (mode='Test') only")
    return EvalModel(dilate_kernel=dilate_kernel, trimap=model_trimap, stage=cfg.TRAIN.STAGE)
