"""Emotion-vector path (row a7): CPU oracle vs goldens from the reference modules; GPU engine vs both."""
import os

import numpy as np
import pytest
import torch

from oracle.emo import EMO_CFG, get_emovec, make_emo_weights, merge_emovec, small_emo_cfg

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,cfg", [("emo_small", small_emo_cfg()), ("emo_full", dict(EMO_CFG))])
def test_oracle_matches_reference_golden(name, cfg):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    w = make_emo_weights(cfg, seed=int(g["seed"]))
    ev = get_emovec(w, cfg, torch.from_numpy(g["feats"]))
    assert np.abs(ev.numpy() - g["emovec"]).max() < 5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name,cfg", [("emo_small", small_emo_cfg()), ("emo_full", dict(EMO_CFG))])
def test_engine_emovec_vs_golden_and_merge(engine, name, cfg):
    """tf32 GEMMs (default): emo_vec within 2e-2 of the reference golden on values of std ~1;
    strict fp32 back end within 2e-3."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    w = make_emo_weights(cfg, seed=int(g["seed"]))
    engine.load_state_dict("gpt.", w)
    engine.emo_init(cfg)
    ev = engine.merge_emovec(g["feats"])
    err = np.abs(ev - g["emovec"]).max()
    print(f"{name}: emovec max err (tf32) {err:.2e}, ref std {g['emovec'].std():.2f}")
    assert err < 2e-2
    engine.set_option("gemm_backend", 1)
    try:
        ev32 = engine.merge_emovec(g["feats"])
    finally:
        engine.set_option("gemm_backend", 0)
    err32 = np.abs(ev32 - g["emovec"]).max()
    print(f"{name}: emovec max err (fp32) {err32:.2e}")
    assert err32 < 2e-3
    # merge with a different emotion reference and alpha (model_v2.py:837)
    gen = torch.Generator().manual_seed(4)
    emo_feats = torch.randn(g["feats"].shape[0] + 9, cfg["idim"], generator=gen)
    ref = merge_emovec(w, cfg, torch.from_numpy(g["feats"]), emo_feats, alpha=0.6).numpy()
    got = engine.merge_emovec(g["feats"], emo_feats.numpy(), alpha=0.6)
    assert np.abs(got - ref).max() < 2e-2
