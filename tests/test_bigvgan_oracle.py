"""CPU: BigVGAN restatement against goldens minted from the reference module."""
import os

import numpy as np
import torch

from oracle.bigvgan import (BIGVGAN_V2_22K, activation1d, bigvgan_forward, kaiser_sinc_filter1d,
                            make_bigvgan_weights, small_config)

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_activation1d_known_answer():
    g = np.load(os.path.join(GOLD, "activation1d_kat.npz"))
    filt = kaiser_sinc_filter1d()
    assert np.array_equal(filt.numpy().reshape(-1), g["filt"])
    # SURVEY.md Appendix A.1 taps
    a1 = [0.0020289647, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832, 0.4432097971]
    assert np.abs(g["filt"][:6] - np.array(a1)).max() < 1e-9 and np.array_equal(g["filt"], g["filt"][::-1])
    y = activation1d(torch.from_numpy(g["x"]), torch.from_numpy(g["alpha"]), torch.from_numpy(g["beta"]), filt)
    assert np.array_equal(y.numpy(), g["y"])
    yr = activation1d(torch.from_numpy(g["xr"]), torch.from_numpy(g["alpha_r"]), torch.from_numpy(g["beta_r"]), filt)
    assert np.array_equal(yr.numpy(), g["yr"])


def test_small_generator_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "bigvgan_small.npz"))
    h = small_config()
    w = make_bigvgan_weights(h, seed=int(g["seed"]))
    wav = bigvgan_forward(h, w, torch.from_numpy(g["mel"]))
    assert wav.shape == (2, 1, 40 * 8)
    assert np.abs(wav.numpy() - g["wav"]).max() < 1e-5


def test_full_generator_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "bigvgan_full_f12.npz"))
    h = dict(BIGVGAN_V2_22K)
    w = make_bigvgan_weights(h, seed=int(g["seed"]))
    assert sum(v.numel() for k, v in w.items() if "filter" not in k) == 112199472
    wav = bigvgan_forward(h, w, torch.from_numpy(g["mel"]))
    assert wav.shape == (1, 1, 12 * 256)
    assert np.abs(wav.numpy() - g["wav"]).max() < 1e-5
