"""CPU: the v1 / v1.5 vocoder-side restatement (oracle/v1.py: ECAPA-TDNN speaker encoder + latent-conditioned BigVGAN,
SURVEY section 8 row a13) against outputs of the reference's own `indextts.BigVGAN.models.BigVGAN`
(tests/golden/v1_vocoder_small.npz, oracle/make_goldens_v1.py).  The CUDA side of a13 is not built yet."""
import os

import numpy as np
import torch

from indextts_b200 import synth
from oracle import v1

GOLD = os.path.join(os.path.dirname(__file__), "golden", "v1_vocoder_small.npz")


def test_ecapa_and_bigvgan_v1_match_reference_module():
    g = np.load(GOLD)
    h = synth.small_v1_config()
    w = synth.make_bigvgan_v1_weights(h, seed=int(g["seed"]))
    latent, mel_ref = torch.from_numpy(g["latent"]), torch.from_numpy(g["mel_ref"])
    spk = v1.ecapa_tdnn(w, mel_ref).numpy()
    assert spk.shape == g["spk"].shape == (1, 1, h["speaker_embedding_dim"])
    assert np.abs(spk - g["spk"]).max() <= 1e-4 * max(1.0, float(np.abs(g["spk"]).max()))
    wav = v1.bigvgan_v1_forward(h, w, latent, mel_ref).numpy()
    assert wav.shape == g["wav"].shape == (1, 1, latent.shape[1] * 8)
    assert np.abs(wav - g["wav"]).max() <= 1e-5
    assert np.abs(wav).max() <= 1.0                      # tanh output (models.py:247)
    # the speaker embedding really conditions the waveform
    wav2 = v1.bigvgan_v1_forward(h, w, latent, mel_ref.flip(1) * 0.5).numpy()
    assert np.abs(wav2 - wav).max() > 1e-3
