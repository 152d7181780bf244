"""The per-segment tail as the reference wires it (indextts/infer_v2_5.py:830-855: codec decode -> length regulator ->
cat with the prompt condition -> cfm.inference -> drop the prompt frames -> BigVGAN -> int16 scaling).  The golden is the
output of those very source lines, executed by oracle/make_goldens_tail.py against the reference's own modules with the seeded
weights.  CPU: the restated oracle chain.  GPU: `idx_codes_to_wav`, the single call bench.py times (runs last in the suite)."""
import os

import numpy as np
import pytest
import torch

from indextts_b200 import synth
from oracle.bigvgan import bigvgan_forward
from oracle.s2mel import cfm_inference, codec_decode, length_regulate
from oracle.s2mel import fold_weight_norm as oracle_fold

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tail_wiring_small.npz")


def _weights():
    c, cc, h = synth.small_s2mel_cfg(), synth.small_codec_cfg(), synth.small_config()
    return c, cc, h, synth.make_s2mel_weights(c, 1234), synth.make_codec_weights(cc, 4321), synth.make_bigvgan_weights(h, 1)


def test_oracle_chain_matches_the_executed_reference_tail():
    g = np.load(GOLD)
    c, cc, h, ws, wc, wb = _weights()
    wsf, wcf = oracle_fold(ws), oracle_fold(wc)
    codes = torch.from_numpy(g["codes"])
    S = codec_decode(wcf, codes)
    F = int(S.shape[1] * 1.72)
    assert F == int(g["F"])
    cond = length_regulate(wsf, S, F)
    mu = torch.cat([torch.from_numpy(g["prompt_condition"]), cond], 1)
    P = g["ref_mel"].shape[-1]
    mel = cfm_inference(wsf, c, mu, torch.LongTensor([mu.size(1)]), torch.from_numpy(g["ref_mel"]), torch.from_numpy(g["style"]),
                        torch.from_numpy(g["z"]), 25, 0.7)
    wav = torch.clamp(32767 * bigvgan_forward(h, wb, mel[:, :, P:].float()).squeeze().unsqueeze(0), -32767.0, 32767.0).numpy()
    assert wav.shape == g["wav"].shape and np.abs(wav - g["wav"]).max() < 2.0          # of +-32767


@pytest.mark.gpu
def test_codes_to_wav_matches_the_executed_reference_tail(engine):
    from indextts_b200.engine import fold_weight_norm
    g = np.load(GOLD)
    c, cc, h, ws, wc, wb = _weights()
    engine.load_state_dict("s2mel.", {k: v for k, v in fold_weight_norm(ws).items() if v.is_floating_point()})
    engine.load_state_dict("codec.", fold_weight_norm(wc))
    engine.load_state_dict("bigvgan.", wb)
    engine.s2mel_init(c); engine.codec_init(cc); engine.bigvgan_init(h)
    ref = g["wav"][0] / 32767.0
    rms = float(np.sqrt((ref ** 2).mean()))
    args = (g["codes"][0].astype(np.int32), g["prompt_condition"][0], g["ref_mel"][0], g["style"][0], g["z"][0], int(g["F"]), 25, 0.7)
    # strict fp32 back end, then tf32 (25 Euler steps of tf32 GEMMs in front of the vocoder: same order as the bounds of
    # tests/test_dropin_gpu.py; a wiring mistake would show as an O(1) error)
    for backend, tol in ((1, 3e-3), (0, 4e-2)):
        engine.set_option("gemm_backend", backend)
        try:
            res = engine.codes_to_wav(*args, want_wav=True, want_pcm16=True)
        finally:
            engine.set_option("gemm_backend", 0)
        wav = np.clip(res["wav"], -1.0, 1.0)
        err = float(np.sqrt(((wav - ref) ** 2).mean())) / rms
        print(f"[backend {backend}] idx_codes_to_wav vs the executed reference tail: relative rms error {err:.2e}")
        assert wav.shape == ref.shape and err < tol
        assert np.abs(res["pcm16"].astype(np.float32) - np.clip(np.round(wav * 32767.0), -32767, 32767)).max() <= 1.0
