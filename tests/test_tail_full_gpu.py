"""North-star parity of the codes -> wav tail at the FULL BASELINE config-2 geometry (256 codes, 10 s reference:
P = 861, F = 880, T = 1741, 25 Euler steps, CFG 0.7, BigVGAN-v2 22 kHz -> 225 280 samples) in the precision that bench.py
times.  The golden (tests/golden/tail_full_cfg2.npz) is the fp32 CPU oracle chain at this size, minted offline by
oracle/make_goldens_tail_full.py; the oracle itself is pinned to the reference's modules and, at the small geometry, to the
executed infer_v2_5.py:827-856 source lines (tests/test_zz_tail_wiring.py).

Bound (BASELINE.json north_star / SURVEY §8c): waveform RMS error <= 1e-3 on the +-1 scale with a shared z.  The relative
figure (error rms / signal rms) is printed and bounded too."""
import os

import numpy as np
import pytest

from indextts_b200 import synth
from indextts_b200.engine import fold_weight_norm
from oracle.make_goldens_tail_full import SEED, make_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tail_full_cfg2.npz")
WAV_RMS_BOUND = 1e-3          # north-star, absolute on the +-1 scale


def _load_full(engine):
    c, cc, h = dict(synth.S2MEL_CFG), dict(synth.CODEC_CFG), dict(synth.BIGVGAN_V2_22K)
    ws = fold_weight_norm(synth.make_s2mel_weights(c, seed=1234))
    engine.load_state_dict("s2mel.", {k: v for k, v in ws.items() if v.is_floating_point()})
    engine.load_state_dict("codec.", fold_weight_norm(synth.make_codec_weights(cc, seed=4321)))
    engine.load_state_dict("bigvgan.", synth.make_bigvgan_weights(h, seed=1234))
    engine.s2mel_init(c)
    engine.codec_init(cc)
    engine.bigvgan_init(h)


def test_golden_is_the_committed_one():
    g = np.load(GOLD)
    assert int(g["seed"]) == SEED and int(g["F"]) == 880 and g["wav"].shape == (880 * 256,) and g["mel"].shape == (80, 880)
    assert 0.05 < float(np.sqrt((g["wav"] ** 2).mean())) < 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("backend,f16,name", [(0, 1, "benchmarked precision (tensor cores, fp16 operands)"),
                                              (0, 0, "tensor cores, tf32 over fp32 storage (round 1)"),
                                              (1, 1, "strict fp32 SIMT back end")])
def test_full_cfg2_tail_within_north_star_rms(engine, backend, f16, name):
    g = np.load(GOLD)
    _load_full(engine)
    codes, pc, ref_mel, style, z, F = make_inputs()
    args = (codes[0].numpy().astype(np.int32), pc[0].numpy(), ref_mel[0].numpy(), style[0].numpy(), z[0].numpy(), F, 25, 0.7)
    engine.set_option("gemm_backend", backend)
    engine.set_option("tail_f16", f16)
    try:
        res = engine.codes_to_wav(*args, want_wav=True, want_pcm16=True, want_mel=True)
        res = engine.codes_to_wav(*args, want_wav=True, want_pcm16=True, want_mel=True)      # second call: warm timings
    finally:
        engine.set_option("gemm_backend", 0)
        engine.set_option("tail_f16", 1)
    wav = np.clip(np.asarray(res["wav"]).reshape(-1), -1.0, 1.0)
    ref = g["wav"]
    assert wav.shape == ref.shape and np.isfinite(wav).all()
    err = float(np.sqrt(((wav - ref) ** 2).mean()))
    rms = float(np.sqrt((ref ** 2).mean()))
    mel_err = float(np.abs(np.asarray(res["mel"]) - g["mel"]).max())
    print(f"[{name}] mel after 25 Euler steps: max abs error {mel_err:.3e} (mel std {g['mel'].std():.2f})")
    print(f"[{name}] full config-2 tail vs fp32 oracle: wav rms error {err:.3e} (signal rms {rms:.3f}, relative {err / rms:.3e}), "
          f"max abs {np.abs(wav - ref).max():.3e}; stage ms {engine.s2mel_last_ms()}, bigvgan {engine.bigvgan_last_ms():.2f}")
    assert err <= (WAV_RMS_BOUND if backend == 0 else 1e-4)
    assert err / rms <= (1e-2 if backend == 0 else 1e-3)
    assert np.abs(res["pcm16"].astype(np.float32).reshape(-1) - np.clip(np.round(wav * 32767.0), -32767, 32767)).max() <= 1.0
