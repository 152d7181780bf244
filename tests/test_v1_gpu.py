"""GPU: the v1 / v1.5 vocoder side (ECAPA-TDNN speaker encoder + latent-conditioned BigVGAN)
through the C-ABI against the golden minted from the reference's own `indextts.BigVGAN.models.BigVGAN`
(tests/golden/v1_vocoder_small.npz) and against the oracle on a second input."""
import os

import numpy as np
import pytest
import torch

from indextts_b200 import synth
from oracle import v1

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "v1_vocoder_small.npz")


def test_v1_vocoder_vs_reference_golden(engine):
    g = np.load(GOLD)
    h = synth.small_v1_config()
    w = synth.make_bigvgan_v1_weights(h, seed=int(g["seed"]))
    engine.load_state_dict("bigvgan_v1.", {k: v for k, v in w.items() if v.is_floating_point()})
    engine.v1_vocoder_init(h)
    scale = float(np.abs(g["spk"]).max())
    for backend, tol_e, tol_w in ((1, 1e-4, 2e-4), (0, 5e-3, 5e-3)):
        engine.set_option("gemm_backend", backend)
        try:
            emb = engine.v1_speaker_embedding(g["mel_ref"][0])
            wav = engine.v1_vocode(g["latent"][0], g["mel_ref"][0])
        finally:
            engine.set_option("gemm_backend", 0)
        e1 = float(np.abs(emb - g["spk"][0, 0]).max()) / scale
        e2 = float(np.abs(wav - g["wav"][0, 0]).max())
        print(f"[backend {backend}] ECAPA embedding rel err {e1:.2e}, wav max err {e2:.2e}")
        assert e1 < tol_e and e2 < tol_w
    assert np.abs(wav).max() <= 1.0
    # a longer, different utterance against the oracle (strict back end)
    gen = torch.Generator().manual_seed(21)
    latent = torch.randn(1, 41, h["gpt_dim"], generator=gen)
    mel_ref = torch.randn(1, 203, h["num_mels"], generator=gen) * 1.5 - 4.0
    ref = v1.bigvgan_v1_forward(h, w, latent, mel_ref)[0, 0].numpy()
    engine.set_option("gemm_backend", 1)
    try:
        got = engine.v1_vocode(latent[0].numpy(), mel_ref[0].numpy())
    finally:
        engine.set_option("gemm_backend", 0)
    assert got.shape == ref.shape and np.abs(got - ref).max() < 2e-4
    with pytest.raises(RuntimeError):
        engine.v1_speaker_embedding(mel_ref[0, :3].numpy())


def test_v1_gpt_side_vs_reference_golden(engine):
    """v1 GPT side on the strict fp32 path against the golden minted from the reference's own v1 `UnifiedVoice`
    (tests/golden/v1_gpt_small.npz): 32-latent prompt encoder, prompt rows, greedy decode with the no-KV-cache position
    rule (infer.py:101), vocoder latents."""
    from oracle.validate_gpt_vs_hf import small_case
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "v1_gpt_small.npz"))
    cfg, _, _, _ = small_case()
    ccfg = synth.small_v1_cond_cfg(cfg["model_dim"])
    w = synth.make_gpt_v1_weights(cfg, ccfg, seed=int(g["seed"]))
    engine.load_state_dict("gpt.", w)
    engine.gpt_init(cfg["layers"], cfg["model_dim"], cfg["heads"], cfg["number_mel_codes"], cfg["start_mel_token"],
                    cfg["stop_mel_token"], cfg["max_mel_positions"], max_prompt=128, max_batch=1, weights_bf16=False)
    engine.v1_cond_init(ccfg, 32)
    mel = g["mel"][0].T.copy()                       # [T, 100]
    engine.set_option("gemm_backend", 1)
    try:
        conds = engine.v1_get_conditioning(mel)
    finally:
        engine.set_option("gemm_backend", 0)
    e0 = float(np.abs(conds - g["conds"]).max())
    print(f"v1 get_conditioning (32 latents) max err {e0:.2e}")
    assert conds.shape == (32, cfg["model_dim"]) and e0 < 2e-3
    prompt = engine.gpt_prepare_inputs_v1(g["conds"], g["text"])
    assert prompt.shape == (32 + len(g["text"]) + 2, cfg["model_dim"])
    n = len(g["codes"])
    (codes,), (logits,) = engine.gpt_generate([prompt], n, 10.0, return_logits=True, mel_pos_mode=1)
    assert codes.tolist() == g["codes"].tolist()
    assert np.abs(logits - g["logits"]).max() < 5e-4
    toks = g["codes"][g["codes"] != cfg["stop_mel_token"]]
    lat = engine.gpt_latents_v1(g["conds"], g["text"], toks)
    e2 = float(np.abs(lat - g["latents"]).max())
    print(f"v1 latents max err {e2:.2e}")
    assert lat.shape == g["latents"].shape and e2 < 1e-3


def test_v1_latents_fused_bf16_path_close_to_reference(engine):
    """The latent pass on the fused bf16 kernel (prefill sweep with the residual-stream dump): within bf16 noise of the fp32
    reference latents."""
    from oracle.validate_gpt_vs_hf import small_case
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "v1_gpt_small.npz"))
    cfg, _, _, _ = small_case()
    ccfg = synth.small_v1_cond_cfg(cfg["model_dim"])
    w = synth.make_gpt_v1_weights(cfg, ccfg, seed=int(g["seed"]))
    engine.load_state_dict("gpt.", w)
    engine.gpt_init(cfg["layers"], cfg["model_dim"], cfg["heads"], cfg["number_mel_codes"], cfg["start_mel_token"],
                    cfg["stop_mel_token"], cfg["max_mel_positions"], max_prompt=128, max_batch=1, weights_bf16=True)
    toks = g["codes"][g["codes"] != cfg["stop_mel_token"]]
    lat = engine.gpt_latents_v1(g["conds"], g["text"], toks)
    ref = g["latents"]
    rel = float(np.sqrt(((lat - ref) ** 2).mean()) / ref.std())
    print(f"v1 latents on the fused bf16 path: relative rms error {rel:.3f}")
    assert lat.shape == ref.shape and rel < 0.05
