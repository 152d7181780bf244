"""CPU: the GPT oracle against the golden vectors minted from stock transformers.GPT2Model
(oracle/validate_gpt_vs_hf.py), plus the index-bookkeeping traps of SURVEY.md §0/A.3."""
import os

import numpy as np
import torch

from oracle.gpt import GptOracle, make_gpt_weights, prepare_gpt_inputs, r16
from oracle.validate_gpt_vs_hf import small_case

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gpt_small.npz")


def test_fp32_oracle_matches_hf_golden():
    g = np.load(GOLD)
    cfg, style, emo, text = small_case()
    assert np.array_equal(text.numpy(), g["text"])
    w = make_gpt_weights(cfg, seed=int(g["weight_seed"]), bf16=False)
    prompt = prepare_gpt_inputs(w, style, emo, text, lang=int(g["lang"]), bf16=False)
    n = int(g["n_steps"])
    codes, logits = GptOracle(cfg, w, bf16=False).generate(prompt, n, 10.0, n)
    assert np.array_equal(codes, g["hf_fp32_codes"])
    assert np.abs(logits - g["hf_fp32_logits"]).max() < 2e-4


def test_bf16_oracle_reproduces_golden_bit_exact():
    g = np.load(GOLD)
    cfg, style, emo, text = small_case()
    w = make_gpt_weights(cfg, seed=int(g["weight_seed"]), bf16=True)
    prompt = prepare_gpt_inputs(w, style, r16(emo), text, lang=int(g["lang"]), bf16=True)
    assert np.array_equal(prompt.numpy(), g["prompt_bf16"])
    n = int(g["n_steps"])
    codes, logits = GptOracle(cfg, w, bf16=True).generate(prompt, n, 10.0, n)
    assert np.array_equal(codes, g["bf16_codes"])
    assert np.array_equal(logits, g["bf16_logits"])
    # and it sits inside the bf16 noise of the HF autocast path
    d = logits - g["hf_autocast_logits"]
    assert np.sqrt((d ** 2).mean()) < 0.03


def test_position_rule_and_repetition_penalty_traps():
    """P1: generated token k is embedded at mel position k+1 (position 1 is never used);
    P2: codes 1 and start_mel are penalised from step 0."""
    cfg, style, emo, text = small_case()
    w = make_gpt_weights(cfg, seed=3, bf16=False)
    prompt = prepare_gpt_inputs(w, style, emo, text, lang=0, bf16=False)
    orc = GptOracle(cfg, w, bf16=False)
    codes, logits = orc.generate(prompt, 4, 10.0, 4)
    # recompute step 2 by hand with position index 3 for the second generated token
    orc.reset()
    first = w["mel_embedding.weight"][cfg["start_mel_token"]] + w["mel_pos_embedding.emb.weight"][0]
    h = orc.forward_rows(torch.cat([prompt, first[None]], 0))[-1:]
    h = orc.forward_rows((w["mel_embedding.weight"][int(codes[0])] + w["mel_pos_embedding.emb.weight"][2])[None])
    h = orc.forward_rows((w["mel_embedding.weight"][int(codes[1])] + w["mel_pos_embedding.emb.weight"][3])[None])
    assert np.allclose(orc.logits(h)[0].numpy(), logits[2], atol=1e-5)
    # P2: make token 1 the raw argmax at step 0 and check the penalty removes it
    w2 = dict(w)
    w2["mel_head.bias"] = w["mel_head.bias"].clone()
    w2["mel_head.bias"][1] += 50.0
    c2, l2 = GptOracle(cfg, w2, bf16=False).generate(prompt, 1, 10.0, 1)
    assert int(np.argmax(l2[0])) == 1 and int(c2[0]) != 1


def test_stop_token_ends_generation():
    cfg, style, emo, text = small_case()
    w = make_gpt_weights(cfg, seed=5, bf16=False)
    w["mel_head.bias"] = w["mel_head.bias"].clone()
    w["mel_head.bias"][cfg["stop_mel_token"]] += 100.0
    prompt = prepare_gpt_inputs(w, style, emo, text, lang=0, bf16=False)
    codes, _ = GptOracle(cfg, w, bf16=False).generate(prompt, 10, 10.0, 3)
    assert len(codes) == 4 and codes[-1] == cfg["stop_mel_token"]


def test_philox_known_answer_and_sampler_properties():
    from oracle.gpt import philox4x32_10, sample_token
    # Random123 known-answer test: counter = 0, key = 0
    assert philox4x32_10(0, 0, 0) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    rng = np.random.default_rng(0)
    s = rng.standard_normal(500).astype(np.float32) * 3
    tok, kept = sample_token(s, top_k=1, top_p=1.0, seed=1, step=0, seq=0)
    assert tok == int(np.argmax(s)) and kept == [tok]
    tok, kept = sample_token(s, top_k=30, top_p=0.8, seed=7, step=3, seq=0)
    top30 = set(np.argsort(-s)[:30].tolist())
    assert tok in top30 and set(kept) <= top30 and 1 <= len(kept) <= 30
    # frequencies follow the renormalised distribution
    s2 = np.log(np.array([0.5, 0.3, 0.15, 0.05], dtype=np.float32))
    cnt = np.zeros(4)
    for i in range(4000):
        cnt[sample_token(s2, 0, 1.0, seed=123, step=i, seq=0)[0]] += 1
    assert np.abs(cnt / 4000 - np.array([0.5, 0.3, 0.15, 0.05])).max() < 0.03


REF_GOLD = os.path.join(os.path.dirname(__file__), "golden", "gpt_ref_wrapper.npz")


def _ref_case():
    g = np.load(REF_GOLD)
    cfg, _, _, _ = small_case()
    cfg = dict(cfg, n_langs=106)
    w = make_gpt_weights(cfg, seed=int(g["seed"]), bf16=False)
    return g, cfg, w


def test_oracle_matches_reference_unifiedvoice_wrapper():
    """Golden minted from the reference's own `UnifiedVoice.inference_speech` (model_v2.py:716-825: its
    prepare_gpt_inputs, its vendored GPT2 blocks, its vendored generate loop) run on CPU in fp32 with these seeded
    weights (oracle/make_goldens_gpt_ref.py): the restated wrapper logic — prompt assembly, the P1 position rule, the
    fake prompt ids in the repetition penalty (P2), the double LayerNorm head (P3) — reproduces it."""
    g, cfg, w = _ref_case()
    style, emo, text = torch.from_numpy(g["style"]), torch.from_numpy(g["emo"]), torch.from_numpy(g["text"])
    prompt = prepare_gpt_inputs(w, style, emo, text, lang=int(g["lang"]), bf16=False).numpy()
    assert prompt.shape == g["prompt"].shape and np.abs(prompt - g["prompt"]).max() < 1e-6
    n = int(g["n_steps"])
    codes, logits = GptOracle(cfg, w, bf16=False).generate(prompt, n, 10.0, 0)
    assert codes.tolist() == g["greedy_codes"][: len(codes)].tolist()
    assert np.abs(logits - g["greedy_logits"][: len(codes)]).max() < 1e-4


def test_oracle_beam_search_matches_reference_beam_search():
    """The reference's `_beam_search` + its own BeamSearchScorer (num_beams=3, do_sample=False) produced
    `beam_codes`; the restated beam logic driven by the restated model returns the same hypothesis."""
    from oracle import beam
    g, cfg, w = _ref_case()
    n = int(g["n_steps"])
    p = dict(num_beams=3, start=cfg["start_mel_token"], stop=cfg["stop_mel_token"], repetition_penalty=10.0,
             temperature=0.8, top_k=30, top_p=0.8, length_penalty=0.0, seed=0, forbid_stop_before=0, do_sample=False)
    out = beam.generate_beam(lambda: GptOracle(cfg, w, bf16=False), g["prompt"], p, n)
    assert out["codes"].tolist() == g["beam_codes"][: len(out["codes"])].tolist()


def test_oracle_beam_sample_matches_reference_beam_sample_with_substituted_rng():
    """Beam-sample (num_beams=3, do_sample=True, the `.infer()` default).  The goldens come from the reference's own
    `_beam_search` + HF logits processors + BeamSearchScorer, with `torch.multinomial` replaced by the documented Philox
    draw applied to the scores the reference computed (oracle/make_goldens_gpt_ref.py): every step's (parent, token)
    choices and the returned hypothesis — two sampling configurations, one sharp, one flat."""
    from oracle import beam
    g, cfg, w = _ref_case()
    n = int(g["n_steps"])
    for tag, kw in (("a", dict(top_p=0.8, top_k=30, temperature=0.8)), ("b", dict(top_p=0.95, top_k=12, temperature=2.5))):
        p = dict(num_beams=3, start=cfg["start_mel_token"], stop=cfg["stop_mel_token"], repetition_penalty=10.0,
                 length_penalty=0.0, seed=int(g["beam_sample_seed"]), forbid_stop_before=0, do_sample=True, **kw)
        out = beam.generate_beam(lambda: GptOracle(cfg, w, bf16=False), g["prompt"], p, n)
        assert out["codes"].tolist() == g[f"beam_sample_{tag}_codes"][: len(out["codes"])].tolist()
        assert [t[0] for t in out["trace"]] == g[f"beam_sample_{tag}_parents"].tolist()
        assert [t[1] for t in out["trace"]] == g[f"beam_sample_{tag}_tokens"].tolist()
    assert g["beam_sample_b_codes"].tolist() != g["beam_codes"].tolist()      # the flat configuration really samples


def test_oracle_sampling_matches_reference_sample_loop_with_substituted_rng():
    """num_beams=1, do_sample=True: the reference's `_sample` loop with HF's RepetitionPenalty / Temperature / TopK / TopP
    processors, `torch.multinomial` replaced by the documented Philox draw (oracle/make_goldens_gpt_ref.py)."""
    g, cfg, w = _ref_case()
    n = int(g["n_steps"])
    codes, _ = GptOracle(cfg, w, bf16=False).generate(g["prompt"], n, 10.0, 0, do_sample=True, seed=int(g["sample_seed"]), seq=0,
                                                      top_p=0.9, top_k=20, temperature=1.3)
    assert codes.tolist() == g["sample_codes"][: len(codes)].tolist()
    assert g["sample_codes"].tolist() != g["greedy_codes"].tolist()
