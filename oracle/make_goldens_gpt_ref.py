"""Mint tests/golden/gpt_ref_wrapper.npz from the reference's own `UnifiedVoice.inference_speech`
(indextts/gpt/model_v2.py:716-825 → prepare_gpt_inputs :648-714 → GPT2InferenceModel.forward :121-198 → the vendored
generate() loop, transformers_generation_utils.py:1869-2385,3123-3297,3325-3609), run on CPU in fp32 with the seeded
weights of oracle.gpt.make_gpt_weights loaded into the reference module (oracle/refimport.py:gpt_module).

Build container only (needs /root/reference).  Stored: the prompt embeddings the reference assembled, its greedy
tokens and per-step logits, and its plain beam-search (num_beams=3, do_sample=False) tokens — everything the restated
oracle (oracle/gpt.py, oracle/beam.py) and the strict-fp32 CUDA path are then checked against.

    python -m oracle.make_goldens_gpt_ref
"""
import os

import numpy as np
import torch

from oracle import beam, refimport
from oracle.gpt import GptOracle, make_gpt_weights, prepare_gpt_inputs
from oracle.validate_gpt_vs_hf import small_case


def main():
    cfg, style, emo, text = small_case()
    cfg = dict(cfg, n_langs=106)      # lang_embedding has len(LANGUAGE_DICT) + 1 = 107 rows (model_v2.py:389-390)
    seed = 4242
    w = make_gpt_weights(cfg, seed=seed, bf16=False)
    g = refimport.gpt_module(cfg, w)
    n = 24
    rec = {"logits": [], "prompt": None}
    orig_store = g.inference_model.store_mel_emb

    def store(emb):
        rec["prompt"] = emb.detach().clone()
        return orig_store(emb)
    g.inference_model.store_mel_emb = store
    g.inference_model.register_forward_hook(
        lambda mod, inp, out: rec["logits"].append(out.logits[:, -1, :].detach().float().clone()))
    common = dict(langs=torch.tensor([1]), emo_speech_condition=torch.zeros(1, 8, 1024), cond_lengths=torch.tensor([8]),
                  emo_cond_lengths=torch.tensor([8]), emo_vec=emo[None], campplus_embedding=style[None],
                  top_p=0.8, top_k=30, temperature=0.8, num_return_sequences=1, length_penalty=0.0,
                  repetition_penalty=10.0, max_generate_length=n)
    with torch.no_grad():
        codes, _ = g.inference_speech(torch.zeros(1, 8, 1024), text[None], do_sample=False, num_beams=1, **common)
    codes = codes[0].numpy().astype(np.int32)
    logits = torch.cat(rec["logits"], 0).numpy()
    prompt_ref = rec["prompt"][0].numpy()
    # the reference appends nothing after max length; cut like infer_v2_5.py:809-821 would
    print("reference greedy:", codes.tolist())

    # ---- the restatement against the reference, here and now ----
    prompt = prepare_gpt_inputs(w, style, emo, text, lang=1, bf16=False).numpy()
    assert prompt.shape == prompt_ref.shape, (prompt.shape, prompt_ref.shape)
    print("prepare_gpt_inputs max |diff| vs reference:", float(np.abs(prompt - prompt_ref).max()))
    assert np.abs(prompt - prompt_ref).max() < 1e-5
    o_codes, o_logits = GptOracle(cfg, w, bf16=False).generate(prompt, n, 10.0, 0)
    k = len(o_codes)
    assert codes[:k].tolist() == o_codes.tolist(), (codes, o_codes)
    err = float(np.abs(o_logits - logits[:k]).max())
    print(f"oracle vs reference: {k} greedy tokens identical, max |logit diff| {err:.2e}")
    assert err < 2e-4

    # ---- plain beam search (num_beams=3, do_sample=False): the reference's _beam_search + its BeamSearchScorer ----
    rec["logits"].clear()
    with torch.no_grad():
        bcodes, _ = g.inference_speech(torch.zeros(1, 8, 1024), text[None], do_sample=False, num_beams=3, **common)
    bcodes = bcodes[0].numpy().astype(np.int32)
    print("reference beam search:", bcodes.tolist())
    p = dict(num_beams=3, start=cfg["start_mel_token"], stop=cfg["stop_mel_token"], repetition_penalty=10.0,
             temperature=0.8, top_k=30, top_p=0.8, length_penalty=0.0, seed=0, forbid_stop_before=0, do_sample=False)
    ob = beam.generate_beam(lambda: GptOracle(cfg, w, bf16=False), prompt, p, n)
    print("oracle    beam search:", ob["codes"].tolist())
    assert ob["codes"].tolist() == bcodes[: len(ob["codes"])].tolist()

    # ---- beam-sample (the .infer() default: num_beams=3, do_sample=True) with the RNG contract substituted ----
    # torch.multinomial is replaced, for this call only, by the documented Philox draw of oracle/beam.py applied to the
    # scores the reference itself computed (its log_softmax, its HF RepetitionPenalty/Temperature/TopK/TopP processors with
    # min_tokens_to_keep=2, its `+ beam_scores`): everything but the random stream is the reference's own code.
    import torch.nn.functional as F
    V, mB, bseed = cfg["number_mel_codes"], 3, 77
    state = {"scores": None, "step": 0}
    real_softmax, real_multinomial = F.softmax, torch.multinomial

    def softmax_spy(x, dim=-1, **kw):
        if x.dim() == 2 and x.shape[-1] == mB * V:
            state["scores"] = x.detach().clone()
        return real_softmax(x, dim=dim, **kw)

    def philox_multinomial(probs, num_samples, **kw):
        sc = state["scores"][0].numpy().astype(np.float32)
        cands = []
        for j in range(mB):
            seg = sc[j * V:(j + 1) * V]
            idx = [int(i) for i in np.lexsort((np.arange(V), -seg)) if np.isfinite(seg[i])]
            cands.append((idx, [np.float32(seg[i]) for i in idx]))
        s6, t6, p6 = beam.beam_draw(cands, mB, V, bseed, state["step"], 0, True)
        state["step"] += 1
        return torch.tensor([[pj * V + tj for tj, pj in zip(t6, p6)]], dtype=torch.long)

    import sys as _sys
    Scorer = _sys.modules["transformers.generation.beam_search"].BeamSearchScorer
    real_process = Scorer.process
    trace = []

    def process_spy(self, *a, **k):
        out = real_process(self, *a, **k)
        trace.append((out["next_beam_indices"].tolist(), out["next_beam_tokens"].tolist(),
                      [float(x) for x in out["next_beam_scores"]]))
        return out

    results = {}
    for tag, kw in (("a", dict(top_p=0.8, top_k=30, temperature=0.8)), ("b", dict(top_p=0.95, top_k=12, temperature=2.5))):
        state["step"] = 0
        trace.clear()
        F.softmax, torch.multinomial, Scorer.process = softmax_spy, philox_multinomial, process_spy
        try:
            with torch.no_grad():
                scodes, _ = g.inference_speech(torch.zeros(1, 8, 1024), text[None], do_sample=True, num_beams=3,
                                               **dict(common, **kw))
        finally:
            F.softmax, torch.multinomial, Scorer.process = real_softmax, real_multinomial, real_process
        scodes = scodes[0].numpy().astype(np.int32)
        ps = dict(p, do_sample=True, seed=bseed, **kw)
        osb = beam.generate_beam(lambda: GptOracle(cfg, w, bf16=False), prompt, ps, n)
        print(f"[{tag}] reference beam-sample (Philox draws):", scodes.tolist())
        print(f"[{tag}] oracle    beam-sample              :", osb["codes"].tolist())
        assert osb["codes"].tolist() == scodes[: len(osb["codes"])].tolist()
        assert len(trace) == len(osb["trace"])
        ndiff = 0
        for k_, ((rp, rt, rs), (op, ot, os_)) in enumerate(zip(trace, osb["trace"])):
            assert rp == op and rt == ot, (tag, k_, rp, op, rt, ot)
            assert np.allclose(rs, os_, rtol=0, atol=2e-4), (tag, k_, rs, os_)
            ndiff += int(rt != sorted(rt))
        explored = len({tuple(t[1]) for t in trace})
        print(f"[{tag}] every (parent, token, score) of {len(trace)} steps identical; {explored} distinct beam-token triples explored")
        results[tag] = (scodes, np.array([t[0] for t in trace], np.int32), np.array([t[1] for t in trace], np.int32))

    # ---- multinomial sampling (num_beams=1, do_sample=True): the reference's _sample + HF warpers, RNG substituted ----
    from oracle.gpt import sample_token
    sseed = 55
    sstate = {"scores": None, "step": 0}

    def softmax_spy1(x, dim=-1, **kw):
        if x.dim() == 2 and x.shape[-1] == V:
            sstate["scores"] = x.detach().clone()
        return real_softmax(x, dim=dim, **kw)

    def philox_multinomial1(probs, num_samples, **kw):
        sc = sstate["scores"][0].numpy().astype(np.float32)
        # the warpers already filtered: kept candidates are the finite scores; the draw contract is oracle.gpt.sample_token
        tok, _ = sample_token(sc, 0, 1.0, sseed, sstate["step"], 0)
        sstate["step"] += 1
        return torch.tensor([[tok]], dtype=torch.long)

    skw = dict(top_p=0.9, top_k=20, temperature=1.3)
    F.softmax, torch.multinomial = softmax_spy1, philox_multinomial1
    try:
        with torch.no_grad():
            scodes1, _ = g.inference_speech(torch.zeros(1, 8, 1024), text[None], do_sample=True, num_beams=1, **dict(common, **skw))
    finally:
        F.softmax, torch.multinomial = real_softmax, real_multinomial
    scodes1 = scodes1[0].numpy().astype(np.int32)
    o_s, _ = GptOracle(cfg, w, bf16=False).generate(prompt, n, 10.0, 0, do_sample=True, seed=sseed, seq=0, **skw)
    print("reference sampling (Philox draws):", scodes1.tolist())
    print("oracle    sampling               :", o_s.tolist())
    assert o_s.tolist() == scodes1[: len(o_s)].tolist()
    assert scodes1.tolist() != codes.tolist()

    # ---- row a7 at the wrapper level: UnifiedVoice.merge_emovec (model_v2.py:827-838, :588-593) ----
    from oracle.emo import make_emo_weights, merge_emovec
    ecfg = dict(idim=1024, odim=32, linear_units=48, heads=2, blocks=1, cnn_kernel=15, p_dim=1024, p_heads=2, p_dim_head=64,
                p_depth=2, p_ff_mult=2, model_dim=cfg["model_dim"])
    we = make_emo_weights(ecfg, seed=31)
    sd = g.state_dict()
    unknown = [k for k in we if k not in sd]
    assert not unknown, unknown
    g.load_state_dict(we, strict=False)
    ge = torch.Generator().manual_seed(9)
    spk_f, emo_f = torch.randn(1, 23, 1024, generator=ge), torch.randn(1, 31, 1024, generator=ge)
    with torch.no_grad():
        ev_ref = g.merge_emovec(spk_f, emo_f, torch.tensor([23]), torch.tensor([31]), alpha=0.6)[0]
    ev = merge_emovec(we, ecfg, spk_f[0], emo_f[0], 0.6)
    eerr = float((ev - ev_ref).abs().max())
    print(f"merge_emovec: oracle vs the reference wrapper max |diff| {eerr:.2e} (std {float(ev_ref.std()):.2f})")
    assert eerr < 5e-4

    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gpt_ref_wrapper.npz")
    np.savez_compressed(out, seed=seed, n_steps=n, style=style.numpy(), emo=emo.numpy(), text=text.numpy(), lang=1,
                        prompt=prompt_ref.astype(np.float32), greedy_codes=codes, greedy_logits=logits.astype(np.float32),
                        beam_codes=bcodes, beam_sample_seed=bseed,
                        beam_sample_a_codes=results['a'][0], beam_sample_a_parents=results['a'][1], beam_sample_a_tokens=results['a'][2],
                        beam_sample_b_codes=results['b'][0], beam_sample_b_parents=results['b'][1], beam_sample_b_tokens=results['b'][2],
                        emo_vec=ev_ref.numpy(), emo_seed=31, emo_feats_seed=9,
                        sample_codes=scodes1, sample_seed=sseed)   # features: torch.Generator(9) -> randn(1,23,1024), randn(1,31,1024)
    print("wrote", out)


if __name__ == "__main__":
    main()
