"""__graft_entry__.smoke(): one tiny invocation of each hot path on cuda:0, checked against the
CPU oracle (the oracle is the checker here, never the thing executed as the product)."""
import numpy as np
import torch


def run():
    from indextts_b200.engine import Engine
    from oracle.bigvgan import bigvgan_forward, make_bigvgan_weights, small_config, synthetic_mel, wav_rms_err
    from oracle.gpt import GptOracle, make_gpt_weights, prepare_gpt_inputs, r16
    from oracle.validate_gpt_vs_hf import small_case
    from tests.gpt_common import check_teacher_forced, load_gpt

    e = Engine(0)
    # GPT: 8 greedy steps of the small geometry
    cfg, style, emo, text = small_case()
    w = make_gpt_weights(cfg, seed=1234, bf16=True)
    load_gpt(e, cfg, w)
    prompt = prepare_gpt_inputs(w, style, r16(emo), text, lang=1, bf16=True).numpy()
    n = 8
    o_codes, o_logits = GptOracle(cfg, w, bf16=True).generate(prompt, n, 10.0, n)
    (codes,), (logits,) = e.gpt_generate([prompt], n, 10.0, forbid_stop_before=n, forced_codes=[o_codes],
                                         return_logits=True)
    check_teacher_forced(cfg, o_codes, o_logits, codes, logits, 10.0, n, max_abs=0.16, max_rms=0.035, tie_tol=0.13)
    # BigVGAN: reduced-channel generator, 16 frames
    h = small_config()
    wv = make_bigvgan_weights(h, seed=1)
    e.load_state_dict("bigvgan.", wv)
    e.bigvgan_init(h)
    mel = synthetic_mel(1, 16, seed=0)
    ref = bigvgan_forward(h, wv, mel).numpy()
    wav = e.bigvgan_forward(mel.numpy())
    err = wav_rms_err(wav, ref)
    assert err <= 1e-3, err
    print(f"smoke ok: gpt {n} steps token/logit parity, bigvgan rms err {err:.2e}, launches {e.launches}")
    e.close()
