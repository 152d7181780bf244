"""Shared helpers of the GPT parity tests (engine = CUDA path through the C-ABI, oracle = CPU)."""
import numpy as np

from oracle.gpt import GptOracle, gpt_config, make_gpt_weights, prepare_gpt_inputs, r16  # noqa: F401


def load_gpt(engine, cfg, w, max_batch=1, max_prompt=128, bf16=True):
    engine.load_state_dict("gpt.", w)
    engine.gpt_init(cfg["layers"], cfg["model_dim"], cfg["heads"], cfg["number_mel_codes"],
                    cfg["start_mel_token"], cfg["stop_mel_token"], cfg["max_mel_positions"],
                    max_prompt=max_prompt, max_batch=max_batch, weights_bf16=bf16)


def processed(logits_row, seen, rep_penalty, stop, forbid):
    s = logits_row.astype(np.float32).copy()
    idx = np.array(sorted(seen))
    sv = s[idx]
    s[idx] = np.where(sv < 0, sv * rep_penalty, sv / rep_penalty)
    if forbid:
        s[stop] = -np.inf
    return s


def check_teacher_forced(cfg, o_codes, o_logits, e_codes, e_logits, rep_penalty, forbid_stop_before,
                         max_abs, max_rms, tie_tol):
    """Engine was run teacher-forced on the oracle's codes.  Every step's raw logits must agree
    within (max_abs, max_rms); the engine's own pick must equal the oracle's unless the oracle's
    processed scores of the two candidates are within tie_tol (a near-tie at the precision of
    the path).  Returns the number of near-tie disagreements."""
    n = len(o_codes)
    assert len(e_codes) == n, (len(e_codes), n)
    seen = {1, cfg["start_mel_token"]}
    ties = 0
    for k in range(n):
        d = e_logits[k] - o_logits[k]
        assert np.isfinite(e_logits[k]).all()
        assert np.abs(d).max() <= max_abs, (k, float(np.abs(d).max()))
        assert np.sqrt((d.astype(np.float64) ** 2).mean()) <= max_rms, (k, float(np.sqrt((d ** 2).mean())))
        if e_codes[k] != o_codes[k]:
            s = processed(o_logits[k], seen, rep_penalty, cfg["stop_mel_token"], k < forbid_stop_before)
            gap = abs(float(s[o_codes[k]]) - float(s[e_codes[k]]))
            assert gap <= tie_tol, f"step {k}: engine {e_codes[k]} vs oracle {o_codes[k]}, gap {gap}"
            ties += 1
        seen.add(int(o_codes[k]))
    return ties
