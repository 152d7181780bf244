"""Mint tests/golden/beam_scorer.json from the reference's own BeamSearchScorer (run in the build container only:
needs /root/reference).  Random candidate streams are pushed through
`indextts/gpt/transformers_beam_search.py:BeamSearchScorer.process/finalize`; inputs and the reference's outputs are
stored so that oracle/beam.py's restatement can be checked anywhere (tests/test_beam_oracle.py).

    python -m oracle.make_goldens_beam
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch


def load_reference_scorer():
    # the vendored file imports transformers.generation.beam_constraints (removed from transformers 5.x); it is only
    # used by ConstrainedBeamSearchScorer
    stub = types.ModuleType("transformers.generation.beam_constraints")
    stub.Constraint = type("Constraint", (), {})
    stub.ConstraintListState = type("ConstraintListState", (), {})
    sys.modules.setdefault("transformers.generation.beam_constraints", stub)
    spec = importlib.util.spec_from_file_location("ref_beam_search", "/root/reference/indextts/gpt/transformers_beam_search.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.BeamSearchScorer


def main():
    Scorer = load_reference_scorer()
    rng = np.random.default_rng(7)
    m, eos, pad, prompt_len = 3, 9, 9, 4
    trials = []
    for trial in range(40):
        lp = float(rng.choice([0.0, 0.0, 1.0, 0.5]))
        max_new = int(rng.integers(3, 14))
        sc = Scorer(batch_size=1, num_beams=m, device=torch.device("cpu"), length_penalty=lp, do_early_stopping=False,
                    max_length=prompt_len + max_new)
        ids = torch.full((m, prompt_len), 1, dtype=torch.long)
        beam_scores = np.array([0.0, -1e9, -1e9], dtype=np.float32)
        steps = []
        p_eos = float(rng.choice([0.05, 0.2, 0.5]))
        for k in range(max_new):
            par = rng.integers(0, m, size=2 * m)
            tok = rng.integers(0, 9, size=2 * m)
            tok[rng.random(2 * m) < p_eos] = eos
            # keep at least m non-eos candidates (HF raises otherwise)
            non = np.flatnonzero(tok != eos)
            if len(non) < m:
                tok[: m] = rng.integers(0, 9, size=m)
            s = (beam_scores[par] + (-rng.random(2 * m) * 3).astype(np.float32)).astype(np.float32)
            s[s < -1e8] = (-1e9 - rng.random((s < -1e8).sum())).astype(np.float32)
            order = np.argsort(-s, kind="stable")
            s, tok, par = s[order], tok[order], par[order]
            out = sc.process(ids, torch.tensor(s)[None], torch.tensor(tok)[None], torch.tensor(par)[None],
                             pad_token_id=pad, eos_token_id=eos, decoder_prompt_len=prompt_len)
            nb_s = out["next_beam_scores"].numpy().copy()
            nb_t = out["next_beam_tokens"].numpy().copy()
            nb_i = out["next_beam_indices"].numpy().copy()
            steps.append(dict(scores=s.copy(), tokens=tok.copy(), parents=par.copy(), out_scores=nb_s, out_tokens=nb_t,
                              out_parents=nb_i, done=bool(sc.is_done)))
            ids = torch.cat([ids[torch.tensor(nb_i)], torch.tensor(nb_t)[:, None]], dim=-1)
            beam_scores = nb_s.astype(np.float32)
            if sc.is_done:
                break
        fin = sc.finalize(ids, torch.tensor(beam_scores), None, None, max_length=prompt_len + max_new, pad_token_id=pad,
                          eos_token_id=eos, decoder_prompt_len=prompt_len)
        seq = fin["sequences"][0, prompt_len:].numpy()
        trials.append(dict(lp=lp, max_new=max_new, steps=steps, final=seq.copy(), final_score=float(fin["sequence_scores"][0])))
    def js(v):
        return v.tolist() if isinstance(v, np.ndarray) else v
    doc = dict(m=m, eos=eos, trials=[dict(lp=t["lp"], max_new=t["max_new"], final=js(t["final"]), final_score=t["final_score"],
                                          steps=[{k: js(np.asarray(v)) for k, v in st.items()} for st in t["steps"]])
                                     for t in trials])
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "beam_scorer.json")
    with open(out, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("wrote", out, "trials", len(trials), "done-early", sum(t["steps"][-1]["done"] for t in trials))


if __name__ == "__main__":
    main()
