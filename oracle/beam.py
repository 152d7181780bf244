"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's default decoding mode,
beam-search multinomial sampling ("beam-sample", `num_beams=3, do_sample=True`).

Follows
  * `GenerationMixin._beam_search`      indextts/gpt/transformers_generation_utils.py:3325-3609
    (log_softmax → processors → + beam_scores → view [B, m*V] → softmax → multinomial(2m) → sort → scorer.process →
     reorder ids and KV cache; `finalize` at the end),
  * processor order / min_tokens_to_keep = 2 with beams    :900-905, 1019-1047,
  * `BeamSearchScorer.process / finalize`                  indextts/gpt/transformers_beam_search.py:215-420,
  * `BeamHypotheses.add / is_done`                          :930-1010.

The scorer logic (process / finalize / hypotheses heap) is pinned against the reference's own `BeamSearchScorer`
(tests/golden/beam_scorer.json, minted by oracle/make_goldens_beam.py).  The random draw is NOT torch.multinomial's
stream: it is the documented Philox contract of the device: 2m successive draws without replacement
by inverse CDF over the union of the beams' kept candidates (beam-major, descending score inside a beam), one
Philox4x32-10 block per draw with counter (step, 0x10000 + 16*utterance + draw).  Successive draws without
replacement have the same distribution as torch.multinomial(probs, 2m) (Plackett-Luce)."""
import math

import numpy as np

from oracle.gpt import CMAX, philox4x32_10

F = np.float32


class BeamHyps:
    """BeamHypotheses (transformers_beam_search.py:930-1010); python floats (double) like `.item()` values."""

    def __init__(self, num_beams, length_penalty, early_stopping=False):
        self.num_beams, self.length_penalty, self.early_stopping = num_beams, float(length_penalty), early_stopping
        self.beams = []            # (score, tokens)
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs, generated_len):
        score = float(sum_logprobs) / (generated_len ** self.length_penalty)
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, list(hyp)))
            if len(self) > self.num_beams:
                srt = sorted([(s, idx) for idx, (s, _) in enumerate(self.beams)])
                del self.beams[srt[0][1]]
                self.worst_score = srt[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_gen_len):
        if len(self) < self.num_beams:
            return False
        if self.early_stopping is True:
            return True
        highest = float(best_sum_logprobs) / (cur_gen_len ** self.length_penalty)
        return self.worst_score >= highest


def scorer_process(hyps, seqs, next_scores, next_tokens, next_indices, eos, gen_len):
    """BeamSearchScorer.process for one batch entry (transformers_beam_search.py:215-320).
    seqs: the m current token lists (generated part only); candidates sorted by descending score.
    Returns (beam_scores[m], beam_tokens[m], beam_parents[m], done)."""
    m = hyps.num_beams
    out_s, out_t, out_p = [], [], []
    for rank, (tok, sc, par) in enumerate(zip(next_tokens, next_scores, next_indices)):
        if tok == eos:
            if rank >= m:
                continue
            hyps.add(seqs[par], float(sc), gen_len)
        else:
            out_s.append(F(sc)); out_t.append(int(tok)); out_p.append(int(par))
        if len(out_s) == m:
            break
    if len(out_s) < m:
        raise ValueError("fewer than num_beams non-eos candidates")
    done = hyps.is_done(float(max(next_scores)), gen_len)
    return out_s, out_t, out_p, done


def scorer_finalize(hyps, done, seqs, beam_scores, eos, max_new):
    """BeamSearchScorer.finalize, num_return_sequences = 1 (transformers_beam_search.py:322-420): best hypothesis,
    eos appended when it is shorter than the longest allowed length."""
    if not done:
        for s, sc in zip(seqs, beam_scores):
            hyps.add(s, float(sc), len(s))
    best = sorted(hyps.beams, key=lambda x: x[0])[-1]
    codes = list(best[1])
    if len(codes) < max_new:
        codes.append(eos)
    return np.array(codes, dtype=np.int32), best[0]


def beam_candidates(logits_row, seen, beam_score, k, p):
    """One beam: log_softmax → RepetitionPenalty → (forbid stop) → Temperature → TopK(min keep 2) → TopP(min keep 2).
    Returns (tokens, scores + beam_score) of the kept candidates in descending score order."""
    lg = np.asarray(logits_row, dtype=F)
    mx = lg.max()
    lse = F(mx + F(np.log(np.exp((lg - mx).astype(F)).astype(F).sum(dtype=F))))
    s = (lg - lse).astype(F)
    if seen:
        idx = np.array(sorted(seen), dtype=np.int64)
        sv = s[idx]
        s[idx] = np.where(sv < 0, sv * F(p["repetition_penalty"]), sv / F(p["repetition_penalty"])).astype(F)
    if k < p.get("forbid_stop_before", 0):
        s[p["stop"]] = -np.inf
    s = (s * F(1.0 / p["temperature"])).astype(F)
    order = np.lexsort((np.arange(len(s)), -s))
    kk = min(max(p["top_k"], 2), CMAX) if p["top_k"] > 0 else CMAX
    cand, kth = [], None
    for i in order:
        if not np.isfinite(s[i]) or len(cand) >= CMAX:
            break
        if len(cand) < kk:
            cand.append(int(i)); kth = s[i]
        elif s[i] == kth:
            cand.append(int(i))
        else:
            break
    cv = np.exp((s[cand] - s[cand[0]]).astype(F)).astype(F)
    tot = F(0)
    for v in cv:
        tot = F(tot + v)
    keep = len(cand)
    if p["top_p"] < 1.0:
        tail = F(0)
        for i in range(len(cand) - 1, 1, -1):          # min_tokens_to_keep = 2 with beams (:1023-1027)
            tail = F(tail + F(cv[i] / tot))
            if tail <= F(1.0 - p["top_p"]):
                keep = i
            else:
                break
    toks = cand[:keep]
    return toks, [F(s[t] + F(beam_score)) for t in toks]


def beam_draw(cands, m, V, seed, step, utt, do_sample=True):
    """Union of the beams' candidates → softmax → 2m draws without replacement → sort by score (descending, stable).
    Returns (scores, tokens, parents) of the 2m candidates (transformers_generation_utils.py:3505-3530)."""
    u_sc, u_tok, u_par = [], [], []
    for j, (toks, scs) in enumerate(cands):
        u_tok += toks; u_sc += scs; u_par += [j] * len(toks)
    u_sc = np.array(u_sc, dtype=F)
    w = np.exp((u_sc - u_sc.max()).astype(F)).astype(F)
    used = np.zeros(len(w), dtype=bool)
    picks = []
    for d in range(2 * m):
        tot = F(0)
        for i in range(len(w)):
            if not used[i]:
                tot = F(tot + w[i])
        pick = -1
        if tot > 0 and do_sample:              # do_sample=False: plain beam search, torch.topk of the union (:3527-3530)
            r0 = philox4x32_10(seed, step, 0x10000 + 16 * utt + d)[0]
            u = F(F(r0 >> 8) * F(1.0 / 16777216.0) * tot)
            acc = F(0)
            last = -1
            for i in range(len(w)):
                if used[i] or w[i] == 0:
                    continue
                acc = F(acc + w[i]); last = i
                if u < acc:
                    pick = i
                    break
            if pick < 0:
                pick = last
        if pick < 0:   # fewer positive-probability candidates than draws (torch.multinomial is undefined here):
            rest = [i for i in range(len(w)) if not used[i]]          # take the best remaining score
            pick = max(rest, key=lambda i: (u_sc[i], -i))
        used[pick] = True
        picks.append(pick)
    picks.sort(key=lambda i: -float(u_sc[i]))      # list.sort is stable
    return [u_sc[i] for i in picks], [u_tok[i] for i in picks], [u_par[i] for i in picks]


def run_beam(logits_fn, p, max_new, utt=0):
    """Drive beam-sample for one utterance.  `logits_fn(step, parents, tokens)` returns the [m, V] logits of the
    current beams (parents/tokens = the reorder that produced them; None at step 0).
    Returns dict(codes, score, trace=[(parents, tokens, scores)], steps)."""
    m, eos = p["num_beams"], p["stop"]
    hyps = BeamHyps(m, p.get("length_penalty", 0.0), False)
    beam_scores = [F(0)] + [F(-1e9)] * (m - 1)
    seqs = [[] for _ in range(m)]
    seen = [{1, p["start"]} for _ in range(m)]
    parents = tokens = None
    done = False
    trace = []
    steps = 0
    do_sample = p.get("do_sample", True)
    pc = p if do_sample else dict(p, temperature=1.0, top_k=0, top_p=1.0)      # warpers exist only when sampling
    for k in range(max_new):
        lg = logits_fn(k, parents, tokens)
        cands = [beam_candidates(lg[j], seen[j], beam_scores[j], k, pc) for j in range(m)]
        sc, tk, pr = beam_draw(cands, m, lg.shape[-1], p["seed"], k, utt, do_sample)
        bs, bt, bp, done = scorer_process(hyps, seqs, sc, tk, pr, eos, k + 1)
        seqs = [seqs[q] + [t] for q, t in zip(bp, bt)]
        seen = [set(seen[q]) | {t} for q, t in zip(bp, bt)]
        beam_scores, parents, tokens = bs, bp, bt
        trace.append((list(bp), list(bt), [float(x) for x in bs]))
        steps = k + 1
        if done:
            break
    codes, score = scorer_finalize(hyps, done, seqs, beam_scores, eos, max_new)
    return dict(codes=codes, score=score, trace=trace, steps=steps)


def generate_beam(make_oracle, prompt_emb, p, max_new):
    """Model-driven beam-sample with m independent KV caches (copied on reorder like `_temporary_reorder_cache`)."""
    import copy

    import torch
    m = p["num_beams"]
    base = make_oracle()
    w, rr, cfg = base.w, base.rr, base.cfg
    base.reset()
    prompt_emb = torch.as_tensor(prompt_emb, dtype=torch.float32)
    first = rr(w["mel_embedding.weight"][cfg["start_mel_token"]] + w["mel_pos_embedding.emb.weight"][0])
    hidden0 = base.forward_rows(torch.cat([prompt_emb, first[None]], dim=0))[-1:]
    state = {"beams": [base] + [copy.deepcopy(base) for _ in range(m - 1)], "hidden": [hidden0] * m}
    dumped = []

    def logits_fn(k, parents, tokens):
        if parents is not None:
            nb, nh = [], []
            for q, t in zip(parents, tokens):
                o = copy.deepcopy(state["beams"][q])
                emb = rr(w["mel_embedding.weight"][t] + w["mel_pos_embedding.emb.weight"][k + 1])   # trap P1
                nh.append(o.forward_rows(emb[None]))
                nb.append(o)
            state["beams"], state["hidden"] = nb, nh
        lg = np.stack([state["beams"][j].logits(state["hidden"][j])[0].numpy() for j in range(m)])
        dumped.append(lg)
        return lg

    out = run_beam(logits_fn, p, max_new)
    out["logits"] = np.stack(dumped)
    return out
