"""Pin oracle/gpt.py against stock `transformers.GPT2Model` — the block arithmetic the reference
executes (indextts/gpt/model_v2.py:259-279 builds exactly this model, with `wpe` replaced by
zeros and `wte` deleted).  Run in the build container (CPU):

    python -m oracle.validate_gpt_vs_hf            # checks + writes tests/golden/gpt_small.npz

Checks
  1. fp32: restated logits == HF logits (prefill + N cached decode steps), max-abs <= 2e-4.
  2. bf16: restated (explicit rounding points) vs HF under torch.autocast('cpu', bfloat16) with
     .bfloat16() weights — the reference's use_bf16 recipe (infer_v2_5.py:143-146,758); agreement
     to bf16 resolution validates the rounding-point model.
  3. greedy tokens with repetition penalty equal between the two fp32 implementations.
The golden file stores the HF-path fp32 logits/tokens and the bf16-policy tokens/logits of the
restatement for the small config used by tests/test_gpt_oracle.py and the GPU parity tests.
"""
import functools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gpt import GptOracle, gpt_config, make_gpt_weights, prepare_gpt_inputs, r16  # noqa: E402


def null_position_embeddings(range_, dim):  # model_v2.py:23-24
    return torch.zeros((range_.shape[0], range_.shape[1], dim), device=range_.device)


def build_hf(cfg, w, dtype=torch.float32):
    from transformers import GPT2Config, GPT2Model
    gc = GPT2Config(vocab_size=256, n_positions=cfg["max_mel_positions"] + cfg["max_text_tokens"] + 2,
                    n_embd=cfg["model_dim"], n_layer=cfg["layers"], n_head=cfg["heads"],
                    use_cache=True)
    gpt = GPT2Model(gc)
    del gpt.wpe
    gpt.wpe = functools.partial(null_position_embeddings, dim=cfg["model_dim"])
    del gpt.wte
    sd = {k[len("gpt."):]: v for k, v in w.items() if k.startswith("gpt.")}
    missing, unexpected = gpt.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("attn.bias" in m or "masked_bias" in m for m in missing), missing
    gpt.eval()
    final_norm = torch.nn.LayerNorm(cfg["model_dim"])
    final_norm.weight.data.copy_(w["final_norm.weight"])
    final_norm.bias.data.copy_(w["final_norm.bias"])
    head = torch.nn.Linear(cfg["model_dim"], cfg["number_mel_codes"])
    head.weight.data.copy_(w["mel_head.weight"])
    head.bias.data.copy_(w["mel_head.bias"])
    lm_head = torch.nn.Sequential(final_norm, head)  # model_v2.py:54
    if dtype != torch.float32:
        # `.bfloat16()` weights; LayerNorm kept fp32 because CUDA autocast runs layer_norm in
        # fp32 (inputs/params upcast) while CPU autocast has no such rule.  Values are identical
        # (bf16-representable), only the container dtype differs.
        gpt = gpt.to(dtype)
        lm_head = lm_head.to(dtype)
        for m in list(gpt.modules()) + list(lm_head.modules()):
            if isinstance(m, torch.nn.LayerNorm):
                m.float()
    return gpt, lm_head


@torch.no_grad()
def hf_generate(cfg, w, prompt_emb, max_new, rep_penalty, forbid_stop_before, autocast_bf16=False,
                forced=None):
    """GPT2InferenceModel.forward + greedy _sample, driven through stock GPT2Model."""
    dt = torch.bfloat16 if autocast_bf16 else torch.float32
    gpt, lm_head = build_hf(cfg, w, dt)
    mel_emb = w["mel_embedding.weight"].to(dt)
    mel_pos = w["mel_pos_embedding.emb.weight"].to(dt)
    start, stop = cfg["start_mel_token"], cfg["stop_mel_token"]
    S = prompt_emb.shape[0]
    ctx = torch.autocast("cpu", dtype=torch.bfloat16) if autocast_bf16 else torch.autocast("cpu", enabled=False)
    with ctx:
        first = mel_emb[start] + mel_pos[0]
        emb = torch.cat([prompt_emb.to(dt), first[None]], 0)[None]
        mask = torch.ones(1, S + 1, dtype=torch.long)
        out = gpt(inputs_embeds=emb, attention_mask=mask, use_cache=True)
        past = out.past_key_values
        hidden = out.last_hidden_state[:, -1:]
        ids = [1] * S + [start]
        codes, logits = [], []
        for k in range(max_new):
            lg = lm_head(hidden)[0, 0].float()
            logits.append(lg.clone())
            s = lg.clone()
            idx = torch.tensor(sorted(set(ids)))
            sv = s[idx]
            s[idx] = torch.where(sv < 0, sv * rep_penalty, sv / rep_penalty)
            if k < forbid_stop_before:
                s[stop] = float("-inf")
            tok = int(torch.argmax(s))
            codes.append(tok)
            feed = tok if forced is None else int(forced[k])
            if (forced is None and tok == stop) or k + 1 >= max_new:
                break
            ids.append(feed)
            mask = torch.cat([mask, torch.ones(1, 1, dtype=torch.long)], 1)   # generation_utils:766-772
            pos = mask.shape[1] - S                                           # model_v2.py:158-161
            emb = (mel_emb[feed] + mel_pos[pos])[None, None]
            out = gpt(inputs_embeds=emb, attention_mask=mask, past_key_values=past, use_cache=True)
            past = out.past_key_values
            hidden = out.last_hidden_state[:, -1:]
    return np.array(codes, dtype=np.int32), torch.stack(logits).numpy()


def small_case(seed=7, n_text=9):
    cfg = gpt_config(layers=2, model_dim=256, heads=4, number_mel_codes=322, start_mel_token=320,
                     stop_mel_token=321, max_mel_tokens=60, max_text_tokens=30,
                     number_text_tokens=100, n_langs=3)
    g = torch.Generator().manual_seed(seed)
    style = torch.randn(192, generator=g)
    emo = torch.randn(cfg["model_dim"], generator=g) * 0.5
    text = torch.randint(2, 100, (n_text,), generator=g)
    return cfg, style, emo, text


def main():
    torch.manual_seed(0)
    cfg, style, emo, text = small_case()
    N = 24
    # ---- 1. fp32 restatement vs HF fp32 ----
    w32 = make_gpt_weights(cfg, seed=1234, bf16=False)
    prompt32 = prepare_gpt_inputs(w32, style, emo, text, lang=1, bf16=False)
    hf_codes, hf_logits = hf_generate(cfg, w32, prompt32, N, 10.0, N)
    orc = GptOracle(cfg, w32, bf16=False)
    o_codes, o_logits = orc.generate(prompt32, N, 10.0, N)
    d = np.abs(hf_logits - o_logits).max()
    print(f"fp32: max|logits_hf - logits_oracle| = {d:.3e}; tokens equal = {np.array_equal(hf_codes, o_codes)}")
    assert d < 2e-4 and np.array_equal(hf_codes, o_codes)

    # ---- 2. bf16 rounding-point model vs HF CPU autocast ----
    w16 = make_gpt_weights(cfg, seed=1234, bf16=True)
    emo16 = r16(emo)
    prompt16 = prepare_gpt_inputs(w16, style, emo16, text, lang=1, bf16=True)
    orc16 = GptOracle(cfg, w16, bf16=True)
    b_codes, b_logits = orc16.generate(prompt16, N, 10.0, N)
    # teacher-force HF on the oracle's tokens so both see identical inputs every step
    h_codes, h_logits = hf_generate(cfg, w16, prompt16, N, 10.0, N, autocast_bf16=True, forced=b_codes)
    f_codes, f_logits = hf_generate(cfg, w16, prompt16, N, 10.0, N, autocast_bf16=False, forced=b_codes)

    def rms(a):
        return float(np.sqrt((a.astype(np.float64) ** 2).mean()))

    # The bf16 path is only defined up to its own rounding noise (logits are bf16 values, and
    # SDPA backends differ in where they round P); the restatement must sit inside that noise:
    noise_hf = rms(h_logits - f_logits)      # HF autocast-bf16 vs HF fp32 (same bf16 weights)
    noise_or = rms(b_logits - f_logits)      # restatement bf16 vs HF fp32
    mutual = rms(b_logits - h_logits)
    print(f"bf16: logit std {b_logits.std():.3f}; rms(hf16-hf32) = {noise_hf:.4f}, "
          f"rms(oracle16-hf32) = {noise_or:.4f}, rms(oracle16-hf16) = {mutual:.4f}, "
          f"bit-identical = {100 * float((b_logits == h_logits).mean()):.1f}%, "
          f"argmax agreement = {int((h_codes == b_codes).sum())}/{len(b_codes)}")
    assert noise_or <= 1.25 * noise_hf and mutual <= 1.5 * noise_hf, \
        "rounding-point model drifted from HF autocast"

    # ---- 3. goldens ----
    out = os.path.join(ROOT, "tests", "golden", "gpt_small.npz")
    np.savez_compressed(
        out, style=style.numpy(), emo=emo.numpy(), text=text.numpy().astype(np.int32), lang=1,
        hf_fp32_codes=hf_codes, hf_fp32_logits=hf_logits.astype(np.float32),
        prompt_bf16=prompt16.numpy(), bf16_codes=b_codes, bf16_logits=b_logits.astype(np.float32),
        hf_autocast_logits=h_logits.astype(np.float32), hf_autocast_codes=h_codes,
        weight_seed=1234, n_steps=N)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
