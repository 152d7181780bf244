"""Full-size GPT decode timing of the other modes: 8 greedy sequences per group and beam-sample (3 beams).
    python -m tests.tools.gpt_modes [steps]"""
import sys

import torch

from indextts_b200.engine import Engine
from tests.gpt_common import gpt_config, load_gpt, make_gpt_weights, prepare_gpt_inputs, r16


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    cfg = gpt_config()
    w = make_gpt_weights(cfg, seed=2025, bf16=True)
    e = Engine(0)
    load_gpt(e, cfg, w, max_batch=8, max_prompt=64)
    g = torch.Generator().manual_seed(11)
    prompts = []
    for b in range(8):
        style = torch.randn(192, generator=g)
        emo = r16(torch.randn(cfg["model_dim"], generator=g) * 0.5)
        text = torch.randint(2, 12000, (32,), generator=g)
        prompts.append(prepare_gpt_inputs(w, style, emo, text, lang=1, bf16=True).numpy())
    e.gpt_generate(prompts, 8, 10.0, forbid_stop_before=8)
    for nb in (1, 8):
        e.gpt_generate(prompts[:nb], steps, 10.0, forbid_stop_before=steps)
        t = e.gpt_last_timing()
        print(f"greedy batch {nb}: {t['decode_ms'] / t['steps'] * 1000:.1f} us/step, {nb * t['steps'] / (t['decode_ms'] * 1e-3):.0f} tokens/s")
    kw = dict(do_sample=True, num_beams=3, top_k=30, top_p=0.8, temperature=0.8, seed=1, forbid_stop_before=steps)
    for nu in (1, 2):
        e.gpt_generate(prompts[:nu], steps, 10.0, **kw)
        t = e.gpt_last_timing()
        print(f"beam-sample 3 beams x {nu} utterance(s): {t['decode_ms'] / t['steps'] * 1000:.1f} us/step, "
              f"{nu * t['steps'] / (t['decode_ms'] * 1e-3):.0f} tokens/s, launches {t['launches']}")
    e.close()


if __name__ == "__main__":
    main()
