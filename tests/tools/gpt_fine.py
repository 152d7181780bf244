"""GPU-box diagnostic: sub-phase timeline of the fused decode kernel for ALL CTAs (middle layer, last step).
    python -m tests.tools.gpt_fine [steps]
Prints, per stamp slot, min / median / max over the CTAs relative to the earliest phase-start stamp, and the per-CTA
durations between consecutive slots."""
import sys

import numpy as np
import torch

from indextts_b200.engine import Engine
from tests.gpt_common import gpt_config, load_gpt, make_gpt_weights, prepare_gpt_inputs, r16

NAMES = {
    0: "layer start", 1: "QKV x polled", 2: "QKV LN done", 3: "QKV xs ready", 4: "QKV w0 ready", 5: "QKV mma done",
    6: "QKV red bar", 7: "QKV epi done", 8: "QKV end", 9: "QKV ret", 10: "ATT q polled", 11: "ATT keys done",
    12: "ATT syncwarp", 13: "ATT smem merged", 14: "ATT end", 16: "OPJ flags ok", 17: "OPJ xs ready", 18: "OPJ w0 ready",
    19: "OPJ mma done", 20: "OPJ red bar", 21: "OPJ epi done", 22: "OPJ end", 23: "OPJ ret", 24: "FC x polled",
    25: "FC LN done", 26: "FC xs ready", 27: "FC w0 ready", 28: "FC mma done(b0)", 29: "FC red bar(b0)", 30: "FC epi done",
    31: "FC end", 32: "FC ret", 33: "PRJ f polled", 34: "PRJ xs ready", 35: "PRJ w0 ready", 36: "PRJ mma done(b0)",
    37: "PRJ red bar(b0)", 38: "PRJ epi done", 39: "PRJ end", 40: "PRJ ret",
}


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cfg = gpt_config()
    w = make_gpt_weights(cfg, seed=2025, bf16=True)
    e = Engine(0)
    load_gpt(e, cfg, w, max_batch=1, max_prompt=64)
    g = torch.Generator().manual_seed(11)
    style = torch.randn(192, generator=g)
    emo = r16(torch.randn(cfg["model_dim"], generator=g) * 0.5)
    text = torch.randint(2, 12000, (32,), generator=g)
    prompts = [prepare_gpt_inputs(w, style, emo, text, lang=1, bf16=True).numpy()]
    e.gpt_generate(prompts, 8, 10.0, forbid_stop_before=8)
    e.gpt_profile(True)
    for rep in range(2):
        e.gpt_generate(prompts, steps, 10.0, forbid_stop_before=steps)
        t = e.gpt_last_timing()
        print(f"decode {t['decode_ms'] / t['steps'] * 1000:.1f} us/step")
    f = e.gpt_profile_fine().astype(np.float64)
    t0 = f[:, 0][f[:, 0] > 0].min()
    used = [i for i in range(64) if (f[:, i] > 0).any()]
    print(f"{'slot':>4} {'name':18s} {'n':>4} {'min':>8} {'med':>8} {'max':>8}   (us after the earliest layer start)")
    for i in used:
        v = f[:, i][f[:, i] > 0]
        r = (v - t0) / 1000.0
        print(f"{i:4d} {NAMES.get(i, ''):18s} {len(v):4d} {r.min():8.2f} {np.median(r):8.2f} {r.max():8.2f}")
    print("per-CTA deltas between consecutive used slots (us): median / p90 / max")
    prev = None
    for i in used:
        if prev is not None:
            m = (f[:, i] > 0) & (f[:, prev] > 0)
            if m.any():
                d = (f[m, i] - f[m, prev]) / 1000.0
                print(f"  {prev:2d}->{i:2d} {NAMES.get(i, ''):18s} {np.median(d):7.2f} {np.percentile(d, 90):7.2f} {d.max():7.2f}")
        prev = i
    np.save("gpurun_out/gpt_fine.npy", f)
    e.close()


if __name__ == "__main__":
    main()
