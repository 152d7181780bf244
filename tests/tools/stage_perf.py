"""GPU-box diagnostic: one full-size codes→wav tail (codec, length regulator, CFM 25 steps, BigVGAN)
with synthetic weights; prints stage times.  Used under ncu for the launch list.
    python -m tests.tools.stage_perf [n_cfm_steps] [reps]"""
import sys
import time

import numpy as np
import torch

from indextts_b200 import synth
from indextts_b200.engine import Engine, fold_weight_norm


def main():
    n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    e = Engine(0)
    c, cc, h = dict(synth.S2MEL_CFG), dict(synth.CODEC_CFG), dict(synth.BIGVGAN_V2_22K)
    ws = fold_weight_norm(synth.make_s2mel_weights(c, seed=1234))
    e.load_state_dict("s2mel.", {k: v for k, v in ws.items() if v.is_floating_point()})
    e.load_state_dict("codec.", fold_weight_norm(synth.make_codec_weights(cc, seed=4321)))
    e.s2mel_init(c)
    e.codec_init(cc)
    e.load_state_dict("bigvgan.", synth.make_bigvgan_weights(h, seed=1234))
    e.bigvgan_init(h)
    g = torch.Generator().manual_seed(0)
    n, P = 256, 861
    F = int(2 * n * 1.72)
    codes = torch.randint(0, 8192, (n,), generator=g).numpy().astype(np.int32)
    pc = torch.randn(P, 512, generator=g).numpy()
    ref_mel = (torch.randn(80, P, generator=g) * 1.5 - 4.0).numpy()
    style = torch.randn(192, generator=g).numpy()
    z = torch.randn(80, P + F, generator=g).numpy()
    for r in range(reps):
        l0 = e.launches
        t0 = time.perf_counter()
        e.codes_to_wav(codes, pc, ref_mel, style, z, F, n_steps, 0.7, want_wav=False, want_pcm16=True)
        dt = time.perf_counter() - t0
        print(f"rep {r}: wall {dt * 1000:.1f} ms, {e.s2mel_last_ms()}, bigvgan {e.bigvgan_last_ms():.2f} ms, "
              f"launches {e.launches - l0}")
    e.close()


if __name__ == "__main__":
    main()
