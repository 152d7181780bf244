"""Mint tests/golden/tail_wiring_small.npz by EXECUTING the reference's own per-segment tail — the source lines of
`IndexTTS2.infer_generator` between `S_infer = self.semantic_codec.decode(codes)` and the int16 clamp
(indextts/infer_v2_5.py:827-856) — read from /root/reference at run time (nothing is copied into the repository) and run
against the reference's own modules (EnhancedCodec, MyModel['length_regulator'|'cfm'], BigVGAN) holding the seeded weights
of index-tts_b200/synth.py.  The restated chain of oracle/ (and of `idx_codes_to_wav`) is compared with its output.
Build container only.
    python -m oracle.make_goldens_tail"""
import os
import time
import types

import numpy as np
import torch

from indextts_b200 import synth
from oracle import refimport
from oracle.bigvgan import bigvgan_forward
from oracle.s2mel import cfm_inference, codec_decode, fold_weight_norm, length_regulate


def reference_tail_source():
    lines = open(os.path.join(refimport.REF, "indextts", "infer_v2_5.py")).read().split("\n")
    a = next(i for i, l in enumerate(lines) if "diffusion_steps = 25" in l)
    b = next(i for i, l in enumerate(lines) if "wav = torch.clamp(32767 * wav" in l)
    body = [l for l in lines[a:b + 1] if "print(" not in l]
    # the slice spans two indentation levels (inside / after the autocast block); every statement is flat, continuation
    # lines sit inside parentheses, so stripping the indentation keeps the code itself untouched
    return "\n".join(l.strip() for l in body), (a + 1, b + 1)


def main():
    refimport.setup()
    c, cc, h = synth.small_s2mel_cfg(), synth.small_codec_cfg(), synth.small_config()
    ws, wc, wb = synth.make_s2mel_weights(c, 1234), synth.make_codec_weights(cc, 4321), synth.make_bigvgan_weights(h, 1)
    s2 = refimport.s2mel_module(refimport.s2mel_args(hidden=c["hidden"], heads=c["heads"], depth=c["depth"], wn_hidden=c["wn_hidden"],
                                                     wn_layers=c["wn_layers"], content_dim=c["content_dim"], lr_in=c["lr_in"],
                                                     style_dim=c["style_dim"]))
    miss, unexp = s2.load_state_dict({"models." + k: v for k, v in ws.items()}, strict=False)
    assert not unexp, unexp
    codec = refimport.codec_module(**cc)
    miss, unexp = codec.load_state_dict(wc, strict=False)
    assert not unexp, unexp
    bv = refimport.bigvgan_module(h)
    miss, unexp = bv.load_state_dict(wb, strict=False)
    assert not unexp, unexp
    src, (l0, l1) = reference_tail_source()
    print(f"executing indextts/infer_v2_5.py:{l0}-{l1} ({len(src.splitlines())} lines)")
    g = torch.Generator().manual_seed(2)
    n, P = 7, 9
    codes = torch.randint(0, cc["codebook_size"], (1, n), generator=g)
    prompt_condition = torch.randn(1, P, c["content_dim"], generator=g)
    ref_mel = torch.randn(1, 80, P, generator=g) * 1.5 - 4.0
    style = torch.randn(1, c["style_dim"], generator=g)
    ns = dict(self=types.SimpleNamespace(semantic_codec=codec, s2mel=s2, bigvgan=bv), codes=codes, duration_factor=1.0,
              prompt_condition=prompt_condition, ref_mel=ref_mel, style=style, torch=torch, time=time, s2mel_time=0.0,
              bigvgan_time=0.0, m_start_time=0.0)
    torch.manual_seed(321)
    with torch.no_grad():
        exec(compile(src, "infer_v2_5_tail", "exec"), ns)
    wav_ref = ns["wav"]
    T = int(ns["cat_condition"].size(1))
    print("reference tail: wav", tuple(wav_ref.shape), "T =", T, "target_lengths =", int(ns["target_lengths"][0]))

    # the restated chain (what idx_codes_to_wav implements)
    wsf, wcf = fold_weight_norm(ws), fold_weight_norm(wc)
    S = codec_decode(wcf, codes)
    F = int(S.shape[1] * 1.72)
    cond = length_regulate(wsf, S, F)
    mu = torch.cat([prompt_condition, cond], 1)
    torch.manual_seed(321)
    z = torch.randn([1, 80, mu.size(1)])
    mel = cfm_inference(wsf, c, mu, torch.LongTensor([mu.size(1)]), ref_mel, style, z, 25, 0.7)
    wav = bigvgan_forward(h, wb, mel[:, :, P:].float()).squeeze().unsqueeze(0)
    wav = torch.clamp(32767 * wav, -32767.0, 32767.0)
    err = float((wav - wav_ref).abs().max())
    print(f"oracle chain vs the executed reference tail: max |diff| {err:.3e} of +-32767 (rms {float(wav_ref.pow(2).mean().sqrt()):.0f})")
    assert wav.shape == wav_ref.shape and err < 2.0
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tail_wiring_small.npz")
    np.savez_compressed(out, codes=codes.numpy(), prompt_condition=prompt_condition.numpy(), ref_mel=ref_mel.numpy(),
                        style=style.numpy(), z=z.numpy(), wav=wav_ref.numpy().astype(np.float32), F=F)
    print("wrote", out)


if __name__ == "__main__":
    main()
