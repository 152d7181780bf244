"""GPU: the drop-in binding (index-tts_b200/dropin.py) on a stand-in for the reference IndexTTS2 object.
The stand-in exposes exactly the attributes `attach` touches (modules with the reference state-dict
names and hyper-parameter attributes); the test then executes the per-segment tail of
`infer_generator` (infer_v2_5.py:771-856) through the rebound callables and checks it against the oracle."""
import types

import numpy as np
import pytest
import torch

from indextts_b200 import synth
from indextts_b200.dropin import attach
from oracle.bigvgan import bigvgan_forward
from oracle import emo as emo_oracle
from oracle.gpt import GptOracle, prepare_gpt_inputs
from oracle.s2mel import cfm_inference, codec_decode, fold_weight_norm, length_regulate
from oracle.validate_gpt_vs_hf import small_case

pytestmark = pytest.mark.gpu


class _Mod:
    def __init__(self, sd, **attrs):
        self._sd = sd
        self.__dict__.update(attrs)

    def state_dict(self):
        return self._sd


def test_attach_rebinds_the_reference_seams(engine):
    cfg, style, emo, text = small_case()
    wg = synth.make_gpt_weights(cfg, seed=1234, bf16=True)
    ec = dict(synth.small_emo_cfg(), model_dim=cfg["model_dim"])
    we = synth.make_emo_weights(ec, seed=77)
    wg_all = dict(wg, **we)
    c, cc, h = synth.small_s2mel_cfg(), synth.small_codec_cfg(), synth.small_config()
    ws, wc, wb = synth.make_s2mel_weights(c, 1234), synth.make_codec_weights(cc, 4321), synth.make_bigvgan_weights(h, 1)
    tts = types.SimpleNamespace()
    tts.gpt = _Mod(wg_all, model_dim=cfg["model_dim"], heads=cfg["heads"], number_mel_codes=cfg["number_mel_codes"],
                   start_mel_token=cfg["start_mel_token"], stop_mel_token=cfg["stop_mel_token"],
                   max_mel_tokens=cfg["max_mel_tokens"], gpt=types.SimpleNamespace(h=[None] * cfg["layers"]))
    cfm = _Mod({}, in_channels=80)
    tts.s2mel = _Mod({"models." + k: v for k, v in ws.items()}, models={"cfm": cfm, "length_regulator": None})
    tts.semantic_codec = _Mod(wc)
    tts.bigvgan = _Mod(wb, h=h)
    attach(tts, engine=engine)

    # --- infer_v2_5.py:759: emotion vector through the rebound merge_emovec
    ge = torch.Generator().manual_seed(5)
    spk_f, emo_f = torch.randn(1, 37, ec["idim"], generator=ge), torch.randn(1, 29, ec["idim"], generator=ge)
    ev = tts.gpt.merge_emovec(spk_f, emo_f, torch.tensor([37]), torch.tensor([29]), alpha=0.6)
    ev_ref = emo_oracle.merge_emovec(we, ec, spk_f[0], emo_f[0], 0.6)
    assert ev.shape == (1, cfg["model_dim"])
    assert np.abs(ev[0].cpu().numpy() - ev_ref.numpy()).max() < 2e-2 * max(1.0, float(ev_ref.abs().max()))

    # --- infer_v2_5.py:771-791: speech tokens
    emo16 = synth.r16(emo)
    n = 12
    codes, _ = tts.gpt.inference_speech(None, text[None], langs=torch.tensor([1]), emo_vec=emo16[None],
                                        campplus_embedding=style[None], do_sample=False, num_beams=1,
                                        repetition_penalty=10.0, max_generate_length=n)
    prompt = prepare_gpt_inputs(wg, style, emo16, text, lang=1, bf16=True)
    o_codes, _ = GptOracle(cfg, wg, bf16=True).generate(prompt, n, 10.0, 0)
    k = min(len(o_codes), codes.shape[1])
    agree = int((codes[0, :k].cpu().numpy() == o_codes[:k]).sum())
    assert agree >= k - 2, (codes, o_codes)          # bf16 near-ties may flip a pick

    # --- :827-850 with the rebound modules (codes clipped to the small codebook)
    cds = torch.from_numpy(o_codes.astype(np.int64) % cc["codebook_size"])[None]
    S = tts.semantic_codec.decode(cds.cuda())
    ylen = int(S.shape[1] * 1.72)
    cond = tts.s2mel.models["length_regulator"](S[:, :, : c["lr_in"]] if S.shape[-1] >= c["lr_in"] else S,
                                                 ylens=torch.LongTensor([ylen]), n_quantizers=3, f0=None)[0]
    wsf, wcf = fold_weight_norm(ws), fold_weight_norm(wc)
    S_ref = codec_decode(wcf, cds)
    assert np.abs(S.cpu().numpy() - S_ref.numpy()).max() < 1e-2
    cond_ref = length_regulate(wsf, S_ref, ylen)
    assert np.abs(cond.cpu().numpy() - cond_ref.numpy()).max() < 1e-2
    g = torch.Generator().manual_seed(2)
    P = 9
    pc = torch.randn(1, P, c["content_dim"], generator=g)
    ref_mel = torch.randn(1, 80, P, generator=g) * 1.5 - 4.0
    sty = torch.randn(1, c["style_dim"], generator=g)
    cat = torch.cat([pc.cuda(), cond], dim=1)
    torch.manual_seed(123)
    mel = tts.s2mel.models["cfm"].inference(cat, torch.LongTensor([cat.size(1)]), ref_mel.cuda(), sty.cuda(), None, 6,
                                            inference_cfg_rate=0.7)
    torch.manual_seed(123)
    z = torch.randn([1, 80, cat.size(1)], device="cuda").cpu()
    mel_ref = cfm_inference(wsf, c, torch.cat([pc, cond_ref], 1), torch.LongTensor([cat.size(1)]), ref_mel, sty, z, 6, 0.7)
    assert np.abs(mel.cpu().numpy() - mel_ref.numpy()).max() < 3e-2
    wav = tts.bigvgan.forward(mel[:, :, P:])
    wav_ref = bigvgan_forward(h, wb, mel_ref[:, :, P:])
    assert wav.shape == wav_ref.shape
    assert float(np.sqrt(((wav.cpu().numpy() - wav_ref.numpy()) ** 2).mean())) < 3e-3
