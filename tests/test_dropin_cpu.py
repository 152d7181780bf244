"""CPU (build container only: needs /root/reference): `dropin.load_reference_weights` against the REAL reference
modules.  The reference's own classes are instantiated (random init, small dims) through oracle/refimport.py, a
recording stand-in replaces the engine, and the test checks that every hyper-parameter the drop-in derives from the
modules' state dicts equals what the modules were constructed with, and that every tensor name the engine will ask for
(`gpt.…`, `s2mel.…`, `codec.…`, `bigvgan.…`) is present.  Skipped where the reference tree is absent (GPU box)."""
import types

import pytest
import torch

from oracle import refimport

pytestmark = pytest.mark.skipif(not refimport.available(), reason="/root/reference not present")


class RecordingEngine:
    device = 0

    def __init__(self):
        self.weights, self.calls = {}, {}

    def load_state_dict(self, prefix, sd):
        for k, v in sd.items():
            self.weights[prefix + k] = tuple(v.shape)

    def gpt_init(self, *a, **k):
        self.calls["gpt_init"] = (a, k)

    def emo_init(self, c):
        self.calls["emo_init"] = dict(c)
        self.emo_cfg = types.SimpleNamespace(**c)

    def s2mel_init(self, c):
        self.calls["s2mel_init"] = dict(c)

    def codec_init(self, c):
        self.calls["codec_init"] = dict(c)

    def bigvgan_init(self, h):
        self.calls["bigvgan_init"] = dict(h)


def test_config_derivation_from_real_reference_modules():
    from indextts_b200 import synth
    from indextts_b200.dropin import load_reference_weights
    from oracle.gpt import make_gpt_weights
    from oracle.validate_gpt_vs_hf import small_case

    cfg, _, _, _ = small_case()
    cfg = dict(cfg, n_langs=106)
    gpt = refimport.gpt_module(cfg, make_gpt_weights(cfg, seed=1, bf16=False))
    s2 = refimport.s2mel_module(refimport.s2mel_args(hidden=64, heads=1, depth=3, wn_hidden=64, wn_layers=2,
                                                     content_dim=64, lr_in=96, style_dim=24))
    codec = refimport.codec_module(codebook_size=64, hidden_size=96, codebook_dim=8, vocos_dim=48,
                                   vocos_intermediate_dim=64, vocos_num_layers=2)
    h = synth.small_config()
    bv = refimport.bigvgan_module(h)
    tts = types.SimpleNamespace(gpt=gpt, s2mel=s2, semantic_codec=codec, bigvgan=bv)
    eng = RecordingEngine()
    load_reference_weights(eng, tts, max_batch=4)

    a, k = eng.calls["gpt_init"]
    assert a[:6] == (cfg["layers"], cfg["model_dim"], cfg["heads"], cfg["number_mel_codes"], cfg["start_mel_token"],
                     cfg["stop_mel_token"])
    assert a[6] == cfg["max_mel_tokens"] + 2 + 1            # mel_pos rows (model_v2.py:398-400)
    assert k["max_batch"] == 4 and k["weights_bf16"] is True
    emo = eng.calls["emo_init"]
    assert (emo["idim"], emo["odim"], emo["linear_units"], emo["heads"], emo["blocks"]) == (1024, 32, 48, 2, 1)
    assert emo["model_dim"] == cfg["model_dim"] and emo["p_dim"] == gpt.emo_perceiver_encoder.latents.shape[-1]
    s = eng.calls["s2mel_init"]
    assert (s["hidden"], s["heads"], s["depth"], s["wn_hidden"], s["wn_layers"], s["wn_kernel"]) == (64, 1, 3, 64, 2, 5)
    assert (s["in_channels"], s["content_dim"], s["style_dim"], s["lr_in"], s["lr_convs"]) == (80, 64, 24, 96, 4)
    c = eng.calls["codec_init"]
    assert c == dict(codebook_size=64, hidden_size=96, codebook_dim=8, vocos_dim=48, vocos_intermediate_dim=64,
                     vocos_num_layers=2)
    assert eng.calls["bigvgan_init"]["upsample_rates"] == list(h["upsample_rates"])
    # the tensors the C++ side looks up by name exist under the names the reference uses
    need = ["gpt.gpt.h.0.attn.c_attn.weight", "gpt.mel_head.weight", "gpt.final_norm.weight", "gpt.spk_emb_proj.weight",
            "gpt.lang_embedding.weight", "gpt.text_pos_embedding.emb.weight", "gpt.emovec_layer.weight",
            "gpt.emo_conditioning_encoder.embed.out.0.weight", "gpt.emo_perceiver_encoder.latents",
            "s2mel.cfm.estimator.cond_projection.weight", "s2mel.cfm.estimator.wavenet.in_layers.0.conv.conv.weight",
            "s2mel.length_regulator.content_in_proj.weight", "codec.quantizer.quantizers.0.codebook.weight",
            "codec.decoder.1.weight", "bigvgan.conv_pre.weight", "bigvgan.ups.0.0.weight",
            "bigvgan.resblocks.0.convs1.0.weight", "bigvgan.conv_post.weight"]
    missing = [n for n in need if n not in eng.weights]
    assert not missing, missing
    assert not any(".weight_g" in n or ".weight_v" in n or "parametrizations" in n for n in eng.weights), \
        "weight-norm pairs must be folded before they reach the engine"


def _reference_call_sites():
    """(positional count, keyword names) of the calls `infer_generator` makes at the six seams (infer_v2_5.py:749-864),
    read from the reference source with ast."""
    import ast
    import os
    src = open(os.path.join(refimport.REF, "indextts", "infer_v2_5.py")).read()
    tree = ast.parse(src)
    want = {"self.gpt.merge_emovec": "merge_emovec", "self.gpt.inference_speech": "inference_speech",
            "self.semantic_codec.decode": "codec_decode", "self.s2mel.models['length_regulator']": "length_regulator",
            "self.s2mel.models['cfm'].inference": "cfm_inference", "self.bigvgan": "bigvgan"}
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call):
            name = ast.unparse(node.func)
            if name in want:
                kws = [k.arg for k in node.keywords if k.arg is not None]
                star = any(k.arg is None for k in node.keywords)
                found.setdefault(want[name], []).append((len(node.args), kws, star))
    return found


def test_rebound_seams_accept_the_reference_call_sites():
    """Every call `infer_v2_5.py` makes at a seam binds to the callable `attach` installs there (same positional
    arity, same keyword names) — the drop-in claim of INTEGRATION.md, checked against the reference source."""
    import inspect

    from indextts_b200 import synth
    from indextts_b200.dropin import attach
    from oracle.gpt import make_gpt_weights
    from oracle.validate_gpt_vs_hf import small_case

    cfg, _, _, _ = small_case()
    cfg = dict(cfg, n_langs=106)
    gpt = refimport.gpt_module(cfg, make_gpt_weights(cfg, seed=1, bf16=False))
    s2 = refimport.s2mel_module(refimport.s2mel_args(hidden=64, heads=1, depth=3, wn_hidden=64, wn_layers=2,
                                                     content_dim=64, lr_in=96, style_dim=24))
    codec = refimport.codec_module(codebook_size=64, hidden_size=96, codebook_dim=8, vocos_dim=48,
                                   vocos_intermediate_dim=64, vocos_num_layers=2)
    bv = refimport.bigvgan_module(synth.small_config())
    tts = types.SimpleNamespace(gpt=gpt, s2mel=s2, semantic_codec=codec, bigvgan=bv)
    eng = RecordingEngine()
    attach(tts, engine=eng)
    seams = {"merge_emovec": tts.gpt.merge_emovec, "inference_speech": tts.gpt.inference_speech,
             "codec_decode": tts.semantic_codec.decode, "length_regulator": tts.s2mel.models["length_regulator"].forward,
             "cfm_inference": tts.s2mel.models["cfm"].inference, "bigvgan": tts.bigvgan.forward}
    sites = _reference_call_sites()
    assert set(sites) == set(seams), (sorted(sites), sorted(seams))
    for name, calls in sites.items():
        sig = inspect.signature(seams[name])
        for npos, kws, star in calls:
            try:
                sig.bind(*([None] * npos), **{k: None for k in kws})
            except TypeError as e:
                raise AssertionError(f"{name}: reference call ({npos} positional, keywords {kws}) does not bind to {sig}: {e}")
    # the rebound objects are the reference's own module objects: infer_generator reaches them unchanged
    assert tts.gpt is gpt and tts.bigvgan is bv and tts._b200_engine is eng


def test_attach_v1_on_real_reference_v1_modules():
    """v1 / v1.5 drop-in (row a13): `attach_v1` on the reference's own v1 `UnifiedVoice` and `BigVGAN` classes — the derived
    prompt-encoder / vocoder configuration equals what the modules were built with, and the calls `indextts/infer.py` makes
    at the three seams bind to the rebound callables."""
    import ast
    import inspect
    import os

    from indextts_b200 import synth
    from indextts_b200.dropin import attach_v1
    from oracle.make_goldens_v1 import reference_module
    from oracle.validate_gpt_vs_hf import small_case

    cfg, _, _, _ = small_case()
    ccfg = synth.small_v1_cond_cfg(cfg["model_dim"])
    gpt = refimport.gpt_module_v1(cfg, ccfg, synth.make_gpt_v1_weights(cfg, ccfg, seed=3), kv_cache=False)
    h = synth.small_v1_config()
    bv = reference_module(h, synth.make_bigvgan_v1_weights(h, seed=5))
    tts = types.SimpleNamespace(gpt=gpt, bigvgan=bv)

    class Rec(RecordingEngine):
        def v1_cond_init(self, c, n):
            self.calls["v1_cond_init"] = (dict(c), n)

        def v1_vocoder_init(self, hh):
            self.calls["v1_vocoder_init"] = dict(hh)

    eng = Rec()
    attach_v1(tts, engine=eng)
    c, n = eng.calls["v1_cond_init"]
    assert n == 32
    for k in ("idim", "odim", "linear_units", "heads", "blocks", "cnn_kernel", "p_dim", "p_heads", "p_dim_head", "p_depth", "p_ff_mult"):
        assert c[k] == ccfg[k], (k, c[k], ccfg[k])
    a, kw = eng.calls["gpt_init"]
    assert a[:3] == (cfg["layers"], cfg["model_dim"], cfg["heads"]) and kw["weights_bf16"] is False
    assert eng.calls["v1_vocoder_init"]["gpt_dim"] == h["gpt_dim"]
    assert "bigvgan_v1.speaker_encoder.blocks.0.conv.conv.weight" in eng.weights and "bigvgan_v1.cond_layer.weight" in eng.weights
    assert "gpt.conditioning_encoder.embed.out.0.weight" in eng.weights and "gpt.perceiver_encoder.latents" in eng.weights
    # call sites of indextts/infer.py
    tree = ast.parse(open(os.path.join(refimport.REF, "indextts", "infer.py")).read())
    want = {"self.gpt.inference_speech": tts.gpt.inference_speech, "self.gpt": tts.gpt.forward, "self.bigvgan": tts.bigvgan.forward}
    seen = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and ast.unparse(node.func) in want:
            name = ast.unparse(node.func)
            kws = [k.arg for k in node.keywords if k.arg is not None]
            inspect.signature(want[name]).bind(*([None] * len(node.args)), **{k: None for k in kws})
            seen.add(name)
    assert seen == set(want)
