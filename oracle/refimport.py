"""Import the reference's own modules from /root/reference (build container only; the GPU box has
no /root/reference — nothing under tests -m gpu, smoke() or bench.py may call this).

Three stubs are needed: matplotlib/librosa (training utilities only,
bigvgan/utils.py:6-12), munch.Munch (attr-dict, s2mel/modules/commons.py:6) and
indextts.s2mel.dac.nn.quantize.VectorQuantize (dac/__init__.py imports audiotools)."""
import os
import sys
import types

REF = os.environ.get("IDX_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "indextts"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    import importlib.machinery
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)     # importlib.util.find_spec(name) must not raise
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def setup():
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?) — goldens are generated in the build container")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for mod in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "librosa", "librosa.util",
                "librosa.filters"):
        try:
            __import__(mod)
        except Exception:
            m = _stub(mod)
            m.use = lambda *a, **k: None
            m.normalize = lambda *a, **k: None
            m.mel = lambda *a, **k: None
    if "munch" not in sys.modules:
        class Munch(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v
        _stub("munch", Munch=Munch)
    # indextts.s2mel.dac imports audiotools at package import; provide the one class used
    import torch
    pkg = "indextts.s2mel.dac"
    if pkg not in sys.modules:
        for name in (pkg, pkg + ".nn", pkg + ".nn.quantize"):
            m = _stub(name)
            m.__path__ = []
        sys.modules[pkg + ".nn.quantize"].VectorQuantize = type("VectorQuantize", (torch.nn.Module,), {})


def bigvgan_module(h):
    setup()
    from indextts.s2mel.modules.bigvgan.bigvgan import BigVGAN
    from indextts.s2mel.modules.bigvgan.env import AttrDict
    m = BigVGAN(AttrDict(dict(h)), use_cuda_kernel=False)
    m.remove_weight_norm()
    m.eval()
    return m


class AttrDict(dict):
    """attribute dictionary with hasattr semantics (missing key → AttributeError)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def __setattr__(self, k, v):
        self[k] = v


def s2mel_args(hidden=512, heads=8, depth=13, wn_hidden=512, wn_layers=8, content_dim=512,
               lr_in=1024, style_dim=192):
    """The s2mel section of checkpoints/config.yaml [ASSUMED dims, SURVEY.md §8] as the reference
    classes read it (commons.py:390-414, diffusion_transformer.py:103-184)."""
    return AttrDict(
        reg_loss_type="l1", dit_type="DiT",
        DiT=dict(hidden_dim=hidden, num_heads=heads, depth=depth, class_dropout_prob=0.1, block_size=8192,
                 in_channels=80, style_condition=True, final_layer_type="wavenet", target="mel",
                 content_dim=content_dim, content_codebook_size=1024, content_type="discrete",
                 f0_condition=False, n_f0_bins=512, content_codebooks=1, is_causal=False,
                 long_skip_connection=True, zero_prompt_speech_token=False, time_as_token=False,
                 style_as_token=False, uvit_skip_connection=True, add_resblock_in_transformer=False),
        wavenet=dict(hidden_dim=wn_hidden, num_layers=wn_layers, kernel_size=5, dilation_rate=1, p_dropout=0.2,
                     style_condition=True),
        style_encoder=dict(dim=style_dim),
        length_regulator=dict(channels=content_dim, is_discrete=False, in_channels=lr_in,
                              content_codebook_size=2048, sampling_ratios=[1, 1, 1, 1],
                              vector_quantize=False, n_codebooks=1, quantizer_dropout=0.0,
                              f0_condition=False, n_f0_bins=512),
    )


def s2mel_module(args):
    setup()
    from indextts.s2mel.modules.commons import MyModel
    m = MyModel(args)
    m.eval()
    m.models["cfm"].estimator.setup_caches(max_batch_size=2, max_seq_length=8192)
    return m


def codec_module(**kw):
    setup()
    from indextts.codec.models import EnhancedCodec
    m = EnhancedCodec(**kw)
    m.eval()
    return m


def _import_gpt(modname):
    """Import `indextts.gpt.<modname>` under transformers 5.x (see gpt_module).  Returns the module."""
    import importlib
    import importlib.util
    import re

    import torch
    # transformers probes optional dependencies at import: it must not meet the librosa stub of setup() (it would then
    # import soxr); park the stubs while its modules load
    parked = {k: sys.modules.pop(k) for k in list(sys.modules)
              if k.split(".")[0] == "librosa" and getattr(sys.modules[k], "__file__", None) is None}
    import transformers  # noqa: F401
    from transformers import GenerationConfig, PretrainedConfig
    from transformers import GPT2Config, GPT2Model  # noqa: F401  (the v1 model builds on the stock classes, model.py:262)
    import transformers.generation  # noqa: F401
    sys.modules.update(parked)
    setup()

    def dummy(name):
        return type(name, (), {"__init__": lambda self, *a, **k: None})

    def stub_module(name):
        mod = types.ModuleType(name)
        mod.__path__ = []

        def _ga(n):
            if n.startswith("__"):
                raise AttributeError(n)
            return dummy(n)
        mod.__getattr__ = _ga
        sys.modules[name] = mod
        parent, _, child = name.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], child, mod)
        return mod

    # the reference's own BeamSearchScorer stands in for the module transformers 5.x dropped
    if "transformers.generation.beam_constraints" not in sys.modules:
        stub_module("transformers.generation.beam_constraints")
    if "transformers.generation.beam_search" not in sys.modules:
        spec = importlib.util.spec_from_file_location("transformers.generation.beam_search",
                                                      os.path.join(REF, "indextts", "gpt", "transformers_beam_search.py"))
        bs = importlib.util.module_from_spec(spec)
        sys.modules["transformers.generation.beam_search"] = bs
        spec.loader.exec_module(bs)
        setattr(sys.modules["transformers.generation"], "beam_search", bs)
    m = None
    for _ in range(100):
        for k in [k for k in sys.modules if k.startswith("indextts.gpt")]:
            del sys.modules[k]
        try:
            m = importlib.import_module("indextts.gpt." + modname)
            break
        except ModuleNotFoundError as e:
            stub_module(e.name)
        except ImportError as e:
            mm = re.search(r"cannot import name '(\w+)' from '([\w\.]+)'", str(e))
            if not mm:
                raise
            setattr(sys.modules[mm.group(2)], mm.group(1), dummy(mm.group(1)))
    assert m is not None, "could not import indextts.gpt." + modname
    tgu = sys.modules["indextts.gpt.transformers_generation_utils"]
    tgu.isin_mps_friendly = lambda elements, test_elements: torch.isin(elements, test_elements)
    if not hasattr(PretrainedConfig, "_get_non_default_generation_parameters"):
        PretrainedConfig._get_non_default_generation_parameters = lambda self: {}
    names = set(re.findall(r"generation_config\.(\w+)", open(os.path.join(REF, "indextts", "gpt", "transformers_generation_utils.py")).read()))
    probe = GenerationConfig()
    for n in sorted(names):
        if not n.startswith("_") and not hasattr(probe, n):
            setattr(GenerationConfig, n, None)
    return m


def _load(g, weights):
    sd = g.state_dict()
    unknown = [k for k in weights if k not in sd]
    assert not unknown, f"oracle weight names unknown to the reference module: {unknown}"
    for k, v in weights.items():
        assert tuple(sd[k].shape) == tuple(v.shape), (k, tuple(sd[k].shape), tuple(v.shape))
    g.load_state_dict(weights, strict=False)
    g.eval()
    return g


def gpt_module(cfg, weights):
    """The reference's own `UnifiedVoice` (indextts/gpt/model_v2.py:305-493, spk_cond_mode="campplus" = IndexTTS-2.5)
    with its vendored GPT2 blocks and generate() loop, running on CPU under transformers 5.x.

    `indextts/gpt/transformers_generation_utils.py` was vendored from transformers 4.52 and imports ~25 names that
    5.x removed.  None of them is on the executed path of greedy / beam decoding: missing classes and constants are
    stubbed with empty types, `isin_mps_friendly` is given its definition (torch.isin), the removed
    `transformers.generation.beam_search` module is replaced by the reference's own vendored
    `indextts/gpt/transformers_beam_search.py`, and three legacy GenerationConfig attributes default to None.
    `weights` (reference state-dict names, oracle.gpt.make_gpt_weights) are loaded into the module."""
    m = _import_gpt("model_v2")
    tiny = dict(output_size=32, linear_units=48, attention_heads=2, num_blocks=1, input_layer="conv2d2", perceiver_mult=2)
    g = m.UnifiedVoice(layers=cfg["layers"], model_dim=cfg["model_dim"], heads=cfg["heads"],
                       max_text_tokens=cfg["max_text_tokens"], max_mel_tokens=cfg["max_mel_tokens"],
                       number_text_tokens=cfg["number_text_tokens"], number_mel_codes=cfg["number_mel_codes"],
                       start_mel_token=cfg["start_mel_token"], stop_mel_token=cfg["stop_mel_token"],
                       condition_type="conformer_perceiver", condition_module=dict(tiny), emo_condition_module=dict(tiny),
                       spk_cond_mode="campplus")
    _load(g, weights)
    g.post_init_gpt2_config(use_deepspeed=False, kv_cache=True, half=False)
    return g


def gpt_module_v1(cfg, ccfg, weights, kv_cache=False):
    """The reference's v1 / v1.5 `UnifiedVoice` (indextts/gpt/model.py:305-440): 32-latent conformer-perceiver prompt
    from a 100-bin mel; `kv_cache=False` is what `infer.py:101` uses on CPU."""
    m = _import_gpt("model")
    g = m.UnifiedVoice(layers=cfg["layers"], model_dim=cfg["model_dim"], heads=cfg["heads"],
                       max_text_tokens=cfg["max_text_tokens"], max_mel_tokens=cfg["max_mel_tokens"],
                       number_text_tokens=cfg["number_text_tokens"], number_mel_codes=cfg["number_mel_codes"],
                       start_mel_token=cfg["start_mel_token"], stop_mel_token=cfg["stop_mel_token"],
                       condition_type="conformer_perceiver",
                       condition_module=dict(output_size=ccfg["odim"], linear_units=ccfg["linear_units"],
                                             attention_heads=ccfg["heads"], num_blocks=ccfg["blocks"], input_layer="conv2d2",
                                             perceiver_mult=ccfg["p_ff_mult"]))
    _load(g, weights)
    g.post_init_gpt2_config(use_deepspeed=False, kv_cache=kv_cache, half=False)
    return g
