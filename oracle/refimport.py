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
    m = types.ModuleType(name)
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
