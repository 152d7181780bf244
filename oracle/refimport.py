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
