"""CPU restatement of the BigVGAN-v2 generator used by IndexTTS-2 / 2.5.

TEST INFRASTRUCTURE ONLY (see oracle/gpt.py header).  Pinned against the reference's own module
(`indextts.s2mel.modules.bigvgan.bigvgan.BigVGAN`, imported from /root/reference with two
sys.modules stubs) by oracle/make_goldens_bigvgan.py; the reference's tests hold no golden
vectors for this path, so the goldens in tests/golden/bigvgan_*.npz are outputs
of that reference module run in the build container.

Restated (reference file:line):
  BigVGAN.forward                  s2mel/modules/bigvgan/bigvgan.py:360-386
  AMPBlock1.forward                bigvgan.py:132-141
  Activation1d                     alias_free_activation/torch/act.py:8-30
  UpSample1d / DownSample1d        alias_free_activation/torch/resample.py:10-58
  LowPassFilter1d / kaiser_sinc    alias_free_activation/torch/filter.py:30-101
  SnakeBeta                        activations.py:104-119
  get_padding                      bigvgan/utils.py:57-58
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from indextts_b200.synth import (BIGVGAN_V2_22K, kaiser_sinc_filter1d, make_bigvgan_weights,  # noqa: F401,E402
                                  small_config, synthetic_mel)











def activation1d(x, alpha, beta, filt, logscale=True):
    """x [B,C,T] → anti-aliased SnakeBeta (act.py:24-30)."""
    C = x.shape[1]
    ratio, ks = 2, filt.shape[-1]
    pad = ks // ratio - 1
    pad_left = pad * ratio + (ks - ratio) // 2
    pad_right = pad * ratio + (ks - ratio + 1) // 2
    y = F.pad(x, (pad, pad), mode="replicate")
    y = ratio * F.conv_transpose1d(y, filt.expand(C, -1, -1), stride=ratio, groups=C)
    y = y[..., pad_left:-pad_right]
    a = alpha[None, :, None]
    b = beta[None, :, None]
    if logscale:
        a, b = torch.exp(a), torch.exp(b)
    y = y + (1.0 / (b + 1e-9)) * torch.pow(torch.sin(y * a), 2)
    y = F.pad(y, (ks // 2 - 1, ks // 2), mode="replicate")
    return F.conv1d(y, filt.expand(C, -1, -1), stride=ratio, groups=C)


@torch.no_grad()
def bigvgan_forward(h, w, mel):
    """mel [B, 80, F] fp32 → wav [B, 1, F*prod(rates)]."""
    ls = h.get("snake_logscale", True)

    def act(name, x):
        return activation1d(x, w[name + ".act.alpha"], w[name + ".act.beta"],
                            w[name + ".upsample.filter"], ls)

    x = F.conv1d(mel, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(x, w[f"ups.{i}.0.weight"], w[f"ups.{i}.0.bias"], stride=u,
                               padding=(k - u) // 2)
        xs = None
        for j, (ks, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            rb = i * nk + j
            xb = x
            for m, d in enumerate(dil):
                xt = act(f"resblocks.{rb}.activations.{2 * m}", xb)
                xt = F.conv1d(xt, w[f"resblocks.{rb}.convs1.{m}.weight"], w[f"resblocks.{rb}.convs1.{m}.bias"],
                              dilation=d, padding=int((ks * d - d) / 2))
                xt = act(f"resblocks.{rb}.activations.{2 * m + 1}", xt)
                xt = F.conv1d(xt, w[f"resblocks.{rb}.convs2.{m}.weight"], w[f"resblocks.{rb}.convs2.{m}.bias"],
                              padding=int((ks - 1) / 2))
                xb = xt + xb
            xs = xb if xs is None else xs + xb
        x = xs / nk
    x = act("activation_post", x)
    x = F.conv1d(x, w["conv_post.weight"], w.get("conv_post.bias"), padding=3)
    return torch.tanh(x) if h.get("use_tanh_at_final", True) else torch.clamp(x, -1.0, 1.0)


def wav_rms_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()))
