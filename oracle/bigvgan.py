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

BIGVGAN_V2_22K = dict(  # s2mel/modules/bigvgan/config.json:11-21
    num_mels=80, upsample_rates=[4, 4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
    upsample_initial_channel=1536, resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], use_tanh_at_final=False,
    use_bias_at_final=False, activation="snakebeta", snake_logscale=True, resblock="1")


def small_config(ch=64, rates=(4, 2), ksz=(8, 4)):
    h = dict(BIGVGAN_V2_22K)
    h.update(upsample_initial_channel=ch, upsample_rates=list(rates), upsample_kernel_sizes=list(ksz))
    return h


def kaiser_sinc_filter1d(cutoff=0.25, half_width=0.3, kernel_size=12):  # filter.py:30-70
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    time = torch.arange(-half_size, half_size) + 0.5
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    filt = filt / filt.sum()
    return filt.view(1, 1, kernel_size)


def make_bigvgan_weights(h, seed=1234):
    """Seeded synthetic generator weights under the reference state-dict names (weight norm
    already removed).  Conv weights are scaled to 1/sqrt(fan_in) so activations neither vanish
    nor explode through the 6 stages (the reference's std 0.01 init, bigvgan/utils.py:45-48,
    would drive everything to ~0); alpha/beta ~ N(0, 0.3) in log scale."""
    g = torch.Generator().manual_seed(seed)
    w = {}

    def conv(name, co, ci, k, bias=True, gain=1.0):
        w[name + ".weight"] = torch.randn(co, ci, k, generator=g) * (gain / math.sqrt(ci * k))
        if bias:
            w[name + ".bias"] = torch.randn(co, generator=g) * 0.05

    def act(name, c):
        w[name + ".act.alpha"] = torch.randn(c, generator=g) * 0.3
        w[name + ".act.beta"] = torch.randn(c, generator=g) * 0.3
        filt = kaiser_sinc_filter1d()
        w[name + ".upsample.filter"] = filt.clone()
        w[name + ".downsample.lowpass.filter"] = filt.clone()

    ch = h["upsample_initial_channel"]
    conv("conv_pre", ch, h["num_mels"], 7)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        # ConvTranspose1d weight is [in, out, k]; each output sees k/u taps
        w[f"ups.{i}.0.weight"] = torch.randn(ch, ch // 2, k, generator=g) * (1.0 / math.sqrt(ch * k / u))
        w[f"ups.{i}.0.bias"] = torch.randn(ch // 2, generator=g) * 0.05
        ch //= 2
        for j, ks in enumerate(h["resblock_kernel_sizes"]):
            rb = i * nk + j
            for m in range(3):
                conv(f"resblocks.{rb}.convs1.{m}", ch, ch, ks, gain=0.7)
                conv(f"resblocks.{rb}.convs2.{m}", ch, ch, ks, gain=0.5)
            for q in range(6):
                act(f"resblocks.{rb}.activations.{q}", ch)
    act("activation_post", ch)
    conv("conv_post", 1, ch, 7, bias=h.get("use_bias_at_final", True), gain=0.06)
    return w


def synthetic_mel(B, F_, seed=0, num_mels=80):
    """log-mel shaped input: N(-5, 2^2) clipped to [-11.5, 2]."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, num_mels, F_, generator=g) * 2.0 - 5.0).clamp(-11.5, 2.0)


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
