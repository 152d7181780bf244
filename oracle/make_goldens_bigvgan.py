"""Mint BigVGAN golden vectors from the REFERENCE module (build container only) and pin the
restatement oracle/bigvgan.py against it.

    python -m oracle.make_goldens_bigvgan

Writes tests/golden/bigvgan_small.npz (reduced-channel generator, F=40, B=2),
tests/golden/bigvgan_full_f12.npz (full bigvgan_v2_22khz_80band_256x geometry, F=12) and
tests/golden/activation1d_kat.npz (Activation1d known-answer test incl. SURVEY.md A.1).
Weights are NOT stored: they are regenerated from the seed by oracle.bigvgan.make_bigvgan_weights
(torch CPU generator, same image on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refimport  # noqa: E402
from oracle.bigvgan import (BIGVGAN_V2_22K, activation1d, bigvgan_forward, kaiser_sinc_filter1d,  # noqa: E402
                            make_bigvgan_weights, small_config, synthetic_mel)

GOLD = os.path.join(ROOT, "tests", "golden")


@torch.no_grad()
def run_reference(h, w, mel):
    m = refimport.bigvgan_module(h)
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m(mel)


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    # ---- Activation1d KAT ----
    refimport.setup()
    from indextts.s2mel.modules.bigvgan.alias_free_activation.torch.act import Activation1d
    from indextts.s2mel.modules.bigvgan import activations
    x = torch.arange(32, dtype=torch.float32).reshape(1, 4, 8) / 10 - 1.5
    sb = activations.SnakeBeta(4, alpha_logscale=True)
    sb.alpha.data = torch.tensor([0.3, -0.2, 0.0, 0.5])
    sb.beta.data = torch.tensor([-0.1, 0.2, 0.0, 0.4])
    a1 = Activation1d(activation=sb)
    y_ref = a1(x)
    kat = np.array([-0.4579573, -0.4191469, -0.3698377, -0.3153958, -0.2515286, -0.1793794, -0.0913750, -0.0037160])
    assert np.abs(y_ref[0, 1].numpy() - kat).max() < 2e-6, y_ref[0, 1]
    filt = kaiser_sinc_filter1d()
    assert torch.equal(filt, a1.upsample.filter) and torch.equal(filt, a1.downsample.lowpass.filter)
    y_or = activation1d(x, sb.alpha.data, sb.beta.data, filt)
    assert torch.equal(y_or, y_ref)
    g = torch.Generator().manual_seed(3)
    xr = torch.randn(2, 37, 53, generator=g) * 1.5
    sb2 = activations.SnakeBeta(37, alpha_logscale=True)
    sb2.alpha.data = torch.randn(37, generator=g) * 0.3
    sb2.beta.data = torch.randn(37, generator=g) * 0.3
    yr = Activation1d(activation=sb2)(xr)
    assert torch.equal(activation1d(xr, sb2.alpha.data, sb2.beta.data, filt), yr)
    np.savez_compressed(os.path.join(GOLD, "activation1d_kat.npz"), x=x.numpy(), alpha=sb.alpha.data.numpy(),
                        beta=sb.beta.data.numpy(), y=y_ref.numpy(), filt=filt.numpy().reshape(-1),
                        xr=xr.numpy(), alpha_r=sb2.alpha.data.numpy(), beta_r=sb2.beta.data.numpy(),
                        yr=yr.numpy())
    print("activation1d: KAT ok, restatement bit-identical to the reference module")

    # ---- small generator ----
    h = small_config()
    w = make_bigvgan_weights(h, seed=1234)
    mel = synthetic_mel(2, 40, seed=1)
    ref = run_reference(h, w, mel)
    orc = bigvgan_forward(h, w, mel)
    d = (ref - orc).abs().max().item()
    print(f"small: reference vs restatement max|diff| = {d:.2e}; wav rms {ref.pow(2).mean().sqrt():.3f}; "
          f"clipped {(ref.abs() >= 1).float().mean():.4f}")
    assert d < 1e-5
    np.savez_compressed(os.path.join(GOLD, "bigvgan_small.npz"), mel=mel.numpy(), wav=ref.numpy(), seed=1234)

    # ---- full geometry, short mel ----
    h = dict(BIGVGAN_V2_22K)
    w = make_bigvgan_weights(h, seed=1234)
    nparam = sum(v.numel() for k, v in w.items() if "filter" not in k)
    mel = synthetic_mel(1, 12, seed=2)
    ref = run_reference(h, w, mel)
    orc = bigvgan_forward(h, w, mel)
    d = (ref - orc).abs().max().item()
    print(f"full: params {nparam}; reference vs restatement max|diff| = {d:.2e}; wav rms "
          f"{ref.pow(2).mean().sqrt():.3f}; clipped {(ref.abs() >= 1).float().mean():.4f}")
    assert d < 1e-5
    np.savez_compressed(os.path.join(GOLD, "bigvgan_full_f12.npz"), mel=mel.numpy(), wav=ref.numpy(), seed=1234)
    print("wrote goldens to", GOLD)


if __name__ == "__main__":
    main()
