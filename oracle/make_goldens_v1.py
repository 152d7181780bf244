"""Mint tests/golden/v1_vocoder_small.npz from the reference's own v1 `BigVGAN` (indextts/BigVGAN/models.py:129-249, which
owns the ECAPA-TDNN speaker encoder, ECAPA_TDNN.py:429-582) with the seeded weights of index-tts_b200/synth.py loaded into
it.  Build container only.
    python -m oracle.make_goldens_v1"""
import os

import numpy as np
import torch

from indextts_b200 import synth
from oracle import refimport, v1


def reference_module(h, w):
    refimport.setup()
    from indextts.BigVGAN.models import BigVGAN
    m = BigVGAN(refimport.AttrDict(dict(h)))
    m.remove_weight_norm()
    m.eval()
    sd = m.state_dict()
    unknown = [k for k in w if k not in sd]
    assert not unknown, unknown
    missing = [k for k in sd if k not in w and not k.endswith("num_batches_tracked")]
    assert not missing, missing
    m.load_state_dict(w, strict=False)
    return m


def main():
    h = synth.small_v1_config()
    seed = 4321
    w = synth.make_bigvgan_v1_weights(h, seed=seed)
    m = reference_module(h, w)
    g = torch.Generator().manual_seed(7)
    latent = torch.randn(1, 9, h["gpt_dim"], generator=g)
    mel_ref = torch.randn(1, 37, h["num_mels"], generator=g) * 1.5 - 4.0
    with torch.no_grad():
        spk_ref = m.speaker_encoder(mel_ref, torch.tensor([1.0]))
        wav_ref, _ = m(latent, mel_ref, torch.tensor([1.0]))
    spk = v1.ecapa_tdnn(w, mel_ref)
    wav = v1.bigvgan_v1_forward(h, w, latent, mel_ref)
    e1, e2 = float((spk - spk_ref).abs().max()), float((wav - wav_ref).abs().max())
    print(f"ECAPA-TDNN embedding max |diff| vs reference {e1:.2e} (std {float(spk_ref.std()):.2f}); wav max |diff| {e2:.2e} "
          f"(rms {float(wav_ref.pow(2).mean().sqrt()):.3f})")
    assert e1 < 1e-4 and e2 < 1e-5
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "v1_vocoder_small.npz")
    np.savez_compressed(out, seed=seed, latent=latent.numpy(), mel_ref=mel_ref.numpy(), spk=spk_ref.numpy(), wav=wav_ref.numpy())
    print("wrote", out)


if __name__ == "__main__":
    main()
