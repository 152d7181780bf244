"""TEST INFRASTRUCTURE ONLY: CPU restatement of the IndexTTS v1 / v1.5 vocoder side (SURVEY.md section 8 row a13,
BASELINE config 1): the ECAPA-TDNN speaker encoder and the latent-conditioned BigVGAN.

  ECAPA_TDNN.forward                 indextts/BigVGAN/ECAPA_TDNN.py:543-582
    TDNNBlock (conv -> ReLU -> BN)   :79-128      Conv1d "same" + reflect padding: indextts/BigVGAN/nnet/CNN.py:411-470
    Res2NetBlock                     :131-191
    SEBlock                          :194-242
    AttentiveStatisticsPooling       :245-338
    SERes2NetBlock                   :341-426
  BigVGAN.forward (v1)               indextts/BigVGAN/models.py:201-249  (speaker embedding -> cond_layer / conds[i],
                                     latents [B, T, gpt_dim] in, tanh out)
  AMPBlock1 / Activation1d           models.py:33-125, alias_free_torch/* (same arithmetic as the v2 copy restated in
                                     oracle/bigvgan.py)

Pinned against the reference's own `indextts.BigVGAN.models.BigVGAN` (which owns the ECAPA encoder) by
oracle/make_goldens_v1.py -> tests/golden/v1_vocoder_small.npz.  The CUDA side of row a13 is NOT built yet
(DESIGN.md section 1): this file and its goldens are the checker it will be built against."""
import torch
import torch.nn.functional as F

from oracle.bigvgan import activation1d


def _conv_same(x, w, b, dilation=1):
    """speechbrain-style Conv1d(padding="same", padding_mode="reflect") for odd kernels (nnet/CNN.py:458-470)."""
    k = w.shape[-1]
    pad = dilation * (k - 1) // 2
    if pad:
        x = F.pad(x, (pad, pad), mode="reflect")
    return F.conv1d(x, w, b, dilation=dilation)


def _bn(x, w, p, eps=1e-5):
    return F.batch_norm(x, w[p + ".norm.running_mean"], w[p + ".norm.running_var"], w[p + ".norm.weight"], w[p + ".norm.bias"],
                        training=False, eps=eps)


def _tdnn(x, w, p, dilation=1):
    return _bn(F.relu(_conv_same(x, w[p + ".conv.conv.weight"], w[p + ".conv.conv.bias"], dilation)), w, p + ".norm")


@torch.no_grad()
def ecapa_tdnn(w, mel, prefix="speaker_encoder.", channels=(512, 512, 512, 512, 1536), kernel_sizes=(5, 3, 3, 3, 1),
               dilations=(1, 2, 3, 4, 1), scale=8):
    """mel [B, T, n_mels] (full-length utterances: lengths = 1) -> speaker embedding [B, 1, lin_neurons]."""
    q = prefix
    x = mel.transpose(1, 2)
    xl = []
    x = _tdnn(x, w, q + "blocks.0", dilations[0])
    xl.append(x)
    for i in range(1, len(channels) - 1):
        p = q + f"blocks.{i}"
        res = x
        if (p + ".shortcut.conv.weight") in w:
            res = _conv_same(x, w[p + ".shortcut.conv.weight"], w[p + ".shortcut.conv.bias"])
        y = _tdnn(x, w, p + ".tdnn1")
        parts, prev = [], None
        for j, c in enumerate(torch.chunk(y, scale, dim=1)):                # Res2NetBlock (:179-191)
            if j == 0:
                prev = c
            elif j == 1:
                prev = _tdnn(c, w, p + f".res2net_block.blocks.{j - 1}", dilations[i])
            else:
                prev = _tdnn(c + prev, w, p + f".res2net_block.blocks.{j - 1}", dilations[i])
            parts.append(prev)
        y = torch.cat(parts, dim=1)
        y = _tdnn(y, w, p + ".tdnn2")
        s = y.mean(dim=2, keepdim=True)                                     # SEBlock (:228-242), lengths = 1
        s = F.relu(F.conv1d(s, w[p + ".se_block.conv1.conv.weight"], w[p + ".se_block.conv1.conv.bias"]))
        s = torch.sigmoid(F.conv1d(s, w[p + ".se_block.conv2.conv.weight"], w[p + ".se_block.conv2.conv.bias"]))
        x = s * y + res
        xl.append(x)
    x = _tdnn(torch.cat(xl[1:], dim=1), w, q + "mfa", dilations[-1])
    # AttentiveStatisticsPooling with global context (:282-338)
    L = x.shape[-1]
    m = torch.full((x.shape[0], 1, L), 1.0 / L)
    mean = (m * x).sum(2)
    std = torch.sqrt((m * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(1e-12))
    attn = torch.cat([x, mean.unsqueeze(2).repeat(1, 1, L), std.unsqueeze(2).repeat(1, 1, L)], dim=1)
    attn = torch.tanh(_tdnn(attn, w, q + "asp.tdnn"))
    attn = F.conv1d(attn, w[q + "asp.conv.conv.weight"], w[q + "asp.conv.conv.bias"])
    attn = F.softmax(attn, dim=2)
    mean = (attn * x).sum(2)
    std = torch.sqrt((attn * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(1e-12))
    pooled = torch.cat((mean, std), dim=1).unsqueeze(2)
    pooled = F.batch_norm(pooled, w[q + "asp_bn.norm.running_mean"], w[q + "asp_bn.norm.running_var"],
                          w[q + "asp_bn.norm.weight"], w[q + "asp_bn.norm.bias"], training=False, eps=1e-5)
    out = F.conv1d(pooled, w[q + "fc.conv.weight"], w[q + "fc.conv.bias"])
    return out.transpose(1, 2)


@torch.no_grad()
def bigvgan_v1_forward(h, w, latent, mel_ref):
    """latent [B, T, gpt_dim], mel_ref [B, Tm, n_mels] -> wav [B, 1, T * prod(rates)]  (models.py:201-249,
    feat_upsample = False, cond_d_vector_in_each_upsampling_layer as configured)."""
    ls = h.get("snake_logscale", True)
    spk = ecapa_tdnn(w, mel_ref).transpose(1, 2)                            # [B, emb, 1]

    def act(name, x):
        return activation1d(x, w[name + ".act.alpha"], w[name + ".act.beta"], w[name + ".upsample.filter"], ls)

    x = latent.transpose(1, 2)
    x = F.conv1d(x, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    x = x + F.conv1d(spk, w["cond_layer.weight"], w["cond_layer.bias"])
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(x, w[f"ups.{i}.0.weight"], w[f"ups.{i}.0.bias"], stride=u, padding=(k - u) // 2)
        if h.get("cond_d_vector_in_each_upsampling_layer", True):
            x = x + F.conv1d(spk, w[f"conds.{i}.weight"], w[f"conds.{i}.bias"])
        xs = None
        for j, (ks, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            rb = i * nk + j
            xb = x
            for mi, d in enumerate(dil):
                xt = act(f"resblocks.{rb}.activations.{2 * mi}", xb)
                xt = F.conv1d(xt, w[f"resblocks.{rb}.convs1.{mi}.weight"], w[f"resblocks.{rb}.convs1.{mi}.bias"],
                              dilation=d, padding=int((ks * d - d) / 2))
                xt = act(f"resblocks.{rb}.activations.{2 * mi + 1}", xt)
                xt = F.conv1d(xt, w[f"resblocks.{rb}.convs2.{mi}.weight"], w[f"resblocks.{rb}.convs2.{mi}.bias"],
                              padding=int((ks - 1) / 2))
                xb = xt + xb
            xs = xb if xs is None else xs + xb
        x = xs / nk
    x = act("activation_post", x)
    x = F.conv1d(x, w["conv_post.weight"], w["conv_post.bias"], padding=3)
    return torch.tanh(x)
