// pipeline.cu — the per-segment "codes → waveform" tail of IndexTTS2.infer as one C-ABI call:
// semantic-codec decode → length regulator → cat(prompt_condition) → CFM solve → crop the prompt
// frames → BigVGAN → clamp/int16.  Replaces indextts/infer_v2_5.py:827-856 (one text segment);
// every intermediate stays in HBM, the host sees only the request and the waveform.
#include "stages.h"

namespace {
__global__ void crop_cols_kernel(const float* src, int ld_src, int col0, float* dst, int ncols, int rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * ncols) return;
  const int r = (int)(i / ncols), c = (int)(i % ncols);
  dst[i] = src[(long long)r * ld_src + col0 + c];
}
__global__ void pcm16_kernel(const float* wav, int16_t* pcm, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // wav = clamp(32767 * wav, -32767, 32767) (infer_v2_5.py:855) ; .type(torch.int16) truncates
  float v = fminf(fmaxf(32767.f * wav[i], -32767.f), 32767.f);
  pcm[i] = (int16_t)v;
}
}  // namespace

extern "C" int idx_codes_to_wav(idx_engine* e, const idx_vocode_request* r, int n_steps, float cfg_rate) {
  IDX_API_BEGIN
  IDX_CHECK(e && r, IDX_ERR_ARG, "null argument");
  IDX_CHECK(s2mel_ready(e->s2mel) && e->bigvgan, IDX_ERR_STATE, "s2mel / codec / bigvgan not initialised");
  IDX_CHECK(r->codes && r->n_codes >= 1 && r->F >= 1 && r->P >= 0 && r->style && r->z, IDX_ERR_ARG, "bad request");
  IDX_CHECK(r->P == 0 || (r->prompt_condition && r->ref_mel), IDX_ERR_ARG, "prompt tensors missing");
  IDX_CUDA(cudaSetDevice(e->device));
  S2melState* s = e->s2mel;
  BigvganState* bv = e->bigvgan;
  const int n = r->n_codes, F = r->F, P = r->P, T = P + F;
  const int Cd = s2mel_content_dim(s), Hs = s2mel_codec_hidden(s), C = 80, up = bigvgan_total_up(bv);
  const int Sd = s2mel_style_dim(s);        // 192 for IndexTTS-2.5 (infer_v2_5.py:218); the caller's buffer holds exactly this many
  const size_t need = codec_arena_bytes(s, n) + lr_arena_bytes(s, 2 * n, F) + cfm_arena_bytes(s, T, n_steps) +
                      bigvgan_arena_bytes(bv, 1, F) +
                      4 * ((size_t)2 * n * Hs + (size_t)T * Cd + (size_t)C * (2 * T + P + F) + Sd + (size_t)F * up * 2) +
                      (8 << 20);
  e->ensure_arena(need);
  e->arena.reset();
  int* d_codes = e->arena.get<int>(n);
  float* d_S = e->arena.get<float>((size_t)2 * n * Hs);
  float* d_mu = e->arena.get<float>((size_t)T * Cd);
  float* d_prompt = e->arena.get<float>((size_t)C * std::max(P, 1));
  float* d_style = e->arena.get<float>(Sd);
  float* d_z = e->arena.get<float>((size_t)C * T);
  float* d_mel = e->arena.get<float>((size_t)C * T);
  float* d_melF = e->arena.get<float>((size_t)C * F);
  float* d_wav = e->arena.get<float>((size_t)F * up);
  int16_t* d_pcm = (int16_t*)e->arena.alloc((size_t)F * up * 2);
  idx_to_device(e, d_codes, r->codes, (size_t)n * 4);
  if (P > 0) {
    idx_to_device(e, d_mu, r->prompt_condition, (size_t)P * Cd * 4);
    idx_to_device(e, d_prompt, r->ref_mel, (size_t)C * P * 4);
  }
  idx_to_device(e, d_style, r->style, (size_t)Sd * 4);
  idx_to_device(e, d_z, r->z, (size_t)C * T * 4);
  auto stamp = [&](int slot) {
    if (!e->events[slot]) IDX_CUDA(cudaEventCreate(&e->events[slot]));
    IDX_CUDA(cudaEventRecord(e->events[slot], e->stream));
  };
  stamp(10);
  codec_decode_dev(e, s, d_codes, n, d_S);                              // infer_v2_5.py:832
  stamp(11);
  length_regulate_dev(e, s, d_S, 2 * n, F, d_mu + (size_t)P * Cd);      // :835-840 (cat with prompt_condition)
  stamp(12);
  cfm_solve_dev(e, s, d_mu, T, d_prompt, P, d_style, d_z, n_steps, cfg_rate, d_mel);   // :841-845
  stamp(13);
  crop_cols_kernel<<<(unsigned)(((long long)C * F + 255) / 256), 256, 0, e->stream>>>(d_mel, T, P, d_melF, F, C);  // :846
  IDX_CUDA(cudaGetLastError()); e->launches++;
  bigvgan_forward_dev(e, bv, d_melF, 1, F, d_wav);                      // :850
  stamp(14);
  if (r->mel_out) idx_from_device(e, r->mel_out, d_melF, (size_t)C * F * 4);
  if (r->wav_out) idx_from_device(e, r->wav_out, d_wav, (size_t)F * up * 4);
  if (r->pcm16_out) {
    pcm16_kernel<<<(unsigned)(((long long)F * up + 255) / 256), 256, 0, e->stream>>>(d_wav, d_pcm, (long long)F * up);
    IDX_CUDA(cudaGetLastError()); e->launches++;
    idx_from_device(e, r->pcm16_out, d_pcm, (size_t)F * up * 2);
  }
  IDX_CUDA(cudaStreamSynchronize(e->stream));
  float m0, m1, m2, m3;
  IDX_CUDA(cudaEventElapsedTime(&m0, e->events[10], e->events[11]));
  IDX_CUDA(cudaEventElapsedTime(&m1, e->events[11], e->events[12]));
  IDX_CUDA(cudaEventElapsedTime(&m2, e->events[12], e->events[13]));
  IDX_CUDA(cudaEventElapsedTime(&m3, e->events[13], e->events[14]));
  s2mel_set_ms(s, m0, m1, m2);
  bigvgan_set_ms(bv, m3);
  e->check_flag("semantic code outside the codebook (codes must be cut before the stop token, infer_v2_5.py:809-821)");
  IDX_API_END(e)
}
