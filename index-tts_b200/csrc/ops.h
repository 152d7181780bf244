// ops.h — generic building blocks shared by the vocoder / s2mel / codec paths.
// Activations are channels-last fp32: a [B][T][C] tensor is a row-major [B*T, C] matrix whose
// rows are time steps, so every Conv1d / Linear is a (multi-tap) GEMM with K = C contiguous.
#pragma once
#include "engine.h"

enum { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_SILU = 2, ACT_MISH = 3, ACT_GELU_TANH = 4, ACT_RELU = 5 };

// D[b][m][j] = epi( sum_{tap} sum_{k} A[b][m + tap*dil - pad][k] * W[tap][k][j] )
// rows of A outside [0, Tin) read as zero (Conv1d zero padding) or are reflected (SConv1d).
struct ConvGemm {
  const float* A = nullptr;   // [B][Tin][K]
  int B = 1, Tin = 0, K = 0;
  long long a_batch_stride = 0;  // elements; 0 → Tin*K
  int lda = 0;                   // row stride of A in elements; 0 → K
  const float* W = nullptr;   // [taps][K][N]  (N contiguous)  — SIMT layout
  const float* Wk = nullptr;  // [N][taps*K]   (K contiguous)  — tensor-core layout (optional)
  int taps = 1, dil = 1, pad = 0, reflect = 0;
  int M = 0;                  // output rows per batch
  int N = 0;
  const float* bias = nullptr;
  int biasN = 0;              // bias index = j % biasN (0 → N)
  int act = ACT_NONE;
  const float* res = nullptr;   // optional residual, indexed like out
  const float* rowscale = nullptr;  // optional per-(b,m) multiplier (masks), [B][M]
  const float* colscale = nullptr;  // optional per-column multiplier applied to (acc+bias) (layer scale)
  int accum = 0;              // out = scale * (v + out_old)
  float scale = 1.f;
  float* out = nullptr;
  long long out_batch_stride = 0;  // elements; 0 → M*N
  long long out_off = 0;           // flat offset added to m*ldo + j (may be negative: ConvTranspose)
  long long out_valid = 0;         // writes only where 0 <= flat < out_valid (0 → M*ldo)
  int ldo = 0;                     // 0 → N
};

void conv_gemm(idx_engine* e, const ConvGemm& g);

// [B][C][T] <-> [B][T][C]
void transpose_bct_to_btc(idx_engine* e, const float* in, float* out, int B, int C, int T);
void transpose_btc_to_bct(idx_engine* e, const float* in, float* out, int B, int T, int C);
