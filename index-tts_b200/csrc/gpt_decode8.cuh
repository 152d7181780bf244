// gpt_decode8.cuh — decode of 2..8 sequences per group on the round-2 tile / ring structure (included by gpt_decode.cu
// after gpt_decode1.cuh, inside its anonymous namespace).
//
// Replaces `gpt_fused_kernel<8, NPL>` for plain (non-beam) multi-sequence decode — BASELINE configs 3 and 5 decode 8
// utterances per group.  The round-1 kernel spent 47 us per layer (profile of round 2: O-proj 11.7 us merging key splits for
// 8 x 20 (row, head) pairs, attention 8.6, FC 6.4, QKV 5.3, barriers 10.9): 1.1 ms per 8-row step.  Here:
//   * the m16n8k16 MMA's N = 8 columns are the 8 sequences (they were 7/8 wasted at batch 1), 16 real weight rows per tile,
//     all tiles of a phase in flight, the same row-granular bulk-copy weight ring and phase barriers as gpt_decode1_kernel;
//   * LayerNorm: one warp per sequence row; the normalised bf16 rows go to shared memory (swizzled for ldmatrix); phases
//     without a LayerNorm (O-proj, PROJ) take their B fragments straight from global memory into registers;
//   * attention: one CTA per (sequence, head PAIR), the two heads on two groups of four warps, every key of the head
//     (no split, no merge pass), the normalised output written once;
//   * hand-overs between phases are grid barriers (the tagged-word protocol of the 1-row kernel would poll 8x the data).
// Same arithmetic, rounding points, sampler contract and KV-cache layout as the other GPT kernels; prefill stays on
// gpt_fused_kernel<8, .>, which leaves the cache and the per-sequence state exactly as this kernel expects them.

constexpr int B8 = 8;                                   // rows of the group (N of the MMA)
constexpr int RED8_FLOATS = MAXIT * NCW * 16 * B8;      // K-split partial sums [item][warp][16 weight rows][8 sequences]

struct Smem8 {
  __nv_bfloat16* ring;  // [R][D]
  __nv_bfloat16* xs;    // [8][D] LayerNorm output rows (bf16, 16-byte chunks XOR-swizzled by the row)
  float* red;           // [RED8_FLOATS]
  float* att;           // [NCW][66] attention merge
  uint64_t* full;       // [NBAR]
  Phase1* pht;          // [5]
  float* bias_s;
  float* xres;          // [8][ocap]
  unsigned* seen_s;     // [(V+31)/32] repetition bitmap of the sequence this CTA samples
  float* lnp;           // [2][2][D]
};

// B fragments (the 8 sequences' activations) of one k-step: from the swizzled bf16 rows in shared memory ...
__device__ __forceinline__ void bfrag_smem(uint32_t xs_base, int D, int kk, int lane, uint32_t& b0, uint32_t& b1) {
  const int n = lane & 7, hi = (lane >> 3) & 1;
  ldmatrix_x2(xs_base + (uint32_t)(n * D * 2) + (uint32_t)((((2 * kk + hi) ^ n)) << 4), b0, b1);
}

template <int D, int NIT, typename BLoad>
__device__ __forceinline__ void mma8_n(const Smem8& sm, const Phase1& ph, int row0, unsigned phase_idx, int R, int warp, int lane,
                                       BLoad&& bload) {
  constexpr int KS = (D / 16) / NCW;
  const uint32_t ring_base = ptx::smem_u32(sm.ring);
  const int g = lane >> 2, t4 = lane & 3;
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, khalf = lane >> 4;
  uint32_t a_base[NIT];
  int key[NIT];
  float acc[NIT][4];
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    const int rr = min(lrow, max(ph.nrows[i] - 1, 0));
    int slot = row0 + ph.off[i] + rr;
    if (slot >= R) slot -= R;
    a_base[i] = ring_base + (uint32_t)(slot * (D * 2));
    key[i] = rr & 7;
  }
  ptx::mbar_wait(&sm.full[phase_idx % NBAR], (phase_idx / NBAR) & 1u);
  uint32_t a[2][NIT][4], b[2][NIT][2];
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    ldmatrix_x4(a_base[i] + (uint32_t)(((2 * (warp * KS) + khalf) ^ key[i]) << 4), a[0][i][0], a[0][i][1], a[0][i][2], a[0][i][3]);
    bload(ph.seg[i], warp * KS, b[0][i][0], b[0][i][1]);
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int kk = warp * KS + ks;
    if (ks + 1 < KS) {
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        ldmatrix_x4(a_base[i] + (uint32_t)(((2 * (kk + 1) + khalf) ^ key[i]) << 4), a[(ks + 1) & 1][i][0], a[(ks + 1) & 1][i][1],
                    a[(ks + 1) & 1][i][2], a[(ks + 1) & 1][i][3]);
        bload(ph.seg[i], kk + 1, b[(ks + 1) & 1][i][0], b[(ks + 1) & 1][i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < NIT; ++i)
      mma_bf16_16816(acc[i], a[ks & 1][i][0], a[ks & 1][i][1], a[ks & 1][i][2], a[ks & 1][i][3], b[ks & 1][i][0], b[ks & 1][i][1]);
  }
  __syncwarp();
  // c0, c1: weight row g, sequences 2 t4, 2 t4 + 1 ; c2, c3: weight row g + 8
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    float* rp = sm.red + ((i * NCW + warp) * 16) * B8;
    *(float2*)(rp + g * B8 + 2 * t4) = make_float2(acc[i][0], acc[i][1]);
    *(float2*)(rp + (g + 8) * B8 + 2 * t4) = make_float2(acc[i][2], acc[i][3]);
  }
  ptx::named_bar_sync(1, NCT);
}

template <int D, typename BLoad>
__device__ __forceinline__ void mma8(const Smem8& sm, const Phase1& ph, int row0, unsigned phase_idx, int R, int warp, int lane, BLoad&& bload) {
  switch (ph.nitems) {
    case 1: mma8_n<D, 1>(sm, ph, row0, phase_idx, R, warp, lane, bload); break;
    case 2: mma8_n<D, 2>(sm, ph, row0, phase_idx, R, warp, lane, bload); break;
    case 3: mma8_n<D, 3>(sm, ph, row0, phase_idx, R, warp, lane, bload); break;
    default: mma8_n<D, 4>(sm, ph, row0, phase_idx, R, warp, lane, bload); break;
  }
}

__device__ __forceinline__ float red8_sum(const float* red, int item, int row, int b) {
  float a = 0.f;
#pragma unroll
  for (int w = 0; w < NCW; ++w) a += red[((item * NCW + w) * 16 + row) * B8 + b];
  return a;
}

template <int NPL>
__global__ void __launch_bounds__(NCT, 1) gpt_decode8_kernel(const GptParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int D = NPL * 32, FF = 4 * D, NSEG = FF / D;
  const int G = p.G, L = p.L, H = p.H, V = p.V, R = p.ring_rows, B = p.B;
  const int cta = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wv = (V + 31) / 32;

  Smem8 sm;
  {
    unsigned char* q = smem_raw;
    sm.ring = (__nv_bfloat16*)q;  q += (size_t)R * D * 2;
    sm.xs = (__nv_bfloat16*)q;    q += (size_t)B8 * D * 2;
    sm.red = (float*)q;           q += sizeof(float) * RED8_FLOATS;
    sm.att = (float*)q;           q += sizeof(float) * NCW * PART_STRIDE;
    sm.full = (uint64_t*)q;       q += sizeof(uint64_t) * NBAR;
    sm.pht = (Phase1*)q;          q += 5 * sizeof(Phase1);
    q = (unsigned char*)(((uintptr_t)q + 15) & ~(uintptr_t)15);
    sm.bias_s = (float*)q;        q += sizeof(float) * (size_t)p.bias_cap;
    sm.xres = (float*)q;          q += sizeof(float) * (size_t)B8 * p.ocap;
    sm.seen_s = (unsigned*)q;     q += sizeof(unsigned) * (size_t)wv;
    q = (unsigned char*)(((uintptr_t)q + 15) & ~(uintptr_t)15);
    sm.lnp = (float*)q;
  }
  __shared__ int s_plen[B8], s_tok[B8], s_fin[B8], s_flag;
  if (tid == 0) {
    for (int s = 0; s < NBAR; ++s) ptx::mbar_init(&sm.full[s], 1);
    ptx::fence_mbar_init();
  }
  const int q0 = col_begin(3 * D, cta, G), q1 = col_begin(3 * D, cta + 1, G);
  const int o0 = col_begin(D, cta, G), o1 = col_begin(D, cta + 1, G);
  const int f0 = col_begin(FF, cta, G), f1 = col_begin(FF, cta + 1, G);
  const int h0 = col_begin(V, cta, G), h1 = col_begin(V, cta + 1, G);
  const int nq = q1 - q0, no = o1 - o0, nf = f1 - f0, nh = h1 - h0;
  Sched1 sc;
  sc.L = L; sc.nseg = NSEG; sc.nq = nq; sc.no = no; sc.nf = nf; sc.nh = nh;
  sc.ntq = (nq + TROWS - 1) / TROWS; sc.ntf = (nf + TROWS - 1) / TROWS; sc.nth = (nh + TROWS - 1) / TROWS;
  const int bstride = nq + 2 * no + nf;
  for (int i = tid; i < L * bstride; i += NCT) {
    const int l = i / bstride, j = i % bstride;
    float v;
    if (j < nq) v = p.qkv_b[(size_t)l * 3 * D + q0 + j];
    else if (j < nq + no) v = p.o_b[(size_t)l * D + o0 + (j - nq)];
    else if (j < nq + no + nf) v = p.fc_b[(size_t)l * FF + f0 + (j - nq - no)];
    else v = p.proj_b[(size_t)l * D + o0 + (j - nq - no - nf)];
    sm.bias_s[i] = v;
  }
  for (int i = tid; i < nh; i += NCT) sm.bias_s[L * bstride + i] = p.head_b[h0 + i];
  if (cta < B)
    for (int i = tid; i < wv; i += NCT) sm.seen_s[i] = p.seen[(size_t)cta * wv + i];
  if (tid < B8) {
    s_plen[tid] = (tid < B) ? p.prompt_len[tid] : 0;
    s_tok[tid] = (tid < B) ? p.tok[tid] : 0;
    s_fin[tid] = (tid < B) ? p.finished[tid] : 1;
  }
  auto make_phase = [&](int ncols, int nt, int nseg) {
    Phase1 ph;
    ph.nitems = nt * nseg;
    int r = 0;
#pragma unroll
    for (int i = 0; i < MAXIT; ++i) {
      const int rows = (i < ph.nitems) ? ((nseg > 1) ? ncols : split_rows(ncols, nt, i)) : 0;
      ph.off[i] = r; ph.nrows[i] = rows; ph.seg[i] = (nseg > 1) ? i : 0;
      r += rows;
    }
    ph.total = r;
    return ph;
  };
  if (tid == 0) {
    sm.pht[0] = make_phase(nq, sc.ntq, 1);
    sm.pht[1] = make_phase(no, 1, 1);
    sm.pht[2] = make_phase(nf, sc.ntf, 1);
    sm.pht[3] = make_phase(no, 1, NSEG);
    sm.pht[4] = make_phase(nh, sc.nth, 1);
  }
  __syncthreads();
  const Phase1 &ph_q = sm.pht[0], &ph_o = sm.pht[1], &ph_f = sm.pht[2], &ph_p = sm.pht[3], &ph_h = sm.pht[4];

  // ---- the weight stream (as in gpt_decode1_kernel: the last thread issues what fits, one poll-free point later) ----
  const bool is_prod = (tid == NCT - 1);
  const uint64_t pol = ptx::policy_evict_first();
  const __nv_bfloat16* wbase = p.wstream1 + (size_t)p.stream_off1[cta] * D;
  const int pps = sc.phases_per_step();
  unsigned tix = 0, cons_tile = 0, bar_target = 0;
  int fill = 0, wpos = 0, pstep = 0, pidx = 0, cons_row = 0;
  size_t uoff = 0;
  auto issue_fitting = [&]() {
    if (!is_prod) return;
    while (pstep < p.nsteps && tix - cons_tile < (unsigned)NBAR) {
      const int n = sc.rows(pidx);
      if (fill + n > R) break;
      uint64_t* bar = &sm.full[tix % NBAR];
      ptx::mbar_arrive_expect_tx(bar, (uint32_t)n * D * 2);
      const int n1 = min(n, R - wpos);
      ptx::bulk_g2s(sm.ring + (size_t)wpos * D, wbase + uoff * D, (uint32_t)n1 * D * 2, bar, pol);
      if (n1 < n) ptx::bulk_g2s(sm.ring, wbase + (uoff + n1) * D, (uint32_t)(n - n1) * D * 2, bar, pol);
      wpos += n;
      if (wpos >= R) wpos -= R;
      fill += n;
      uoff += n;
      ++tix;
      if (++pidx == pps) { pidx = 0; uoff = 0; ++pstep; }
    }
  };
  auto advance = [&](const Phase1& ph) {
    const int tot = ph.total;
    cons_row += tot;
    if (cons_row >= R) cons_row -= R;
    ++cons_tile;
    fill -= tot;
    if (tix == cons_tile) issue_fitting();
  };
  auto gsync = [&]() {
    grid_sync(p.barrier, bar_target, G, 1);
    issue_fitting();                 // refills go out while the phase that follows the barrier computes
  };
  auto prefetch_ln = [&](int buf, const float* w, const float* bb) {
    float* dst = sm.lnp + (size_t)buf * 2 * D;
    const int n4 = D / 4;
    for (int i = tid; i < 2 * n4; i += NCT) {
      const int which = i / n4, off = (i % n4) * 4;
      cp_async16(dst + which * D + off, (which ? bb : w) + off);
    }
  };
  const float* lnA = sm.lnp;
  const float* lnB = sm.lnp + 2 * D;
  const int rr = p.round_bf16;
  const uint32_t xs_base = ptx::smem_u32(sm.xs);
  auto b_from_xs = [&](int seg, int kk, uint32_t& b0, uint32_t& b1) { (void)seg; bfrag_smem(xs_base, D, kk, lane, b0, b1); };
  // LayerNorm of row `warp` of xg (one warp per sequence) into the swizzled bf16 rows; second = the head's double LayerNorm
  auto ln_rows = [&](const float* w1, const float* b1, const float* w2, const float* b2) {
    float v[NPL], o[NPL];
    if (warp < B) {
      load_row<NPL>(v, p.xg + (size_t)warp * D, lane);
      ln_row<NPL>(v, o, w1, b1, lane);
      if (w2) { ln_row<NPL>(o, v, w2, b2, lane); }
      else {
#pragma unroll
        for (int j = 0; j < NPL; ++j) v[j] = o[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < NPL; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int k = lane + 32 * j;
      sm.xs[(size_t)warp * D + ((((k >> 3) ^ (warp & 7)) << 3) | (k & 7))] = __float2bfloat16_rn(v[j]);
    }
  };

  issue_fitting();
  prefetch_ln(0, p.ln1_w, p.ln1_b);
  bool alldone = true;
  for (int b = 0; b < B; ++b) alldone &= (s_fin[b] != 0);

  for (int step = 0; step < p.nsteps && !alldone; ++step) {
    const int k = p.step0 + step;
    const int posidx = (k == 0 || p.pos_plain) ? k : k + 1;
    // ---- input rows: every CTA writes its own O-proj column slice of the 8 rows ----
    for (int idx = tid; idx < B8 * no; idx += NCT) {
      const int b = idx / no, c = o0 + idx % no;
      float v = 0.f;
      if (b < B) v = rnd(__ldg(p.mel_emb + (size_t)s_tok[b] * D + c) + __ldg(p.mel_pos + (size_t)posidx * D + c), rr);
      p.xg[(size_t)b * D + c] = v;
      sm.xres[b * p.ocap + (c - o0)] = v;
    }
    gsync();

    for (int l = 0; l < L; ++l) {
      // ---------------- P1: LN1 -> QKV ----------------
      cp_async_wait_all();
      ptx::named_bar_sync(1, NCT);                 // LN1 parameters (buffer A) visible
      ln_rows(lnA, lnA + D, nullptr, nullptr);
      prefetch_ln(1, p.ln2_w + (size_t)l * D, p.ln2_b + (size_t)l * D);
      ptx::named_bar_sync(1, NCT);
      mma8<D>(sm, ph_q, cons_row, cons_tile, R, warp, lane, b_from_xs);
      advance(ph_q);
      for (int idx = tid; idx < nq * B8; idx += NCT) {
        const int cl = idx / B8, b = idx % B8;
        if (b >= B) continue;
        int j = 0;
        while (j + 1 < sc.ntq && cl >= split_begin(nq, sc.ntq, j + 1)) ++j;
        const float v = rnd(red8_sum(sm.red, j, cl - split_begin(nq, sc.ntq, j), b) + sm.bias_s[l * bstride + cl], rr);
        const int c = q0 + cl;
        if (c < D) {
          p.qg[(size_t)b * D + c] = v;
        } else {
          const size_t base = (((size_t)l * p.nseq + b) * p.maxpos + (s_plen[b] + k)) * D;
          const __nv_bfloat16 kvb = __float2bfloat16_rn(v);
          if (c < 2 * D) p.kc[base + (c - D)] = kvb;
          else p.vc[base + (c - 2 * D)] = kvb;
        }
      }
      gsync();

      // ---------------- P2: attention: (sequence, head pair) = CTA, one head per group of four warps ----------------
      if (cta < B * (H / 2)) {
        const int b = cta / (H / 2), h = 2 * (cta % (H / 2)) + (warp >> 2);
        const int w4 = warp & 3;
        const int ctx = s_plen[b] + k + 1;
        const int g4 = lane >> 3, sub = lane & 7;
        const size_t cbase = ((size_t)l * p.nseq + b) * p.maxpos;
        const size_t coff = (size_t)h * HD + sub * 8;
        float qv[8];
        {
          const float* qp = p.qg + (size_t)b * D + h * HD + sub * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) qv[i] = __ldcg(qp + i);
        }
        constexpr int PF = 4, STRIDE = 16;               // 4 warps x 4 key groups
        const int jbase = w4 * 4 + g4;
        uint4 kb[PF], vb[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          kb[u] = make_uint4(0, 0, 0, 0);
          vb[u] = make_uint4(0, 0, 0, 0);
          const int ju = jbase + STRIDE * u;
          if (ju < ctx) {
            kb[u] = __ldcg((const uint4*)(p.kc + (cbase + ju) * D + coff));
            vb[u] = __ldcg((const uint4*)(p.vc + (cbase + ju) * D + coff));
          }
        }
        float m = -INFINITY, lsum = 0.f, ov[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ov[i] = 0.f;
        const int span = ctx - w4 * 4;
        const int niter = span > 0 ? (span + STRIDE - 1) / STRIDE : 0;
        for (int it0 = 0; it0 < niter; it0 += PF) {
#pragma unroll
          for (int u = 0; u < PF; ++u) {
            const int it = it0 + u;
            if (it < niter) {
              const int j = jbase + STRIDE * it;
              const bool valid = j < ctx;
              const uint4 kk = kb[u], vv = vb[u];
              const int jn = j + STRIDE * PF;
              kb[u] = make_uint4(0, 0, 0, 0);
              vb[u] = make_uint4(0, 0, 0, 0);
              if (jn < ctx) {
                kb[u] = __ldcg((const uint4*)(p.kc + (cbase + jn) * D + coff));
                vb[u] = __ldcg((const uint4*)(p.vc + (cbase + jn) * D + coff));
              }
              float s = qv[0] * lo_bf(kk.x) + qv[1] * hi_bf(kk.x) + qv[2] * lo_bf(kk.y) + qv[3] * hi_bf(kk.y) +
                        qv[4] * lo_bf(kk.z) + qv[5] * hi_bf(kk.z) + qv[6] * lo_bf(kk.w) + qv[7] * hi_bf(kk.w);
              s += __shfl_xor_sync(0xffffffffu, s, 1);
              s += __shfl_xor_sync(0xffffffffu, s, 2);
              s += __shfl_xor_sync(0xffffffffu, s, 4);
              if (valid) {
                s *= 0.125f;
                const float mn = fmaxf(m, s);
                const float corr = __expf(m - mn);
                const float pr = __expf(s - mn);
                lsum = lsum * corr + pr;
                const float vf[8] = {lo_bf(vv.x), hi_bf(vv.x), lo_bf(vv.y), hi_bf(vv.y),
                                     lo_bf(vv.z), hi_bf(vv.z), lo_bf(vv.w), hi_bf(vv.w)};
#pragma unroll
                for (int i = 0; i < 8; ++i) ov[i] = ov[i] * corr + pr * vf[i];
                m = mn;
              }
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int xo = 8; xo <= 16; xo <<= 1) {
          const float m2 = __shfl_xor_sync(0xffffffffu, m, xo);
          const float l2 = __shfl_xor_sync(0xffffffffu, lsum, xo);
          const float mn = fmaxf(m, m2);
          const float c1 = (m == -INFINITY) ? 0.f : __expf(m - mn);
          const float c2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
          lsum = lsum * c1 + l2 * c2;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float o2 = __shfl_xor_sync(0xffffffffu, ov[i], xo);
            ov[i] = ov[i] * c1 + o2 * c2;
          }
          m = mn;
        }
        if (lane < 8) {
          float* rw = sm.att + warp * PART_STRIDE;
          if (lane == 0) { rw[0] = m; rw[1] = lsum; }
#pragma unroll
          for (int i = 0; i < 8; ++i) rw[2 + lane * 8 + i] = ov[i];
        }
        ptx::named_bar_sync(1, NCT);
        if (tid < 2 * HD) {
          const int hg = tid >> 6, d = tid & 63;          // head group (warps 4 hg .. 4 hg + 3), dim
          const float* base = sm.att + hg * 4 * PART_STRIDE;
          float mm = -INFINITY;
#pragma unroll
          for (int w = 0; w < 4; ++w) mm = fmaxf(mm, base[w * PART_STRIDE]);
          float lt = 0.f, oa = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float mw = base[w * PART_STRIDE];
            const float c = (mw == -INFINITY) ? 0.f : __expf(mw - mm);
            lt += base[w * PART_STRIDE + 1] * c;
            oa += base[w * PART_STRIDE + 2 + d] * c;
          }
          const int hh = 2 * (cta % (H / 2)) + hg;
          p.part[(size_t)b * D + hh * HD + d] = oa * ((lt > 0.f) ? 1.0f / lt : 0.f);     // `part` holds the [8][D] attention output here
        }
      }
      gsync();

      // ---------------- P3: O-proj + residual (B fragments straight from the attention output in global memory) ----------------
      {
        const int g = lane >> 2, t4 = lane & 3;
        const float* orow = p.part + (size_t)g * D;
        auto b_from_o = [&](int seg, int kk, uint32_t& b0, uint32_t& b1) {
          (void)seg;
          const float2 x0 = __ldcg((const float2*)(orow + kk * 16 + 2 * t4));
          const float2 x1 = __ldcg((const float2*)(orow + kk * 16 + 8 + 2 * t4));
          __nv_bfloat162 h0 = __floats2bfloat162_rn(x0.x, x0.y), h1 = __floats2bfloat162_rn(x1.x, x1.y);
          b0 = *(uint32_t*)&h0;
          b1 = *(uint32_t*)&h1;
        };
        mma8<D>(sm, ph_o, cons_row, cons_tile, R, warp, lane, b_from_o);
        advance(ph_o);
        for (int idx = tid; idx < no * B8; idx += NCT) {
          const int cl = idx / B8, b = idx % B8;
          if (b >= B) continue;
          const float o = rnd(red8_sum(sm.red, 0, cl, b) + sm.bias_s[l * bstride + nq + cl], rr);
          const float xn = sm.xres[b * p.ocap + cl] + o;
          sm.xres[b * p.ocap + cl] = xn;
          p.xg[(size_t)b * D + o0 + cl] = xn;
        }
      }
      if (l + 1 < L) prefetch_ln(0, p.ln1_w + (size_t)(l + 1) * D, p.ln1_b + (size_t)(l + 1) * D);
      else prefetch_ln(0, p.lnf_w, p.lnf_b);
      gsync();

      // ---------------- P4: LN2 -> FC + gelu_new ----------------
      cp_async_wait_all();
      ptx::named_bar_sync(1, NCT);
      ln_rows(lnB, lnB + D, nullptr, nullptr);
      ptx::named_bar_sync(1, NCT);
      mma8<D>(sm, ph_f, cons_row, cons_tile, R, warp, lane, b_from_xs);
      advance(ph_f);
      for (int idx = tid; idx < nf * B8; idx += NCT) {
        const int cl = idx / B8, b = idx % B8;
        if (b >= B) continue;
        int j = 0;
        while (j + 1 < sc.ntf && cl >= split_begin(nf, sc.ntf, j + 1)) ++j;
        const float f = rnd(red8_sum(sm.red, j, cl - split_begin(nf, sc.ntf, j), b) + sm.bias_s[l * bstride + nq + no + cl], rr);
        p.fg[(size_t)b * FF + f0 + cl] = __float2bfloat16_rn(gelu_new(f, rr));
      }
      if (l + 1 == L) prefetch_ln(1, p.fn_w, p.fn_b);
      gsync();

      // ---------------- P5: PROJ + residual (B fragments straight from gelu(fc) in global memory) ----------------
      {
        const int g = lane >> 2, t4 = lane & 3;
        const uint32_t* frow = (const uint32_t*)(p.fg + (size_t)g * FF);
        auto b_from_f = [&](int seg, int kk, uint32_t& b0, uint32_t& b1) {
          const uint32_t* q = frow + seg * (D / 2) + kk * 8 + t4;
          b0 = __ldcg(q);
          b1 = __ldcg(q + 4);
        };
        mma8<D>(sm, ph_p, cons_row, cons_tile, R, warp, lane, b_from_f);
        advance(ph_p);
        for (int idx = tid; idx < no * B8; idx += NCT) {
          const int cl = idx / B8, b = idx % B8;
          if (b >= B) continue;
          float a = 0.f;
#pragma unroll
          for (int s = 0; s < NSEG; ++s) a += red8_sum(sm.red, s, cl, b);
          const float o = rnd(a + sm.bias_s[l * bstride + nq + no + nf + cl], rr);
          const float xn = sm.xres[b * p.ocap + cl] + o;
          sm.xres[b * p.ocap + cl] = xn;
          p.xg[(size_t)b * D + o0 + cl] = xn;
        }
      }
      gsync();
    }

    // ---------------- head: ln_f -> final_norm -> mel_head ----------------
    cp_async_wait_all();
    ptx::named_bar_sync(1, NCT);
    ln_rows(lnA, lnA + D, lnB, lnB + D);
    ptx::named_bar_sync(1, NCT);
    prefetch_ln(0, p.ln1_w, p.ln1_b);
    mma8<D>(sm, ph_h, cons_row, cons_tile, R, warp, lane, b_from_xs);
    advance(ph_h);
    for (int idx = tid; idx < nh * B8; idx += NCT) {
      const int cl = idx / B8, b = idx % B8;
      if (b >= B) continue;
      int j = 0;
      while (j + 1 < sc.nth && cl >= split_begin(nh, sc.nth, j + 1)) ++j;
      const float lg = rnd(red8_sum(sm.red, j, cl - split_begin(nh, sc.nth, j), b) + sm.bias_s[L * bstride + cl], rr);
      p.logits[(size_t)b * V + h0 + cl] = lg;
      if (p.logits_dump) p.logits_dump[((size_t)b * p.max_new + k) * V + h0 + cl] = lg;
    }
    gsync();

    // ---------------- sampling: CTA b decides for sequence b ----------------
    if (cta < B) {
      const int b = cta;
      int tk;
      if (p.do_sample) {
        SampleArgs sa;
        sa.V = V; sa.stop_tok = p.stop_tok; sa.forbid_stop_before = p.forbid_stop_before; sa.top_k = p.top_k;
        sa.seq_base = p.seq_base; sa.rep_penalty = p.rep_penalty; sa.temperature = p.temperature; sa.top_p = p.top_p;
        sa.seed = p.seed;
        tk = sample_block(sa, sm.red, sm.seen_s, p.logits + (size_t)b * V, k, b, tid, lane, warp);
      } else {
        // greedy: RepetitionPenalty -> (forbid stop) -> argmax, lowest index among ties
        float best = -INFINITY;
        int bi = 0x7fffffff;
        const float* lg = p.logits + (size_t)b * V;
        for (int i = tid; i < V; i += NCT) {
          float s = __ldcg(lg + i);
          if ((sm.seen_s[i >> 5] >> (i & 31)) & 1u) s = (s < 0.f) ? s * p.rep_penalty : s / p.rep_penalty;
          if (i == p.stop_tok && k < p.forbid_stop_before) s = -INFINITY;
          if (s > best) { best = s; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float b2 = __shfl_xor_sync(0xffffffffu, best, o);
          const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
          if (b2 > best || (b2 == best && i2 < bi)) { best = b2; bi = i2; }
        }
        if (lane == 0) { sm.att[warp * 2] = best; ((int*)sm.att)[warp * 2 + 1] = bi; }
        ptx::named_bar_sync(1, NCT);
        for (int w = 0; w < NCW; ++w) {
          const float b2 = sm.att[w * 2];
          const int i2 = ((int*)sm.att)[w * 2 + 1];
          if (w == 0 || b2 > best || (b2 == best && i2 < bi)) { best = b2; bi = i2; }
        }
        tk = bi;
      }
      if (tid == 0 && !s_fin[b]) {
        p.codes[(size_t)b * p.max_new + k] = tk;
        p.nout[b] = k + 1;
        int feed = tk;
        int fin = 0;
        if (p.forced) feed = p.forced[(size_t)b * p.max_new + k];
        else if (tk == p.stop_tok) fin = 1;
        if (k + 1 >= p.max_new) fin = 1;
        p.tok[b] = feed;
        if (fin) p.finished[b] = 1;
        sm.seen_s[feed >> 5] |= 1u << (feed & 31);
        p.seen[(size_t)b * wv + (feed >> 5)] |= 1u << (feed & 31);
      }
    }
    gsync();
    if (tid < B8) {
      s_tok[tid] = (tid < B) ? __ldcg(p.tok + tid) : 0;
      s_fin[tid] = (tid < B) ? __ldcg(p.finished + tid) : 1;
    }
    ptx::named_bar_sync(1, NCT);
    alldone = true;
    for (int b = 0; b < B; ++b) alldone &= (s_fin[b] != 0);
    if (alldone && cta == 0 && tid == 0) *p.done = 1;
  }
  // drain: bulk copies issued beyond what was consumed must land before the CTA exits
  if (is_prod)
    for (unsigned n = cons_tile; n < tix; ++n) ptx::mbar_wait(&sm.full[n % NBAR], (n / NBAR) & 1u);
  __syncthreads();
  (void)s_flag;
}
