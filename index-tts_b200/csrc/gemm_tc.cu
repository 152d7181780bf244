// gemm_tc.cu — tcgen05 implicit-GEMM back end of conv_gemm() for sm_100a.
//
//   D[b][m][n] = epi( sum_tap sum_k A[b][m + tap*dil - pad][k] * Wk[n][tap*K + k] )
//
// One CTA computes a 128 x BN output tile (M = time rows, N = output channels):
//   warp 0      TMA producer: per (tap, k-chunk) one 3-D box of the activations [B][T][K]
//               (32 fp32 x 128 rows, SWIZZLE_128B; rows outside [0,T) and columns >= K are
//               zero-filled by TMA = Conv1d zero padding / ragged K for free) and one 2-D box of the
//               K-major weights, into a ring of shared-memory stages (mbarrier full/empty);
//   warp 1      allocates TMEM, issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8 per instruction,
//               4 per stage) from shared-memory descriptors, commits to the stage's empty barrier;
//               the fp32 accumulator tile lives in TMEM (BN columns x 128 lanes);
//   warps 2..5  epilogue: tcgen05.ld the accumulator (each warp its 32-lane quarter), apply
//               bias / activation / layer-scale / residual / accumulate / scale, store with the
//               generic (out_off, ldo, out_valid) mapping that also serves ConvTranspose1d.
// Two operand formats share the kernel (template EB = operand element bytes):
//   EB = 4: fp32 in HBM, tensor maps TFLOAT32 (the TMA unit rounds to tf32 on load), tcgen05.mma kind::tf32, K = 8 per MMA;
//   EB = 2: fp16 in HBM (activations written as fp16 by their producing kernel, weights converted once at init),
//           kind::f16, K = 16 per MMA: the same 10-bit mantissa as tf32 at half the bytes per operand — the tf32 tiles of
//           this kernel are bound by L2 -> shared-memory operand traffic, so halving the bytes is what counts.
// A 128-byte swizzle row holds 32 fp32 or 64 fp16 along K; a stage is always A 16 KB + B BN x 128 B.
#include "ops.h"
#include "ptx.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <map>
#include <mutex>
#include <tuple>
#include <cstdlib>

namespace {

constexpr int BM = 128;
constexpr int BKB = 128; // bytes along K per stage row = one SWIZZLE_128B row (32 tf32 or 64 fp16 elements)

__device__ __forceinline__ float apply_act_tc(float v, int act) {
  switch (act) {
    case ACT_GELU_ERF: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    case ACT_SILU: return v / (1.f + __expf(-v));
    case ACT_MISH: {
      float sp = (v > 20.f) ? v : log1pf(__expf(v));
      return v * tanhf(sp);
    }
    case ACT_GELU_TANH: {
      float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
      return 0.5f * v * (1.f + tanhf(u));
    }
    case ACT_RELU: return v > 0.f ? v : 0.f;
    default: return v;
  }
}

struct TcParams {
  int M, N, K, taps, dil, pad, a_bcast, w_batched;
  int n_kchunks;        // ceil(K / (BKB / EB))
  const float* bias; int biasN; int act;
  const float* res; const float* rowscale; const float* colscale;
  int accum; float scale;
  float* out; long long out_batch_stride, out_off, out_valid; int ldo;
  int stages;
  int epi; __half* out16; const float* aux; int aux_stride;
};

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start address (>>4), LBO (unused for
// swizzled K-major, set to 1), SBO = 1024 B (8 rows x 128 B), version = 1 (sm_100), layout = 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, int EB>
__global__ void __launch_bounds__(192, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: stages x (A 16 KB | B BN*128 B), then barriers
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int BK = BKB / EB;       // elements along K per stage
  constexpr int A_BYTES = BM * BKB;
  constexpr int B_BYTES = BN * BKB;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full = (uint64_t*)(base + (size_t)p.stages * STAGE_BYTES);
  uint64_t* empty = full + p.stages;
  uint64_t* accum_full = empty + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, b = blockIdx.z;
  const int n_iters = p.taps * p.n_kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(accum_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlapped with the previous kernel of the stream (programmatic dependent launch); from here on
  // this grid reads what that kernel wrote
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tensormap(&tmA);
      ptx::prefetch_tensormap(&tmB);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1u;
        ptx::mbar_wait(&empty[s], ph ^ 1u);
        const int tap = it / p.n_kchunks, kc = it % p.n_kchunks;
        unsigned char* sa = base + (size_t)s * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        ptx::mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        ptx::tma_load_3d(sa, &tmA, &full[s], kc * BK, m0 + tap * p.dil - p.pad, p.a_bcast ? 0 : b);
        ptx::tma_load_3d(sb, &tmB, &full[s], tap * p.K + kc * BK, n0, p.w_batched ? b : 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32 (1<<4), A=B format (tf32 = 2, f16 = 0) at bits 7 and 10, K-major both, N>>3, M>>4
      constexpr uint32_t FMT = (EB == 4) ? 2u : 0u;
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1u;
        ptx::mbar_wait(&full[s], ph);
        ptx::tcgen05_fence_after();
        const uint32_t sa = ptx::smem_u32(base + (size_t)s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
        const uint64_t da = make_desc(sa), db = make_desc(sb);
#pragma unroll
        for (int k = 0; k < BKB / 32; ++k) {
          // one MMA consumes 32 bytes of K (8 tf32 / 16 fp16) inside the 128-byte swizzle row: +2 in 16-byte units
          if (EB == 4) ptx::umma_tf32(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
          else ptx::umma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty[s]);
      }
      ptx::umma_commit(accum_full);
    }
  } else {
    // ---------------- epilogue warps (2..5): TMEM lane quarter = warp % 4 ----------------
    // Each warp owns 32 accumulator rows.  A 32x32 chunk is read from TMEM (lane = row), transposed
    // through a private shared-memory tile (the operand ring is dead by now) so that lane = column:
    // every global load/store of the epilogue is then a coalesced 128-byte row segment.
    const int q = warp & 3;
    ptx::mbar_wait(accum_full, 0);
    ptx::tcgen05_fence_after();
    float* tile = (float*)base + (size_t)q * 32 * 33;
    const long long obs = p.out_batch_stride;
    const int biasN = p.biasN ? p.biasN : p.N;
    const int mrow0 = m0 + q * 32;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
      if (p.epi == EPI_NONE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(r[j]);
      }
      __syncwarp();
      const int n = n0 + c0 + lane;
      if (p.epi != EPI_NONE) {
        // fused pair epilogues, thread = accumulator row (the TMEM lane it just read): both columns of every pair are in this
        // thread's registers, no transpose and no shuffle; each thread writes 16 or 32 fp16 values = whole 32-byte sectors.
        // (N is a multiple of 32 for these GEMMs: whole 32-column chunks only.)
        const int row = mrow0 + lane;
        if (row < p.M) {
          const int nb = n0 + c0;                                // first column of this chunk
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + nb + j) : 0.f);
          if (p.epi == EPI_ROPE) {
            const int HD = 64, Hh = p.aux_stride, HW = Hh * HD;
            const int which = nb / HW, hh = (nb % HW) / HD, d0 = nb % HD;      // q | k | v, head, first dim of the chunk
            __half* dst = p.out16 + (size_t)which * ((size_t)gridDim.z * Hh * p.M * HD) + (((size_t)b * Hh + hh) * (size_t)p.M + row) * HD + d0;
            uint32_t o[16];
            if (which < 2) {
              const float4* tab = (const float4*)(p.aux + ((size_t)row * (HD / 2) + (d0 >> 1)) * 2);   // (cos, sin) pairs
              const float sc = (which == 0) ? p.scale : 1.0f;      // 1/8 (x log2 e when the tcgen05 flash kernel consumes q)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 t4 = __ldg(tab + i);
                const float a0 = v[4 * i], a1 = v[4 * i + 1], b0 = v[4 * i + 2], b1 = v[4 * i + 3];
                __half2 h0 = __floats2half2_rn((a0 * t4.x - a1 * t4.y) * sc, (a1 * t4.x + a0 * t4.y) * sc);
                __half2 h1 = __floats2half2_rn((b0 * t4.z - b1 * t4.w) * sc, (b1 * t4.z + b0 * t4.w) * sc);
                o[2 * i] = *(uint32_t*)&h0;
                o[2 * i + 1] = *(uint32_t*)&h1;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                o[i] = *(uint32_t*)&h;
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) ((uint4*)dst)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          } else {
            const int NH = p.N >> 1, j0 = nb >> 1;
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float r0, r1;
              if (p.epi == EPI_SWIGLU) {                          // F.silu(w1 x) * (w3 x)
                r0 = __fdividef(v[4 * i], 1.f + __expf(-v[4 * i])) * v[4 * i + 1];
                r1 = __fdividef(v[4 * i + 2], 1.f + __expf(-v[4 * i + 2])) * v[4 * i + 3];
              } else {                                            // fused_add_tanh_sigmoid_multiply
                const float* ga = p.aux + (size_t)b * p.aux_stride + j0 + 2 * i;
                r0 = tanhf(v[4 * i] + __ldg(ga)) * __fdividef(1.f, 1.f + __expf(-(v[4 * i + 1] + __ldg(ga + NH))));
                r1 = tanhf(v[4 * i + 2] + __ldg(ga + 1)) * __fdividef(1.f, 1.f + __expf(-(v[4 * i + 3] + __ldg(ga + NH + 1))));
              }
              __half2 h = __floats2half2_rn(r0, r1);
              o[i] = *(uint32_t*)&h;
            }
            uint4* dst = (uint4*)(p.out16 + ((size_t)b * p.M + row) * NH + j0);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
      } else if (n < p.N) {
        const float bv = p.bias ? __ldg(p.bias + (n % biasN)) : 0.f;
        const float cs = (p.colscale ? __ldg(p.colscale + n) : 1.f);
        const long long flat0 = p.out_off + (long long)mrow0 * p.ldo + n;     // row rr adds rr*ldo
        const int rows = min(32, p.M - mrow0);
        // fast path: the 32-row x 32-col chunk lies wholly inside the valid output range
        const bool interior = rows == 32 && (p.out_off + (long long)mrow0 * p.ldo + n0 + c0) >= 0 &&
                              (p.out_off + (long long)(mrow0 + 31) * p.ldo + n0 + c0 + 31) < p.out_valid;
        float* op = p.out + (long long)b * obs + flat0;
        const float* rp = p.res ? p.res + (long long)b * obs + flat0 : nullptr;
        if (interior && p.act == ACT_NONE && !p.rowscale) {
          const float sc = p.scale;
          if (!rp && !p.accum) {
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) op[(long long)rr * p.ldo] = (tile[rr * 33 + lane] + bv) * cs * sc;
          } else {
            // res may alias out (in-place residual): fetch the whole 32-row column into registers first so the
            // 32 L2 round trips overlap instead of serialising behind the stores
            float addv[32];
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
              float a = 0.f;
              if (rp) a = rp[(long long)rr * p.ldo];
              if (p.accum) a += op[(long long)rr * p.ldo];
              addv[rr] = a;
            }
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              op[(long long)rr * p.ldo] = ((tile[rr * 33 + lane] + bv) * cs + addv[rr]) * sc;
          }
        } else {
          for (int rr = 0; rr < rows; ++rr) {
            const long long flat = flat0 + (long long)rr * p.ldo;
            if (flat < 0 || flat >= p.out_valid) continue;
            float v = tile[rr * 33 + lane] + bv;
            v = apply_act_tc(v, p.act) * cs;
            if (p.rowscale) v *= __ldg(p.rowscale + (long long)b * p.M + mrow0 + rr);
            if (rp) v += rp[(long long)rr * p.ldo];
            if (p.accum) v += op[(long long)rr * p.ldo];
            op[(long long)rr * p.ldo] = v * p.scale;
          }
        }
      }
      __syncwarp();
    }
    ptx::tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

PFN_cuTensorMapEncodeTiled get_encode() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess || !f)
      throw IdxError(IDX_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    fn = (PFN_cuTensorMapEncodeTiled)f;
  }
  return fn;
}

using MapKey = std::tuple<const void*, long long, long long, long long, long long, int, int>;
std::map<MapKey, CUtensorMap>& map_cache() {
  static std::map<MapKey, CUtensorMap> c;
  return c;
}

CUtensorMap make_map(const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, bool half = false) {
  static const bool plain_f32 = getenv("IDX_TMA_F32") != nullptr;
  MapKey key(ptr, (long long)dims[0], (long long)dims[1], rank > 2 ? (long long)dims[2] : 0,
             (long long)strides_bytes[0] ^ ((rank > 2 ? (long long)strides_bytes[1] : 0) << 20), (int)box[1] | ((int)box[0] << 12), rank | (half ? 16 : 0));
  // process-wide cache (keys carry the unique UVA address): engines may be driven from different threads
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  auto& cache = map_cache();
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap m;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(&m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : (plain_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32),
                            (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw IdxError(IDX_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  if (cache.size() > 20000) cache.clear();
  cache.emplace(key, m);
  return m;
}

template <int BN, int EB>
void launch_bn(idx_engine* e, const ConvGemm& g, const CUtensorMap& tmA, const CUtensorMap& tmB) {
  constexpr int BK = BKB / EB;
  TcParams p;
  p.M = g.M; p.N = g.N; p.K = g.K; p.taps = g.taps; p.dil = g.dil; p.pad = g.pad; p.a_bcast = g.a_bcast;
  p.w_batched = g.w_batch_stride != 0;
  p.n_kchunks = (g.K + BK - 1) / BK;
  p.bias = g.bias; p.biasN = g.biasN; p.act = g.act;
  p.res = g.res; p.rowscale = g.rowscale; p.colscale = g.colscale;
  p.accum = g.accum; p.scale = g.scale; p.out = g.out;
  p.ldo = g.ldo ? g.ldo : g.N;
  p.out_batch_stride = g.out_batch_stride ? g.out_batch_stride : (long long)g.M * g.N;
  p.out_off = g.out_off;
  p.out_valid = g.out_valid ? g.out_valid : (long long)g.M * p.ldo;
  p.epi = g.epi; p.out16 = g.out16; p.aux = g.aux; p.aux_stride = g.aux_stride;
  if (g.epi != EPI_NONE)
    IDX_CHECK(g.out16 && g.N % 32 == 0 && !g.res && !g.accum && g.act == ACT_NONE && (g.epi == EPI_SWIGLU || g.aux), IDX_ERR_ARG,
              "conv_gemm: bad fused-epilogue arguments");
  constexpr int STAGE = BM * BKB + BN * BKB;
  const int n_iters = p.taps * p.n_kchunks;
  static const int occ = getenv("IDX_GEMM_OCC") ? atoi(getenv("IDX_GEMM_OCC")) : 2;
  int stages = ((occ == 1 ? 208 : 104) * 1024) / STAGE;   // two CTAs per SM: one's epilogue overlaps the other's mainloop
  if (stages > 8) stages = 8;
  if (stages > n_iters) stages = n_iters < 2 ? 2 : n_iters;
  p.stages = stages;
  const size_t smem = (size_t)stages * STAGE + 1024 + (2 * stages + 1) * 8 + 16;
  const unsigned bit = (BN == 32 ? 1u : (BN == 64 ? 2u : 4u)) << (EB == 2 ? 8 : 0);
  if (!(e->attr_done & bit)) {     // per engine = per device: function attributes live in the device's context
    IDX_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    e->attr_done |= bit;
  }
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.B);
  launch_pdl(e, gemm_tc_kernel<BN, EB>, grid, dim3(192), smem, tmA, tmB, p);
  e->launches++;
}


// ================================================================================================================
// tcgen05 flash attention for the DiT (full attention, head dim 64, fp16 operands, fp32 softmax and accumulation).
//   gpt_fast/model.py:293-303 (F.scaled_dot_product_attention over all T keys; q arrives pre-scaled by 1/8 and
//   rotated — EPI_ROPE above).
// One CTA = 128 queries of one (batch, head).  TMEM: S [128 lanes x 128 cols fp32] | O [x 64] | P [x 64: 128 fp16 per
// row packed two per column] = 256 columns, so two CTAs share an SM and fill each other's bubbles.
//   warp 0      TMA: Q tile once, then K_j / V_j tiles of 128 keys into a 2-stage ring (SWIZZLE_128B rows of 64 fp16);
//   warp 1      one thread issues S = Q K_j^T (4 x tcgen05.mma M128 N128 K16, both operands K-major in shared memory) and,
//               once the softmax warps have stored P_j, O (+)= P_j V_j (8 x M128 N64 K16: A = P from TENSOR MEMORY,
//               B = V straight from its [key][dim] rows = MN-major operand);
//   warps 2..5  thread = query row = TMEM lane: online softmax in the log2 domain (two passes over S in TMEM: row max,
//               then exp2 / row sum / fp16 pack -> tcgen05.st P), rescale of O when the row max moved, final O / l.
// The legacy mma.sync flash kernel (nn_ops.cu) ran at the mma.sync ceiling of this part (~170 TFLOP/s); here the exp
// throughput (MUFU) is the bound, as in every Blackwell attention kernel.
constexpr int FA_Q = 128, FA_K = 128, FA_D = 64;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
struct FaParams { int T, H; float* out; __half* out16; };

__global__ void __launch_bounds__(192, 2)
fa5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
           const FaParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int TILE = FA_K * FA_D * 2;                  // 16 KB: 128 rows x 128 bytes
  unsigned char* sQ = base;
  unsigned char* sK = base + TILE;                       // [2][TILE]
  unsigned char* sV = base + 3 * TILE;                   // [2][TILE]
  uint64_t* bars = (uint64_t*)(base + 5 * TILE);
  uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = bars + 3, *s_full = bars + 5, *p_full = bars + 6, *o_full = bars + 7;
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, q0 = blockIdx.x * FA_Q;
  const int T = p.T, ntiles = (T + FA_K - 1) / FA_K;
  if (threadIdx.x == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(o_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 256);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t tS = tm, tO = tm + 128, tP = tm + 192;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tensormap(&tmQ); ptx::prefetch_tensormap(&tmK); ptx::prefetch_tensormap(&tmV);
      ptx::mbar_arrive_expect_tx(q_full, TILE);
      ptx::tma_load_3d(sQ, &tmQ, q_full, 0, q0, bh);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1;
        ptx::mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&kv_full[s], 2 * TILE);
        ptx::tma_load_3d(sK + s * TILE, &tmK, &kv_full[s], 0, j * FA_K, bh);
        ptx::tma_load_3d(sV + s * TILE, &tmV, &kv_full[s], 0, j * FA_K, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = f32, A = B = f16; S: N = 128 keys, both K-major; PV: N = 64 dims, B (V) MN-major (bit 16)
      const uint32_t idesc_s = (1u << 4) | ((uint32_t)(FA_K >> 3) << 17) | ((uint32_t)(FA_Q >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 16) | ((uint32_t)(FA_D >> 3) << 17) | ((uint32_t)(FA_Q >> 4) << 24);
      ptx::mbar_wait(q_full, 0);
      const uint64_t dq = make_desc(ptx::smem_u32(sQ));
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1;
        ptx::mbar_wait(&kv_full[s], (j >> 1) & 1u);
        ptx::tcgen05_fence_after();
        const uint64_t dk = make_desc(ptx::smem_u32(sK + s * TILE));
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) ptx::umma_f16(tS, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(s_full);                       // tracks every MMA issued so far: S_j ready AND O += P_{j-1} V_{j-1} done
        ptx::mbar_wait(p_full, j & 1u);                 // P_j is in tensor memory, O has been rescaled
        ptx::tcgen05_fence_after();
        const uint64_t dv = make_desc(ptx::smem_u32(sV + s * TILE));
#pragma unroll
        for (int k = 0; k < FA_K / 16; ++k)             // 16 keys per MMA: 8 packed P columns, 16 V rows = 2048 bytes
          ptx::umma_f16_ts(tO, tP + (uint32_t)(8 * k), dv + (uint64_t)(128 * k), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        ptx::umma_commit(&kv_empty[s]);                 // K_j / V_j consumed
      }
      ptx::umma_commit(o_full);
    }
  } else {
    const int qd = warp & 3;                            // TMEM lane quarter of this warp
    const int row = qd * 32 + lane;                     // query row inside the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float m = -INFINITY, l = 0.f;                       // running row max (log2 domain) and row sum
    for (int j = 0; j < ntiles; ++j) {
      ptx::mbar_wait(s_full, j & 1u);
      ptx::tcgen05_fence_after();
      const int kbase = j * FA_K;
      // the whole S row of this thread (128 keys) in registers: four tensor-memory loads in flight, one wait
      uint32_t r[FA_K];
#pragma unroll
      for (int c = 0; c < FA_K; c += 32) ptx::tmem_ld_32x32b_x32(tS + lane_off + (uint32_t)c, r + c);
      ptx::tmem_ld_wait();
      // (ncu of the first version: 2200 warp instructions per tile per softmax warp — per element a masked max, an FMA, the
      // accurate exp2f sequence, two adds — made the kernel issue-bound at the speed of the mma.sync one.  Now the scores
      // arrive in the log2 domain (q is pre-scaled by log2(e)/8 in EPI_ROPE), exp2 is the bare MUFU, and only the last
      // key tile pays for masking.)
      const int nvalid = T - kbase;                     // keys beyond T (zero-filled rows of the last tile) are masked
      if (nvalid < FA_K) {
#pragma unroll
        for (int i = 0; i < FA_K; ++i)
          if (i >= nvalid) r[i] = 0xff800000u;          // -inf
      }
      float mx = __uint_as_float(r[0]);
#pragma unroll
      for (int i = 1; i < FA_K; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
      const float mn = fmaxf(m, mx);
      const float corr = ex2_approx(m - mn);            // 0 at the first tile (m = -inf)
      float rs0 = 0.f, rs1 = 0.f;
      // p = 2^(s - mn), packed to fp16 pairs in place: r[i] <- (p[2i], p[2i+1])
#pragma unroll
      for (int i = 0; i < FA_K / 2; ++i) {
        const float p0 = ex2_approx(__uint_as_float(r[2 * i]) - mn);
        const float p1 = ex2_approx(__uint_as_float(r[2 * i + 1]) - mn);
        rs0 += p0;
        rs1 += p1;
        __half2 h = __floats2half2_rn(p0, p1);
        r[i] = *(uint32_t*)&h;
      }
      const float rs = rs0 + rs1;
      ptx::tmem_st_32x32b_x32(tP + lane_off, r);
      ptx::tmem_st_32x32b_x32(tP + lane_off + 32u, r + 32);
      l = l * corr + rs;
      if (j > 0 && __any_sync(0xffffffffu, corr != 1.f)) {     // warp-uniform: tcgen05.ld / st are warp-collective
        // the row max moved: O (complete up to tile j-1: s_full tracks that MMA too) is rescaled in tensor memory
#pragma unroll 1
        for (int c = 0; c < FA_D; c += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(tO + lane_off + (uint32_t)c, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
          ptx::tmem_st_32x32b_x32(tO + lane_off + (uint32_t)c, r);
        }
      }
      m = mn;
      ptx::tmem_st_wait();
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive(p_full);
    }
    // epilogue: O / l -> fp16 (operand of the output projection) and / or fp32, [B][T][H*64]
    ptx::mbar_wait(o_full, 0);
    ptx::tcgen05_fence_after();
    const int t = q0 + row;
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const int b = bh / p.H, h = bh % p.H;
    const size_t o = ((size_t)b * T + t) * (size_t)p.H * FA_D + (size_t)h * FA_D;
#pragma unroll 1
    for (int c = 0; c < FA_D; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(tO + lane_off + (uint32_t)c, r);
      ptx::tmem_ld_wait();
      if (t < T) {
        if (p.out16) {
          uint32_t hh[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            __half2 h2 = __floats2half2_rn(__uint_as_float(r[2 * i]) * inv, __uint_as_float(r[2 * i + 1]) * inv);
            hh[i] = *(uint32_t*)&h2;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) ((uint4*)(p.out16 + o + c))[i] = make_uint4(hh[4 * i], hh[4 * i + 1], hh[4 * i + 2], hh[4 * i + 3]);
        }
        if (p.out) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            ((float4*)(p.out + o + c))[i] = make_float4(__uint_as_float(r[4 * i]) * inv, __uint_as_float(r[4 * i + 1]) * inv,
                                                        __uint_as_float(r[4 * i + 2]) * inv, __uint_as_float(r[4 * i + 3]) * inv);
        }
      }
    }
    ptx::tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tm, 256);
  }
}

}  // namespace

// tcgen05 flash attention on the rotated / split fp16 tensors Qr | Kr | Vb [B*H][T][64] (see fa5_kernel)
void flash_attention_tc5(idx_engine* e, const __half* Qr, const __half* Kr, const __half* Vb, float* out, __half* out16,
                         int B, int T, int H) {
  const int BH = B * H;
  cuuint64_t dims[3] = {(cuuint64_t)FA_D, (cuuint64_t)T, (cuuint64_t)BH};
  cuuint64_t str[2] = {(cuuint64_t)FA_D * 2, (cuuint64_t)T * FA_D * 2};
  cuuint32_t box[3] = {FA_D, FA_K, 1};
  CUtensorMap tq = make_map(Qr, 3, dims, str, box, true), tk = make_map(Kr, 3, dims, str, box, true), tv = make_map(Vb, 3, dims, str, box, true);
  FaParams p;
  p.T = T; p.H = H; p.out = out; p.out16 = out16;
  const size_t smem = 5 * (size_t)(FA_K * FA_D * 2) + 1024 + 128;
  if (!(e->attr_done & (1u << 20))) {
    IDX_CUDA(cudaFuncSetAttribute(fa5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    e->attr_done |= 1u << 20;
  }
  launch_pdl(e, fa5_kernel, dim3((T + FA_Q - 1) / FA_Q, BH), dim3(192), smem, tq, tk, tv, p);
  e->launches++;
}

bool gemm_tc_supported(const ConvGemm& g) {
  static const bool off = getenv("IDX_NO_TC") != nullptr;
  const bool half = g.A16 && g.Wk16;
  if (off || g.reflect || (!g.Wk && !half)) return false;
  const int al = half ? 8 : 4;                                // 16-byte global strides: 8 fp16 / 4 fp32 elements
  if (g.K % al != 0) return false;
  const int lda = g.lda ? g.lda : g.K;
  if (lda % al != 0) return false;
  if (half ? (((uintptr_t)g.A16 & 15) || ((uintptr_t)g.Wk16 & 15)) : (((uintptr_t)g.A & 15) || ((uintptr_t)g.Wk & 15))) return false;
  const long long abs_ = g.a_batch_stride ? g.a_batch_stride : (long long)g.Tin * lda;
  if (abs_ % al != 0) return false;
  if ((g.ldw && g.ldw % al) || (g.w_batch_stride % al)) return false;
  if (!half && (long long)g.M * g.N * g.K * g.taps < (1 << 18)) return false;   // tiny problems: SIMT (fp16 operands have no SIMT twin)
  return true;
}

void gemm_tc_launch(idx_engine* e, const ConvGemm& g) {
  const bool half = g.A16 && g.Wk16;
  const int EBh = half ? 2 : 4;
  const int lda = g.lda ? g.lda : g.K;
  const long long abs_ = g.a_bcast ? (long long)g.Tin * lda : (g.a_batch_stride ? g.a_batch_stride : (long long)g.Tin * lda);
  // A: [B][Tin][K], dims (K, Tin, B)
  cuuint64_t adims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.Tin, (cuuint64_t)(g.a_bcast ? 1 : g.B)};
  cuuint64_t astr[2] = {(cuuint64_t)lda * EBh, (cuuint64_t)abs_ * EBh};
  cuuint32_t abox[3] = {(cuuint32_t)(BKB / EBh), BM, 1};
  CUtensorMap tmA = make_map(half ? (const void*)g.A16 : (const void*)g.A, 3, adims, astr, abox, half);
  // B: Wk [nb][N][taps*K], dims (taps*K, N, nb)   (nb = 1 for shared weights)
  const int ldw = g.ldw ? g.ldw : g.taps * g.K;
  const long long wbs = g.w_batch_stride ? g.w_batch_stride : (long long)g.N * ldw;
  cuuint64_t bdims[3] = {(cuuint64_t)g.taps * g.K, (cuuint64_t)g.N, (cuuint64_t)(g.w_batch_stride ? g.B : 1)};
  cuuint64_t bstr[2] = {(cuuint64_t)ldw * EBh, (cuuint64_t)wbs * EBh};
  int BN = g.N <= 32 ? 32 : (g.N <= 64 ? 64 : 128);
  {
    // fewer 128-wide tiles than SMs: halve the tile so every SM gets work (2 CTAs/SM are resident anyway)
    static const int bn64 = getenv("IDX_GEMM_BN64") ? atoi(getenv("IDX_GEMM_BN64")) : 1;
    const long long tiles128 = (long long)((g.N + 127) / 128) * ((g.M + BM - 1) / BM) * g.B;
    if (bn64 && BN == 128 && tiles128 < 148) BN = 64;
  }
  cuuint32_t bbox[3] = {(cuuint32_t)(BKB / EBh), (cuuint32_t)BN, 1};
  CUtensorMap tmB = make_map(half ? (const void*)g.Wk16 : (const void*)g.Wk, 3, bdims, bstr, bbox, half);
  if (half) {
    if (BN == 32) launch_bn<32, 2>(e, g, tmA, tmB);
    else if (BN == 64) launch_bn<64, 2>(e, g, tmA, tmB);
    else launch_bn<128, 2>(e, g, tmA, tmB);
  } else {
    if (BN == 32) launch_bn<32, 4>(e, g, tmA, tmB);
    else if (BN == 64) launch_bn<64, 4>(e, g, tmA, tmB);
    else launch_bn<128, 4>(e, g, tmA, tmB);
  }
}
