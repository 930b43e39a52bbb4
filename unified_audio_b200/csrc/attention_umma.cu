// Non-causal multi-head self-attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), head_dim 64 / 128, with the library's
// two precision policies: single-pass fp16 operands, or the fp16 hi + lo split (3 tensor-core passes, fp32-grade) for both
// contractions.  Callers: the LSTM-transformers of the codecs (HCodec-2.0/vq/encoder_modules/transformer.py:134-215), the SSL front
// ends, and the 96 mimi transformer layers of H-Codec-1.5 (HCodec-1.5/adaptive/model_blocks/mimi/transformer.py:377-424), where the
// fp32 SIMT kernel of attention.cu was 54 % of the step (profiles/r02_h15_launches.md).
//
// Two launches:
//  1. fa5_prep_kernel: qkv fp32 [B*L, 3*H*D] (the in_proj GEMM's output) -> RoPE (rotate-half tables) + 1/sqrt(D) on q -> fp16 hi (+ lo)
//     planes in the operand layouts the MMAs want, all K-major:  Q [plane*B*H + bh][L][D],  K likewise,  V TRANSPOSED [..][D][Lp]
//     (P.V contracts over keys, so V is the B operand [N = D rows] x [K = keys]).
//  2. fa5_kernel: one CTA = 128 queries of one (batch, head).  warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, warps 2-5 =
//     softmax (thread = query row = TMEM lane).  Per 64-key tile j:
//        S_j   = Q K_j^T            UMMA M128 N64  K=D   -> TMEM columns [0, 64)          (3 chains when split: hh + lh + hl)
//        P_j   = exp2(S_j - m)      softmax threads: tcgen05.ld, online max / sum, fp16 hi (+ lo) written to shared memory in the
//                                   128-byte-swizzled K-major tile layout TMA would have produced (A operand of the next MMA)
//        O_j   = P_j V_j            UMMA M128 N=D  K=64  -> TMEM columns [64, 64 + D), NOT accumulated across tiles:
//        o     = o * corr + O_j     in the softmax threads' registers (fp32) - no in-TMEM rescale pass when the running max moves.
//     K and V each live in ONE shared-memory slot: K is only read by the S phase and V only by the P.V phase, so the TMA refill of
//     one overlaps the other's phase (k_empty / v_empty are signalled by tcgen05.commit).  S_{j+1} is issued right after P.V_j, so it
//     runs under the softmax threads' O_j accumulation.  Overlap across query tiles comes from 2-4 resident CTAs per SM (D = 64).
#include <atomic>
#include <cuda.h>

#include "common.cuh"
#include "quark_b200.h"

namespace qb {
extern std::atomic<long long> g_launches;

constexpr int F5_BQ = 128, F5_BK = 64, F5_THREADS = 192;

// ---------------------------------------------------------------------------------------------- prep
// grid (ceil(L / 32), H, B), 256 threads.  q16 / k16: [(plane * BH + bh) * L + t] * D + d;  vT: [(plane * BH + bh) * D + d] * Lp + t.
template <int D>
__global__ void __launch_bounds__(256)
fa5_prep_kernel(const float* __restrict__ qkv, int L, int Lp, int H, const float* __restrict__ rcos, const float* __restrict__ rsin,
                float scale, __half* __restrict__ q16, __half* __restrict__ k16, __half* __restrict__ vT, int planes) {
  constexpr int HD = D / 2;
  __shared__ float vs[32][D + 1];
  const int t0 = blockIdx.x * 32, h = blockIdx.y, b = blockIdx.z;
  const long long BH = (long long)gridDim.z * H, bh = (long long)b * H + h;
  const long long pitch = 3LL * H * D;
  const float* base = qkv + (long long)b * L * pitch;
  for (int e = threadIdx.x; e < 32 * HD; e += 256) {
    const int tt = e / HD, d = e - tt * HD, t = t0 + tt;
    if (t >= L) continue;
    const float* row = base + (long long)t * pitch;
    const float c1 = rcos[(long long)t * D + d], s1 = rsin[(long long)t * D + d];
    const float c2 = rcos[(long long)t * D + d + HD], s2 = rsin[(long long)t * D + d + HD];
    const float q1 = row[h * D + d], q2 = row[h * D + d + HD];
    const float k1 = row[(H + h) * D + d], k2 = row[(H + h) * D + d + HD];
    const float qa = (q1 * c1 - q2 * s1) * scale, qb_ = (q2 * c2 + q1 * s2) * scale;
    const float ka = k1 * c1 - k2 * s1, kb = k2 * c2 + k1 * s2;
    const long long o = (bh * L + t) * D + d, po = BH * L * D;
    __half hh, ll;
    split_f16(qa, hh, ll); q16[o] = hh; if (planes == 2) q16[po + o] = ll;
    split_f16(qb_, hh, ll); q16[o + HD] = hh; if (planes == 2) q16[po + o + HD] = ll;
    split_f16(ka, hh, ll); k16[o] = hh; if (planes == 2) k16[po + o] = ll;
    split_f16(kb, hh, ll); k16[o + HD] = hh; if (planes == 2) k16[po + o + HD] = ll;
  }
  for (int e = threadIdx.x; e < 32 * D; e += 256) {
    const int tt = e / D, d = e - tt * D, t = t0 + tt;
    vs[tt][d] = t < L ? base[(long long)t * pitch + (2 * H + h) * D + d] : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * D; e += 256) {
    const int d = e >> 5, tt = e & 31, t = t0 + tt;
    if (t >= L) continue;
    __half hh, ll;
    split_f16(vs[tt][d], hh, ll);
    const long long o = (bh * D + d) * Lp + t;
    vT[o] = hh;
    if (planes == 2) vT[BH * D * Lp + o] = ll;
  }
}

// ---------------------------------------------------------------------------------------------- attention
template <int D, bool SPLIT>
struct F5Cfg {
  static constexpr int NPL = SPLIT ? 2 : 1, KBQ = D / 64;
  static constexpr uint32_t Q_KB = F5_BQ * 128, K_KB = F5_BK * 128;          // bytes of one 64-wide K-block of Q / of K
  static constexpr uint32_t Q_PLANE = KBQ * Q_KB, K_PLANE = KBQ * K_KB, V_PLANE = D * 128, P_PLANE = F5_BQ * 128;
  static constexpr uint32_t OFF_Q = 0, OFF_K = OFF_Q + NPL * Q_PLANE, OFF_V = OFF_K + NPL * K_PLANE, OFF_P = OFF_V + NPL * V_PLANE;
  static constexpr uint32_t OFF_BAR = OFF_P + NPL * P_PLANE;
  static constexpr uint32_t SMEM = OFF_BAR + 128 + 1024;                      // + barriers + alignment slack
  static constexpr uint32_t TCOLS = D == 64 ? 128 : 256;                      // S: 64 columns, O tile: D columns
  static constexpr int MIN_CTAS = D == 64 ? 2 : 1;
};

template <int D, bool SPLIT>
__global__ void __launch_bounds__(F5_THREADS, F5Cfg<D, SPLIT>::MIN_CTAS)
fa5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
           int L, int H, int BH, __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  using C = F5Cfg<D, SPLIT>;
  constexpr int NPL = C::NPL, KBQ = C::KBQ;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + C::OFF_BAR);
  uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 2, *v_full = bars + 3, *v_empty = bars + 4, *s_full = bars + 5,
           *p_full = bars + 6, *o_full = bars + 7;
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * F5_BQ, h = blockIdx.y, b = blockIdx.z;
  const int bh = b * H + h;
  const int n_tiles = (L + F5_BK - 1) / F5_BK;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, C::TCOLS); tmem_relinquish(); }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 64;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, NPL * C::Q_PLANE);
      for (int pl = 0; pl < NPL; ++pl)
        for (int kb = 0; kb < KBQ; ++kb)
          tma_load_3d(smem + C::OFF_Q + pl * C::Q_PLANE + kb * C::Q_KB, &tmQ, q_full, kb * 64, q0, pl * BH + bh);
      for (int j = 0; j < n_tiles; ++j) {
        const int k0 = j * F5_BK;
        if (j > 0) mbar_wait(k_empty, (j - 1) & 1);
        mbar_arrive_expect_tx(k_full, NPL * C::K_PLANE);
        for (int pl = 0; pl < NPL; ++pl)
          for (int kb = 0; kb < KBQ; ++kb)
            tma_load_3d(smem + C::OFF_K + pl * C::K_PLANE + kb * C::K_KB, &tmK, k_full, kb * 64, k0, pl * BH + bh);
        if (j > 0) mbar_wait(v_empty, (j - 1) & 1);
        mbar_arrive_expect_tx(v_full, NPL * C::V_PLANE);
        for (int pl = 0; pl < NPL; ++pl) tma_load_3d(smem + C::OFF_V + pl * C::V_PLANE, &tmV, v_full, k0, 0, pl * BH + bh);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(F5_BQ, F5_BK), idesc_o = make_idesc_f16(F5_BQ, D);
      const uint32_t sQ = smem_u32(smem + C::OFF_Q), sK = smem_u32(smem + C::OFF_K), sV = smem_u32(smem + C::OFF_V),
                     sP = smem_u32(smem + C::OFF_P);
      auto issue_s = [&](int j) {
        mbar_wait(k_full, j & 1);
        tc_fence_after();
        uint32_t acc = 0;
#pragma unroll
        for (int term = 0; term < (SPLIT ? 3 : 1); ++term) {
          const uint32_t qa = sQ + (term == 1 ? C::Q_PLANE : 0), ka = sK + (term == 2 ? C::K_PLANE : 0);
#pragma unroll
          for (int kb = 0; kb < KBQ; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(tmem_S, make_sw128_kmajor_desc(qa + kb * C::Q_KB + k * 32), make_sw128_kmajor_desc(ka + kb * C::K_KB + k * 32),
                       idesc_s, acc);
              acc = 1;
            }
        }
        umma_commit(k_empty);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(v_full, j & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        uint32_t acc = 0;
#pragma unroll
        for (int term = 0; term < (SPLIT ? 3 : 1); ++term) {
          const uint32_t pa = sP + (term == 1 ? C::P_PLANE : 0), va = sV + (term == 2 ? C::V_PLANE : 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(tmem_O, make_sw128_kmajor_desc(pa + k * 32), make_sw128_kmajor_desc(va + k * 32), idesc_o, acc);
            acc = 1;
          }
        }
        umma_commit(v_empty);
        umma_commit(o_full);
        if (j + 1 < n_tiles) issue_s(j + 1);     // p_full(j) also says the softmax threads have drained S_j from TMEM
      }
    }
  } else {
    // ===================== softmax / accumulate (thread = query row) =====================
    const int quad = warp & 3, r = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = smem + C::OFF_P + (r >> 3) * 1024 + (r & 7) * 128;
    for (int j = 0; j < n_tiles; ++j) {
      const int k0 = j * F5_BK;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row maximum of the tile
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < F5_BK; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_S + lane_addr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (k0 + c0 + c < L) mx = fmaxf(mx, __uint_as_float(v[c]));
      }
      const float m_new = fmaxf(m, mx);
      const float corr = exp2f((m - m_new) * LOG2E);          // m = -inf on the first tile -> 0
      const float mb = m_new * LOG2E;
      float rs = 0.f;
      // pass 2: P = exp2(S - m) as fp16 hi (+ lo) into the swizzled K-major A tile: 16-byte chunk cc of row r sits at cc ^ (r % 8)
#pragma unroll
      for (int c0 = 0; c0 < F5_BK; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_S + lane_addr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = g * 8 + e * 2;
            const float p0 = (k0 + c0 + c < L) ? exp2f(fmaf(__uint_as_float(v[c]), LOG2E, -mb)) : 0.f;
            const float p1 = (k0 + c0 + c + 1 < L) ? exp2f(fmaf(__uint_as_float(v[c + 1]), LOG2E, -mb)) : 0.f;
            rs += p0 + p1;
            const __half2 hh = __floats2half2_rn(p0, p1);
            hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
            if (SPLIT) {
              const float2 back = __half22float2(hh);
              const __half2 ll = __floats2half2_rn(p0 - back.x, p1 - back.y);
              lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
            }
          }
          const int cc = (c0 >> 3) + g;
          uint8_t* dst = prow + ((cc ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(dst + C::P_PLANE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      l = l * corr + rs;
      m = m_new;
      tc_fence_before();
      fence_proxy_async();                       // generic-proxy stores of P -> visible to the tensor core's async proxy
      mbar_arrive(p_full);
      // O_j (unscaled by the running max of later tiles): o = o * corr + O_j
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_O + lane_addr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c0 + c] = fmaf(o[c0 + c], corr, __uint_as_float(v[c]));
      }
      tc_fence_before();
    }
    const int tq = q0 + r;
    if (tq < L) {
      const float inv = 1.f / l;
      const long long ob = ((long long)b * L + tq) * (long long)(H * D) + h * D;
#pragma unroll
      for (int d0 = 0; d0 < D; d0 += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = o[d0 + 2 * e] * inv, c = o[d0 + 2 * e + 1] * inv;
          const __half2 hh = __floats2half2_rn(a, c);
          hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
          const float2 back = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(a - back.x, c - back.y);
          lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(out_hi + ob + d0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (out_lo) *reinterpret_cast<uint4*>(out_lo + ob + d0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*F5EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static F5EncodeFn f5_encode() {
  static F5EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (F5EncodeFn)p;
  }
  return fn;
}
// 3-D fp16 map {inner, rows, z}, 128-byte swizzle, out-of-range rows / columns read as zero
static int f5_map(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t z, uint64_t row_stride_elems, uint32_t box_inner,
                  uint32_t box_rows) {
  F5EncodeFn enc = f5_encode();
  QB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  const cuuint64_t dims[3] = {inner, rows, z};
  const cuuint64_t strides[2] = {row_stride_elems * 2, row_stride_elems * rows * 2};
  const cuuint32_t box[3] = {box_inner, box_rows, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QB_REQUIRE(r == CUDA_SUCCESS, "attention_umma: cuTensorMapEncodeTiled failed: %d (dims %llu %llu %llu)", (int)r, (unsigned long long)inner,
             (unsigned long long)rows, (unsigned long long)z);
  return 0;
}

template <int D, bool SPLIT>
static int f5_launch(const float* qkv, int64_t B, int64_t L, int32_t H, const float* rc, const float* rs, __half* out_hi, __half* out_lo,
                     void* workspace, cudaStream_t st) {
  using C = F5Cfg<D, SPLIT>;
  constexpr int NPL = C::NPL;
  const int64_t BH = B * H, Lp = (L + 7) / 8 * 8;
  __half* q16 = (__half*)workspace;
  __half* k16 = q16 + NPL * BH * L * D;
  __half* vT = k16 + NPL * BH * L * D;
  fa5_prep_kernel<D><<<dim3((unsigned)ceil_div(L, 32), (unsigned)H, (unsigned)B), 256, 0, st>>>(qkv, (int)L, (int)Lp, H, rc, rs,
                                                                                                1.0f / sqrtf((float)D), q16, k16, vT, NPL);
  g_launches++;
  QB_CHECK_CUDA(cudaGetLastError());
  CUtensorMap tmQ, tmK, tmV;
  if (int e = f5_map(&tmQ, q16, D, L, NPL * BH, D, 64, F5_BQ)) return e;
  if (int e = f5_map(&tmK, k16, D, L, NPL * BH, D, 64, F5_BK)) return e;
  if (int e = f5_map(&tmV, vT, L, D, NPL * BH, Lp, F5_BK, D)) return e;
  QB_CHECK_CUDA(cudaFuncSetAttribute(fa5_kernel<D, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  fa5_kernel<D, SPLIT><<<dim3((unsigned)ceil_div(L, F5_BQ), (unsigned)H, (unsigned)B), F5_THREADS, C::SMEM, st>>>(tmQ, tmK, tmV, (int)L, H,
                                                                                                                (int)BH, out_hi, out_lo);
  g_launches++;
  QB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace qb
using namespace qb;

extern "C" int64_t qb_attention_umma_workspace_bytes(int64_t B, int64_t L, int32_t heads, int32_t head_dim, int32_t split) {
  const int64_t planes = split ? 2 : 1, Lp = (L + 7) / 8 * 8;
  return planes * B * heads * head_dim * (2 * L + Lp) * 2 + 1024;
}

extern "C" int qb_attention_umma(const float* qkv, int64_t B, int64_t L, int32_t heads, int32_t head_dim, const float* rope_cos,
                                 const float* rope_sin, qb_half* out_hi, qb_half* out_lo, int32_t split, void* workspace, void* stream) {
  QB_REQUIRE(qkv && rope_cos && rope_sin && out_hi && workspace && B > 0 && L > 0 && heads > 0, "attention_umma: bad args");
  QB_REQUIRE(head_dim == 64 || head_dim == 128, "attention_umma: head_dim must be 64 or 128 (got %d)", head_dim);
  QB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127) == 0, "attention_umma: workspace must be 128-byte aligned");
  QB_REQUIRE(B <= 65535 && heads <= 65535, "attention_umma: grid limits (B, heads <= 65535)");
  cudaStream_t st = (cudaStream_t)stream;
  __half *oh = (__half*)out_hi, *ol = (__half*)out_lo;
  if (head_dim == 64)
    return split ? f5_launch<64, true>(qkv, B, L, heads, rope_cos, rope_sin, oh, ol, workspace, st)
                 : f5_launch<64, false>(qkv, B, L, heads, rope_cos, rope_sin, oh, ol, workspace, st);
  return split ? f5_launch<128, true>(qkv, B, L, heads, rope_cos, rope_sin, oh, ol, workspace, st)
               : f5_launch<128, false>(qkv, B, L, heads, rope_cos, rope_sin, oh, ol, workspace, st);
}
