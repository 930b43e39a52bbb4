// LSTM recurrence on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), persistent cooperative kernel.
// Reference: nn.LSTM(H,H,1,batch_first), HCodec-2.0/vq/encoder_modules/transformer.py:115,133.
//
// Measured (profiles/): the mma.sync version (lstm.cu) spends 15.4k of its 26k cycles per step in legacy
// HMMA issue - mma.sync runs at a fraction of the tcgen05 rate on sm_100.  Here each CTA keeps its W_hh
// slice [4U gate rows x H] resident in shared memory as the UMMA *B* operand (K-major, 128B swizzle, loaded
// once by TMA), streams h_{t-1} [batch x H] (fp16, published by all CTAs) through a 4-stage TMA ring as
// the *A* operand (M = 128 batch rows), and accumulates gates[batch, 4U] in TMEM: 96 tcgen05.mma of
// 128 x 4U x 16 per step (24 cycles each) instead of 2304 HMMA.  8 epilogue warps read TMEM (lane = batch
// row), add the precomputed input projection, do the cell update (c in registers) and publish h_t; a
// flag-per-CTA grid barrier separates the steps.
// W rows are pre-permuted by the host to unit-major order: row (4*j + g) of CTA c = gate g of unit c*U + j.
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "quark_b200.h"

namespace qb {
extern std::atomic<long long> g_launches;

constexpr int LT_EPI_WARPS = 8, LT_THREADS = (LT_EPI_WARPS + 2) * 32, LT_MAX_STAGES = 8;
constexpr int LT_REP = 1;                        // replicas of the published h (readers pick blockIdx % LT_REP)
constexpr int LT_KG = 4;                         // K-blocks fetched by one TMA instruction (4-D box)
constexpr uint32_t LT_RING_BYTES = 72 * 1024;   // 8 x 8 KB slots + 8 KB pad (batch 64) or 4 x 16 KB (+ pad)

__device__ __forceinline__ unsigned lt_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int x, int y, int z, int w) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(w)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}

template <int U>
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
               const float* __restrict__ xp, int B, int T, int H, __half* __restrict__ out_hi,
               __half* __restrict__ out_lo, __half* hbuf, unsigned* flags, int Bp, int box_rows, long long* prof) {
  constexpr int N = 4 * U;                      // gate rows of this CTA = UMMA N
  constexpr int HALF = U / 2;                   // units per epilogue thread
  static_assert(N % 16 == 0 && N <= 256 && HALF * 4 % 8 == 0, "unsupported slice width");
  constexpr uint32_t WBLK = N * 128;            // bytes of one [N x 64] K-block of W
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int KB = H / 64;
  uint8_t* Wsm = smem;
  uint8_t* ring = smem + (size_t)KB * WBLK;     // KB*WBLK is a multiple of 1024 when N % 8 == 0
  // ring slots hold box_rows (64 or 128) rows; the UMMA A tile always spans 128 rows, so with 64-row slots
  // the upper half of a tile aliases the next slot (finite values -> garbage only in the unused D rows)
  const uint32_t kblk_bytes = (uint32_t)box_rows * 128;              // one K-block of h in the ring
  const uint32_t slot_bytes = kblk_bytes * LT_KG;                      // one TMA instruction
  const int n_stages = (box_rows == 64 ? 8 : 4) / LT_KG;
  uint64_t* full = (uint64_t*)(ring + LT_RING_BYTES);
  uint64_t* empty = full + LT_MAX_STAGES;
  uint64_t* wbar = empty + LT_MAX_STAGES;
  uint64_t* tfull = wbar + 1;
  uint32_t* tmem_slot = (uint32_t*)(tfull + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = blockIdx.x * U, G = gridDim.x;
  const int passes = (Bp + 127) / 128;
  constexpr uint32_t TCOLS = 2 * N <= 64 ? 64 : (2 * N <= 128 ? 128 : (2 * N <= 256 ? 256 : 512));

  if (tid == 0) {
    for (int s = 0; s < LT_MAX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(wbar, 1);
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  // zero the ring once: rows beyond the TMA box (batch < 128) must stay finite
  for (int i = tid; i < (int)(LT_RING_BYTES / 16); i += LT_THREADS) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == LT_EPI_WARPS + 1) { tmem_alloc(tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == LT_EPI_WARPS) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      prefetch_tmap(&tmW);
      prefetch_tmap(&tmH);
      mbar_arrive_expect_tx(wbar, (uint32_t)KB * WBLK);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(Wsm + (size_t)kb * WBLK, &tmW, wbar, kb * 64, blockIdx.x * N);
    }
    uint32_t stage = 0, phase = 0;
    long long pw = 0, pl = 0;
    for (int t = 1; t < T; ++t) {
      // wait until every CTA has published h_{t-1}
      long long c0 = clock64();
      unsigned spins = 0;
      for (;;) {
        bool ok = true;
        for (int c = lane; c < G; c += 32) ok = ok && (lt_ld_acquire(flags + c) >= (unsigned)t);
        if (__all_sync(0xffffffffu, ok)) break;
        if (++spins > (1u << 26)) asm volatile("trap;");
      }
      long long c1 = clock64();
      pw += c1 - c0;
      if (lane == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy writes of h -> async-proxy (TMA) reads
        const int buf = (t + 1) & 1;                            // h_{t-1} lives in buffer (t-1)&1
        for (int ps = 0; ps < passes; ++ps)
          for (int kb = 0; kb < KB; kb += LT_KG) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], slot_bytes);
            tma_load_4d(ring + stage * slot_bytes, &tmH, &full[stage], 0, ps * 128, kb, buf);
            if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
          }
      }
      __syncwarp();
      pl += clock64() - c1;
    }
    if (prof && lane == 0 && blockIdx.x == 0) { prof[0] = pw; prof[1] = pl; }
  } else if (warp == LT_EPI_WARPS + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, N);
      mbar_wait(wbar, 0);
      uint32_t stage = 0, phase = 0;
      long long mw = 0, mfirst = 0;
      for (int t = 1; t < T; ++t) {
        long long m0 = clock64();
        for (int ps = 0; ps < passes; ++ps) {
          const uint32_t d_tmem = tmem_base + ps * N;
          for (int kb0 = 0; kb0 < KB; kb0 += LT_KG) {
            mbar_wait(&full[stage], phase);
            if (ps == 0 && kb0 == 0) { const long long m1 = clock64(); mfirst += m1 - m0; m0 = m1; }
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < LT_KG; ++g) {
              const int kb = kb0 + g;
              const uint32_t sa = smem_u32(ring + stage * slot_bytes + g * kblk_bytes), sw = smem_u32(Wsm + (size_t)kb * WBLK);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(d_tmem, make_sw128_kmajor_desc(sa + k * 32), make_sw128_kmajor_desc(sw + k * 32), idesc,
                         (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty[stage]);
            if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(tfull);
        mw += clock64() - m0;
      }
      if (prof && blockIdx.x == 0) { prof[2] = mfirst; prof[3] = mw; }
    }
  } else {
    // ===================== epilogue: cell update =====================
    const int q = warp & 3, half = warp >> 2;            // TMEM lane quadrant, unit half
    float c[2][HALF];
#pragma unroll
    for (int ps = 0; ps < 2; ++ps)
#pragma unroll
      for (int i = 0; i < HALF; ++i) c[ps][i] = 0.f;
    long long ew = 0, ec = 0, ep = 0;
    for (int t = 0; t < T; ++t) {
      long long e0 = clock64(), e1 = e0;
      __half* hcur = hbuf + (size_t)(t & 1) * LT_REP * Bp * H;
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        if (ps < passes) {
          const int n = ps * 128 + q * 32 + lane;
          const bool act = n < B;
          // input projection for (row n, units u0 + half*HALF .. +HALF), all 4 gates
          float xg[4][HALF];
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < HALF; ++i)
              xg[g][i] = act ? xp[((long long)n * T + t) * 4 * H + (long long)g * H + u0 + half * HALF + i] : 0.f;
          float acc[HALF * 4];
#pragma unroll
          for (int i = 0; i < HALF * 4; ++i) acc[i] = 0.f;
          if (t > 0) {
            if (ps == 0) { mbar_wait(tfull, (t - 1) & 1); tc_fence_after(); e1 = clock64(); ew += e1 - e0; }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ps * N + half * HALF * 4;
#pragma unroll
            for (int cc = 0; cc < HALF * 4; cc += 8) {
              uint32_t r[8];
              tmem_ld_32x32b_x8(taddr + cc, r);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[cc + e] = __uint_as_float(r[e]);
            }
          }
          if (act) {
            __half hv[HALF], lv[HALF];
#pragma unroll
            for (int i = 0; i < HALF; ++i) {
              const float gi = acc[4 * i] + xg[0][i], gf = acc[4 * i + 1] + xg[1][i];
              const float gg = acc[4 * i + 2] + xg[2][i], go = acc[4 * i + 3] + xg[3][i];
              const float ig = sigmoid_acc(gi), fg = sigmoid_acc(gf), cg = tanhf(gg), og = sigmoid_acc(go);
              const float cn = fg * c[ps][i] + ig * cg;
              c[ps][i] = cn;
              split_f16(og * tanhf(cn), hv[i], lv[i]);
            }
            // packed stores: HALF halves are contiguous (4-byte aligned: u0 and HALF are even)
            const int u = u0 + half * HALF;
            const long long o = ((long long)n * T + t) * H + u;
#pragma unroll
            for (int i = 0; i < HALF; i += 2) {
              const __half2 h2 = __halves2half2(hv[i], hv[i + 1]);
#pragma unroll
              for (int rp = 0; rp < LT_REP; ++rp) *reinterpret_cast<__half2*>(hcur + ((size_t)rp * Bp + n) * H + u + i) = h2;
              *reinterpret_cast<__half2*>(out_hi + o + i) = h2;
              if (out_lo) *reinterpret_cast<__half2*>(out_lo + o + i) = __halves2half2(lv[i], lv[i + 1]);
            }
          }
        }
      }
      // publish h_t: all epilogue threads done (TMEM reads + h stores) -> flag
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(LT_EPI_WARPS * 32) : "memory");
      const long long e2 = clock64();
      ec += e2 - e1;
      if (tid == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"((unsigned)(t + 1)) : "memory");
        ep += clock64() - e2;
      }
    }
    if (prof && tid == 0 && blockIdx.x == 0) { prof[4] = ew; prof[5] = ec; prof[6] = ep; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == LT_EPI_WARPS + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn lt_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace qb
using namespace qb;

extern "C" int32_t qb_lstm_tc_units(int64_t H) {
  // units per CTA: H/U CTAs must be co-resident (<= SM count) and 4U a multiple of 16
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    sms = 148;
  for (int U : {4, 8, 12}) if (H % U == 0 && H / U <= sms) return U;
  return 0;
}

extern "C" int64_t qb_lstm_tc_workspace_bytes(int64_t B, int64_t H) {
  const int64_t Bp = B <= 64 ? 64 : ceil_div(B, 128) * 128;
  return 2 * LT_REP * Bp * H * 2 + 4096;
}

extern "C" int qb_lstm_tc(const float* xp, const qb_half* whh_perm, int32_t units, int64_t B, int64_t T, int64_t H,
                          qb_half* out_hi, qb_half* out_lo, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  QB_REQUIRE(xp && whh_perm && out_hi && workspace, "lstm_tc: bad args");
  QB_REQUIRE(H % (64 * LT_KG) == 0 && (units == 4 || units == 8 || units == 12) && H % units == 0, "lstm_tc: unsupported H / units");
  QB_REQUIRE(B >= 1 && B <= 256, "lstm_tc: batch must be 1..256 per call (got %lld)", (long long)B);
  const int U = units, N = 4 * U, grid = (int)(H / U), KB = (int)(H / 64);
  const int Bp = (int)(B <= 64 ? 64 : ceil_div(B, 128) * 128);
  const int box_rows = Bp < 128 ? Bp : 128;
  EncodeTiledFn enc = lt_encode();
  QB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available");
  CUtensorMap tmW, tmH;
  cuuint32_t es[3] = {1, 1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)(4 * H)};
    cuuint64_t str[1] = {(cuuint64_t)H * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)N};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)whh_perm, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    QB_REQUIRE(r == CUDA_SUCCESS, "lstm_tc: W tensor map failed (%d)", (int)r);
  }
  __half* hbuf = (__half*)workspace;
  {
    // 4-D view (k within block, batch row, K-block, buffer): one box = LT_KG consecutive K-block tiles
    cuuint64_t dims[4] = {64, (cuuint64_t)Bp, (cuuint64_t)(H / 64), 2};
    cuuint64_t str[3] = {(cuuint64_t)H * 2, 128, (cuuint64_t)Bp * H * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, LT_KG, 1};
    cuuint32_t es4[4] = {1, 1, 1, 1};
    (void)es;
    CUresult r = enc(&tmH, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)hbuf, dims, str, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    QB_REQUIRE(r == CUDA_SUCCESS, "lstm_tc: h tensor map failed (%d)", (int)r);
  }
  const size_t smem = (size_t)KB * N * 128 + (size_t)LT_RING_BYTES + 1024 + 256;
  QB_REQUIRE(smem <= 227 * 1024, "lstm_tc: shared memory budget exceeded (%zu)", smem);
  QB_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)qb_lstm_tc_workspace_bytes(B, H), st));
  unsigned* flags = (unsigned*)((uint8_t*)workspace + (size_t)2 * LT_REP * Bp * H * 2);
  __half* oh = (__half*)out_hi;
  __half* ol = (__half*)out_lo;
  int Bi = (int)B, Ti = (int)T, Hi = (int)H, Bpi = Bp, br = box_rows;
  static long long* prof = nullptr;
  if (!prof && getenv("QB_LSTM_PROF")) { cudaMalloc(&prof, 64); cudaMemset(prof, 0, 64); }
  void* args[] = {&tmW, &tmH, &xp, &Bi, &Ti, &Hi, &oh, &ol, &hbuf, &flags, &Bpi, &br, &prof};
  const void* fn = U == 4 ? (const void*)lstm_tc_kernel<4> : U == 8 ? (const void*)lstm_tc_kernel<8> : (const void*)lstm_tc_kernel<12>;
  QB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  QB_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(LT_THREADS), args, smem, st));
  g_launches++;
  if (prof) {
    long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, prof, 64, cudaMemcpyDeviceToHost);
    const double d = (double)Ti;
    fprintf(stderr, "[lstm_tc prof] cycles/step: producer flag-wait %.0f tma-issue %.0f | mma first-block wait %.0f rest %.0f | "
            "epilogue wait-acc %.0f compute %.0f publish %.0f\n", h[0] / d, h[1] / d, h[2] / d, h[3] / d, h[4] / d, h[5] / d, h[6] / d);
  }
  return 0;
}
