// Residual vector quantiser (reference call sites HCodec-2.0/vq/codec.py:81-82, 94-95; arithmetic
// template HCodec-2.0/vq/core_vq.py:223-238, 394-412; upstream vector-quantize-pytorch 1.22.15).
//
// Encode, per layer q (strictly sequential - the residual chain):
//   1. scores[m,j] = |e_j|^2 - 2 r_m.e_j  on the tensor cores: the tcgen05 GEMM with 3-term fp16
//      split operands (~2^-21 relative), bias = -|e|^2/2 and gamma = -2 folded into its epilogue;
//   2. rvq_select (this file): warp per token - arg-min over the K scores; every candidate whose
//      score lies within `tol` of the minimum is re-ranked by its EXACT squared distance in fp64
//      (lowest index wins exact ties), so the chosen index equals the exact-arithmetic arg-min of
//      the fp32 residual path irrespective of tensor-core rounding;
//   3. the same kernel updates the residual in fp32 (r -= e_idx, as the reference), accumulates
//      the quantised sum and emits the next layer's fp16 planes.
#include <atomic>

#include "common.cuh"
#include "quark_b200.h"

namespace qb {
extern std::atomic<long long> g_launches;

__global__ void rvq_select_kernel(const float* __restrict__ scores, float* __restrict__ resid,
                                  const float* __restrict__ cb /*[K,D] layer q*/, long long M, int D, int K,
                                  float tol_rel, float e2max, int64_t* __restrict__ idx, int nq, int q,
                                  float* __restrict__ quant, __half* __restrict__ hi, __half* __restrict__ lo) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const float* sr = scores + m * K;
  float* r = resid + m * D;
  // |r|^2 for the tolerance scale
  float r2 = 0.f;
  for (int d = lane; d < D; d += 32) r2 = fmaf(r[d], r[d], r2);
  r2 = warp_sum(r2);
  // pass 1: fp32 arg-min, lowest index on ties
  float best = INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < K; j += 32) {
    const float s = sr[j];
    if (s < best) { best = s; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  // pass 2: near-ties -> exact fp64 re-rank
  const float thr = best + tol_rel * (r2 + e2max);
  int cnt = 0;
  for (int j = lane; j < K; j += 32) cnt += (sr[j] <= thr) ? 1 : 0;
  cnt = (int)warp_sum((float)cnt);
  if (cnt > 1) {
    double dbest = 1e300;
    int di = 0x7fffffff;
    for (int j0 = 0; j0 < K; j0 += 32) {
      const int j = j0 + lane;
      unsigned mask = __ballot_sync(0xffffffffu, j < K && sr[j] <= thr);
      while (mask) {
        const int cand = j0 + __ffs(mask) - 1;
        mask &= mask - 1;
        const float* e = cb + (long long)cand * D;
        double acc = 0.0;
        for (int d = lane; d < D; d += 32) {
          const double df = (double)r[d] - (double)e[d];
          acc += df * df;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (acc < dbest) { dbest = acc; di = cand; }  // candidates visited in ascending index order
      }
    }
    bi = di;
  }
  if (lane == 0) idx[m * nq + q] = (int64_t)bi;
  const float* e = cb + (long long)bi * D;
  for (int d = lane; d < D; d += 32) {
    const float ev = e[d];
    const float rn = r[d] - ev;
    r[d] = rn;
    if (quant) quant[m * D + d] = (q == 0 ? 0.f : quant[m * D + d]) + ev;
    if (hi) {
      __half h, l;
      split_f16(rn, h, l);
      hi[m * D + d] = h;
      lo[m * D + d] = l;
    }
  }
}

__global__ void rvq_init_kernel(const float* __restrict__ x, float* __restrict__ resid, __half* __restrict__ hi,
                                __half* __restrict__ lo, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  resid[i] = v;
  __half h, l;
  split_f16(v, h, l);
  hi[i] = h;
  lo[i] = l;
}

__global__ void rvq_decode_kernel(const int64_t* __restrict__ idx, const float* __restrict__ cb, int D, int K, int nq,
                                  float* __restrict__ out, long long out_ld, long long col_off) {
  const long long m = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int q = 0; q < nq; ++q) {
      const int64_t i = idx[m * nq + q];
      if (i >= 0) acc += cb[((long long)q * K + i) * D + d];
    }
    out[m * out_ld + col_off + d] = acc;
  }
}

}  // namespace qb
using namespace qb;

extern "C" int64_t qb_rvq_workspace_bytes(int64_t M, int32_t D, int32_t K) {
  return M * D * 4 + 2 * M * D * 2 + M * (int64_t)K * 4 + 1024;
}

extern "C" int qb_rvq_encode(const float* x, const float* codebooks, const qb_half* cb_hi, const qb_half* cb_lo,
                             const float* neg_half_e2, float e2max, int64_t M, int32_t D, int32_t K, int32_t nq,
                             int64_t* idx, float* quantized, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  QB_REQUIRE(x && codebooks && cb_hi && cb_lo && neg_half_e2 && idx && workspace, "rvq_encode: bad args");
  QB_REQUIRE(D % 64 == 0, "rvq_encode: D must be a multiple of 64");
  if (M == 0) return 0;
  uint8_t* ws = (uint8_t*)workspace;
  float* resid = (float*)ws; ws += (size_t)M * D * 4;
  __half* hi = (__half*)ws; ws += (size_t)M * D * 2;
  __half* lo = (__half*)ws; ws += (size_t)M * D * 2;
  ws = (uint8_t*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  float* scores = (float*)ws;
  const long long n = (long long)M * D;
  rvq_init_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(x, resid, hi, lo, n);
  g_launches++;
  for (int q = 0; q < nq; ++q) {
    qb_gemm_desc g = {};
    g.a_hi = (const qb_half*)hi; g.a_lo = (const qb_half*)lo;
    g.a_batch = 1; g.a_rows_per_batch = M; g.a_ld = D; g.taps = 1; g.stride = 1; g.m_per_batch = M;
    g.w_hi = cb_hi + (size_t)q * K * D; g.w_lo = cb_lo + (size_t)q * K * D; g.n = K;
    g.bias = neg_half_e2 + (size_t)q * K;            // v = (acc - |e|^2/2) ...
    g.gamma = neg_half_e2 + (size_t)nq * K;           // ... * (-2): K-vector of -2 appended by the caller
    g.act = QB_ACT_NONE; g.act2 = QB_ACT_NONE;
    g.out_f32.ptr = scores; g.out_f32.ld = K; g.out_f32.rows_per_batch = M; g.out_f32.row_off = 0;
    if (int e = qb_gemm(&g, stream)) return e;
    rvq_select_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(scores, resid, codebooks + (size_t)q * K * D, M, D, K,
                                                                1e-4f, e2max, idx, nq, q, quantized,
                                                                q + 1 < nq ? hi : nullptr, lo);
    g_launches++;
  }
  QB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int qb_rvq_decode(const int64_t* idx, const float* codebooks, int64_t M, int32_t D, int32_t K, int32_t nq,
                             float* out, int64_t out_ld, int64_t col_off, void* stream) {
  QB_REQUIRE(idx && codebooks && out, "rvq_decode: bad args");
  if (M == 0) return 0;
  rvq_decode_kernel<<<(unsigned)M, 128, 0, (cudaStream_t)stream>>>(idx, codebooks, D, K, nq, out, out_ld, col_off);
  g_launches++;
  QB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
