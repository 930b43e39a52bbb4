// Non-causal multi-head self-attention with fused RoPE, head_dim 64
// (reference: HCodec-2.0/vq/encoder_modules/transformer.py:134-215).
// v1: fp32 SIMT flash-style kernel - one thread per query row (q, o in registers), K/V tiles of 32
// keys staged in shared memory and read as warp-wide broadcasts.  Attention is ~0.5 % of the path's
// FLOPs; the tensor-core version is a later optimisation.
#include <atomic>

#include "common.cuh"
#include "quark_b200.h"

namespace qb {
extern std::atomic<long long> g_launches;

constexpr int ATT_D = 64, ATT_KT = 32, ATT_THREADS = 128;

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const float* __restrict__ qkv, int T, int H, const float* __restrict__ rcos,
                 const float* __restrict__ rsin, __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  __shared__ __align__(16) float ks[ATT_KT][ATT_D];
  __shared__ __align__(16) float vs[ATT_KT][ATT_D];
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tq = qt * ATT_THREADS + threadIdx.x;
  const long long pitch = 3LL * H * ATT_D;
  const float* base = qkv + (long long)b * T * pitch;
  const bool active = tq < T;

  float q[ATT_D], o[ATT_D];
  if (active) {
    const float* qp = base + (long long)tq * pitch + h * ATT_D;
    const float* c = rcos + (long long)tq * ATT_D;
    const float* s = rsin + (long long)tq * ATT_D;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      float x1 = qp[d], x2 = qp[d + 32];
      q[d] = (x1 * c[d] - x2 * s[d]) * 0.125f;            // q*cos + rotate_half(q)*sin, then * head_dim^-0.5
      q[d + 32] = (x2 * c[d + 32] + x1 * s[d + 32]) * 0.125f;
    }
  } else {
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) q[d] = 0.f;
  }
#pragma unroll
  for (int d = 0; d < ATT_D; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;

  for (int k0 = 0; k0 < T; k0 += ATT_KT) {
    __syncthreads();
    // stage K (with RoPE) and V: 32 keys x 64 dims, 16 elements per thread
    for (int e = threadIdx.x; e < ATT_KT * 32; e += ATT_THREADS) {
      const int j = e >> 5, d = e & 31, tk = k0 + j;
      float k1 = 0.f, k2 = 0.f, v1 = 0.f, v2 = 0.f;
      if (tk < T) {
        const float* kp = base + (long long)tk * pitch + (H + h) * ATT_D;
        const float* vp = base + (long long)tk * pitch + (2 * H + h) * ATT_D;
        const float x1 = kp[d], x2 = kp[d + 32];
        const float* c = rcos + (long long)tk * ATT_D;
        const float* s = rsin + (long long)tk * ATT_D;
        k1 = x1 * c[d] - x2 * s[d];
        k2 = x2 * c[d + 32] + x1 * s[d + 32];
        v1 = vp[d];
        v2 = vp[d + 32];
      }
      ks[j][d] = k1; ks[j][d + 32] = k2;
      vs[j][d] = v1; vs[j][d + 32] = v2;
    }
    __syncthreads();
    float sc[ATT_KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < ATT_KT; ++j) {
      float acc = 0.f;
      const float4* kr = reinterpret_cast<const float4*>(ks[j]);
#pragma unroll
      for (int d4 = 0; d4 < ATT_D / 4; ++d4) {
        float4 kk = kr[d4];
        acc = fmaf(q[4 * d4], kk.x, acc);
        acc = fmaf(q[4 * d4 + 1], kk.y, acc);
        acc = fmaf(q[4 * d4 + 2], kk.z, acc);
        acc = fmaf(q[4 * d4 + 3], kk.w, acc);
      }
      sc[j] = (k0 + j < T) ? acc : -INFINITY;
      tmax = fmaxf(tmax, sc[j]);
    }
    const float m_new = fmaxf(m, tmax);
    const float corr = expf(m - m_new);   // m = -inf on the first tile -> 0
    l *= corr;
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < ATT_KT; ++j) {
      const float pj = expf(sc[j] - m_new);
      l += pj;
      const float4* vr = reinterpret_cast<const float4*>(vs[j]);
#pragma unroll
      for (int d4 = 0; d4 < ATT_D / 4; ++d4) {
        float4 vv = vr[d4];
        o[4 * d4] = fmaf(pj, vv.x, o[4 * d4]);
        o[4 * d4 + 1] = fmaf(pj, vv.y, o[4 * d4 + 1]);
        o[4 * d4 + 2] = fmaf(pj, vv.z, o[4 * d4 + 2]);
        o[4 * d4 + 3] = fmaf(pj, vv.w, o[4 * d4 + 3]);
      }
    }
    m = m_new;
  }
  if (active) {
    const float inv = 1.f / l;
    const long long ob = ((long long)b * T + tq) * (long long)(H * ATT_D) + h * ATT_D;
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) {
      __half hh, ll;
      split_f16(o[d] * inv, hh, ll);
      out_hi[ob + d] = hh;
      if (out_lo) out_lo[ob + d] = ll;
    }
  }
}

}  // namespace qb
using namespace qb;

extern "C" int qb_attention(const float* qkv, int64_t B, int64_t T, int32_t heads, const float* rope_cos,
                            const float* rope_sin, qb_half* out_hi, qb_half* out_lo, void* stream) {
  QB_REQUIRE(qkv && rope_cos && rope_sin && out_hi && T > 0 && heads > 0, "attention: bad args");
  dim3 grid((unsigned)ceil_div(T, ATT_THREADS), (unsigned)heads, (unsigned)B);
  attention_kernel<<<grid, ATT_THREADS, 0, (cudaStream_t)stream>>>(qkv, (int)T, heads, rope_cos, rope_sin,
                                                                  (__half*)out_hi, (__half*)out_lo);
  g_launches++;
  QB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
