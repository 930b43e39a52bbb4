// Probe: tcgen05.cp.32x128b.warpx4 (smem -> TMEM, 32 rows broadcast to the 4 lane quadrants) feeding tcgen05.mma with the A operand
// in TMEM (kind::f16, M = 128, N = 48, K = 64 = 4 x k16).  Checks D = A B^T against the host for rows 0..31 of every quadrant.
// A in shared memory: core-matrix layout, no swizzle: [kc = k/8][rb = r/8][r%8][k%8] (8 x 8 halves = 128 B per core matrix).
// B in shared memory: K-major rows of 128 B with the 128-byte swizzle (as TMA writes the W slice).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I unified_audio_b200/csrc -I include profiles/experiments/ts_mma_probe.cu -o gpurun_out/ts_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "common.cuh"
using namespace qb;
namespace qb { void set_error(const char*, ...) {} }

__device__ __forceinline__ uint64_t desc_noswizzle(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;      // layout_type 0 = no swizzle
}
__device__ __forceinline__ void cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}

constexpr int N = 48, K = 64;
__global__ void __launch_bounds__(128) probe(const __half* A /*[32][K]*/, const __half* B /*[N][K]*/, float* D /*[128][N]*/, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __half* sA = (__half*)smem;                 // 32 x 64 halves = 4 KB, core-matrix layout
  __half* sB = (__half*)(smem + 4096);        // 64 rows x 128 B (48 used), SW128
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 32 * K; i += 128) {
    const int r = i / K, k = i % K;
    sA[(((k >> 3) * 4 + (r >> 3)) * 8 + (r & 7)) * 8 + (k & 7)] = A[i];
  }
  for (int i = tid; i < 64 * K; i += 128) {
    const int n = i / K, k = i % K;
    const __half v = n < N ? B[n * K + k] : __float2half(0.f);
    sB[n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7))] = v;
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async();
  if (warp == 0) { tmem_alloc(&tslot, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  const uint32_t a_col = 64;                  // A: columns 64..95 (K = 64 halves = 32 columns), D: columns 0..47
  if (tid == 0) {
    // 8 chunks of 8 halves (128 bit): chunk kc -> TMEM columns a_col + 4 kc; source = 4 core matrices (rb = 0..3) 128 B apart
    for (int kc = 0; kc < 8; ++kc) {
      const uint32_t src = smem_u32(sA) + kc * 512;
      const uint64_t d = variant == 0 ? desc_noswizzle(src, 128, 128) : desc_noswizzle(src, 512, 128);
      cp_32x128b_warpx4(tb + a_col + kc * 4, d);
    }
    const uint32_t idesc = make_idesc_f16(128, N);
    for (int k = 0; k < 4; ++k)
      umma_ts_f16(tb, tb + a_col + k * 8, make_sw128_kmajor_desc(smem_u32(sB) + k * 32), idesc, k ? 1u : 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    tmem_ld_x16(tb + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 128); }
}

int main(int argc, char** argv) {
  std::vector<__half> hA(32 * K), hB(N * K);
  std::vector<float> fA(32 * K), fB(N * K);
  srand(1);
  for (int i = 0; i < 32 * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dD, 0, 128 * N * 4);
    probe<<<1, 128, 4096 + 8192 + 1024>>>(dA, dB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> hD(128 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    int bad_q = -1;
    for (int q = 0; q < 4; ++q)
      for (int r = 0; r < 32; ++r)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)fA[r * K + k] * fB[n * K + k];
          const double err = fabs(ref - hD[(q * 32 + r) * N + n]);
          if (err > worst) { worst = err; bad_q = q; }
        }
    printf("variant %d: %s  max |D - A B^T| over 4 quadrants = %.3e (quadrant %d)  D[0][0..3] = %.4f %.4f %.4f %.4f\n", variant,
           cudaGetErrorString(e), worst, bad_q, hD[0], hD[1], hD[2], hD[3]);
  }
  return 0;
}
