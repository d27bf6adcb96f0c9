// Probe: tcgen05.mma with the A operand in tensor memory (bf16 pairs packed per 32-bit column, lane = row).
// D[128 x 64] = A[128 x 64] * B[64 x 64]^T, A written with tcgen05.st, B K-major in 128-byte-swizzled shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I plbert_b200/csrc -I include ...
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "common.cuh"
using namespace plb;

__device__ __forceinline__ void umma_bf16_tmem_a(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* Dout) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, row = threadIdx.x;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  // A row -> 32 packed words -> TMEM columns [64, 96)
  uint32_t a[32];
  for (int i = 0; i < 32; ++i) a[i] = reinterpret_cast<const uint32_t*>(A + row * 64)[i];
  tmem_st_32x32(tmem + lane_addr + 64, a);
  tmem_st_wait();
  // B[n][k] (64 x 64), K-major, 128-byte swizzle: 16-byte chunk c of row n at n*128 + ((c ^ (n & 7)) << 4)
  if (row < 64) {
    for (int c = 0; c < 8; ++c) {
      const uint4 v = reinterpret_cast<const uint4*>(B + row * 64)[c];
      *reinterpret_cast<uint4*>(smem + row * 128 + ((c ^ (row & 7)) << 4)) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (warp == 1 && elect_one_sync()) {
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const uint64_t dB = umma_smem_desc(smem_u32(smem), 16, 1024);
    for (int kk = 0; kk < 4; ++kk) umma_bf16_tmem_a(tmem, tmem + 64 + kk * 8, dB + 2 * kk, idesc, kk > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t d[32];
  for (int q = 0; q < 2; ++q) {
    tmem_ld_32x32(tmem + lane_addr + q * 32, d);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) Dout[row * 64 + q * 32 + i] = __uint_as_float(d[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
  std::vector<__nv_bfloat16> hA(128 * 64), hB(64 * 64);
  std::vector<float> fA(128 * 64), fB(64 * 64), ref(128 * 64), out(128 * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      float s = 0;
      for (int k = 0; k < 64; ++k) s += fA[m * 64 + k] * fB[n * 64 + k];
      ref[m * 64 + n] = s;
    }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  probe<<<1, 128, 16384>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (size_t i = 0; i < out.size(); ++i) maxerr = fmax(maxerr, fabs(out[i] - ref[i]));
  printf("tmem-A MMA: %s, max |err| = %g (ref[0]=%g out[0]=%g, ref[100]=%g out[100]=%g)\n", cudaGetErrorString(e), maxerr, ref[0], out[0],
         ref[100], out[100]);
  return 0;
}
