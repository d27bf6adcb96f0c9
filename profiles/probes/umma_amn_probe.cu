// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, N = 64, K = 16, B MN-major) as a function of the A operand:
// K-major shared memory, MN-major shared memory (two 64-row blocks 16 KB apart: the transposed read of dS^T that the fused
// attention backward uses for dQ = dS K), or tensor memory.  One CTA per SM, one issuing thread, 128-byte swizzle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I plbert_b200/csrc -I include -o umma_amn_probe umma_amn_probe.cu -lcuda
#include <cstdio>
#include "common.cuh"
using namespace plb;

template <int A_MODE, int N, int B_MN, int NOISE = 0>
__global__ void __launch_bounds__(384, 1) probe(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (NOISE && warp >= 4) {
    // the softmax warps of the fused attention backward: eight warps pulling S^T / dP^T out of tensor memory at full rate
    // (NOISE 2: and writing packed results back)
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t acc = 0, col = 0;
    while (!stop) {
      uint32_t v[4][16];
#pragma unroll
      for (int k = 0; k < 4; ++k) tmem_ld_32x16(tmem + lane_addr + 128 + ((col + 16 * k) & 127), v[k]);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= v[k][i];
      if (NOISE == 2) {
        tmem_st_32x16(tmem + lane_addr + 320 + (col & 63), v[0]);
        tmem_st_wait();
      }
      col += 64;
    }
    if (acc == 0x12345678) out[1] = acc;
  }
  if (warp == 1 && elect_one_sync()) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, A_MODE == 1 ? 1 : 0, B_MN);
    const uint64_t dA = A_MODE == 1 ? umma_smem_desc(smem_u32(smem), 16384, 1024) : umma_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t dB = B_MN ? umma_smem_desc(smem_u32(smem + 32768), 16384, 1024) : umma_smem_desc(smem_u32(smem + 32768), 16, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (A_MODE == 2) umma_bf16_tmem_a(tmem + (it & 1) * 64, tmem + 256 + 8 * kk, dB + (B_MN ? 128 * kk : 2 * (kk & 3)), idesc, 1);
        else umma_bf16(tmem + (it & 1) * 64, dA + (A_MODE == 1 ? 128 * kk : (kk >> 2) * 1024 + 2 * (kk & 3)),
                       dB + (B_MN ? 128 * kk : 2 * (kk & 3)), idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) out[0] = t2 - t0;
    stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int A_MODE, int N, int B_MN, int NOISE = 0>
void run(const char* name) {
  long long* out; cudaMalloc(&out, 16);
  const int smem = 32768 + 4 * 16384 + 1024, iters = 256;
  cudaFuncSetAttribute(probe<A_MODE, N, B_MN, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<A_MODE, N, B_MN, NOISE><<<148, 384, smem>>>(out, iters);
  long long h[1]; cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-40s N=%3d B %s: %.1f clk/MMA (math floor %d)  %s\n", name, N, B_MN ? "MN-major" : "K-major ", (double)h[0] / (iters * 8), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}
int main() {
  run<0, 64, 1>("A K-major smem");
  run<1, 64, 1>("A MN-major smem (2 x 64 rows, LBO 16 KB)");
  run<2, 64, 1>("A tensor memory");
  run<0, 128, 0>("A K-major smem");
  run<1, 128, 0>("A MN-major smem (2 x 64 rows, LBO 16 KB)");
  run<0, 128, 1>("A K-major smem");
  run<1, 128, 1>("A MN-major smem (2 x 64 rows, LBO 16 KB)");
  run<1, 256, 1>("A MN-major smem (2 x 64 rows, LBO 16 KB)");
  run<0, 64, 1, 1>("A K-major smem + 8 warps tcgen05.ld");
  run<1, 64, 1, 1>("A MN-major smem + 8 warps tcgen05.ld");
  run<2, 64, 1, 1>("A tensor memory + 8 warps tcgen05.ld");
  run<0, 128, 0, 1>("A K-major smem + 8 warps tcgen05.ld");
  run<0, 64, 1, 2>("A K-major smem + 8 warps ld + st");
  run<2, 64, 1, 2>("A tensor memory + 8 warps ld + st");
  return 0;
}
