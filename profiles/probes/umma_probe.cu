// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N and of the B-operand layout,
// operands in 128-byte-swizzled shared memory (contents irrelevant).  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I plbert_b200/csrc -I include -o umma_probe umma_probe.cu
#include <cstdio>
#include "common.cuh"
using namespace plb;

template <int N, int B_MN, int ISSUERS = 1, int NOISE = 0>
__global__ void __launch_bounds__(256, 1) probe(long long* out, int iters) {
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
  __shared__ uint64_t bar2;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) { mbar_init(&bar2, 1); fence_mbar_init(); stop = 0; }
  __syncthreads();
  if (NOISE && warp >= 4) {
    // softmax-like background traffic: TMEM loads + 16-byte shared-memory stores while the MMAs run
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t v[32];
    uint32_t acc = 0;
    while (!stop) {
      tmem_ld_32x32(tmem + lane_addr + 384, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        sts128(smem_u32(smem) + 5 * 16384 + (threadIdx.x - 128) * 16, make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        acc += v[i];
      }
    }
    if (acc == 0x12345678) out[3] = acc;
  }
  if (warp == 2 && ISSUERS == 2 && elect_one_sync()) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, B_MN);
    const uint64_t dA = umma_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t dB = B_MN ? umma_smem_desc(smem_u32(smem + 16384), 16384, 1024) : umma_smem_desc(smem_u32(smem + 16384), 16, 1024);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem + 128 + (it & 1) * 256, dA + 2 * kk, dB + (B_MN ? 128 * kk : 2 * kk), idesc, 1);
    }
    umma_commit(&bar2);
    mbar_wait(&bar2, 0);
  }
  if (warp == 1 && elect_one_sync()) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, B_MN);
    const uint64_t dA = umma_smem_desc(smem_u32(smem), 16, 1024);                       // A: [128 x 64] K-major, 16 KB
    const uint64_t dB = B_MN ? umma_smem_desc(smem_u32(smem + 16384), 16384, 1024)      // B: MN-major, 64-element blocks 16 KB apart
                             : umma_smem_desc(smem_u32(smem + 16384), 16, 1024);        // B: [N x 64] K-major
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem + (it & 1) * 256, dA + 2 * kk, dB + (B_MN ? 128 * kk : 2 * kk), idesc, 1);
    }
    umma_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int B_MN, int ISSUERS = 1, int NOISE = 0>
void run(const char* name) {
  long long* out; cudaMalloc(&out, 16);
  const int smem = 16384 + 4 * 16384 + 4096, iters = 512;
  cudaFuncSetAttribute(probe<N, B_MN, ISSUERS, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<N, B_MN, ISSUERS, NOISE><<<148, 256, smem>>>(out, iters);
  long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-34s N=%3d : issue %.1f clk/MMA, complete %.1f clk/MMA per issuer, %.1f clk/MMA overall (math floor %d)  %s\n", name, N,
         (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), (double)h[1] / (iters * 4 * ISSUERS), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}
int main() {
  run<64, 0>("B K-major");
  run<128, 0>("B K-major");
  run<256, 0>("B K-major");
  run<64, 1>("B MN-major");
  run<128, 1>("B MN-major");
  run<64, 0, 2>("two issuers");
  run<128, 0, 2>("two issuers");
  run<64, 0, 1, 1>("one issuer + ld/st noise");
  run<64, 0, 2, 1>("two issuers + ld/st noise");
  run<128, 0, 2, 1>("two issuers + ld/st noise");
  return 0;
}
