// Micro-benchmark: tensor-memory -> register bandwidth of tcgen05.ld (and register -> tensor memory of tcgen05.st) per SM as a
// function of the number of warps and the vector width.  One CTA per SM; every warp reads its own 32-lane quadrant.
//   nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -I plbert_b200/csrc -I include -o tmem_bw_probe tmem_bw_probe.cu ../../plbert_b200/csrc/abi.cu -lcuda
#include <cstdio>
#include "common.cuh"
using namespace plb;

template <int X>
__device__ __forceinline__ uint32_t ld_once(uint32_t taddr) {
  uint32_t acc = 0;
  if (X == 32) {
    uint32_t v[32];
    tmem_ld_32x32(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= v[i];
  } else {
    uint32_t v[16];
    tmem_ld_32x16(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= v[i];
  }
  return acc;
}

template <int X, int MODE>   // MODE 0: loads with a wait after each, 1: four loads in flight per wait, 2: stores
__global__ void __launch_bounds__(512, 1) probe(long long* out, int iters, int active_warps) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < active_warps) {
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = (it * X) & 255;
      if (MODE == 0) {
        acc ^= ld_once<X>(tmem + col);
      } else if (MODE == 1) {
        uint32_t v[4][16];
#pragma unroll
        for (int k = 0; k < 4; ++k) tmem_ld_32x16(tmem + ((col + 16 * k) & 255), v[k]);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int i = 0; i < 16; ++i) acc ^= v[k][i];
      } else {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = acc + i;
        tmem_st_32x32(tmem + (col & 255), v);
        tmem_st_wait();
        acc += it;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

template <int X, int MODE>
void run(const char* name, int warps) {
  long long* out; cudaMalloc(&out, 16);
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) probe<X, MODE><<<148, 512>>>(out, iters, warps);
  long long h[1]; cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  const double bytes = (double)iters * warps * 32 * 4 * (MODE == 1 ? 64 : X);
  printf("%-44s %2d warps: %7.1f B/clk per SM  (%.0f clk per warp instruction)  %s\n", name, warps, bytes / h[0],
         (double)h[0] / (iters * (MODE == 1 ? 4 : 1)), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}
int main() {
  for (int w : {1, 4, 8, 16}) run<32, 0>("tcgen05.ld 32x32b.x32, wait after each", w);
  for (int w : {1, 4, 8, 16}) run<16, 0>("tcgen05.ld 32x32b.x16, wait after each", w);
  for (int w : {4, 8, 16}) run<16, 1>("tcgen05.ld 32x32b.x16, 4 in flight", w);
  for (int w : {4, 8, 16}) run<32, 2>("tcgen05.st 32x32b.x32, wait after each", w);
  return 0;
}
