// Which (lane, column) does each register of tcgen05.ld.16x256b / 16x128b hold?  A [128 lanes x 64 columns] pattern
// value = lane * 1000 + column is written with the 32x32b shape (thread i of warp w = lane 32 w + i), then read back by warp 0
// (and warp 1) with the matrix-fragment shapes; the store shapes are checked the other way round.
//   nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -I plbert_b200/csrc -I include -o tmem_shape_probe tmem_shape_probe.cu ../../plbert_b200/csrc/abi.cu -lcuda
#include <cstdio>
#include "common.cuh"
using namespace plb;

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (warp * 32 + lane) * 1000 + half * 32 + i;
    tmem_st_32x32(tmem + half * 32, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 2) {
    // 16x256b.x2: 8 registers per thread, lanes [base, base + 16)
    for (int lo = 0; lo < 2; ++lo) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tmem + (static_cast<uint32_t>(lo * 16) << 16)));
      tmem_ld_wait();
      for (int i = 0; i < 8; ++i) out[((warp * 2 + lo) * 32 + lane) * 8 + i] = r[i];
    }
    // 16x128b.x2: 4 registers per thread
    for (int lo = 0; lo < 2; ++lo) {
      uint32_t r[4];
      asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                   : "r"(tmem + (static_cast<uint32_t>(lo * 16) << 16)));
      tmem_ld_wait();
      for (int i = 0; i < 4; ++i) out[1024 + ((warp * 2 + lo) * 32 + lane) * 4 + i] = r[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 64);
}

int main() {
  uint32_t* out; cudaMalloc(&out, 2048 * 4); cudaMemset(out, 0xff, 2048 * 4);
  probe<<<1, 128>>>(out);
  uint32_t h[2048]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  for (int w = 0; w < 2; ++w)
    for (int lo = 0; lo < 2; ++lo) {
      printf("16x256b.x2 warp %d lane-base %d: thread -> (lane:col) x 8\n", w, lo * 16);
      for (int t = 0; t < 32; ++t) {
        printf("  t%2d:", t);
        for (int i = 0; i < 8; ++i) { uint32_t x = h[((w * 2 + lo) * 32 + t) * 8 + i]; printf(" %3u:%-2u", x / 1000, x % 1000); }
        printf("\n");
      }
    }
  for (int w = 0; w < 1; ++w)
    for (int lo = 0; lo < 2; ++lo) {
      printf("16x128b.x2 warp %d lane-base %d: thread -> (lane:col) x 4\n", w, lo * 16);
      for (int t = 0; t < 32; ++t) {
        printf("  t%2d:", t);
        for (int i = 0; i < 4; ++i) { uint32_t x = h[1024 + ((w * 2 + lo) * 32 + t) * 4 + i]; printf(" %3u:%-2u", x / 1000, x % 1000); }
        printf("\n");
      }
    }
  return 0;
}
