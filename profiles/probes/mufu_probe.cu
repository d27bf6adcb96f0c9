// Micro-benchmark: MUFU.EX2 issue interval per warp instruction on one SM sub-partition, alone and mixed with FMA-pipe work,
// and the cost of a Cody-Waite + cubic polynomial exp2 on the FMA pipe.  nvcc -arch=sm_100a -O3 -o mufu_probe mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.0555041087f, 0.2402265069f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
template <int MODE>
__global__ void probe(float* out, long long* clk, float a, float b, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i * 0.01f;
  float s = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float v = fmaf(x[i], a, b);
      if (MODE == 0) v = ex2(v);                               // all MUFU
      if (MODE == 1) v = (i & 3) == 3 ? ex2_poly(v) : ex2(v);  // 25 % polynomial
      if (MODE == 2) v = (i & 1) ? ex2_poly(v) : ex2(v);       // 50 % polynomial
      if (MODE == 3) v = ex2_poly(v);                          // all polynomial
      s += v;
      x[i] = v;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 1024;
  probe<MODE><<<148, threads>>>(out, clk, -0.5f, -0.1f, iters);
  probe<MODE><<<148, threads>>>(out, clk, -0.5f, -0.1f, iters);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SMSP %d : %.2f clk per exp2 warp-instruction per warp, %.2f per SMSP\n", name, threads / 128,
         (double)h / (iters * 16), (double)h / (iters * 16) / (threads / 128));
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("MUFU only", threads);
    run<1>("25% polynomial", threads);
    run<2>("50% polynomial", threads);
    run<3>("polynomial only", threads);
  }
  return 0;
}
