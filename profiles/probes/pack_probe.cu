// Micro-benchmark: does the fp32 -> bf16x2 pack (F2FP) share the MUFU pipe?  Per element: FFMA + MUFU.EX2 + FADD, per pair one
// pack, done three ways.  nvcc -arch=sm_100a -O3 -o pack_probe pack_probe.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack_cvt(float lo, float hi) {
  unsigned r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
__device__ __forceinline__ unsigned pack_trunc(float lo, float hi) { return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632); }
__device__ __forceinline__ unsigned pack_round(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
template <int MODE>
__global__ void probe(unsigned* out, long long* clk, float a, float b, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i * 0.01f;
  float s = 0.f;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float v0 = fmaf(x[i], a, b), v1 = fmaf(x[i + 1], a, b);
      if (MODE != 4) { v0 = ex2(v0); v1 = ex2(v1); }
      s += v0 + v1;
      if (MODE == 1 || MODE == 4) acc ^= pack_cvt(v0, v1);
      if (MODE == 2) acc ^= pack_trunc(v0, v1);
      if (MODE == 3) acc ^= pack_round(v0, v1);
      x[i] = v0; x[i + 1] = v1;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(s + x[3]);
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE>
void run(const char* name, int threads) {
  unsigned* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 1024;
  probe<MODE><<<148, threads>>>(out, clk, -0.5f, -0.1f, iters);
  probe<MODE><<<148, threads>>>(out, clk, -0.5f, -0.1f, iters);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-34s warps/SMSP %d : %.2f clk per element per SMSP\n", name, threads / 128, (double)h / (iters * 16) / (threads / 128));
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int threads : {128, 256}) {
    run<0>("exp2 only", threads);
    run<1>("exp2 + cvt.rn.bf16x2 pack", threads);
    run<2>("exp2 + PRMT truncating pack", threads);
    run<3>("exp2 + IADD round + PRMT pack", threads);
    run<4>("cvt.rn.bf16x2 pack only", threads);
  }
  return 0;
}
