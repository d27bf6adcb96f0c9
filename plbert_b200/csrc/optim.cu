// plbert_b200 — fused AdamW over the flat fp32 parameter arena, plus the fp32 -> bf16 shadow cast.
//
// Replaces torch.optim.AdamW as constructed at R/train.py:272 (one param group: lr, betas (0.9, 0.999), eps 1e-8,
// weight_decay 0.01 on every tensor) — update rule torch/optim/adam.py:417-419,457,476,531-547:
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// One pass: reads p, g, m, v and writes p, m, v (28 B/param) + the bf16 shadow the next forward's GEMMs read
// (2 B/param) and, optionally, zeroes g for the next step's in-place gradient accumulation (4 B/param).
#include "common.cuh"

namespace plb {
namespace {

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ shadow, long n, float beta1, float beta2, float one_minus_beta1,
             float one_minus_beta2, float eps, float decay, float step_size, float inv_bc2_sqrt, float grad_scale,
             int zero_grad) {
  pdl_grid_sync();
  const long n4 = n >> 2;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    float* ga = reinterpret_cast<float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * grad_scale;
      pa[k] *= decay;
      ma[k] = ma[k] * beta1 + gk * one_minus_beta1;
      va[k] = va[k] * beta2 + gk * gk * one_minus_beta2;
      const float denom = sqrtf(va[k]) * inv_bc2_sqrt + eps;
      pa[k] -= step_size * (ma[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) {
      uint2 o;
      o.x = pack_bf16x2(pa[0], pa[1]);
      o.y = pack_bf16x2(pa[2], pa[3]);
      reinterpret_cast<uint2*>(shadow)[i] = o;
    }
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long n) {
  const long n4 = n >> 2;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 s = reinterpret_cast<const float4*>(src)[i];
    uint2 o;
    o.x = pack_bf16x2(s.x, s.y);
    o.y = pack_bf16x2(s.z, s.w);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
}

int flat_grid(long n4) {
  const long need = (n4 + 255) / 256;
  const long cap = static_cast<long>(sm_count()) * 8;
  return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace
}  // namespace plb

using namespace plb;

extern "C" int plb_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* bf16_shadow,
                              long n, int step, double lr, double beta1, double beta2, double eps, double weight_decay,
                              double grad_scale, int zero_grad, plb_stream_t stream) {
  PLB_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, PLB_ERR_ARG, "adamw: bad arguments");
  PLB_REQUIRE(n % 4 == 0, PLB_ERR_ARG, "adamw: arena length must be a multiple of 4 (pad the arena)");
  PLB_REQUIRE((reinterpret_cast<uintptr_t>(params) % 16) == 0 && (reinterpret_cast<uintptr_t>(grads) % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(exp_avg) % 16) == 0 && (reinterpret_cast<uintptr_t>(exp_avg_sq) % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(bf16_shadow) % 8) == 0,
              PLB_ERR_ARG, "adamw: arenas must be 16-byte aligned");
  // bias corrections in double on the host, exactly like torch's python-scalar path (adam.py:531-547)
  // every derived scalar is formed in double from the caller's doubles and rounded once, like torch's python scalars
  const double bc1 = 1.0 - pow(beta1, step);
  const double bc2 = 1.0 - pow(beta2, step);
  const float step_size = static_cast<float>(lr / bc1);
  const float inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
  const float decay = static_cast<float>(1.0 - lr * weight_decay);
  cudaError_t e = launch_kernel(adamw_kernel, dim3(flat_grid(n / 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), params, grads,
                                exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(bf16_shadow), n, static_cast<float>(beta1),
                                static_cast<float>(beta2), static_cast<float>(1.0 - beta1), static_cast<float>(1.0 - beta2),
                                static_cast<float>(eps), decay, step_size, inv_bc2_sqrt, static_cast<float>(grad_scale), zero_grad);
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "plb_adamw_step: launch failed: %s", cudaGetErrorString(e));
  return check_launch("plb_adamw_step");
}

extern "C" int plb_cast_bf16(const float* src, void* dst, long n, plb_stream_t stream) {
  PLB_REQUIRE(src && dst && n > 0 && n % 4 == 0, PLB_ERR_ARG, "cast_bf16: bad arguments (n must be a multiple of 4)");
  PLB_REQUIRE((reinterpret_cast<uintptr_t>(src) % 16) == 0 && (reinterpret_cast<uintptr_t>(dst) % 8) == 0, PLB_ERR_ARG,
              "cast_bf16: misaligned");
  cast_bf16_kernel<<<flat_grid(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), n);
  return check_launch("plb_cast_bf16");
}
