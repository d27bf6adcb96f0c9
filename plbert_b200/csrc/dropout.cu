// plbert_b200 — the dropout kernel (north star (c); HF modeling_albert.py:109 and :201, nn.Dropout(hidden_dropout_prob)):
//     y[r, c] = (residual ? residual[r, c] : 0) + x[r, c] * m(r * cols + c),   m = 0 or 1 / (1 - p)
// bf16 in / out (in place allowed), 16-byte vectors, mask from the counter-based generator of rng.cuh — so the same call
// with the same (seed, stream) is also the BACKWARD of the op (dx = dy * m) and no mask tensor exists.  The reference runs
// with p = 0 (its `dropout: 0.1` is not an AlbertConfig field); the engine only launches this for p > 0.
#include "common.cuh"
#include "rng.cuh"

namespace plb {
namespace {

__global__ void __launch_bounds__(256) dropout_kernel(const __nv_bfloat16* __restrict__ x, int ld_x,
                                                       const __nv_bfloat16* __restrict__ res, int ld_res,
                                                       __nv_bfloat16* __restrict__ y, int ld_y, int rows, int cols,
                                                       const DropoutParams d) {
  const int vec_per_row = cols >> 3;
  const long long total = static_cast<long long>(rows) * vec_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / vec_per_row), c = static_cast<int>(i - static_cast<long long>(r) * vec_per_row) * 8;
    const uint4 xv = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * ld_x + c);
    uint4 rv = make_uint4(0, 0, 0, 0);
    if (res != nullptr) rv = *reinterpret_cast<const uint4*>(res + static_cast<size_t>(r) * ld_res + c);
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
    uint32_t ow[4];
    const uint32_t pair0 = static_cast<uint32_t>((static_cast<long long>(r) * cols + c) >> 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float m0, m1;
      dropout_pair(d, pair0 + k, m0, m1);
      ow[k] = pack_bf16x2(fmaf(bf16_lo(xw[k]), m0, bf16_lo(rw[k])), fmaf(bf16_hi(xw[k]), m1, bf16_hi(rw[k])));
    }
    *reinterpret_cast<uint4*>(y + static_cast<size_t>(r) * ld_y + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

}  // namespace
}  // namespace plb

using namespace plb;

extern "C" int plb_dropout(const void* x, int ld_x, const void* residual, int ld_res, void* y, int ld_y, int rows, int cols,
                           float p, unsigned long long seed, unsigned int stream, plb_stream_t stream_) {
  PLB_REQUIRE(x && y, PLB_ERR_ARG, "dropout: null pointer");
  PLB_REQUIRE(rows > 0 && cols > 0 && (cols % 8) == 0 && (ld_x % 8) == 0 && (ld_y % 8) == 0 && (residual == nullptr || (ld_res % 8) == 0),
              PLB_ERR_ARG, "dropout: cols and leading dimensions must be multiples of 8");
  PLB_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)residual % 16) == 0, PLB_ERR_ARG,
              "dropout: operands must be 16-byte aligned");
  PLB_REQUIRE(p >= 0.f && p < 1.f, PLB_ERR_ARG, "dropout: p = %f outside [0, 1)", p);
  PLB_REQUIRE(static_cast<long long>(rows) * cols < (1ll << 33), PLB_ERR_UNSUPPORTED, "dropout: more than 2^33 elements");
  DropoutParams d = make_dropout(p, seed, stream);
  if (p == 0.f) d.thresh16 = 65536u;   // identity through the same code path (keep everything, scale 1)
  const long long vecs = static_cast<long long>(rows) * (cols / 8);
  const long long want = (vecs + 255) / 256;
  const int grid = static_cast<int>(want < sm_count() * 16ll ? want : sm_count() * 16ll);
  dropout_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(x), ld_x, static_cast<const __nv_bfloat16*>(residual), ld_res,
      static_cast<__nv_bfloat16*>(y), ld_y, rows, cols, d);
  return check_launch("plb_dropout");
}
