// plbert_b200 — masked-row loss path and small reductions.
//
//   gather_rows / scatter_add_rows : only the rows named by masked_indices enter the heads
//                                    (R/train.py:125 `_s2s_pred[:len][masked_indices]`), so logits for the other
//                                    ~86 % of positions are never computed (R/model.py:15-16 computes all).
//   masked_ce                      : row-wise log-softmax cross-entropy with the per-row weight
//                                    1 / (count * n_b) of R/train.py:123-129 (mean of per-sample means); writes the
//                                    logits gradient in place.  nn.CrossEntropyLoss at R/train.py:215.
//   colsum                         : bias gradients (sum over rows of dY) for the Linear layers.
#include "common.cuh"

namespace plb {
namespace {

// dst[i, :] = src[rows[i], :]   (bf16, H % 8 == 0)
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, int ld_src, const int32_t* __restrict__ rows,
                                   int M, int H, __nv_bfloat16* __restrict__ dst, int ld_dst) {
  const int chunks = H >> 3;
  const long total = static_cast<long>(M) * chunks;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / chunks), c = static_cast<int>(i % chunks);
    const uint4 v = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(rows[m]) * ld_src + c * 8);
    *reinterpret_cast<uint4*>(dst + static_cast<size_t>(m) * ld_dst + c * 8) = v;
  }
}

// dst[rows[i], :] += src[i, :]   (bf16x2 atomics: duplicate rows accumulate like autograd's index backward)
__global__ void scatter_add_rows_kernel(const __nv_bfloat16* __restrict__ src, int ld_src,
                                        const int32_t* __restrict__ rows, int M, int H, __nv_bfloat16* __restrict__ dst,
                                        int ld_dst) {
  const int pairs = H >> 1;
  const long total = static_cast<long>(M) * pairs;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / pairs), c = static_cast<int>(i % pairs);
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + static_cast<size_t>(m) * ld_src + c * 2);
    atomicAdd(reinterpret_cast<__nv_bfloat162*>(dst + static_cast<size_t>(rows[m]) * ld_dst + c * 2), v);
  }
}

// One group of kThreads threads per row.  row_loss[m] = w * (lse - z[y]); logits <- w * (softmax - onehot)
template <int kThreads>
__global__ void masked_ce_kernel(__nv_bfloat16* __restrict__ logits, int ld, int M, int V,
                                 const int32_t* __restrict__ labels, const float* __restrict__ weights,
                                 float* __restrict__ row_loss) {
  constexpr int kGroups = 256 / kThreads;
  __shared__ float s_red[kGroups][kThreads / 32 + 1];
  const int group = threadIdx.x / kThreads, tid = threadIdx.x % kThreads;
  const int m = blockIdx.x * kGroups + group;
  const bool active = m < M;
  __nv_bfloat16* z = logits + static_cast<size_t>(active ? m : 0) * ld;

  auto group_reduce = [&](float v, bool is_max) -> float {
    v = is_max ? warp_max(v) : warp_sum(v);
    if (kThreads == 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) s_red[group][tid >> 5] = v;
    __syncthreads();
    float r = s_red[group][0];
    for (int i = 1; i < kThreads / 32; ++i) r = is_max ? fmaxf(r, s_red[group][i]) : r + s_red[group][i];
    return r;
  };

  float mx = -INFINITY;
  if (active)
    for (int j = tid; j < V; j += kThreads) mx = fmaxf(mx, __bfloat162float(z[j]));
  mx = group_reduce(mx, true);
  float se = 0.f;
  if (active)
    for (int j = tid; j < V; j += kThreads) se += __expf(__bfloat162float(z[j]) - mx);
  se = group_reduce(se, false);
  if (!active) return;
  const int y = labels[m];
  const bool y_ok = y >= 0 && y < V;
  // a label outside [0, V) contributes neither loss nor gradient (the whole row is ignored)
  const float w = y_ok ? weights[m] : 0.f;
  const float lse = __logf(se) + mx;
  if (!y_ok && tid == 0) row_loss[m] = 0.f;
  const float inv = w / se;
  for (int j = tid; j < ld; j += kThreads) {
    float g = 0.f;
    if (j < V) {
      const float zj = __bfloat162float(z[j]);
      g = __expf(zj - mx) * inv;
      if (j == y) {  // the thread that owns the label column reads z[y] before it overwrites it: no cross-thread ordering
        row_loss[m] = w * (lse - zj);
        g -= w;
      }
    }
    z[j] = __float2bfloat16(g);  // padding columns [V, ld) are zeroed so the dgrad/wgrad GEMMs may read them
  }
}

// Large vocabularies (the 64 k token head): one CTA per row, 16-byte accesses.  Pass 1 streams the row from HBM, passes 2
// and 3 re-read it from L1/L2 (a 64 k row is 128 KB), so the traffic is one read and one write of the logits.
__global__ void __launch_bounds__(256)
masked_ce_wide_kernel(__nv_bfloat16* __restrict__ logits, int ld, int M, int V, const int32_t* __restrict__ labels,
                      const float* __restrict__ weights, float* __restrict__ row_loss) {
  __shared__ float s_red[8];
  const int m = blockIdx.x, tid = threadIdx.x;
  __nv_bfloat16* z = logits + static_cast<size_t>(m) * ld;
  const int nvec = V >> 3;  // full 8-element chunks; the tail (V % 8) and the padding columns are handled scalar
  auto block_reduce = [&](float v, bool is_max) -> float {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    float r = s_red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, s_red[i]) : r + s_red[i];
    return r;
  };
  float mx = -INFINITY;
  for (int c = tid; c < nvec; c += 256) {
    const uint4 v = *reinterpret_cast<const uint4*>(z + c * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) mx = fmaxf(mx, fmaxf(bf16_lo(w[i]), bf16_hi(w[i])));
  }
  for (int j = nvec * 8 + tid; j < V; j += 256) mx = fmaxf(mx, __bfloat162float(z[j]));
  mx = block_reduce(mx, true);
  const float mx2 = mx * 1.4426950408889634f;
  float se = 0.f;
  for (int c = tid; c < nvec; c += 256) {
    const uint4 v = *reinterpret_cast<const uint4*>(z + c * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      se += ex2_fast(fmaf(bf16_lo(w[i]), 1.4426950408889634f, -mx2)) + ex2_fast(fmaf(bf16_hi(w[i]), 1.4426950408889634f, -mx2));
  }
  for (int j = nvec * 8 + tid; j < V; j += 256) se += ex2_fast(fmaf(__bfloat162float(z[j]), 1.4426950408889634f, -mx2));
  se = block_reduce(se, false);
  const int y = labels[m];
  const bool y_ok = y >= 0 && y < V;
  const float wgt = y_ok ? weights[m] : 0.f;  // a label outside [0, V): no loss and no gradient for the row
  if (tid == 0) row_loss[m] = y_ok ? wgt * (__logf(se) + mx - __bfloat162float(z[y])) : 0.f;
  __syncthreads();  // z[y] has been read before anyone overwrites it
  const float inv = wgt / se;
  for (int c = tid; c < nvec; c += 256) {
    const uint4 v = *reinterpret_cast<const uint4*>(z + c * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float g[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      g[2 * i] = ex2_fast(fmaf(bf16_lo(w[i]), 1.4426950408889634f, -mx2)) * inv;
      g[2 * i + 1] = ex2_fast(fmaf(bf16_hi(w[i]), 1.4426950408889634f, -mx2)) * inv;
    }
    if (y >= c * 8 && y < c * 8 + 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c * 8 + i == y) g[i] -= wgt;
    }
    uint4 o;
    o.x = pack_bf16x2(g[0], g[1]); o.y = pack_bf16x2(g[2], g[3]);
    o.z = pack_bf16x2(g[4], g[5]); o.w = pack_bf16x2(g[6], g[7]);
    *reinterpret_cast<uint4*>(z + c * 8) = o;
  }
  for (int j = nvec * 8 + tid; j < ld; j += 256) {
    float g = 0.f;
    if (j < V) {
      g = ex2_fast(fmaf(__bfloat162float(z[j]), 1.4426950408889634f, -mx2)) * inv;
      if (j == y) g -= wgt;
    }
    z[j] = __float2bfloat16(g);
  }
}

// deterministic sum of row_loss -> out[0] (+= when accumulate)
__global__ void sum_rows_kernel(const float* __restrict__ v, int n, float* __restrict__ out, int accumulate) {
  __shared__ float s[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += v[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + r : r;
  }
}

// out[n] += sum_t x[t, n]   (bf16 in, fp32 accumulate; one 16-byte chunk per lane)
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, int ld, int T, int N,
                                                     float* __restrict__ out) {
  __shared__ float s[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    const int stride = gridDim.y * 8;
    int t = blockIdx.y * 8 + warp;
    // four independent 16-byte loads in flight per lane
    for (; t + 3 * stride < T; t += 4 * stride) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(t + u * stride) * ld + col);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] += bf16_lo(w[i]);
          acc[2 * i + 1] += bf16_hi(w[i]);
        }
      }
    }
    for (; t < T; t += stride) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(t) * ld + col);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += bf16_lo(w[i]);
        acc[2 * i + 1] += bf16_hi(w[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float r = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) r += s[wv][threadIdx.x];
    atomicAdd(out + c, r);
  }
}

}  // namespace
}  // namespace plb

using namespace plb;

extern "C" int plb_gather_rows(const void* src, int ld_src, const int32_t* rows, int M, int H, void* dst, int ld_dst,
                               plb_stream_t stream) {
  PLB_REQUIRE(src && rows && dst && M >= 0 && H > 0, PLB_ERR_ARG, "gather_rows: bad arguments");
  PLB_REQUIRE(H % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0, PLB_ERR_ARG, "gather_rows: H and lds must be multiples of 8");
  if (M == 0) return PLB_OK;
  const long total = static_cast<long>(M) * (H / 8);
  const int grid = static_cast<int>(total / 256 + 1 < sm_count() * 8L ? total / 256 + 1 : sm_count() * 8L);
  gather_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(src), ld_src, rows, M,
                                                                         H, static_cast<__nv_bfloat16*>(dst), ld_dst);
  return check_launch("plb_gather_rows");
}

extern "C" int plb_scatter_add_rows(const void* src, int ld_src, const int32_t* rows, int M, int H, void* dst, int ld_dst,
                                    plb_stream_t stream) {
  PLB_REQUIRE(src && rows && dst && M >= 0 && H > 0, PLB_ERR_ARG, "scatter_add_rows: bad arguments");
  PLB_REQUIRE(H % 2 == 0 && ld_src % 2 == 0 && ld_dst % 2 == 0, PLB_ERR_ARG, "scatter_add_rows: H and lds must be even");
  if (M == 0) return PLB_OK;
  const long total = static_cast<long>(M) * (H / 2);
  const int grid = static_cast<int>(total / 256 + 1 < sm_count() * 8L ? total / 256 + 1 : sm_count() * 8L);
  scatter_add_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(src), ld_src,
                                                                              rows, M, H, static_cast<__nv_bfloat16*>(dst), ld_dst);
  return check_launch("plb_scatter_add_rows");
}

extern "C" int plb_masked_ce(void* logits, int ld, int M, int V, const int32_t* labels, const float* weights,
                             float* row_loss, float* loss_out, int accumulate, plb_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLB_REQUIRE(logits && labels && weights && row_loss && loss_out && M >= 0 && V > 0 && ld >= V, PLB_ERR_ARG,
              "masked_ce: bad arguments");
  if (M == 0) {
    if (!accumulate) {
      cudaError_t e = cudaMemsetAsync(loss_out, 0, sizeof(float), stream);  // R/train.py:129: empty -> 0
      PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "masked_ce: memset failed: %s", cudaGetErrorString(e));
    }
    return PLB_OK;
  }
  __nv_bfloat16* z = static_cast<__nv_bfloat16*>(logits);
  if (V <= 2048)
    masked_ce_kernel<32><<<(M + 7) / 8, 256, 0, stream>>>(z, ld, M, V, labels, weights, row_loss);
  else
    masked_ce_wide_kernel<<<M, 256, 0, stream>>>(z, ld, M, V, labels, weights, row_loss);
  int rc = check_launch("plb_masked_ce");
  if (rc) return rc;
  sum_rows_kernel<<<1, 1024, 0, stream>>>(row_loss, M, loss_out, accumulate);
  return check_launch("plb_masked_ce(sum)");
}

extern "C" int plb_colsum(const void* x, int ld, int T, int N, float* out_accum, plb_stream_t stream) {
  PLB_REQUIRE(x && out_accum && T > 0 && N > 0, PLB_ERR_ARG, "colsum: bad arguments");
  PLB_REQUIRE(ld % 8 == 0 && ld >= (N + 7) / 8 * 8 && (reinterpret_cast<uintptr_t>(x) % 16) == 0, PLB_ERR_ARG,
              "colsum: ld must be a multiple of 8 covering N rounded up to 8, x 16-byte aligned");
  const int gx = (N + 255) / 256;
  int gy = (8 * sm_count() + gx - 1) / gx;
  if (gy > (T + 7) / 8) gy = (T + 7) / 8;
  colsum_kernel<<<dim3(gx, gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), ld, T, N,
                                                                            out_accum);
  return check_launch("plb_colsum");
}
