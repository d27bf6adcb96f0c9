// plbert_b200 — HBM-bound row kernels: LayerNorm forward/backward over bf16 activations and the fused
// embedding gather + LayerNorm (forward) / LayerNorm-backward + scatter-add (backward).
//
// One warp per row, 16-byte vector accesses (lane i owns 8-element chunks i, i+32, ...), fp32 statistics with a
// two-pass variance, warp-shuffle reductions.  Replaces torch.nn.LayerNorm in HF modeling_albert.py:108
// (embeddings), :202 (attention.LayerNorm) and :235 (full_layer_layer_norm) and nn.Embedding x3 at :101-106.
// Backward formulas: SURVEY.md appendix A.
#include "common.cuh"

namespace plb {
namespace {

constexpr int WARPS = 8;
constexpr int LN_MAXC = 4;  // up to 4 chunks of 8 elements per lane => row width <= 1024

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(w[i]);
    f[2 * i + 1] = bf16_hi(w[i]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// y = LN(z) * w + b ; stats[row] = (mean, rstd)
template <int MAXC>
__global__ void __launch_bounds__(WARPS * 32) ln_fwd_kernel(const __nv_bfloat16* __restrict__ z, int T, int H,
                                                            const float* __restrict__ w, const float* __restrict__ b,
                                                            float eps, __nv_bfloat16* __restrict__ y,
                                                            float* __restrict__ stats) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunk = H >> 3;
  for (int row = blockIdx.x * WARPS + warp; row < T; row += gridDim.x * WARPS) {
    const uint4* zr = reinterpret_cast<const uint4*>(z + static_cast<size_t>(row) * H);
    float x[MAXC][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunk) {
        unpack8(zr[c], x[i]);
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += x[i][k];
      }
    }
    const float mean = warp_sum(sum) / H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i)
      if (lane + 32 * i < nchunk) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = x[i][k] - mean;
          sq += d * d;
        }
      }
    const float rstd = rsqrtf(warp_sum(sq) / H + eps);
    uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * H);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunk) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w + c * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + c * 8 + 4));
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (x[i][k] - mean) * rstd * ww[k] + bb[k];
        yr[c] = pack8(o);
      }
    }
    if (lane == 0) {
      stats[2 * row] = mean;
      stats[2 * row + 1] = rstd;
    }
  }
}

// dz = LN'(dy); dw += sum dy*xhat; db += sum dy; colsum += sum dz (bias gradient of the producing Linear)
// A row is split over WPR warps (each lane owns ONE 16-byte chunk), so the per-thread column partials are only
// 3 x 8 registers and two 12-warp CTAs fit per SM; the two row statistics cross the WPR warps through shared memory.
constexpr int LNB_ROWS = 4;  // rows in flight per CTA
template <int WPR>
__global__ void __launch_bounds__(LNB_ROWS * WPR * 32, (LNB_ROWS * WPR * 32 <= 384) ? 2 : 1)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
              const float* __restrict__ stats, const float* __restrict__ w, int T, int H,
              __nv_bfloat16* __restrict__ dz, float* __restrict__ dw_acc, float* __restrict__ db_acc,
              float* __restrict__ colsum_acc) {
  extern __shared__ float s_red[];                      // [3][H] column partials of the CTA
  __shared__ float s_row[2][LNB_ROWS][WPR][2];          // per-row partial statistics (double buffered)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rslot = warp / WPR, wr = warp % WPR;        // row slot of this warp, position within the row
  const int c = wr * 32 + lane;                         // 16-byte chunk of the row owned by this thread
  const int nchunk = H >> 3;
  const bool active = c < nchunk;
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();

  float ww[8], pdw[8], pdb[8], pcs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    ww[k] = active ? __ldg(w + c * 8 + k) : 0.f;
    pdw[k] = pdb[k] = pcs[k] = 0.f;
  }
  const float invH = 1.f / H;
  int it = 0;
  // software pipeline: the loads of the next row are in flight while the current row is reduced
  uint4 nz = make_uint4(0, 0, 0, 0), nd = make_uint4(0, 0, 0, 0);
  float nmean = 0.f, nrstd = 0.f;
  auto prefetch = [&](int row0) {
    const int row = row0 + rslot;
    if (row < T) {
      nmean = stats[2 * row];
      nrstd = stats[2 * row + 1];
      if (active) {
        nz = *reinterpret_cast<const uint4*>(z + static_cast<size_t>(row) * H + c * 8);
        nd = *reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * H + c * 8);
      }
    }
  };
  prefetch(blockIdx.x * LNB_ROWS);
  for (int row0 = blockIdx.x * LNB_ROWS; row0 < T; row0 += gridDim.x * LNB_ROWS, ++it) {
    const int row = row0 + rslot;
    const bool rv = row < T;
    float xh[8], g[8];
    float s1 = 0.f, s2 = 0.f;
    const float mean = nmean, rstd = nrstd;
    const uint4 cz = nz, cd = nd;
    if (row0 + gridDim.x * LNB_ROWS < T) prefetch(row0 + gridDim.x * LNB_ROWS);
    if (rv && active) {
      float zz[8], dd[8];
      unpack8(cz, zz);
      unpack8(cd, dd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xh[k] = (zz[k] - mean) * rstd;
        pdw[k] += dd[k] * xh[k];
        pdb[k] += dd[k];
        g[k] = dd[k] * ww[k];
        s1 += g[k];
        s2 += g[k] * xh[k];
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (WPR > 1) {
      if (lane == 0) {
        s_row[it & 1][rslot][wr][0] = s1;
        s_row[it & 1][rslot][wr][1] = s2;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + rslot), "n"(WPR * 32) : "memory");
      s1 = 0.f;
      s2 = 0.f;
#pragma unroll
      for (int q = 0; q < WPR; ++q) {
        s1 += s_row[it & 1][rslot][q][0];
        s2 += s_row[it & 1][rslot][q][1];
      }
    }
    s1 *= invH;
    s2 *= invH;
    if (rv && active) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rstd * (g[k] - s1 - xh[k] * s2);
      const uint4 packed = pack8(o);
      *reinterpret_cast<uint4*>(dz + static_cast<size_t>(row) * H + c * 8) = packed;
      // the bias gradient must see exactly the values the GEMMs will read -> sum the rounded outputs
      float r[8];
      unpack8(packed, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) pcs[k] += r[k];
    }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&s_red[c * 8 + k], pdw[k]);
      atomicAdd(&s_red[H + c * 8 + k], pdb[k]);
      atomicAdd(&s_red[2 * H + c * 8 + k], pcs[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(dw_acc + i, s_red[i]);
    atomicAdd(db_acc + i, s_red[H + i]);
    if (colsum_acc) atomicAdd(colsum_acc + i, s_red[2 * H + i]);
  }
}

// e = LN(word[id] + type[0] + pos[p]) -> bf16, stats   (chunks of 4 floats; E <= 512)
constexpr int EMAXC = 4;
__global__ void __launch_bounds__(WARPS * 32)
embed_ln_fwd_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ pos, int T, int E,
                    const float* __restrict__ word, const float* __restrict__ posemb, const float* __restrict__ type0,
                    const float* __restrict__ w, const float* __restrict__ b, float eps, __nv_bfloat16* __restrict__ out,
                    float* __restrict__ stats) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunk = E >> 2;
  for (int row = blockIdx.x * WARPS + warp; row < T; row += gridDim.x * WARPS) {
    const float4* wr = reinterpret_cast<const float4*>(word + static_cast<size_t>(ids[row]) * E);
    const float4* pr = reinterpret_cast<const float4*>(posemb + static_cast<size_t>(pos[row]) * E);
    const float4* tr = reinterpret_cast<const float4*>(type0);
    float x[EMAXC][4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < EMAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunk) {
        const float4 a = __ldg(wr + c), t4 = __ldg(tr + c), p4 = __ldg(pr + c);
        x[i][0] = (a.x + t4.x) + p4.x; x[i][1] = (a.y + t4.y) + p4.y;
        x[i][2] = (a.z + t4.z) + p4.z; x[i][3] = (a.w + t4.w) + p4.w;
        sum += x[i][0] + x[i][1] + x[i][2] + x[i][3];
      }
    }
    const float mean = warp_sum(sum) / E;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < EMAXC; ++i)
      if (lane + 32 * i < nchunk) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = x[i][k] - mean;
          sq += d * d;
        }
      }
    const float rstd = rsqrtf(warp_sum(sq) / E + eps);
#pragma unroll
    for (int i = 0; i < EMAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunk) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + c), b4 = __ldg(reinterpret_cast<const float4*>(b) + c);
        uint2 o;
        o.x = pack_bf16x2((x[i][0] - mean) * rstd * w4.x + b4.x, (x[i][1] - mean) * rstd * w4.y + b4.y);
        o.y = pack_bf16x2((x[i][2] - mean) * rstd * w4.z + b4.z, (x[i][3] - mean) * rstd * w4.w + b4.w);
        *reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * E + c * 4) = o;
      }
    }
    if (lane == 0) {
      stats[2 * row] = mean;
      stats[2 * row + 1] = rstd;
    }
  }
}

// de (bf16 grad of the LN output) -> LN backward -> scatter-add into the three tables.
// kSmemWord: the word-table gradient (V x E floats; 96 KB for the 188 x 128 table) is accumulated in shared memory per
// CTA and flushed once, because a handful of rows (mask, separator) receive ~30 % of all tokens and would otherwise
// serialise in the L2 atomic units.
template <bool kSmemWord>
__global__ void __launch_bounds__(WARPS * 32)
embed_ln_bwd_kernel(const __nv_bfloat16* __restrict__ de, const int32_t* __restrict__ ids,
                    const int32_t* __restrict__ pos, int T, int E, const float* __restrict__ word,
                    const float* __restrict__ posemb, const float* __restrict__ type0, const float* __restrict__ w,
                    const float* __restrict__ stats, int pad_id, int V, float* __restrict__ dword,
                    float* __restrict__ dpos, float* __restrict__ dtype0, float* __restrict__ dw_acc,
                    float* __restrict__ db_acc) {
  extern __shared__ float s_red[];  // [3][E]: dw, db, dtype0; then (kSmemWord) [V][E] word-table gradient
  float* s_word = s_red + 3 * E;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunk = E >> 2;
  for (int i = threadIdx.x; i < 3 * E + (kSmemWord ? V * E : 0); i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  float ww[EMAXC][4], pdw[EMAXC][4], pdb[EMAXC][4], pdt[EMAXC][4];
#pragma unroll
  for (int i = 0; i < EMAXC; ++i) {
    const int c = lane + 32 * i;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ww[i][k] = (c < nchunk) ? __ldg(w + c * 4 + k) : 0.f;
      pdw[i][k] = pdb[i][k] = pdt[i][k] = 0.f;
    }
  }
  const float invE = 1.f / E;
  for (int row = blockIdx.x * WARPS + warp; row < T; row += gridDim.x * WARPS) {
    const int id = ids[row], ps = pos[row];
    const float4* wr = reinterpret_cast<const float4*>(word + static_cast<size_t>(id) * E);
    const float4* pr = reinterpret_cast<const float4*>(posemb + static_cast<size_t>(ps) * E);
    const float4* tr = reinterpret_cast<const float4*>(type0);
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float xh[EMAXC][4], g[EMAXC][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < EMAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunk) {
        const float4 a = __ldg(wr + c), t4 = __ldg(tr + c), p4 = __ldg(pr + c);
        const float x[4] = {(a.x + t4.x) + p4.x, (a.y + t4.y) + p4.y, (a.z + t4.z) + p4.z, (a.w + t4.w) + p4.w};
        const uint2 dv = *reinterpret_cast<const uint2*>(de + static_cast<size_t>(row) * E + c * 4);
        const float dd[4] = {bf16_lo(dv.x), bf16_hi(dv.x), bf16_lo(dv.y), bf16_hi(dv.y)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          xh[i][k] = (x[k] - mean) * rstd;
          pdw[i][k] += dd[k] * xh[i][k];
          pdb[i][k] += dd[k];
          g[i][k] = dd[k] * ww[i][k];
          s1 += g[i][k];
          s2 += g[i][k] * xh[i][k];
        }
      }
    }
    s1 = warp_sum(s1) * invE;
    s2 = warp_sum(s2) * invE;
#pragma unroll
    for (int i = 0; i < EMAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunk) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = rstd * (g[i][k] - s1 - xh[i][k] * s2);
          pdt[i][k] += d;
          if (id != pad_id) {  // padding_idx row gets no grad
            if (kSmemWord) atomicAdd(s_word + id * E + c * 4 + k, d);
            else atomicAdd(dword + static_cast<size_t>(id) * E + c * 4 + k, d);
          }
          atomicAdd(dpos + static_cast<size_t>(ps) * E + c * 4 + k, d);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < EMAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunk) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(&s_red[c * 4 + k], pdw[i][k]);
        atomicAdd(&s_red[E + c * 4 + k], pdb[i][k]);
        atomicAdd(&s_red[2 * E + c * 4 + k], pdt[i][k]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    atomicAdd(dw_acc + i, s_red[i]);
    atomicAdd(db_acc + i, s_red[E + i]);
    atomicAdd(dtype0 + i, s_red[2 * E + i]);
  }
  if (kSmemWord) {
    for (int i = threadIdx.x; i < V * E; i += blockDim.x) {
      const float v = s_word[i];
      if (v != 0.f) atomicAdd(dword + i, v);
    }
  }
}

int row_grid(int T, int blocks_per_sm) {
  const int need = (T + WARPS - 1) / WARPS;
  const int cap = sm_count() * blocks_per_sm;
  return need < cap ? need : cap;
}

}  // namespace
}  // namespace plb

using namespace plb;

extern "C" int plb_ln_fwd(const void* z, int T, int H, const float* w, const float* b, float eps, void* y, float* stats,
                          plb_stream_t stream) {
  PLB_REQUIRE(z && w && b && y && stats && T > 0, PLB_ERR_ARG, "ln_fwd: bad arguments");
  PLB_REQUIRE(H % 8 == 0 && H >= 8 && H <= 8 * 32 * LN_MAXC, PLB_ERR_UNSUPPORTED, "ln_fwd: H=%d must be a multiple of 8, <= %d", H,
              8 * 32 * LN_MAXC);
  const int nc = (H / 8 + 31) / 32;
  auto kern = nc == 1 ? ln_fwd_kernel<1> : nc == 2 ? ln_fwd_kernel<2> : nc == 3 ? ln_fwd_kernel<3> : ln_fwd_kernel<4>;
  kern<<<row_grid(T, 8), WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(z), T, H, w, b, eps, static_cast<__nv_bfloat16*>(y), stats);
  return check_launch("plb_ln_fwd");
}

extern "C" int plb_ln_bwd(const void* dy, const void* z, const float* stats, const float* w, int T, int H, void* dz,
                          float* dw_accum, float* db_accum, float* colsum_accum, plb_stream_t stream) {
  PLB_REQUIRE(dy && z && stats && w && dz && dw_accum && db_accum && T > 0, PLB_ERR_ARG, "ln_bwd: bad arguments");
  PLB_REQUIRE(H % 8 == 0 && H >= 8 && H <= 8 * 32 * LN_MAXC, PLB_ERR_UNSUPPORTED, "ln_bwd: H=%d must be a multiple of 8, <= %d", H,
              8 * 32 * LN_MAXC);
  const int nc = (H / 8 + 31) / 32;
  auto kern = nc == 1 ? ln_bwd_kernel<1> : nc == 2 ? ln_bwd_kernel<2> : nc == 3 ? ln_bwd_kernel<3> : ln_bwd_kernel<4>;
  const int need = (T + LNB_ROWS - 1) / LNB_ROWS, cap = sm_count() * 4;
  kern<<<need < cap ? need : cap, LNB_ROWS * nc * 32, 3 * H * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(z), stats, w, T, H,
      static_cast<__nv_bfloat16*>(dz), dw_accum, db_accum, colsum_accum);
  return check_launch("plb_ln_bwd");
}

extern "C" int plb_embed_ln_fwd(const int32_t* ids, const int32_t* pos, int T, int E, const float* word,
                                const float* posemb, const float* type0, const float* ln_w, const float* ln_b,
                                float eps, void* out, float* stats, plb_stream_t stream) {
  PLB_REQUIRE(ids && pos && word && posemb && type0 && ln_w && ln_b && out && stats && T > 0, PLB_ERR_ARG,
              "embed_ln_fwd: bad arguments");
  PLB_REQUIRE(E % 4 == 0 && E >= 4 && E <= 4 * 32 * EMAXC, PLB_ERR_UNSUPPORTED, "embed_ln_fwd: E=%d unsupported", E);
  embed_ln_fwd_kernel<<<row_grid(T, 8), WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, pos, T, E, word, posemb, type0, ln_w, ln_b, eps, static_cast<__nv_bfloat16*>(out), stats);
  return check_launch("plb_embed_ln_fwd");
}

extern "C" int plb_embed_ln_bwd(const void* de, const int32_t* ids, const int32_t* pos, int T, int E, int V,
                                const float* word, const float* posemb, const float* type0, const float* ln_w,
                                const float* stats, int pad_id, float* dword, float* dpos, float* dtype0, float* dln_w,
                                float* dln_b, plb_stream_t stream) {
  PLB_REQUIRE(de && ids && pos && word && posemb && type0 && ln_w && stats && dword && dpos && dtype0 && dln_w && dln_b &&
                  T > 0,
              PLB_ERR_ARG, "embed_ln_bwd: bad arguments");
  PLB_REQUIRE(E % 4 == 0 && E >= 4 && E <= 4 * 32 * EMAXC, PLB_ERR_UNSUPPORTED, "embed_ln_bwd: E=%d unsupported", E);
  PLB_REQUIRE(V > 0, PLB_ERR_ARG, "embed_ln_bwd: V must be the number of rows of the word table");
  const size_t small = 3 * E * sizeof(float), big = small + static_cast<size_t>(V) * E * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (big <= 200 * 1024) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(embed_ln_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "embed_ln_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      configured = true;
    }
    const int need = (T + WARPS - 1) / WARPS;
    embed_ln_bwd_kernel<true><<<need < sm_count() ? need : sm_count(), WARPS * 32, big, st>>>(
        static_cast<const __nv_bfloat16*>(de), ids, pos, T, E, word, posemb, type0, ln_w, stats, pad_id, V, dword, dpos,
        dtype0, dln_w, dln_b);
  } else {
    embed_ln_bwd_kernel<false><<<row_grid(T, 4), WARPS * 32, small, st>>>(
        static_cast<const __nv_bfloat16*>(de), ids, pos, T, E, word, posemb, type0, ln_w, stats, pad_id, V, dword, dpos,
        dtype0, dln_w, dln_b);
  }
  return check_launch("plb_embed_ln_bwd");
}
