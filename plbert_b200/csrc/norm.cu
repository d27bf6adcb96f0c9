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
// Every warp keeps FWD_DEPTH rows in flight: lane 0 streams the warp's next rows into a private shared-memory ring with
// 1-D bulk copies (cp.async.bulk + mbarrier), so the bytes in flight per SM (~150 KB) no longer depend on registers or on
// how many warps are resident — the register-resident version waited out a full HBM round trip per row and reached 5 TB/s.
constexpr int FWD_DEPTH = 4;
template <int MAXC, bool kFull>   // kFull: H == MAXC * 256, every lane owns MAXC whole chunks (no per-chunk predicates)
__global__ void __launch_bounds__(WARPS * 32) ln_fwd_kernel(const __nv_bfloat16* __restrict__ z, int T, int H,
                                                            const float* __restrict__ w, const float* __restrict__ b,
                                                            float eps, __nv_bfloat16* __restrict__ y,
                                                            float* __restrict__ stats) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  pdl_grid_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunk = H >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(H) * 2;
  uint8_t* ring = ln_smem + static_cast<size_t>(warp) * FWD_DEPTH * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + static_cast<size_t>(WARPS) * FWD_DEPTH * row_bytes) + warp * FWD_DEPTH;
  const int first = blockIdx.x * WARPS + warp, stride = gridDim.x * WARPS;
  auto issue = [&](int k) {  // lane 0 only
    const long row = first + static_cast<long>(k) * stride;
    if (row < T) {
      const int st = k % FWD_DEPTH;
      mbar_arrive_expect_tx(&bars[st], row_bytes);
      bulk_load_1d(ring + st * row_bytes, z + row * H, row_bytes, &bars[st]);
    }
  };
  if (lane == 0) {
    for (int i = 0; i < FWD_DEPTH; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    for (int k = 0; k < FWD_DEPTH; ++k) issue(k);
  }
  __syncwarp();
  // weight / bias of this lane's columns stay in registers for the whole kernel (per-row __ldg's were the top stall:
  // with the ring there are few resident warps to hide their latency), as fp32 pairs for the packed FFMA2 math
  f32x2 ww[MAXC][4], bb[MAXC][4];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + 32 * i;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool ok = kFull || c < nchunk;
      ww[i][q] = f2_pack(ok ? __ldg(w + c * 8 + 2 * q) : 0.f, ok ? __ldg(w + c * 8 + 2 * q + 1) : 0.f);
      bb[i][q] = f2_pack(ok ? __ldg(b + c * 8 + 2 * q) : 0.f, ok ? __ldg(b + c * 8 + 2 * q + 1) : 0.f);
    }
  }
  const float invH = 1.f / H;
  int k = 0;
  for (int row = first; row < T; row += stride, ++k) {
    const int st = k % FWD_DEPTH;
    mbar_wait(&bars[st], (k / FWD_DEPTH) & 1);
    const uint32_t src = smem_u32(ring + st * row_bytes);
    f32x2 x[MAXC][4];
    f32x2 sum2 = f2_splat(0.f);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (kFull || c < nchunk) {
        const uint4 v = lds128(src + c * 16);
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          x[i][q] = f2_from_bf16x2(wv[q]);
          sum2 = f2_add(sum2, x[i][q]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) x[i][q] = f2_splat(0.f);
      }
    }
    __syncwarp();                    // every lane has its chunks in registers: the slot can be refilled
    if (lane == 0) issue(k + FWD_DEPTH);
    float s_lo, s_hi;
    f2_unpack(sum2, s_lo, s_hi);
    const float mean = warp_sum(s_lo + s_hi) * invH;
    const f32x2 nmean = f2_splat(-mean);
    f32x2 sq2 = f2_splat(0.f);
#pragma unroll
    for (int i = 0; i < MAXC; ++i)
      if (kFull || lane + 32 * i < nchunk) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const f32x2 d = f2_add(x[i][q], nmean);
          sq2 = f2_fma(d, d, sq2);
        }
      }
    f2_unpack(sq2, s_lo, s_hi);
    const float rstd = rsqrtf(warp_sum(s_lo + s_hi) * invH + eps);
    const f32x2 rstd2 = f2_splat(rstd), nmr = f2_splat(-mean * rstd);
    uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * H);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (kFull || c < nchunk) {
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)   // ((x - mean) rstd) w + b, the normalised value first like the reference
          o[q] = f2_to_bf16x2(f2_fma(f2_fma(x[i][q], rstd2, nmr), ww[i][q], bb[i][q]));
        yr[c] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    if (lane == 0) {
      stats[2 * row] = mean;
      stats[2 * row + 1] = rstd;
    }
  }
}

// dz = LN'(dy); dw += sum dy*xhat; db += sum dy; colsum += sum dz (bias gradient of the producing Linear)
// One warp per row, like the forward kernel: each warp streams its rows (z and dy) through a private BWD_DEPTH-deep
// shared-memory ring filled by 1-D bulk copies, so nothing synchronises across warps inside the loop.  The first
// version split a row over three warps (one 16-byte chunk per lane) and spent 270 instructions per 8 elements on
// barriers, statistics exchange and predication (68 % issue-slot utilisation at 4.4 TB/s); here a lane owns all its
// chunks, the math runs on fp32 pairs, and phase two re-derives xhat / g from the raw bf16 words instead of keeping them.
constexpr int BWD_WARPS = 6;  // per CTA, two CTAs per SM
constexpr int BWD_DEPTH = 4;  // rows in flight per warp
template <int MAXC, bool kFull, bool kStatsRing>
__global__ void __launch_bounds__(BWD_WARPS * 32, 2)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
              const float* __restrict__ stats, const float* __restrict__ w, int T, int H,
              __nv_bfloat16* __restrict__ dz, float* __restrict__ dw_acc, float* __restrict__ db_acc,
              float* __restrict__ colsum_acc) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  pdl_grid_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunk = H >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(H) * 2;
  // slot = z row | dy row (| the 16-byte aligned pair of (mean, rstd) entries that holds the row's, kStatsRing: T even, stats
  // 16-byte aligned).  The statistics used to be a plain 8-byte load per row, requested one ring depth ahead and queued in
  // registers: its first use carried 15 % of the kernel's stall samples (the request queues behind ~150 KB of bulk copies per SM).
  const uint32_t slot_bytes = 2 * row_bytes + (kStatsRing ? 16 : 0);
  uint8_t* ring = ln_smem + static_cast<size_t>(warp) * BWD_DEPTH * slot_bytes;
  uint8_t* after = ln_smem + static_cast<size_t>(BWD_WARPS) * BWD_DEPTH * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(after) + warp * BWD_DEPTH;
  float* s_red = reinterpret_cast<float*>(after + BWD_WARPS * BWD_DEPTH * 8);          // [3][H] column partials of the CTA
  const long first = static_cast<long>(blockIdx.x) * BWD_WARPS + warp, stride = static_cast<long>(gridDim.x) * BWD_WARPS;
  auto issue = [&](long k) {  // lane 0 only: slot k % BWD_DEPTH <- the warp's k-th row
    const long row = first + k * stride;
    if (row < T) {
      const int st = static_cast<int>(k % BWD_DEPTH);
      uint8_t* dst = ring + st * slot_bytes;
      mbar_arrive_expect_tx(&bars[st], slot_bytes);
      bulk_load_1d(dst, z + row * H, row_bytes, &bars[st]);
      bulk_load_1d(dst + row_bytes, dy + row * H, row_bytes, &bars[st]);
      if (kStatsRing) bulk_load_1d(dst + 2 * row_bytes, stats + 2 * (row & ~1L), 16, &bars[st]);
    }
  };
  if (lane == 0) {
    for (int i = 0; i < BWD_DEPTH; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    for (int k = 0; k < BWD_DEPTH; ++k) issue(k);
  }
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();

  f32x2 ww[MAXC][4], pdw[MAXC][4], pdb[MAXC][4], pcs[MAXC][4];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + 32 * i;
    const bool ok = kFull || c < nchunk;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ww[i][q] = f2_pack(ok ? __ldg(w + c * 8 + 2 * q) : 0.f, ok ? __ldg(w + c * 8 + 2 * q + 1) : 0.f);
      pdw[i][q] = pdb[i][q] = pcs[i][q] = f2_splat(0.f);
    }
  }
  const float invH = 1.f / H;
  // (mean, rstd) of a row is 8 bytes: a plain load BWD_DEPTH rows ahead, queued in registers
  auto load_stats = [&](long row) -> float2 {   // volatile asm: issued here, not sunk to its use four rows later
    float2 v = make_float2(0.f, 0.f);
    if (row < T) asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(stats + 2 * row));
    return v;
  };
  float2 sq[kStatsRing ? 1 : BWD_DEPTH];
  if (!kStatsRing) {
#pragma unroll
    for (int u = 0; u < BWD_DEPTH; ++u) sq[kStatsRing ? 0 : u] = load_stats(first + u * stride);
  }

  for (long k0 = 0; first + k0 * stride < T; k0 += BWD_DEPTH) {
#pragma unroll
    for (int u = 0; u < BWD_DEPTH; ++u) {
      const long k = k0 + u, row = first + k * stride;
      if (row < T) {
        float mean, rstd;
        if (!kStatsRing) {
          mean = sq[kStatsRing ? 0 : u].x, rstd = sq[kStatsRing ? 0 : u].y;
          sq[kStatsRing ? 0 : u] = load_stats(row + BWD_DEPTH * stride);
        }
        mbar_wait(&bars[u], static_cast<uint32_t>(k0 / BWD_DEPTH) & 1);
        const uint32_t src = smem_u32(ring + u * slot_bytes);
        if (kStatsRing) {
          const float2 st2 = lds64_f2(src + 2 * row_bytes + static_cast<uint32_t>(row & 1) * 8);   // broadcast read
          mean = st2.x, rstd = st2.y;
        }
        uint4 rz[MAXC], rd[MAXC];
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
          const int c = lane + 32 * i;
          if (kFull || c < nchunk) {
            rz[i] = lds128(src + c * 16);
            rd[i] = lds128(src + row_bytes + c * 16);
          } else {
            rz[i] = rd[i] = make_uint4(0, 0, 0, 0);   // zero dy: no contribution to any sum
          }
        }
        __syncwarp();                    // the slot is in registers: refill it
        if (lane == 0) issue(k + BWD_DEPTH);
        const f32x2 rstd2 = f2_splat(rstd), nmr = f2_splat(-mean * rstd);
        // ---- phase 1: column partials and the two row sums ----
        f32x2 s1v = f2_splat(0.f), s2v = f2_splat(0.f);
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
          const uint32_t zw[4] = {rz[i].x, rz[i].y, rz[i].z, rz[i].w}, dw_[4] = {rd[i].x, rd[i].y, rd[i].z, rd[i].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const f32x2 df = f2_from_bf16x2(dw_[q]);
            const f32x2 xh = f2_fma(f2_from_bf16x2(zw[q]), rstd2, nmr);
            pdw[i][q] = f2_fma(df, xh, pdw[i][q]);
            pdb[i][q] = f2_add(pdb[i][q], df);
            const f32x2 g = f2_mul(df, ww[i][q]);
            s1v = f2_add(s1v, g);
            s2v = f2_fma(g, xh, s2v);
          }
        }
        float a0, a1, b0, b1;
        f2_unpack(s1v, a0, a1);
        f2_unpack(s2v, b0, b1);
        float s1 = a0 + a1, s2 = b0 + b1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // the two reductions interleaved
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        // dz = rstd (g - s1/H - xh s2/H)  =  (g - xh * s2/H) * rstd - rstd * s1/H
        const f32x2 ns2 = f2_splat(-s2 * invH), c0 = f2_splat(-s1 * invH * rstd);
        // ---- phase 2: xh and g again from the raw words (cheaper than 48 live registers), outputs ----
        uint4* out = reinterpret_cast<uint4*>(dz + static_cast<size_t>(row) * H);
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
          const int c = lane + 32 * i;
          if (kFull || c < nchunk) {
            const uint32_t zw[4] = {rz[i].x, rz[i].y, rz[i].z, rz[i].w}, dw_[4] = {rd[i].x, rd[i].y, rd[i].z, rd[i].w};
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const f32x2 xh = f2_fma(f2_from_bf16x2(zw[q]), rstd2, nmr);
              const f32x2 g = f2_mul(f2_from_bf16x2(dw_[q]), ww[i][q]);
              o[q] = f2_to_bf16x2(f2_fma(f2_fma(xh, ns2, g), rstd2, c0));
              // the bias gradient must see exactly the values the GEMMs will read -> sum the rounded outputs
              pcs[i][q] = f2_add(pcs[i][q], f2_from_bf16x2(o[q]));
            }
            out[c] = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + 32 * i;
    if (kFull || c < nchunk) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float lo, hi;
        f2_unpack(pdw[i][q], lo, hi);
        atomicAdd(&s_red[c * 8 + 2 * q], lo);
        atomicAdd(&s_red[c * 8 + 2 * q + 1], hi);
        f2_unpack(pdb[i][q], lo, hi);
        atomicAdd(&s_red[H + c * 8 + 2 * q], lo);
        atomicAdd(&s_red[H + c * 8 + 2 * q + 1], hi);
        f2_unpack(pcs[i][q], lo, hi);
        atomicAdd(&s_red[2 * H + c * 8 + 2 * q], lo);
        atomicAdd(&s_red[2 * H + c * 8 + 2 * q + 1], hi);
      }
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
  // the table rows are gathers that depend on ids[row] / pos[row]: the indices of the warp's NEXT row are requested one row ahead
  const int row0 = blockIdx.x * WARPS + warp, rstride = gridDim.x * WARPS;
  int id_nxt = row0 < T ? __ldg(ids + row0) : 0, ps_nxt = row0 < T ? __ldg(pos + row0) : 0;
  for (int row = row0; row < T; row += rstride) {
    const int id_cur = id_nxt, ps_cur = ps_nxt;
    if (row + rstride < T) {
      id_nxt = __ldg(ids + row + rstride);
      ps_nxt = __ldg(pos + row + rstride);
    }
    const float4* wr = reinterpret_cast<const float4*>(word + static_cast<size_t>(id_cur) * E);
    const float4* pr = reinterpret_cast<const float4*>(posemb + static_cast<size_t>(ps_cur) * E);
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
__global__ void __launch_bounds__(WARPS * 32, 2)   // two CTAs per SM: the kernel waits on dependent gathers (ids -> table rows)
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
  // indices and statistics of the warp's NEXT row are requested one row ahead (the table rows are dependent gathers)
  const int row0 = blockIdx.x * WARPS + warp, rstride = gridDim.x * WARPS;
  int id_nxt = 0, ps_nxt = 0;
  float mean_nxt = 0.f, rstd_nxt = 0.f;
  if (row0 < T) {
    id_nxt = __ldg(ids + row0); ps_nxt = __ldg(pos + row0);
    mean_nxt = __ldg(stats + 2 * row0); rstd_nxt = __ldg(stats + 2 * row0 + 1);
  }
  for (int row = row0; row < T; row += rstride) {
    const int id = id_nxt, ps = ps_nxt;
    const float mean = mean_nxt, rstd = rstd_nxt;
    if (row + rstride < T) {
      const int nr = row + rstride;
      id_nxt = __ldg(ids + nr); ps_nxt = __ldg(pos + nr);
      mean_nxt = __ldg(stats + 2 * nr); rstd_nxt = __ldg(stats + 2 * nr + 1);
    }
    const float4* wr = reinterpret_cast<const float4*>(word + static_cast<size_t>(id) * E);
    const float4* pr = reinterpret_cast<const float4*>(posemb + static_cast<size_t>(ps) * E);
    const float4* tr = reinterpret_cast<const float4*>(type0);
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
  PLB_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, PLB_ERR_ARG,
              "ln_fwd: z / y must be 16-byte aligned");
  const bool full = H == nc * 256;
  auto kern = nc == 1 ? (full ? ln_fwd_kernel<1, true> : ln_fwd_kernel<1, false>)
            : nc == 2 ? (full ? ln_fwd_kernel<2, true> : ln_fwd_kernel<2, false>)
            : nc == 3 ? (full ? ln_fwd_kernel<3, true> : ln_fwd_kernel<3, false>)
                      : (full ? ln_fwd_kernel<4, true> : ln_fwd_kernel<4, false>);
  // per-warp ring of FWD_DEPTH rows + its mbarriers
  const int smem = WARPS * FWD_DEPTH * (H * 2 + 8);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  // idempotent
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "ln_fwd: cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
  int per_sm = 1;   // resident CTAs per SM (registers and shared memory): the grid is persistent
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
  PLB_REQUIRE(e == cudaSuccess && per_sm >= 1, PLB_ERR_CUDA, "ln_fwd: occupancy query failed: %s", cudaGetErrorString(e));
  e = launch_kernel(kern, dim3(row_grid(T, per_sm)), dim3(WARPS * 32), smem, static_cast<cudaStream_t>(stream),
                    static_cast<const __nv_bfloat16*>(z), T, H, w, b, eps, static_cast<__nv_bfloat16*>(y), stats);
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "plb_ln_fwd: launch failed: %s", cudaGetErrorString(e));
  return check_launch("plb_ln_fwd");
}

extern "C" int plb_ln_bwd(const void* dy, const void* z, const float* stats, const float* w, int T, int H, void* dz,
                          float* dw_accum, float* db_accum, float* colsum_accum, plb_stream_t stream) {
  PLB_REQUIRE(dy && z && stats && w && dz && dw_accum && db_accum && T > 0, PLB_ERR_ARG, "ln_bwd: bad arguments");
  PLB_REQUIRE(H % 8 == 0 && H >= 8 && H <= 8 * 32 * LN_MAXC, PLB_ERR_UNSUPPORTED, "ln_bwd: H=%d must be a multiple of 8, <= %d", H,
              8 * 32 * LN_MAXC);
  const int nc = (H / 8 + 31) / 32;
  PLB_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dz) & 15) == 0, PLB_ERR_ARG, "ln_bwd: dy / z / dz must be 16-byte aligned");
  const bool full = H == nc * 256;
#ifndef PLB_LN_STATS_RING
#define PLB_LN_STATS_RING 1
#endif
  // the statistics ride in the ring when the 16-byte pair around every row's entry exists and is aligned
  const bool sring = PLB_LN_STATS_RING && full && (T % 2 == 0) && (reinterpret_cast<uintptr_t>(stats) & 15) == 0;
  auto kern = nc == 1 ? (full ? (sring ? ln_bwd_kernel<1, true, true> : ln_bwd_kernel<1, true, false>) : ln_bwd_kernel<1, false, false>)
            : nc == 2 ? (full ? (sring ? ln_bwd_kernel<2, true, true> : ln_bwd_kernel<2, true, false>) : ln_bwd_kernel<2, false, false>)
            : nc == 3 ? (full ? (sring ? ln_bwd_kernel<3, true, true> : ln_bwd_kernel<3, true, false>) : ln_bwd_kernel<3, false, false>)
                      : (full ? (sring ? ln_bwd_kernel<4, true, true> : ln_bwd_kernel<4, true, false>) : ln_bwd_kernel<4, false, false>);
  // per-warp rings of BWD_DEPTH (z row | dy row [| statistics pair]) slots + their mbarriers + the CTA's [3][H] column partials
  const int smem = BWD_WARPS * BWD_DEPTH * (2 * H * 2 + (sring ? 16 : 0) + 8) + 3 * H * static_cast<int>(sizeof(float));
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  // idempotent
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "ln_bwd: cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
  int per_sm = 1;   // resident CTAs per SM: the grid is persistent
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BWD_WARPS * 32, smem);
  PLB_REQUIRE(e == cudaSuccess && per_sm >= 1, PLB_ERR_CUDA, "ln_bwd: occupancy query failed: %s", cudaGetErrorString(e));
  const int need = (T + BWD_WARPS - 1) / BWD_WARPS, cap = sm_count() * per_sm;
  e = launch_kernel(kern, dim3(need < cap ? need : cap), dim3(BWD_WARPS * 32), smem, static_cast<cudaStream_t>(stream),
                    static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(z), stats, w, T, H,
                    static_cast<__nv_bfloat16*>(dz), dw_accum, db_accum, colsum_accum);
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "plb_ln_bwd: launch failed: %s", cudaGetErrorString(e));
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
    // one private copy of the word-table gradient per CTA; two CTAs per SM when two copies fit (the 188 x 128 table: 2 x 98 KB)
    const int per_sm = 2 * (big + 1024) <= 227 * 1024 ? 2 : 1;
    const int cap = sm_count() * per_sm;
    embed_ln_bwd_kernel<true><<<need < cap ? need : cap, WARPS * 32, big, st>>>(
        static_cast<const __nv_bfloat16*>(de), ids, pos, T, E, word, posemb, type0, ln_w, stats, pad_id, V, dword, dpos,
        dtype0, dln_w, dln_b);
  } else {
    embed_ln_bwd_kernel<false><<<row_grid(T, 4), WARPS * 32, small, st>>>(
        static_cast<const __nv_bfloat16*>(de), ids, pos, T, E, word, posemb, type0, ln_w, stats, pad_id, V, dword, dpos,
        dtype0, dln_w, dln_b);
  }
  return check_launch("plb_embed_ln_bwd");
}
