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
  pdl_grid_sync();
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
  pdl_grid_sync();
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
  pdl_grid_sync();
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

// Large vocabularies, ld > 8192 (the 64 k token head): ONE ROW PER CLUSTER of 8 CTAs x 4 warps, persistent.  Every warp owns
// a contiguous piece of the row (kVec x 512 bytes) and streams its pieces of the cluster's rows through a private ring of
// CE_BUFS shared-memory slots: 1-D bulk copies in (cp.async.bulk + mbarrier, issued by lane 0 one row ahead), three passes
// over the slot from shared memory (maximum, sum of exponentials, gradient written back in place), one bulk copy out — so
// the logits cross HBM exactly once each way, nothing is re-read from L2, and the bytes in flight do not depend on
// registers.  Per row every warp reduces its own (max, sum) with shuffles, lane 0 sends the pair into slot [rank * 4 + warp] of
// all eight CTAs with st.async (the store completes 8 bytes of the receiver's mbarrier: no cluster barrier per row — with
// barrier.cluster.arrive.release / wait.acquire around plain st.shared::cluster stores the kernel took 594 us instead of 533),
// and every warp merges the 32 pairs itself (one pair per lane).  The pair slots and their barriers alternate between two
// sets by row parity: a CTA can be at most one row ahead of a peer, because it needs the peer's pair of row r + 1, which the
// peer sends only after it has merged row r.
// What this replaces: one CTA per row, three passes with one 16-byte load per thread in flight (816 us for the 2.92 GB of
// the 64 000-wide head = 3.6 TB/s).  Measured on the way (profiles/ce_bench.py, -D variants, same box): a register-resident
// one-shot cluster per row 580 - 645 us (each generation of CTAs loads, computes and stores in lock-step; made persistent
// its spills fell out of L1: 1.1 ms).  With the cluster barrier per row: this kernel 594 us; refilling the ring at the top of
// the row / before the barrier instead of after it 641 / 690; 256-thread CTAs 641; three slots with 4 resident CTAs 720;
// 1 / 2 / 3 of 4 gradient exponentials on the FMA pipe 583 / 595 / 627; with NO exponentials at all 530 — the skeleton, not
// the MUFU, set the time.  With the st.async exchange: 533 us = 5.5 TB/s (7 / 5 resident CTAs 541 / 551, three slots 550,
// 256-thread CTAs 646, 1 / 2 of 4 gradient exponentials on the FMA pipe 520 / 529).
#ifndef PLB_CE_THREADS
#define PLB_CE_THREADS 128
#endif
constexpr int CE_CLUSTER = 8, CE_THREADS = PLB_CE_THREADS, CE_WARPS = CE_THREADS / 32;
#ifndef PLB_CE_BUFS
#define PLB_CE_BUFS 2      // ring slots per warp (16 KB per CTA and slot at kVec = 8)
#endif
#ifndef PLB_CE_XCHG
#define PLB_CE_XCHG 1      // 1: pairs travel with st.async + mbarrier complete_tx (no cluster barrier per row), 0: barrier.cluster per row
#endif
#ifndef PLB_CE_ISSUE_AT
#define PLB_CE_ISSUE_AT 2  // where lane 0 refills the ring: 0 = top of the row, 1 / 2 = after the pairs are sent (before / after the cluster barrier when PLB_CE_XCHG == 0)
#endif
#ifndef PLB_CE_MINB
#define PLB_CE_MINB 6      // resident CTAs per SM the register budget is cut for (6 x 33 KB of shared memory fit)
#endif
constexpr int CE_BUFS = PLB_CE_BUFS;
__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
#ifndef PLB_CE_POLY
#define PLB_CE_POLY 0      // pairs per 16-byte word (of 4) whose GRADIENT exponentials run on the FMA pipe instead of the MUFU
#endif
// 2^x on a pair without the MUFU: 2^n p(f), n = round(x), f = x - n in [-0.5, 0.5], cubic p (relative error 8e-5; the
// result is rounded to bf16, 2e-3); x clamped to >= -126 so that the exponent arithmetic stays in range
__device__ __forceinline__ f32x2 ex2_poly3_x2(f32x2 x) {
  float lo, hi;
  f2_unpack(x, lo, hi);
  const f32x2 xc = f2_pack(fmaxf(lo, -126.f), fmaxf(hi, -126.f));
  const f32x2 tt = f2_add(xc, f2_splat(12582912.f));                 // low mantissa bits = n
  const f32x2 ff = f2_add(xc, f2_mul(f2_add(tt, f2_splat(-12582912.f)), f2_splat(-1.f)));
  f32x2 pp = f2_fma(ff, f2_splat(0.05508868396282196f), f2_splat(0.24260404706001282f));
  pp = f2_fma(pp, ff, f2_splat(0.6932762265205383f));
  pp = f2_fma(pp, ff, f2_splat(0.9999289512634277f));
  float p0, p1, t0, t1;
  f2_unpack(pp, p0, p1);
  f2_unpack(tt, t0, t1);
  return f2_pack(__uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23)),
                 __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23)));
}
// e[q] = 2^(z log2e + nm2) on pairs; the last kPoly pairs of the word on the FMA pipe
template <int kPoly = 0>
__device__ __forceinline__ void ce_word_exp(const uint4& v, f32x2 l2e, f32x2 nm2, f32x2 (&e)[4]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const f32x2 x = f2_fma(f2_from_bf16x2(w[q]), l2e, nm2);
    if (q >= 4 - kPoly) {
      e[q] = ex2_poly3_x2(x);
    } else {
      float lo, hi;
      f2_unpack(x, lo, hi);
      e[q] = f2_pack(ex2_fast(lo), ex2_fast(hi));
    }
  }
}

template <int kVec>   // 16-byte words per lane and row
__global__ void __cluster_dims__(CE_CLUSTER, 1, 1) __launch_bounds__(CE_THREADS, kVec * CE_WARPS <= 32 ? PLB_CE_MINB : 3)
masked_ce_cluster_kernel(__nv_bfloat16* __restrict__ logits, int ld, int M, int V, const int32_t* __restrict__ labels,
                         const float* __restrict__ weights, float* __restrict__ row_loss) {
  pdl_grid_sync();
  extern __shared__ __align__(128) uint8_t ce_smem[];             // [warp][slot][kVec * 512 bytes]
  __shared__ uint64_t bars[CE_WARPS][CE_BUFS];
  __shared__ float2 s_part[2][CE_CLUSTER * CE_WARPS];
  __shared__ uint64_t xbar[2];                                    // exchange barriers by row parity (PLB_CE_XCHG)
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr int kPiece = kVec * 32;                               // 16-byte words per warp and row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int nclusters = gridDim.x / CE_CLUSTER;
  const int first_row = blockIdx.x / CE_CLUSTER;
  const int nvec = ld >> 3;                                       // ld % 8 == 0 (checked by the caller)
  const int w0 = (static_cast<int>(rank) * CE_WARPS + warp) * kPiece;            // the warp's first word of a row
  const int nwords = nvec - w0 < 0 ? 0 : (nvec - w0 < kPiece ? nvec - w0 : kPiece);   // ... and how many exist (<= kPiece)
  const int j_first = w0 * 8;                                     // first column of the piece
  const bool full = j_first + kPiece * 8 <= V;                    // no column of the piece is padding (warp-uniform)
  const uint32_t ring = smem_u32(ce_smem) + static_cast<uint32_t>(warp) * CE_BUFS * kPiece * 16;
  uint64_t* bar = bars[warp];
  auto issue = [&](int it) {   // lane 0: slot it % CE_BUFS <- the warp's piece of the cluster's it-th row
    const long m = first_row + static_cast<long>(it) * nclusters;
    if (m < M && nwords > 0) {
      const int st = it % CE_BUFS;
      mbar_arrive_expect_tx(&bar[st], static_cast<uint32_t>(nwords) * 16);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       ring + st * kPiece * 16),
                   "l"(reinterpret_cast<const uint4*>(logits + m * ld) + w0), "r"(nwords * 16), "r"(smem_u32(&bar[st]))
                   : "memory");
    }
  };
  if (lane == 0) {
    for (int i = 0; i < CE_BUFS; ++i) mbar_init(&bar[i], 1);
    if (PLB_CE_XCHG && warp == 0) {
      mbar_init(&xbar[0], 1);
      mbar_init(&xbar[1], 1);
    }
    fence_mbar_init();
    for (int i = 0; i < CE_BUFS - 1; ++i) issue(i);
  }
  __syncwarp();
  if (PLB_CE_XCHG) cluster_sync_all();   // once: the peers' exchange barriers exist before anybody stores into this CTA
  const f32x2 l2e = f2_splat(kLog2e);
  int it = 0;
  for (int m = first_row; m < M; m += nclusters, ++it) {
    const int st = it % CE_BUFS;
    const uint32_t piece = ring + st * kPiece * 16;               // column k of the piece: piece + 2 k
    const uint32_t slot = piece + lane * 16;                      // word i of this lane: slot + i * 512 (conflict-free)
    int y = labels[m];
    const bool y_ok = y >= 0 && y < V;
    const float wgt = y_ok ? weights[m] : 0.f;   // a label outside [0, V): no loss and no gradient for the row
    if (!y_ok && rank == 0 && tid == 0) row_loss[m] = 0.f;
    const int ky = y_ok ? y - j_first : -1;      // the label's column inside this warp's piece, if it is there
    const bool mine = ky >= 0 && ky < nwords * 8;
    auto refill = [&]() {
      // the slot the row after this one goes to was handed to a bulk store CE_BUFS - 1 iterations ago: wait until it has been read
      if (lane == 0) {
        bulk_wait_read<CE_BUFS - 2>();
        issue(it + CE_BUFS - 1);
      }
      __syncwarp();
    };
    if (PLB_CE_ISSUE_AT == 0) refill();
    if (nwords > 0) mbar_wait(&bar[st], static_cast<uint32_t>(it / CE_BUFS) & 1);
    float zy = 0.f;
    if (mine && lane == 0) {
      uint16_t bits;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(bits) : "r"(piece + 2 * ky));
      zy = __uint_as_float(static_cast<uint32_t>(bits) << 16);
    }
    // ---- warp-local maximum ----
    float mx = -INFINITY;
    if (full) {
#pragma unroll
      for (int i = 0; i < kVec; ++i) {
        const uint4 v = lds128(slot + i * 512);
        mx = max3(mx, bf16_lo(v.x), bf16_hi(v.x));
        mx = max3(mx, bf16_lo(v.y), bf16_hi(v.y));
        mx = max3(mx, bf16_lo(v.z), bf16_hi(v.z));
        mx = max3(mx, bf16_lo(v.w), bf16_hi(v.w));
      }
    } else {
      for (int i = 0; i < kVec; ++i) {
        const int j0 = j_first + (lane + 32 * i) * 8;
        if (j0 >= V) break;
        const uint4 v = lds128(slot + i * 512);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (j0 + 2 * q < V) mx = fmaxf(mx, bf16_lo(w[q]));
          if (j0 + 2 * q + 1 < V) mx = fmaxf(mx, bf16_hi(w[q]));
        }
      }
    }
    mx = warp_max(mx);
    const float nm2s = mx == -INFINITY ? 0.f : -mx * kLog2e;       // a warp without a valid column contributes (-inf, 0)
    // ---- warp-local sum of exponentials (relative to the warp's maximum) ----
    float se;
    {
      const f32x2 nm2 = f2_splat(nm2s);
      f32x2 acc[2] = {f2_splat(0.f), f2_splat(0.f)};
      if (full) {
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
          f32x2 e[4];
          ce_word_exp(lds128(slot + i * 512), l2e, nm2, e);
          acc[0] = f2_add(acc[0], f2_add(e[0], e[1]));
          acc[1] = f2_add(acc[1], f2_add(e[2], e[3]));
        }
      } else {
        for (int i = 0; i < kVec; ++i) {
          const int j0 = j_first + (lane + 32 * i) * 8;
          if (j0 >= V) break;
          f32x2 e[4];
          ce_word_exp(lds128(slot + i * 512), l2e, nm2, e);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float lo, hi;
            f2_unpack(e[q], lo, hi);
            acc[q & 1] = f2_add(acc[q & 1], f2_pack(j0 + 2 * q < V ? lo : 0.f, j0 + 2 * q + 1 < V ? hi : 0.f));
          }
        }
      }
      float s0, s1;
      f2_unpack(f2_add(acc[0], acc[1]), s0, s1);
      se = warp_sum(s0 + s1);
    }
    if (PLB_CE_XCHG) {
      // Every warp sends its pair to all eight CTAs with st.async: the store lands in the peer's s_part and completes 8 bytes
      // of the peer's exchange barrier; a CTA's barrier of the row completes when its own arrive.expect_tx (thread 0) and all
      // 32 pairs are in.  No cluster-wide barrier: a CTA can only be one row ahead of a peer (it needs the peer's pair of
      // row r + 1, which the peer sends after it has merged row r), which is exactly what the two parity sets allow.
      if (tid == 0) mbar_arrive_expect_tx(&xbar[it & 1], CE_CLUSTER * CE_WARPS * 8);
      if (lane == 0) {
        const uint32_t own = smem_u32(&s_part[it & 1][rank * CE_WARPS + warp]), bar_own = smem_u32(&xbar[it & 1]);
#pragma unroll
        for (uint32_t r = 0; r < CE_CLUSTER; ++r)
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(
                           mapa_u32(own, r)),
                       "f"(mx), "f"(se), "r"(mapa_u32(bar_own, r))
                       : "memory");
      }
      if (PLB_CE_ISSUE_AT != 0) refill();
      mbar_wait(&xbar[it & 1], static_cast<uint32_t>(it >> 1) & 1);
    } else {
      if (lane == 0) {
        const uint32_t own = smem_u32(&s_part[it & 1][rank * CE_WARPS + warp]);
#pragma unroll
        for (uint32_t r = 0; r < CE_CLUSTER; ++r) st_cluster_f32x2(mapa_u32(own, r), mx, se);
      }
      if (PLB_CE_ISSUE_AT == 1) refill();
      cluster_sync_all();   // release / acquire: the 32 pairs of this row are visible everywhere
      if (PLB_CE_ISSUE_AT == 2) refill();
    }
    float row_max = -INFINITY, row_se = 0.f;
    {
      float2 part[CE_CLUSTER * CE_WARPS / 32];
#pragma unroll
      for (int i = 0; i < CE_CLUSTER * CE_WARPS / 32; ++i) {
        part[i] = s_part[it & 1][lane + 32 * i];
        row_max = fmaxf(row_max, part[i].x);
      }
      row_max = warp_max(row_max);                                 // finite: V >= 1
#pragma unroll
      for (int i = 0; i < CE_CLUSTER * CE_WARPS / 32; ++i) row_se += part[i].y * ex2_fast((part[i].x - row_max) * kLog2e);
      row_se = warp_sum(row_se);
    }
    if (nwords > 0) {
      // ---- gradient, in place in the slot: w * (softmax - onehot); columns [V, ld) are zeroed for the GEMMs that read them ----
      const float invs = wgt / row_se, nrm2s = -row_max * kLog2e;
      const f32x2 inv = f2_splat(invs), nrm2 = f2_splat(nrm2s);
      if (full) {
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
          f32x2 e[4];
          ce_word_exp<PLB_CE_POLY>(lds128(slot + i * 512), l2e, nrm2, e);
          sts128(slot + i * 512, make_uint4(f2_to_bf16x2(f2_mul(e[0], inv)), f2_to_bf16x2(f2_mul(e[1], inv)),
                                            f2_to_bf16x2(f2_mul(e[2], inv)), f2_to_bf16x2(f2_mul(e[3], inv))));
        }
      } else {
        for (int i = 0; i < kVec; ++i) {
          const int j0 = j_first + (lane + 32 * i) * 8;
          if (lane + 32 * i >= nwords) break;
          f32x2 e[4];
          ce_word_exp(lds128(slot + i * 512), l2e, nrm2, e);
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float lo, hi;
            f2_unpack(f2_mul(e[q], inv), lo, hi);
            o[q] = pack_bf16x2(j0 + 2 * q < V ? lo : 0.f, j0 + 2 * q + 1 < V ? hi : 0.f);
          }
          sts128(slot + i * 512, make_uint4(o[0], o[1], o[2], o[3]));
        }
      }
      __syncwarp();
      if (mine && lane == 0) {   // the label's entry: w * (softmax - 1); the loss from the logit read before the overwrite
        row_loss[m] = wgt * (__logf(row_se) + row_max - zy);
        const float gy = ex2_fast(fmaf(zy, kLog2e, nrm2s)) * invs - wgt;
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(piece + 2 * ky), "h"(__bfloat16_as_ushort(__float2bfloat16(gy))) : "memory");
      }
      fence_proxy_async_smem();   // every lane: its writes to the slot become visible to the bulk-copy engine
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                         reinterpret_cast<uint4*>(logits + static_cast<size_t>(m) * ld) + w0),
                     "r"(piece), "r"(nwords * 16)
                     : "memory");
        bulk_commit();
      }
    }
  }
  if (lane == 0) bulk_wait<0>();   // the last stores have left shared memory and are performed before the CTA retires
}

// Wide rows the cluster kernel does not take (ld <= 8192, ld > 131072 or an unaligned buffer): one CTA per row, 16-byte accesses.  Pass 1 streams the row from HBM, passes 2
// and 3 re-read it from L1/L2 (a 64 k row is 128 KB), so the traffic is one read and one write of the logits.
__global__ void __launch_bounds__(256)
masked_ce_wide_kernel(__nv_bfloat16* __restrict__ logits, int ld, int M, int V, const int32_t* __restrict__ labels,
                      const float* __restrict__ weights, float* __restrict__ row_loss) {
  pdl_grid_sync();
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
  pdl_grid_sync();
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
  pdl_grid_sync();
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
  cudaError_t e = launch_kernel(gather_rows_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream),
                                static_cast<const __nv_bfloat16*>(src), ld_src, rows, M, H, static_cast<__nv_bfloat16*>(dst), ld_dst);
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "plb_gather_rows: launch failed: %s", cudaGetErrorString(e));
  return check_launch("plb_gather_rows");
}

extern "C" int plb_scatter_add_rows(const void* src, int ld_src, const int32_t* rows, int M, int H, void* dst, int ld_dst,
                                    plb_stream_t stream) {
  PLB_REQUIRE(src && rows && dst && M >= 0 && H > 0, PLB_ERR_ARG, "scatter_add_rows: bad arguments");
  PLB_REQUIRE(H % 2 == 0 && ld_src % 2 == 0 && ld_dst % 2 == 0, PLB_ERR_ARG, "scatter_add_rows: H and lds must be even");
  if (M == 0) return PLB_OK;
  const long total = static_cast<long>(M) * (H / 2);
  const int grid = static_cast<int>(total / 256 + 1 < sm_count() * 8L ? total / 256 + 1 : sm_count() * 8L);
  cudaError_t e = launch_kernel(scatter_add_rows_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream),
                                static_cast<const __nv_bfloat16*>(src), ld_src, rows, M, H, static_cast<__nv_bfloat16*>(dst), ld_dst);
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "plb_scatter_add_rows: launch failed: %s", cudaGetErrorString(e));
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
  cudaError_t le = cudaSuccess;
  if (V <= 2048)
    le = launch_kernel(masked_ce_kernel<32>, dim3((M + 7) / 8), dim3(256), 0, stream, z, ld, M, V, labels, weights, row_loss);
  else if (ld > 8192 && ld <= 16 * 8192 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0) {
    const int per_thread = (ld / 8 + CE_CLUSTER * CE_THREADS - 1) / (CE_CLUSTER * CE_THREADS);   // 16-byte words per thread
    auto kern = per_thread <= 2 ? masked_ce_cluster_kernel<2> : per_thread <= 4 ? masked_ce_cluster_kernel<4>
              : per_thread <= 8 ? masked_ce_cluster_kernel<8> : masked_ce_cluster_kernel<16>;
    const int kv = per_thread <= 2 ? 2 : per_thread <= 4 ? 4 : per_thread <= 8 ? 8 : 16;
    const int smem = CE_WARPS * CE_BUFS * kv * 512;
    static int cached[16][4] = {};   // resident clusters per (device, kernel instantiation)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "masked_ce: cudaGetDevice: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  // per device, idempotent
    PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "masked_ce: cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
    int uncached = 0;
    int& resident = (dev >= 0 && dev < 16) ? cached[dev][kv == 2 ? 0 : kv == 4 ? 1 : kv == 8 ? 2 : 3] : uncached;
    if (resident == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(CE_CLUSTER * sm_count(), 1, 1);
      cfg.blockDim = dim3(CE_THREADS, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = CE_CLUSTER;
      attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      PLB_REQUIRE(e == cudaSuccess && n >= 1, PLB_ERR_CUDA, "masked_ce: cluster occupancy query failed: %s", cudaGetErrorString(e));
      resident = n;
    }
    int nclusters = resident;
    if (nclusters > M) nclusters = M;
    le = launch_kernel(kern, dim3(nclusters * CE_CLUSTER), dim3(CE_THREADS), smem, stream, z, ld, M, V, labels, weights, row_loss);
  } else
    le = launch_kernel(masked_ce_wide_kernel, dim3(M), dim3(256), 0, stream, z, ld, M, V, labels, weights, row_loss);
  PLB_REQUIRE(le == cudaSuccess, PLB_ERR_CUDA, "plb_masked_ce: launch failed: %s", cudaGetErrorString(le));
  int rc = check_launch("plb_masked_ce");
  if (rc) return rc;
  le = launch_kernel(sum_rows_kernel, dim3(1), dim3(1024), 0, stream, static_cast<const float*>(row_loss), M, loss_out, accumulate);
  PLB_REQUIRE(le == cudaSuccess, PLB_ERR_CUDA, "plb_masked_ce(sum): launch failed: %s", cudaGetErrorString(le));
  return check_launch("plb_masked_ce(sum)");
}

extern "C" int plb_colsum(const void* x, int ld, int T, int N, float* out_accum, plb_stream_t stream) {
  PLB_REQUIRE(x && out_accum && T > 0 && N > 0, PLB_ERR_ARG, "colsum: bad arguments");
  PLB_REQUIRE(ld % 8 == 0 && ld >= (N + 7) / 8 * 8 && (reinterpret_cast<uintptr_t>(x) % 16) == 0, PLB_ERR_ARG,
              "colsum: ld must be a multiple of 8 covering N rounded up to 8, x 16-byte aligned");
  const int gx = (N + 255) / 256;
  int gy = (8 * sm_count() + gx - 1) / gx;
  if (gy > (T + 7) / 8) gy = (T + 7) / 8;
  cudaError_t e = launch_kernel(colsum_kernel, dim3(gx, gy), dim3(256), 0, static_cast<cudaStream_t>(stream),
                                static_cast<const __nv_bfloat16*>(x), ld, T, N, out_accum);
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "plb_colsum: launch failed: %s", cudaGetErrorString(e));
  return check_launch("plb_colsum");
}
