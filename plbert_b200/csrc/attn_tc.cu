// plbert_b200 — fused key-padded attention forward on the 5th-generation tensor cores (tcgen05.mma, S / O
// accumulators in TMEM), head_dim 64, sequences up to 512 rows.
//
// Replaces HF modeling_albert.py:184-197 (SDPA with the [B,1,L,L] mask of masking_utils.py:495-520).  One CTA owns one
// (sequence, head): K and V of that head (<= 2 x 64 KB bf16) are TMA-loaded ONCE into 128-byte-swizzled shared memory
// and stay resident while the CTA walks the query tiles, so K/V leave HBM exactly once per head.
//
// Warp roles (12 warps = 3 warpgroups): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer (warps 2,3
// idle; the warpgroup gives its registers away with setmaxnreg), warps 4..7 = softmax group 0, warps 8..11 = softmax
// group 1 (each softmax thread keeps a whole 128-score row in registers).  The two groups work on DIFFERENT query tiles (tile t goes to group t & 1), so
// while one group is in its MUFU-bound exp2 phase the other is in its latency-bound phase (TMEM load, row maximum,
// barriers) and the tensor pipe serves whichever group is ready ("ping-pong").  Thread r of a group owns row r of its
// tile = TMEM lane r.  Per 128-query tile and 128-key chunk c:
//     S_c  = Q K_c^T                 tcgen05.mma  M128 N128 K64  -> TMEM (one S buffer per group)
//     regs = tcgen05.ld(S_c)         the S buffer is released right here, so S_{c+1} is computed during the exp2 phase
//     P_c  = exp2(S_c*k - m)         fp32 softmax, bf16 P_c -> swizzled smem (A operand, one buffer per group)
//     O   += P_c V_c                 tcgen05.mma  M128 N64 K128  (V as MN-major B: no transpose anywhere)
// The running maximum uses the lazy-rescale rule (only when it grows by more than 2^8); the O accumulator is then
// rescaled in TMEM by the row's own thread (tcgen05.ld/st).
#include "common.cuh"

namespace plb {
namespace {

constexpr int D = 64;
constexpr int QT = 128;                     // query rows per tile
constexpr int KC = 128;                     // keys per chunk = one TMA box of K / V
constexpr int MAX_BOXES = 4;                // K / V arrive as 128-row TMA boxes: sequences up to 512 keys
constexpr int TILE16K = 128 * 64 * 2;       // one [128 x 64] bf16 swizzled tile
constexpr int NUM_THREADS = 384;
constexpr int REGS_CONTROL = 48;            // setmaxnreg budgets: TMA / MMA warpgroup ...
constexpr int REGS_SOFTMAX = 224;           // ... and the two softmax warpgroups (128*48 + 256*224 <= 64 K registers)
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units

struct FwdSmem {
  static constexpr int kK = 0;
  static constexpr int kV = kK + MAX_BOXES * TILE16K;
  static constexpr int kQ = kV + MAX_BOXES * TILE16K;   // one Q tile per group
  static constexpr int kP = kQ + 2 * TILE16K;           // one [128 x 128] P buffer (two 16 KB blocks) per group
  static constexpr int kBar = kP + 2 * 2 * TILE16K;
  // k_full[4], v_full[4], then per group: q_full, q_empty, s_full, s_free, p_ready, pv_done, o_free
  static constexpr int kNumBars = 2 * MAX_BOXES + 2 * 7;
  static constexpr int kBytes = kBar + kNumBars * 8 + 16;
  static constexpr int kAlloc = kBytes;  // no slack: the dynamic window is 1024-byte aligned (checked at run time)
};
static_assert(FwdSmem::kAlloc <= 232448, "forward attention exceeds the 227 KB shared-memory limit");

struct AttnTcParams {
  const int32_t* seq_start;
  const int32_t* q_len;
  const int32_t* kv_len;
  int heads;
  int total_rows;
  float scale_log2;  // log2(e) / sqrt(d)
  float scale;       // 1 / sqrt(d)
};

enum { Q_FULL = 0, Q_EMPTY, S_FULL, S_FREE, P_READY, PV_DONE, O_FREE, BARS_PER_GROUP };

__device__ __forceinline__ void group_bar_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); }

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnTcParams p, __nv_bfloat16* __restrict__ ctx,
                   int ld_ctx, float* __restrict__ lse2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();  // 128-byte swizzle needs 1024-byte aligned tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBar);
  uint64_t* k_full = bars;                      // [4]
  uint64_t* v_full = bars + MAX_BOXES;          // [4]
  uint64_t* gbar = bars + 2 * MAX_BOXES;        // [2][BARS_PER_GROUP]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + FwdSmem::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int row_base = p.seq_start[b], nq = p.q_len[b], nk = p.kv_len[b];
  const int H = p.heads * D;
  const int n_chunks = (nk + KC - 1) / KC;
  const int n_boxes = n_chunks;
  const int n_tiles = (nq + QT - 1) / QT;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < 2 * MAX_BOXES; ++i) mbar_init(&bars[i], 1);
    for (int g = 0; g < 2; ++g) {
      uint64_t* gb = gbar + g * BARS_PER_GROUP;
      mbar_init(&gb[Q_FULL], 1);
      mbar_init(&gb[Q_EMPTY], 1);
      mbar_init(&gb[S_FULL], 1);
      mbar_init(&gb[S_FREE], 4);    // one arrive per softmax warp
      mbar_init(&gb[P_READY], 4);
      mbar_init(&gb[PV_DONE], 1);
      mbar_init(&gb[O_FREE], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // TMEM columns: S of group g at [g*128, +128), O of group g at [256 + g*64, +64)

  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONTROL));
   if (warp == 0) {
    // ====================================== TMA producer ======================================
    if (lane == 0) {
      for (int g = 0; g < 2 && g < n_tiles; ++g) {
        uint64_t* gb = gbar + g * BARS_PER_GROUP;
        mbar_arrive_expect_tx(&gb[Q_FULL], TILE16K);
        tma_load_2d(smem + FwdSmem::kQ + g * TILE16K, &tm_qkv, &gb[Q_FULL], h * D, row_base + g * QT);
        if (g == 0) {  // first K box right behind the first Q tile so the first S MMA starts early
          mbar_arrive_expect_tx(&k_full[0], TILE16K);
          tma_load_2d(smem + FwdSmem::kK, &tm_qkv, &k_full[0], H + h * D, row_base);
        }
      }
      mbar_arrive_expect_tx(&v_full[0], TILE16K);
      tma_load_2d(smem + FwdSmem::kV, &tm_qkv, &v_full[0], 2 * H + h * D, row_base);
      for (int j = 1; j < n_boxes; ++j) {
        mbar_arrive_expect_tx(&k_full[j], TILE16K);
        tma_load_2d(smem + FwdSmem::kK + j * TILE16K, &tm_qkv, &k_full[j], H + h * D, row_base + j * 128);
        mbar_arrive_expect_tx(&v_full[j], TILE16K);
        tma_load_2d(smem + FwdSmem::kV + j * TILE16K, &tm_qkv, &v_full[j], 2 * H + h * D, row_base + j * 128);
      }
      for (int t = 2; t < n_tiles; ++t) {
        const int g = t & 1, k = t >> 1;  // k-th tile of group g
        uint64_t* gb = gbar + g * BARS_PER_GROUP;
        mbar_wait(&gb[Q_EMPTY], (k - 1) & 1);  // every S MMA of the group's previous tile has read Q
        mbar_arrive_expect_tx(&gb[Q_FULL], TILE16K);
        tma_load_2d(smem + FwdSmem::kQ + g * TILE16K, &tm_qkv, &gb[Q_FULL], h * D, row_base + t * QT);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, KC, 0, 0);   // S = Q K^T : A K-major, B K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, D, 0, 1);    // O = P V   : A K-major, B MN-major
      const uint32_t sK = smem_u32(smem + FwdSmem::kK), sV = smem_u32(smem + FwdSmem::kV);
      // Two op streams per group: S (at most one chunk ahead: one S buffer) and PV.
      //   ks/js/cs : tile, chunk-in-tile and running count of the next S;  kp/jp/cp : same for the next PV
      int ks[2] = {0, 0}, js[2] = {0, 0}, kp[2] = {0, 0}, jp[2] = {0, 0};
      uint32_t cs[2] = {0, 0}, cp[2] = {0, 0};
      const int tiles_of[2] = {(n_tiles + 1) >> 1, n_tiles >> 1};
      while (kp[0] < tiles_of[0] || kp[1] < tiles_of[1]) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint64_t* gb = gbar + g * BARS_PER_GROUP;
          // ---- O += P_j V_j ----
          if (kp[g] < tiles_of[g] && mbar_test(&gb[P_READY], cp[g] & 1) &&
              (jp[g] > 0 || mbar_test(&gb[O_FREE], (kp[g] & 1) ^ 1))) {
            const int j = jp[g];
            mbar_wait(&v_full[j], 0);
            tc_fence_after();
            const uint32_t sP = smem_u32(smem + FwdSmem::kP + g * 2 * TILE16K);
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk)
              umma_bf16(tmem_base + 256 + g * D, umma_smem_desc(sP + (kk >> 2) * TILE16K + (kk & 3) * 32, 16, 1024),
                        umma_smem_desc(sV + j * TILE16K + kk * 2048, TILE16K, 1024), idesc_o, (j > 0 || kk > 0));
            umma_commit(&gb[PV_DONE]);
            ++cp[g];
            if (++jp[g] == n_chunks) { jp[g] = 0; ++kp[g]; }
          }
          // ---- S_j = Q K_j^T as soon as the group has pulled the previous S out of TMEM ----
          if (ks[g] < tiles_of[g] &&
              (cs[g] == 0 || mbar_test(&gb[S_FREE], (cs[g] - 1) & 1)) &&
              (js[g] > 0 || mbar_test(&gb[Q_FULL], ks[g] & 1))) {
            const int j = js[g];
            mbar_wait(&k_full[j], 0);
            tc_fence_after();
            const uint32_t sQ = smem_u32(smem + FwdSmem::kQ + g * TILE16K);
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)
              umma_bf16(tmem_base + g * KC, umma_smem_desc(sQ + kk * 32, 16, 1024),
                        umma_smem_desc(sK + j * TILE16K + kk * 32, 16, 1024), idesc_s, kk > 0);
            umma_commit(&gb[S_FULL]);
            ++cs[g];
            if (++js[g] == n_chunks) {
              umma_commit(&gb[Q_EMPTY]);
              js[g] = 0;
              ++ks[g];
            }
          }
        }
      }
    }
    __syncwarp();
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
    // ====================================== softmax + epilogue (two groups) ======================================
    const int grp = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int gtid = threadIdx.x - 128 - grp * 128;  // 0..127 within the group
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_addr + grp * KC;
    const uint32_t tmem_o = tmem_base + lane_addr + 256 + grp * D;
    uint64_t* gb = gbar + grp * BARS_PER_GROUP;
    const uint32_t sPg = smem_u32(smem + FwdSmem::kP + grp * 2 * TILE16K);
    const uint32_t sP_row = sPg + row * 128;
    uint32_t c = 0;  // chunk counter of this group (phase of s_full / s_free / p_ready / pv_done)
    for (int t = grp; t < n_tiles; t += 2) {
      float m_used = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_chunks; ++j, ++c) {
        mbar_wait(&gb[S_FULL], c & 1);
        tc_fence_after();
        // pull the 128 scores of this row out of TMEM and hand the S buffer straight back to the MMA warp, so that
        // S_{c+1} is computed while this group is busy with the exponentials of chunk c
        uint32_t sv[KC];
#pragma unroll
        for (int q = 0; q < KC / 32; ++q) tmem_ld_32x32(tmem_s + q * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[q * 32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gb[S_FREE]);
        const int valid = nk - j * KC;  // only the last chunk can hold keys >= kv_len -> -inf
        if (valid < KC) {
#pragma unroll
          for (int i = 0; i < KC; ++i)
            if (i >= valid) sv[i] = 0xff800000u;
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < KC; i += 4) {
          mx[0] = fmaxf(mx[0], __uint_as_float(sv[i]));
          mx[1] = fmaxf(mx[1], __uint_as_float(sv[i + 1]));
          mx[2] = fmaxf(mx[2], __uint_as_float(sv[i + 2]));
          mx[3] = fmaxf(mx[3], __uint_as_float(sv[i + 3]));
        }
        const float cmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        const float m_new = fmaxf(m_used, cmax * p.scale_log2);
        // lazy rescale: keep the old reference maximum unless some row of this warp outgrew it by 2^8
        const bool grow = (m_new > m_used + RESCALE_THRESHOLD);
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = ex2_fast(m_used - m_new);  // first chunk: exp2(-inf) = 0
          m_used = m_new;
          l_run *= alpha;
          if (j > 0) {
            mbar_wait(&gb[PV_DONE], (c - 1) & 1);  // O += P_{j-1} V_{j-1} has landed
            tc_fence_after();
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t o[32];
              tmem_ld_32x32(tmem_o + q * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32(tmem_o + q * 32, o);
            }
            tmem_st_wait();
          }
        }
        // P = exp2(S*k - m) -> bf16 -> swizzled smem (A operand of the PV MMA).  One P buffer per group: the previous
        // PV (issued about one chunk of work ago) must have finished reading it.
        if (c >= 1) mbar_wait(&gb[PV_DONE], (c - 1) & 1);
        float ls[2] = {0.f, 0.f};
        const float mneg = -m_used;
#pragma unroll
        for (int cc = 0; cc < KC / 8; ++cc) {  // sixteen 16-byte pieces = 128 keys (masked keys: exp2(-inf) = 0)
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = ex2_fast(fmaf(__uint_as_float(sv[cc * 8 + i]), p.scale_log2, mneg));
          ls[cc & 1] += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
          uint4 o;
          o.x = pack_bf16x2(e[0], e[1]); o.y = pack_bf16x2(e[2], e[3]);
          o.z = pack_bf16x2(e[4], e[5]); o.w = pack_bf16x2(e[6], e[7]);
          sts128(sP_row + (cc >> 3) * TILE16K + (((cc & 7) ^ (row & 7)) << 4), o);
        }
        l_run += ls[0] + ls[1];
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gb[P_READY]);
      }
      // ---- epilogue of the tile: O / l -> bf16 -> staging (the group's P buffer) -> global; LSE ----
      mbar_wait(&gb[PV_DONE], (c - 1) & 1);  // all PV MMAs of the tile are complete: O final, P buffer dead
      tc_fence_after();
      const float inv = 1.f / l_run;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t o[32];
        tmem_ld_32x32(tmem_o + q * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[cc * 8 + 0]) * inv, __uint_as_float(o[cc * 8 + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(o[cc * 8 + 2]) * inv, __uint_as_float(o[cc * 8 + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(o[cc * 8 + 4]) * inv, __uint_as_float(o[cc * 8 + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(o[cc * 8 + 6]) * inv, __uint_as_float(o[cc * 8 + 7]) * inv);
          sts128(sP_row + (((q * 4 + cc) ^ (row & 7)) << 4), v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&gb[O_FREE]);
      const int qrow = t * QT + row;
      if (qrow < nq) lse2[static_cast<size_t>(h) * p.total_rows + row_base + qrow] = m_used + log2f(l_run);
      group_bar_sync(grp);
      // coalesced copy-out: 128 rows x 8 pieces, guarded by the sequence's row count
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = gtid + i * 128;
        const int r = idx >> 3, cc = idx & 7;
        if (t * QT + r < nq) {
          const uint4 v = lds128(sPg + r * 128 + ((cc ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(ctx + static_cast<size_t>(row_base + t * QT + r) * ld_ctx + h * D + cc * 8) = v;
        }
      }
      group_bar_sync(grp);  // staging fully read before the next tile's P overwrites it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace plb

using namespace plb;

extern "C" int plb_attn_fwd_tc(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                               const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                               void* ctx, int ld_ctx, float* lse2, plb_stream_t stream) {
  PLB_REQUIRE(qkv && seq_start && q_len && kv_len && ctx && lse2, PLB_ERR_ARG, "attn_fwd_tc: null pointer");
  PLB_REQUIRE(head_dim == D, PLB_ERR_UNSUPPORTED, "attn_fwd_tc: head_dim %d not supported (only 64)", head_dim);
  PLB_REQUIRE(B > 0 && heads > 0 && total_rows > 0, PLB_ERR_ARG, "attn_fwd_tc: bad sizes");
  PLB_REQUIRE(max_q_len > 0 && max_q_len <= MAX_BOXES * 128, PLB_ERR_UNSUPPORTED,
              "attn_fwd_tc: sequences longer than %d rows are not supported (got %d)", MAX_BOXES * 128, max_q_len);
  PLB_REQUIRE(ld_qkv % 8 == 0 && ld_ctx % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(ctx) % 16) == 0,
              PLB_ERR_ARG, "attn_fwd_tc: operands must be 16-byte aligned with ld %% 8 == 0");
  CUtensorMap tm;
  if (!tensor_map_bf16(&tm, qkv, total_rows, 3 * heads * D, ld_qkv, 128, 64)) return PLB_ERR_CUDA;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kAlloc);
    PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  AttnTcParams p{seq_start, q_len, kv_len, heads, total_rows, 1.4426950408889634f / sqrtf((float)head_dim),
                 1.f / sqrtf((float)head_dim)};
  attn_fwd_tc_kernel<<<B * heads, NUM_THREADS, FwdSmem::kAlloc, static_cast<cudaStream_t>(stream)>>>(
      tm, p, static_cast<__nv_bfloat16*>(ctx), ld_ctx, lse2);
  return check_launch("plb_attn_fwd_tc");
}
