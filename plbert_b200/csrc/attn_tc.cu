// plbert_b200 — fused key-padded attention forward on the 5th-generation tensor cores (tcgen05.mma, S / O
// accumulators in TMEM), head_dim 64, sequences up to 512 rows.
//
// Replaces HF modeling_albert.py:184-197 (SDPA with the [B,1,L,L] mask of masking_utils.py:495-520).  One CTA owns one
// (sequence, head): K and V of that head (<= 2 x 64 KB bf16) are TMA-loaded ONCE into 128-byte-swizzled shared memory
// and stay resident while the CTA walks the query tiles, so K/V leave HBM exactly once per head.
//
// Warp roles (12 warps = 3 warpgroups): warp 0 = TMA producer, warps 1 / 2 = MMA issuer of softmax group 0 / 1 (one elected
// thread each, blocking on the barrier that is due next; warp 3 idle; the warpgroup gives its registers away with
// setmaxnreg), warps 4..7 = softmax group 0, warps 8..11 = softmax group 1 (each softmax thread keeps a whole 128-score
// row in registers).  The two groups work on DIFFERENT query tiles (tile t goes to group t & 1), so one group's exponentials
// overlap the other's latency-bound phases (TMEM load, row maximum, barriers) and MMAs.  Thread r of a group owns row r of its
// tile = TMEM lane r.  Per 128-query tile and 128-key chunk c:
//     S_c  = Q K_c^T                 tcgen05.mma  M128 N128 K64  -> TMEM (one S buffer per group)
//     regs = tcgen05.ld(S_c)         the S buffer is released right here, so S_{c+1} is computed during the exp2 phase
//     P_c  = exp2(S_c*k - m)         fp32 softmax, bf16 pairs -> TENSOR MEMORY (tcgen05.st), no shared-memory round trip
//     O   += P_c V_c                 tcgen05.mma  M128 N64 K128, A = P_c from TMEM, B = V as MN-major smem (no transpose)
// The running maximum uses the lazy-rescale rule (only when it grows by more than 2^8); the O accumulator is then
// rescaled in TMEM by the row's own thread (tcgen05.ld/st).  A finished tile's O is taken out of TMEM inside the group's NEXT
// chunk (deferred epilogue) and leaves the SM as one 4 KB TMA store per warp.
#include "common.cuh"
#include "rng.cuh"

#include <cstdlib>

namespace plb {
namespace {

constexpr int D = 64;
constexpr int QT = 128;                     // query rows per tile
constexpr int KC = 128;                     // keys per chunk = one TMA box of K / V
constexpr int MAX_BOXES = 4;                // K / V arrive as 128-row TMA boxes: sequences up to 512 keys
constexpr int TILE16K = 128 * 64 * 2;       // one [128 x 64] bf16 swizzled tile
constexpr int NUM_THREADS = 384;
constexpr int BWD_THREADS = 640;            // backward kernels: control warpgroup + 2 groups x 8 warps (two column halves per row)
constexpr int BWD_REGS_CONTROL = 56;        // setmaxnreg only moves registers inside the launch allocation (640 x 96 = 61440):
constexpr int BWD_REGS_SOFTMAX = 104;       // 128*56 + 512*104 = 60416
constexpr int REGS_CONTROL = 56;            // setmaxnreg budgets: TMA / MMA warpgroup ...
constexpr int REGS_SOFTMAX = 224;           // ... and the two softmax warpgroups (128*56 + 256*224 <= 64 K registers)
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units

constexpr int BARS_PER_GROUP_FWD = 9;  // q_full[2], q_empty[2], s_full, s_free, p_ready, pv_done, o_free
struct FwdSmem {
  static constexpr int kK = 0;
  static constexpr int kV = kK + MAX_BOXES * TILE16K;
  static constexpr int kQ = kV + MAX_BOXES * TILE16K;   // two Q tiles per group (the next tile is prefetched)
  static constexpr int kP = kQ + 4 * TILE16K;           // output staging [128 x 64] per group (P itself lives in TMEM)
  static constexpr int kBar = kP + 2 * TILE16K;
  // k_full[4], v_full[4], then per group: q_full, q_empty, s_full, s_free, p_ready, pv_done, o_free
  static constexpr int kNumBars = 2 * MAX_BOXES + 2 * BARS_PER_GROUP_FWD + MAX_BOXES;   // ... then kv_free[4]
  static constexpr int kBytes = kBar + kNumBars * 8 + 16;  // (+ the TMEM base address word)
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

enum { Q_FULL = 0, Q_FULL1, Q_EMPTY, Q_EMPTY1, S_FULL, S_FREE, P_READY, PV_DONE, O_FREE, BARS_PER_GROUP };  // Q: two slots per group
static_assert(BARS_PER_GROUP == BARS_PER_GROUP_FWD, "barrier layout");

__device__ __forceinline__ void group_bar_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); }
__device__ __forceinline__ void group_bar_sync256(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory"); }

// Turn-taking on the exponential phase.  The two groups' warps of one TMEM lane quadrant share an SM sub-partition and
// its MUFU; left alone the groups run in lock-step (both in their TMEM / barrier / store phase, then both fighting for
// the MUFU, with their MMA bursts colliding too).  Each quadrant pair alternates through two 64-thread named barriers:
// group 0 runs step c after group 1 finished step c-1, group 1 runs step c after group 0 finished step c.  ptxas
// schedules freely around bar.sync / bar.arrive unless data flows through them, so acquire returns a volatile zero
// from shared memory (add it to an input of the phase) and release stores a value derived from the phase's results.
//   smem: one zero word at `turn_addr`, then one sink word per softmax warp.
//   kPair = threads of one quadrant pair (64: one warp per group and quadrant; 128: two warps — the column halves).
#ifndef PLB_TURN_BWD
#define PLB_TURN_BWD 1
#endif
template <int kPair = 64>
__device__ __forceinline__ float turn_acquire(int grp, int quad, uint32_t c, uint32_t steps1, uint32_t turn_addr) {
  float z = 0.f;
  if (kPair == 128 && !PLB_TURN_BWD) return z;
  if (grp == 0 ? (c >= 1 && c - 1 < steps1) : true)
    asm volatile("bar.sync %1, %3;\n\tld.volatile.shared.f32 %0, [%2];" : "=f"(z) : "r"(3 + grp * 4 + quad), "r"(turn_addr), "n"(kPair) : "memory");
  return z;
}
template <int kPair = 64>
__device__ __forceinline__ void turn_release(int grp, int quad, uint32_t c, uint32_t steps0, uint32_t steps1, uint32_t turn_addr,
                                             uint32_t result_bits) {
  if (kPair == 128 && !PLB_TURN_BWD) return;
  if (grp == 0 ? (c < steps1) : (c + 1 < steps0))
    asm volatile("st.volatile.shared.b32 [%0], %1;\n\tbar.arrive %2, %3;" ::"r"(turn_addr + 4 + ((threadIdx.x >> 5) - 4) * 4), "r"(result_bits),
                 "r"(3 + (grp ^ 1) * 4 + quad), "n"(kPair) : "memory");
}

// Optional timeline instrumentation (profiles/attn_trace.py builds a separate library with -DPLB_ATTN_TRACE): lane 0 of
// every warp of CTA 0 appends (clock, warp, code, index) records.  Compiled out of the product library.
#ifdef PLB_ATTN_TRACE
#ifndef PLB_TRACE_BLOCK
#define PLB_TRACE_BLOCK 0   // which CTA records (a mid-run block shows the steady state, block 0 the cold first wave)
#endif
constexpr unsigned TRACE_SLOTS = 512;  // records per warp; plain stores, no atomics, so the timeline stays undisturbed
constexpr unsigned TRACE_CAP = 3 * 32 * TRACE_SLOTS;  // kernel id (0 fwd, 1 dQ, 2 dK/dV) x warp x slot
__device__ unsigned long long g_trace[TRACE_CAP];
#define PLB_TRACE_INIT(kid) unsigned trace_k = 0; const unsigned trace_w = (kid) * 32 + (threadIdx.x >> 5)
#define PLB_TRACE(code, idx)                                                                                          \
  do {                                                                                                                \
    if (blockIdx.x == PLB_TRACE_BLOCK && (threadIdx.x & 31) == 0 && trace_k < TRACE_SLOTS)                                          \
      g_trace[trace_w * TRACE_SLOTS + trace_k++] = (static_cast<unsigned long long>(clock64()) << 24) |               \
                                                   (trace_w << 16) | ((code) << 8) | ((idx) & 255);                   \
  } while (0)
#else
#define PLB_TRACE_INIT(kid) do { } while (0)
#define PLB_TRACE(code, idx) do { } while (0)
#endif

// Work item of a persistent forward CTA: one (sequence, head).
struct FwdItem {
  int h, row_base, nq, nk, n_chunks, n_tiles;
};
__device__ __forceinline__ FwdItem fwd_item(const AttnTcParams& p, int item) {
  FwdItem it;
  const int b = item / p.heads;
  it.h = item % p.heads;
  it.row_base = p.seq_start[b];
  it.nq = p.q_len[b];
  it.nk = p.kv_len[b];
  // (lengths are >= 1 by contract; the clamps only keep the roles' barrier bookkeeping consistent on bad input)
  it.n_chunks = max(1, (it.nk + KC - 1) / KC);  // = K / V boxes of the item
  it.n_tiles = max(1, (it.nq + QT - 1) / QT);
  return it;
}

// PERSISTENT: the grid is one CTA per SM and every CTA walks the items blockIdx.x, blockIdx.x + gridDim.x, ...  All
// barriers keep running across items, and K / V box j is refilled with the NEXT item's rows as soon as both groups'
// last tiles of the current item have passed chunk j (kv_free[j]), so the load latency, the pipeline fill and the
// epilogue of one item overlap the tail / head of its neighbours instead of idling the SM once per (sequence, head).
// kDrop: nn.Dropout on the attention probabilities (HF :134 / SDPA dropout_p, HF :194): the P that multiplies V is
// P o m with m = 0 or 1 / (1 - p) from the counter-based generator (rng.cuh; pair (k, k + 1) of query row q); the row sum
// that normalises it is the undropped one, exactly like dropout(softmax(S)) V.
template <bool kDrop>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_ctx, const AttnTcParams p, int n_items,
                   __nv_bfloat16* __restrict__ ctx, int ld_ctx, float* __restrict__ lse2, const DropoutParams drop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();  // 128-byte swizzle needs 1024-byte aligned tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBar);
  uint64_t* k_full = bars;                      // [4]
  uint64_t* v_full = bars + MAX_BOXES;          // [4]
  uint64_t* gbar = bars + 2 * MAX_BOXES;        // [2][BARS_PER_GROUP]
  uint64_t* kv_free = gbar + 2 * BARS_PER_GROUP;  // [4]: both groups are done with K_j / V_j of the current item
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + FwdSmem::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PLB_TRACE_INIT(0);
  PLB_TRACE(0, 0);
  const int H = p.heads * D;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_ctx);
    for (int i = 0; i < 2 * MAX_BOXES; ++i) mbar_init(&bars[i], 1);
    for (int g = 0; g < 2; ++g) {
      uint64_t* gb = gbar + g * BARS_PER_GROUP;
      for (int sl = 0; sl < 2; ++sl) {
        mbar_init(&gb[Q_FULL + sl], 1);
        mbar_init(&gb[Q_EMPTY + sl], 1);
      }
      mbar_init(&gb[S_FULL], 1);
      mbar_init(&gb[S_FREE], 4);    // one arrive per softmax warp
      mbar_init(&gb[P_READY], 4);
      mbar_init(&gb[PV_DONE], 1);
      mbar_init(&gb[O_FREE], 4);
    }
    for (int j = 0; j < MAX_BOXES; ++j) mbar_init(&kv_free[j], 2);  // one arrive per group's MMA issuer
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_grid_sync();   // the prologue above is CTA-local; everything below reads / writes global memory
  // TMEM columns per group g (stride 256): S [+0, +128), O [+128, +64), P as packed bf16 pairs [+192, +64)

  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONTROL));
   if (warp == 0) {
    // ====================================== TMA producer ======================================
    if (elect_one_sync()) {
      uint32_t box_loads[MAX_BOXES] = {0, 0, 0, 0};  // how often box j has been filled so far
      uint32_t q_loads[2] = {0, 0};                   // Q tiles handed to each group so far
      FwdItem nxt = blockIdx.x < n_items ? fwd_item(p, blockIdx.x) : FwdItem{};
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const FwdItem it = nxt;
        if (item + static_cast<int>(gridDim.x) < n_items) nxt = fwd_item(p, item + gridDim.x);
        auto load_q = [&](int t) {
          const int g = t & 1;
          uint64_t* gb = gbar + g * BARS_PER_GROUP;
          const uint32_t n = q_loads[g], sl = n & 1;  // two Q slots per group: the tile after the current one is prefetched
          if (n >= 2) mbar_wait(&gb[Q_EMPTY + sl], ((n >> 1) - 1) & 1);  // every S MMA of the slot's previous tile has read Q
          mbar_arrive_expect_tx(&gb[Q_FULL + sl], TILE16K);
          tma_load_2d(smem + FwdSmem::kQ + (g * 2 + sl) * TILE16K, &tm_qkv, &gb[Q_FULL + sl], it.h * D, it.row_base + t * QT);
          ++q_loads[g];
        };
        auto wait_box = [&](int j) {  // both groups are done with the previous item's K_j / V_j
          if (box_loads[j] >= 1) mbar_wait(&kv_free[j], (box_loads[j] - 1) & 1);
          ++box_loads[j];
        };
        auto load_k = [&](int j) {
          mbar_arrive_expect_tx(&k_full[j], TILE16K);
          tma_load_2d(smem + FwdSmem::kK + j * TILE16K, &tm_qkv, &k_full[j], H + it.h * D, it.row_base + j * 128);
        };
        auto load_v = [&](int j) {
          mbar_arrive_expect_tx(&v_full[j], TILE16K);
          tma_load_2d(smem + FwdSmem::kV + j * TILE16K, &tm_qkv, &v_full[j], 2 * H + it.h * D, it.row_base + j * 128);
        };
        // first Q tile and first K box first so that the first S MMA starts early
        load_q(0);
        wait_box(0);
        load_k(0);
        if (it.n_tiles > 1) load_q(1);
        load_v(0);
        for (int j = 1; j < it.n_chunks; ++j) {
          wait_box(j);
          load_k(j);
          load_v(j);
        }
        for (int t = 2; t < it.n_tiles; ++t) load_q(t);
      }
    }
    __syncwarp();
   } else if ((warp == 1 || warp == 2) && elect_one_sync()) {
    // ====================================== MMA issuers: one thread per group ======================================
    // A group's MMAs alternate deterministically — S of chunk c+1 becomes issuable when the group has pulled S_c out of
    // TMEM, P V of chunk c when P_c is in tensor memory — so the issuer blocks on exactly the barrier that is due next.
    const int g = warp - 1;
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, KC, 0, 0);   // S = Q K^T : A K-major, B K-major
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, D, 0, 1);    // O = P V   : A from TMEM, B MN-major
    uint64_t* gb = gbar + g * BARS_PER_GROUP;
    const uint32_t tmem_sg = tmem_base + g * 256, tmem_og = tmem_sg + 128, tmem_pg = tmem_sg + 192;
    // descriptor bases; the start-address field counts 16-byte units, so advancing is a plain add
    const uint64_t dQ = umma_smem_desc(smem_u32(smem + FwdSmem::kQ + g * 2 * TILE16K), 16, 1024);  // slot 0; slot 1 is 16 KB further
    const uint64_t dK = umma_smem_desc(smem_u32(smem + FwdSmem::kK), 16, 1024);
    const uint64_t dV = umma_smem_desc(smem_u32(smem + FwdSmem::kV), TILE16K, 1024);
    uint32_t box_uses[MAX_BOXES] = {0, 0, 0, 0};  // items that have used box j so far (phase of k_full / v_full)
    uint32_t kt = 0;                              // tiles of this group so far (phase of Q_FULL / O_FREE)
    uint32_t c = 0;                               // chunks of this group so far (phase of S_FREE / P_READY)
    // position of the NEXT S to issue (runs one chunk ahead of the PV stream, across tiles and items)
    int s_item = blockIdx.x, s_k = 0, s_j = 0;
    FwdItem s_it = s_item < n_items ? fwd_item(p, s_item) : FwdItem{};
    uint32_t s_kt = 0;
    uint32_t s_box_uses[MAX_BOXES] = {0, 0, 0, 0};
    auto s_tiles = [&]() { return g == 0 ? (s_it.n_tiles + 1) >> 1 : s_it.n_tiles >> 1; };
    uint32_t s_issued = 0;                        // S chunks issued so far (the PV stream is at chunk c <= s_issued)
    bool s_skipped = false;                       // the cursor's last move passed items in which this group has no tile
    auto s_skip_empty = [&]() {  // advance the S cursor past items in which this group has no tile
      s_skipped = false;
      while (s_item < n_items && s_tiles() == 0) {
        for (int j = 0; j < s_it.n_chunks; ++j) ++s_box_uses[j];
        s_item += gridDim.x;
        if (s_item < n_items) s_it = fwd_item(p, s_item);
        s_skipped = true;
      }
    };
    auto issue_s = [&]() {  // S_j = Q K_j^T at the cursor, then advance the cursor
      const uint32_t sl = s_kt & 1;
      if (s_j == 0) mbar_wait(&gb[Q_FULL + sl], (s_kt >> 1) & 1);
      if (s_k == 0) mbar_wait(&k_full[s_j], s_box_uses[s_j] & 1);
      tc_fence_after();
      const uint64_t kb = dK + static_cast<uint64_t>(s_j * (TILE16K >> 4)), qb = dQ + static_cast<uint64_t>(sl * (TILE16K >> 4));
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) umma_bf16(tmem_sg, qb + 2 * kk, kb + 2 * kk, idesc_s, kk > 0);
      umma_commit(&gb[S_FULL]);
      ++s_issued;
      if (s_j == s_it.n_chunks - 1) umma_commit(&gb[Q_EMPTY + sl]);
      PLB_TRACE(1, s_j);
      if (++s_j == s_it.n_chunks) {
        s_j = 0;
        ++s_kt;
        if (++s_k == s_tiles()) {
          s_k = 0;
          for (int j = 0; j < s_it.n_chunks; ++j) ++s_box_uses[j];
          s_item += gridDim.x;
          if (s_item < n_items) s_it = fwd_item(p, s_item);
          s_skip_empty();
        }
      }
    };
    s_skip_empty();
    FwdItem nxt = blockIdx.x < n_items ? fwd_item(p, blockIdx.x) : FwdItem{};
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const FwdItem it = nxt;
      if (item + static_cast<int>(gridDim.x) < n_items) nxt = fwd_item(p, item + gridDim.x);
      const int my_tiles = g == 0 ? (it.n_tiles + 1) >> 1 : it.n_tiles >> 1;
      if (my_tiles == 0) {
        // this group has nothing to do in the item: release its share of the K / V boxes — but only once the item's boxes
        // have LANDED.  The producer refills box j only after both groups' arrivals of the previous item, so waiting for the
        // fill keeps this group's arrivals in item order; arriving right away let a group without tiles in several
        // consecutive short items run ahead, complete kv_free phases on its own (the other group's K / V overwritten under
        // it) and push the barrier two phases past the producer's parity wait (deadlock -> trap).
        for (int j = 0; j < it.n_chunks; ++j) {
          mbar_wait(&k_full[j], box_uses[j] & 1);
          mbar_wait(&v_full[j], box_uses[j] & 1);
          mbar_arrive(&kv_free[j]);
          ++box_uses[j];
        }
        continue;
      }
      for (int k = 0; k < my_tiles; ++k, ++kt) {
        for (int j = 0; j < it.n_chunks; ++j, ++c) {
          const bool next_in_tile = j + 1 < it.n_chunks;
          if (s_issued == c) {
            // S of THIS chunk is still to be issued: the very first chunk of the CTA, or the cursor had passed items without
            // a tile for this group — their K / V boxes had to be released (above, in item order) before blocking on this
            // item's loads, because the producer refills a box only after both groups' releases of its previous item
            if (c > 0) mbar_wait(&gb[S_FREE], (c - 1) & 1);
            issue_s();
          }
          if (next_in_tile) {
            mbar_wait(&gb[S_FREE], c & 1);
            issue_s();
          }
          // ---- O += P_j V_j ----
          mbar_wait(&gb[P_READY], c & 1);
          if (j == 0) mbar_wait(&gb[O_FREE], (kt & 1) ^ 1);
          if (k == 0) mbar_wait(&v_full[j], box_uses[j] & 1);
          tc_fence_after();
          const uint64_t vb = dV + static_cast<uint64_t>(j * (TILE16K >> 4));
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk)
            umma_bf16_tmem_a(tmem_og, tmem_pg + 8 * kk, vb + 128 * kk, idesc_o, (j > 0 || kk > 0));
          umma_commit(&gb[PV_DONE]);
          if (k == my_tiles - 1) umma_commit(&kv_free[j]);  // this group's last use of K_j / V_j in the item
          PLB_TRACE(2, c);
          if (!next_in_tile && s_item < n_items && !s_skipped) {  // next tile (possibly of the next item): its Q arrives after this tile's last S
            mbar_wait(&gb[S_FREE], c & 1);
            issue_s();
          }
        }
      }
      for (int j = 0; j < it.n_chunks; ++j) ++box_uses[j];
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
    // ====================================== softmax + epilogue (two groups) ======================================
    const int grp = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_addr + grp * 256;
    const uint32_t tmem_o = tmem_s + 128, tmem_p = tmem_s + 192;
    uint64_t* gb = gbar + grp * BARS_PER_GROUP;
    uint32_t c = 0;  // chunk counter of this group over all items (phase of s_full / s_free / p_ready / pv_done)
    // Deferred tile epilogue: the finished tile's O stays in TMEM while the group already runs the load / maximum / exponentials of
    // its NEXT tile's first chunk (whose S the issuer computed right after the tile's last PV); O is pulled out just before that
    // chunk's P goes to tensor memory and is converted / stored while the chunk's PV MMA runs.
    bool pend = false;
    int pend_grow = -1, pend_h = 0;  // this thread's output row (global row index, -1: beyond the sequence) and head
    bool pend_wfull = false;         // all 32 rows of the warp are rows of the sequence: they leave with one TMA store
    float pend_inv = 0.f, pend_lse = 0.f;
    const uint32_t sP_warp = smem_u32(smem + FwdSmem::kP + grp * TILE16K + quad * 4096);  // the warp's 32 staging rows (1024-byte aligned)
    auto pend_take = [&](uint32_t (&o)[64]) {  // the caller has waited for PV_DONE of the tile's last chunk
      tc_fence_after();
      tmem_ld_32x32(tmem_o, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
      tmem_ld_32x32(tmem_o + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&gb[O_FREE]);
    };
    auto pend_store = [&](const uint32_t (&o)[64]) {
      pend = false;
      if (pend_wfull) {
        // bf16 rows -> the warp's swizzled staging rows -> ONE 4 KB TMA store (32 rows x 128 bytes).  Per-lane global stores of the
        // rows (32 different lines per instruction) keep the SM's load/store pipe busy for ~300 clk each and every mbarrier operation
        // of the CTA queues behind them: 253 -> 240 us per launch at 192 x 512 x 12 (profiles/r2_attn_fwd.md)
        if (lane == 0) bulk_wait_read<0>();  // the warp's previous store has read the staging rows
        __syncwarp();
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[cc * 8 + 0]) * pend_inv, __uint_as_float(o[cc * 8 + 1]) * pend_inv);
          v.y = pack_bf16x2(__uint_as_float(o[cc * 8 + 2]) * pend_inv, __uint_as_float(o[cc * 8 + 3]) * pend_inv);
          v.z = pack_bf16x2(__uint_as_float(o[cc * 8 + 4]) * pend_inv, __uint_as_float(o[cc * 8 + 5]) * pend_inv);
          v.w = pack_bf16x2(__uint_as_float(o[cc * 8 + 6]) * pend_inv, __uint_as_float(o[cc * 8 + 7]) * pend_inv);
          sts128(sP_warp + lane * 128 + ((cc ^ (lane & 7)) << 4), v);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {  // lane 0's row is the warp's first
          tma_store_2d(&tm_ctx, smem + FwdSmem::kP + grp * TILE16K + quad * 4096, pend_h * D, pend_grow);
          bulk_commit();
        }
        lse2[static_cast<size_t>(pend_h) * p.total_rows + pend_grow] = pend_lse;
      } else if (pend_grow >= 0) {
        // ragged end of a sequence: the thread writes its own row (128 contiguous bytes) as four 32-byte stores
        __nv_bfloat16* orow = ctx + static_cast<size_t>(pend_grow) * ld_ctx + pend_h * D;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            w[i] = pack_bf16x2(__uint_as_float(o[cc * 16 + 2 * i]) * pend_inv, __uint_as_float(o[cc * 16 + 2 * i + 1]) * pend_inv);
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + cc * 16), "r"(w[0]), "r"(w[1]),
                       "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
        }
        lse2[static_cast<size_t>(pend_h) * p.total_rows + pend_grow] = pend_lse;
      }
    };
    FwdItem nxt = blockIdx.x < n_items ? fwd_item(p, blockIdx.x) : FwdItem{};
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
     const FwdItem it = nxt;
     if (item + static_cast<int>(gridDim.x) < n_items) nxt = fwd_item(p, item + gridDim.x);  // its three scalar loads are in flight during this item
     const int h = it.h, row_base = it.row_base, nq = it.nq, nk = it.nk, n_chunks = it.n_chunks, n_tiles = it.n_tiles;
     for (int t = grp; t < n_tiles; t += 2) {
      float m_used = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_chunks; ++j, ++c) {
        mbar_wait(&gb[S_FULL], c & 1);
        tc_fence_after();
        PLB_TRACE(10, c);
        // pull the 128 scores of this row out of TMEM and hand the S buffer straight back to the MMA warp, so that
        // S_{c+1} is computed while this group is busy with the exponentials of chunk c
        uint32_t sv[KC];
        const int valid = nk - j * KC;  // only the last chunk can hold keys >= kv_len -> -inf
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int q = 0; q < KC / 32; ++q) tmem_ld_32x32(tmem_s + q * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[q * 32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gb[S_FREE]);
        PLB_TRACE(11, c);
        if (valid < KC) {
#pragma unroll
          for (int i = 0; i < KC; ++i)
            if (i >= valid) sv[i] = 0xff800000u;
        }
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
        // P = exp2(S*k - m) -> bf16, packed in registers FIRST: the exponentials then overlap the previous PV MMA,
        // which still reads the group's P columns; only the TMEM stores wait for it.
        float ls[2] = {0.f, 0.f};
        const float mneg = -m_used;
        PLB_TRACE(13, c);
        uint32_t pk[KC / 2];
        if constexpr (!kDrop) {
          // on register pairs (FFMA2 / FADD2: half the issue slots of the scale and of the row sum)
          const f32x2 sc2 = f2_splat(p.scale_log2), mn2 = f2_splat(mneg);
          f32x2 acc[4] = {f2_splat(0.f), f2_splat(0.f), f2_splat(0.f), f2_splat(0.f)};
#pragma unroll
          for (int i = 0; i < KC; i += 2) {  // masked keys: exp2(-inf) = 0
            float lo, hi, e0, e1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sc2, mn2), lo, hi);
            if (((i >> 1) & 3) == 3) {
              // Every 4th pair on the FMA pipe instead of the MUFU (16 ex2 per clock and SM, the busiest pipe of this kernel):
              // 2^x = 2^n p(f), n = round(x), f = x - n in [-0.5, 0.5], cubic p with relative error 8e-5 (bf16 rounding of P: 2e-3);
              // x >= -126 keeps the exponent arithmetic in range (masked keys: 2^-126 instead of 0, far below one bf16 ulp of O).
              // Measured at 192 x 512 x 12: 0 / 12.5 / 25 / 37.5 / 50 % of the pairs -> 235 / 228 / 217 / 234 / 260 us.
              const f32x2 xc = f2_pack(fmaxf(lo, -126.f), fmaxf(hi, -126.f));
              const f32x2 tt = f2_add(xc, f2_splat(12582912.f));                 // low mantissa bits = n
              const f32x2 ff = f2_add(xc, f2_mul(f2_add(tt, f2_splat(-12582912.f)), f2_splat(-1.f)));
              f32x2 pp = f2_fma(ff, f2_splat(0.05508868396282196f), f2_splat(0.24260404706001282f));
              pp = f2_fma(pp, ff, f2_splat(0.6932762265205383f));
              pp = f2_fma(pp, ff, f2_splat(0.9999289512634277f));
              float p0, p1, t0, t1;
              f2_unpack(pp, p0, p1);
              f2_unpack(tt, t0, t1);
              e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
              e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
            } else {
              e0 = ex2_fast(lo);
              e1 = ex2_fast(hi);
            }
            acc[(i >> 1) & 3] = f2_add(acc[(i >> 1) & 3], f2_pack(e0, e1));
            pk[i >> 1] = pack_bf16x2(e0, e1);
          }
          f2_unpack(f2_add(f2_add(acc[0], acc[1]), f2_add(acc[2], acc[3])), ls[0], ls[1]);
        } else {
#pragma unroll
          for (int cc = 0; cc < KC / 8; ++cc) {  // sixteen 16-byte pieces = 128 keys
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = ex2_fast(fmaf(__uint_as_float(sv[cc * 8 + i]), p.scale_log2, mneg));
            ls[cc & 1] += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
            const uint32_t ctr0 = attn_dropout_ctr(static_cast<uint32_t>(item), static_cast<uint32_t>(t * QT + row),
                                                   static_cast<uint32_t>(j * KC + cc * 8));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float m0, m1;
              dropout_pair(drop, ctr0 + i, m0, m1);
              e[2 * i] *= m0;
              e[2 * i + 1] *= m1;
            }
            pk[cc * 4 + 0] = pack_bf16x2(e[0], e[1]); pk[cc * 4 + 1] = pack_bf16x2(e[2], e[3]);
            pk[cc * 4 + 2] = pack_bf16x2(e[4], e[5]); pk[cc * 4 + 3] = pack_bf16x2(e[6], e[7]);
          }
        }
        PLB_TRACE(14, c);
        if (c >= 1) mbar_wait(&gb[PV_DONE], (c - 1) & 1);  // the previous PV has finished reading the P columns
        PLB_TRACE(15, c);
        // P goes to tensor memory (the A operand of O += P V is read from there): no shared-memory round trip
        auto store_p = [&]() {
          tc_fence_after();
          tmem_st_32x32(tmem_p, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
          tmem_st_32x32(tmem_p + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
          l_run += ls[0] + ls[1];
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&gb[P_READY]);
          PLB_TRACE(16, c);
        };
        if (pend) {  // (only at j == 0: the wait above covered the previous tile's last PV, its O is final)
          uint32_t o[64];
          pend_take(o);
          store_p();
          pend_store(o);
          PLB_TRACE(21, t);
        } else {
          store_p();
        }
      }
      // the tile's epilogue (O / l -> bf16 -> global, LSE) is deferred into the group's next chunk
      const int qrow = t * QT + row;
      pend = true;
      pend_grow = qrow < nq ? row_base + qrow : -1;
      pend_wfull = t * QT + quad * 32 + 32 <= nq;
      pend_h = h;
      pend_inv = 1.f / l_run;
      pend_lse = m_used + log2f(l_run);
     }
    }
    if (pend) {  // the group's last tile
      mbar_wait(&gb[PV_DONE], (c - 1) & 1);
      PLB_TRACE(20, 0);
      uint32_t o[64];
      pend_take(o);
      pend_store(o);
    }
    if (lane == 0) bulk_wait_read<0>();  // staging rows read before the CTA's shared memory goes away
  }
  tc_fence_before();
  __syncthreads();
  PLB_TRACE(99, 0);
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace plb

using namespace plb;

static int attn_fwd_launch(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len, const int32_t* kv_len,
                           int B, int heads, int head_dim, int max_q_len, int total_rows, void* ctx, int ld_ctx, float* lse2,
                           const DropoutParams& drop, plb_stream_t stream) {
  PLB_REQUIRE(qkv && seq_start && q_len && kv_len && ctx && lse2, PLB_ERR_ARG, "attn_fwd_tc: null pointer");
  PLB_REQUIRE(head_dim == D, PLB_ERR_UNSUPPORTED, "attn_fwd_tc: head_dim %d not supported (only 64)", head_dim);
  PLB_REQUIRE(B > 0 && heads > 0 && total_rows > 0, PLB_ERR_ARG, "attn_fwd_tc: bad sizes");
  PLB_REQUIRE(max_q_len > 0 && max_q_len <= MAX_BOXES * 128, PLB_ERR_UNSUPPORTED,
              "attn_fwd_tc: sequences longer than %d rows are not supported (got %d)", MAX_BOXES * 128, max_q_len);
  PLB_REQUIRE(ld_qkv % 8 == 0 && ld_ctx % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(ctx) % 16) == 0,
              PLB_ERR_ARG, "attn_fwd_tc: operands must be 16-byte aligned with ld %% 8 == 0");
  PLB_REQUIRE(drop.thresh16 == 0 || static_cast<long long>(B) * heads <= 32768, PLB_ERR_UNSUPPORTED,
              "attn_fwd_tc: dropout counters cover at most 32768 (sequence, head) items");
  CUtensorMap tm;
  if (!tensor_map_bf16(&tm, qkv, total_rows, 3 * heads * D, ld_qkv, 128, 64)) return PLB_ERR_CUDA;
  CUtensorMap tm_ctx;   // output rows leave as [32 rows x 64 columns] boxes, one per softmax warp
  if (!tensor_map_bf16(&tm_ctx, ctx, total_rows, heads * D, ld_ctx, 32, 64)) return PLB_ERR_CUDA;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kAlloc);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kAlloc);
    PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  AttnTcParams p{seq_start, q_len, kv_len, heads, total_rows, 1.4426950408889634f / sqrtf((float)head_dim),
                 1.f / sqrtf((float)head_dim)};
  const int n_items = B * heads, sms = sm_count();
  const int grid = n_items < sms ? n_items : sms;
  cudaError_t le = launch_kernel(drop.thresh16 != 0 ? attn_fwd_tc_kernel<true> : attn_fwd_tc_kernel<false>, dim3(grid),
                                 dim3(NUM_THREADS), FwdSmem::kAlloc, static_cast<cudaStream_t>(stream), tm, tm_ctx, p, n_items,
                                 static_cast<__nv_bfloat16*>(ctx), ld_ctx, lse2, drop);
  PLB_REQUIRE(le == cudaSuccess, PLB_ERR_CUDA, "plb_attn_fwd_tc: launch failed: %s", cudaGetErrorString(le));
  return check_launch("plb_attn_fwd_tc");
}

extern "C" int plb_attn_fwd_tc(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                               const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                               void* ctx, int ld_ctx, float* lse2, plb_stream_t stream) {
  return attn_fwd_launch(qkv, ld_qkv, seq_start, q_len, kv_len, B, heads, head_dim, max_q_len, total_rows, ctx, ld_ctx, lse2,
                         make_dropout(0.f, 0, 0), stream);
}

extern "C" int plb_attn_fwd_tc_dropout(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                                       const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                                       void* ctx, int ld_ctx, float* lse2, float p_drop, unsigned long long seed,
                                       unsigned int rng_stream, plb_stream_t stream) {
  PLB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, PLB_ERR_ARG, "attn_fwd_tc_dropout: p = %f outside [0, 1)", p_drop);
  return attn_fwd_launch(qkv, ld_qkv, seq_start, q_len, kv_len, B, heads, head_dim, max_q_len, total_rows, ctx, ld_ctx, lse2,
                         make_dropout(p_drop, seed, rng_stream), stream);
}

// =================================================================================================================
// Backward.  dS = P o (dP - delta), dQ = scale dS K, dK = scale dS^T Q, dV = P^T dO  (SURVEY.md appendix A), with
// P recomputed from the saved log-sum-exp.  Two kernels so that every accumulator has ONE owner and no atomics are
// needed: the dQ kernel owns query tiles, the dK/dV kernel owns key tiles; one CTA per (sequence, head), 64-wide chunks
// along the contracted sequence axis.  20 warps: a control warpgroup (TMA, elected MMA issuers that block on the barrier
// due next) and two softmax groups of EIGHT warps — two threads per row, 32 columns each, because the backward pass has
// no cross-column reduction — which take turns on the exp2 phase.  dS (dQ kernel) and P^T / dS^T (dK/dV kernel) are
// written to tensor memory as packed bf16 and consumed as the TMEM A operand of the accumulating MMAs.
// =================================================================================================================
namespace plb {
namespace {

constexpr int BC = 64;                       // chunk width (keys in the dQ kernel, queries in the dK/dV kernel)
constexpr int CHUNK8K = BC * 128;            // 64 rows of a swizzled [rows x 64] bf16 tile

// delta[h, row] = sum_d dO[row, h*64 + d] * O[row, h*64 + d]   (one warp per row, 16-byte loads)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dctx,
                                                         const __nv_bfloat16* __restrict__ ctx, int ld, int T, int heads,
                                                         int total_rows, float* __restrict__ delta) {
  const int lane = threadIdx.x & 31;
  const int chunks = heads * 8;  // 16-byte chunks per row; 8 chunks per head
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < T; row += gridDim.x * 8) {
    for (int c0 = 0; c0 < chunks; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      if (c < chunks) {
        const uint4 a = *reinterpret_cast<const uint4*>(dctx + static_cast<size_t>(row) * ld + c * 8);
        const uint4 o = *reinterpret_cast<const uint4*>(ctx + static_cast<size_t>(row) * ld + c * 8);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) acc += bf16_lo(aw[i]) * bf16_lo(ow[i]) + bf16_hi(aw[i]) * bf16_hi(ow[i]);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if ((lane & 7) == 0 && c < chunks) delta[static_cast<size_t>(c >> 3) * total_rows + row] = acc;
    }
  }
}


// Bias gradient of the QKV projection fused into the backward kernels: column sums of a staged [128 x 64] bf16 tile
// (rows < n_rows only), added to dbias[col] with atomics.  One column PAIR per lane (32-bit loads), kRowsPerWarp rows per warp.
// The sums run over the ROUNDED values, i.e. exactly what the weight-gradient GEMM reads.
template <int kRowsPerWarp = 32>
__device__ __forceinline__ void tile_colsum_atomic(uint32_t tile_smem, int n_rows, int gtid, float* dbias_cols) {
  const int cp = gtid & 31, quarter = gtid >> 5;
  float a0 = 0.f, a1 = 0.f;
  const int r_end = min(n_rows, quarter * kRowsPerWarp + kRowsPerWarp);
  for (int r = quarter * kRowsPerWarp; r < r_end; ++r) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tile_smem + r * 128 + (((cp >> 2) ^ (r & 7)) << 4) + (cp & 3) * 4));
    a0 += bf16_lo(v);
    a1 += bf16_hi(v);
  }
  if (n_rows > quarter * kRowsPerWarp) {
    atomicAdd(dbias_cols + 2 * cp, a0);
    atomicAdd(dbias_cols + 2 * cp + 1, a1);
  }
}

// ------------------------------------------------------------------------------------------------ dQ kernel
struct DqSmem {
  static constexpr int kK = 0;                               // K of the head, resident: 4 boxes of [128 x 64]
  static constexpr int kV = kK + MAX_BOXES * TILE16K;        // V of the head, resident
  static constexpr int kQ = kV + MAX_BOXES * TILE16K;        // Q tile per group
  static constexpr int kDO = kQ + 2 * TILE16K;               // dO tile per group
  static constexpr int kDS = kDO + 2 * TILE16K;              // dQ staging [128 x 64] per group (dS itself lives in TMEM)
  static constexpr int kBar = kDS + 2 * TILE16K;
  static constexpr int kNumBars = 2 * MAX_BOXES + 2 * 7;
  static constexpr int kTurn = kBar + kNumBars * 8 + 16;     // 17 words for the exponential-phase turn-taking
  static constexpr int kAlloc = kTurn + 80;
};
static_assert(DqSmem::kAlloc <= 232448, "dQ kernel exceeds the 227 KB shared-memory limit");
enum { QDO_FULL = 0, QDO_EMPTY, SD_FULL, SD_FREE, DS_READY, DQ_DONE, DQ_FREE };

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_dctx,
                      const AttnTcParams p, const float* __restrict__ lse2, const float* __restrict__ delta,
                      __nv_bfloat16* __restrict__ dqkv, int ld_dqkv, float* __restrict__ dbias) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DqSmem::kBar);
  uint64_t* k_full = bars;
  uint64_t* v_full = bars + MAX_BOXES;
  uint64_t* gbar = bars + 2 * MAX_BOXES;  // [2][7]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + DqSmem::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PLB_TRACE_INIT(1);
  PLB_TRACE(0, 0);
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int row_base = p.seq_start[b], nq = p.q_len[b], nk = p.kv_len[b];
  const int H = p.heads * D;
  const int n_chunks = (nk + BC - 1) / BC;
  const int n_boxes = (nk + 127) / 128;
  const int n_tiles = (nq + QT - 1) / QT;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_dctx);
    for (int i = 0; i < 2 * MAX_BOXES; ++i) mbar_init(&bars[i], 1);
    for (int g = 0; g < 2; ++g) {
      uint64_t* gb = gbar + g * 7;
      mbar_init(&gb[QDO_FULL], 1);
      mbar_init(&gb[QDO_EMPTY], 1);
      mbar_init(&gb[SD_FULL], 1);
      mbar_init(&gb[SD_FREE], 8);   // one arrive per softmax warp (8 per group)
      mbar_init(&gb[DS_READY], 8);
      mbar_init(&gb[DQ_DONE], 1);
      mbar_init(&gb[DQ_FREE], 8);
    }
    *reinterpret_cast<volatile float*>(smem + DqSmem::kTurn) = 0.f;
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // TMEM columns per group g (stride 224): S [+0, +64), dP [+64, +64), dQ [+128, +64), dS as packed bf16 pairs [+192, +32)

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BWD_REGS_CONTROL));
    if (warp == 0 && elect_one_sync()) {
      // ====================================== TMA producer ======================================
      auto load_tile = [&](int g, int t) {
        uint64_t* gb = gbar + g * 7;
        mbar_arrive_expect_tx(&gb[QDO_FULL], 2 * TILE16K);
        tma_load_2d(smem + DqSmem::kQ + g * TILE16K, &tm_qkv, &gb[QDO_FULL], h * D, row_base + t * QT);
        tma_load_2d(smem + DqSmem::kDO + g * TILE16K, &tm_dctx, &gb[QDO_FULL], h * D, row_base + t * QT);
      };
      load_tile(0, 0);
      mbar_arrive_expect_tx(&k_full[0], TILE16K);
      tma_load_2d(smem + DqSmem::kK, &tm_qkv, &k_full[0], H + h * D, row_base);
      mbar_arrive_expect_tx(&v_full[0], TILE16K);
      tma_load_2d(smem + DqSmem::kV, &tm_qkv, &v_full[0], 2 * H + h * D, row_base);
      if (n_tiles > 1) load_tile(1, 1);
      for (int j = 1; j < n_boxes; ++j) {
        mbar_arrive_expect_tx(&k_full[j], TILE16K);
        tma_load_2d(smem + DqSmem::kK + j * TILE16K, &tm_qkv, &k_full[j], H + h * D, row_base + j * 128);
        mbar_arrive_expect_tx(&v_full[j], TILE16K);
        tma_load_2d(smem + DqSmem::kV + j * TILE16K, &tm_qkv, &v_full[j], 2 * H + h * D, row_base + j * 128);
      }
      for (int t = 2; t < n_tiles; ++t) {
        const int g = t & 1, k = t >> 1;
        mbar_wait(&gbar[g * 7 + QDO_EMPTY], (k - 1) & 1);
        load_tile(g, t);
      }
    } else if ((warp == 1 || warp == 2) && elect_one_sync()) {
      // ====================================== MMA issuers: one thread per group ======================================
      // A group's MMAs alternate deterministically — S/dP of chunk c+1 becomes issuable when the group has pulled chunk c
      // out of TMEM (early in its softmax pass), dQ of chunk c when its dS is in shared memory (end of the pass) — so
      // the issuer blocks on exactly the barrier that is due next; there is no polling loop.
      const int g = warp - 1;
      const int my_tiles = g == 0 ? (n_tiles + 1) >> 1 : n_tiles >> 1;
      constexpr uint32_t idesc_sd = umma_idesc_bf16(128, BC, 0, 0);  // S = Q K^T, dP = dO V^T : all K-major
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, D, 0, 1);   // dQ = dS K : A K-major, B MN-major
      uint64_t* gb = gbar + g * 7;
      const uint32_t tg = tmem_base + g * 224;
      // descriptor bases; the start-address field counts 16-byte units, so advancing is a plain add
      const uint64_t dQt = umma_smem_desc(smem_u32(smem + DqSmem::kQ + g * TILE16K), 16, 1024);
      const uint64_t dDO = umma_smem_desc(smem_u32(smem + DqSmem::kDO + g * TILE16K), 16, 1024);
      const uint64_t dKk = umma_smem_desc(smem_u32(smem + DqSmem::kK), 16, 1024);        // K-major view (S)
      const uint64_t dVk = umma_smem_desc(smem_u32(smem + DqSmem::kV), 16, 1024);        // K-major view (dP)
      const uint64_t dKm = umma_smem_desc(smem_u32(smem + DqSmem::kK), TILE16K, 1024);   // MN-major view (dQ)
      const uint32_t total = my_tiles * n_chunks;
      auto issue_sd = [&](int k, int j) {  // S_j = Q K_j^T and dP_j = dO V_j^T of the group's k-th tile
        if (j == 0) mbar_wait(&gb[QDO_FULL], k & 1);
        if (k == 0 && (j & 1) == 0) {
          mbar_wait(&k_full[j >> 1], 0);
          mbar_wait(&v_full[j >> 1], 0);
        }
        tc_fence_after();
        const uint64_t kb = dKk + static_cast<uint64_t>(j * (CHUNK8K >> 4)), vb = dVk + static_cast<uint64_t>(j * (CHUNK8K >> 4));
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) umma_bf16(tg, dQt + 2 * kk, kb + 2 * kk, idesc_sd, kk > 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) umma_bf16(tg + 64, dDO + 2 * kk, vb + 2 * kk, idesc_sd, kk > 0);
        umma_commit(&gb[SD_FULL]);
        if (j == n_chunks - 1) umma_commit(&gb[QDO_EMPTY]);
        PLB_TRACE(1, j);
      };
      if (total > 0) issue_sd(0, 0);
      uint32_t c = 0;
      for (int k = 0; k < my_tiles; ++k) {
        for (int j = 0; j < n_chunks; ++j, ++c) {
          const bool has_next = c + 1 < total;
          const bool next_in_tile = j + 1 < n_chunks;
          if (has_next && next_in_tile) {
            mbar_wait(&gb[SD_FREE], c & 1);
            issue_sd(k, j + 1);
          }
          // ---- dQ += dS_j K_j ----
          mbar_wait(&gb[DS_READY], c & 1);
          if (j == 0) mbar_wait(&gb[DQ_FREE], (k & 1) ^ 1);
          tc_fence_after();
          const uint64_t kb = dKm + static_cast<uint64_t>(j * (CHUNK8K >> 4));
#pragma unroll
          for (int kk = 0; kk < BC / 16; ++kk) umma_bf16_tmem_a(tg + 128, tg + 192 + 8 * kk, kb + 128 * kk, idesc_dq, (j > 0 || kk > 0));
          umma_commit(&gb[DQ_DONE]);
          PLB_TRACE(2, c);
          if (has_next && !next_in_tile) {  // next tile: its Q / dO arrive only after this tile's last S / dP has run
            mbar_wait(&gb[SD_FREE], c & 1);
            issue_sd(k + 1, 0);
          }
        }
      }
    }
    __syncwarp();
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BWD_REGS_SOFTMAX));
    // ====================================== dS producers + dQ epilogue (two groups x eight warps) ======================================
    // Two threads per query row (same TMEM lane quadrant, different warps), each owning 32 of the chunk's 64 key columns:
    // there is no cross-column reduction in the backward pass, so the split is free and halves every per-chunk latency
    // (TMEM load, exponentials, stores) on the group's critical path; four softmax warps per scheduler hide the rest.
    const int grp = (warp - 4) >> 3;
    const int half = ((warp - 4) >> 2) & 1;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int gtid = threadIdx.x - 128 - grp * 256;  // 0..255 within the group
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tg = tmem_base + lane_addr + grp * 224;
    uint64_t* gb = gbar + grp * 7;
    const uint32_t sDS = smem_u32(smem + DqSmem::kDS + grp * TILE16K);
    const uint32_t sDS_row = sDS + row * 128;
    uint32_t c = 0;
    const uint32_t turn_addr = smem_u32(smem + DqSmem::kTurn);
    const uint32_t steps0 = ((n_tiles + 1) >> 1) * n_chunks, steps1 = (n_tiles >> 1) * n_chunks;  // chunks per group
    for (int t = grp; t < n_tiles; t += 2) {
      const int qrow = t * QT + row;
      const size_t stat_idx = static_cast<size_t>(h) * p.total_rows + row_base + qrow;
      // rows beyond the sequence hold another sequence's data: lse = +inf makes their P (and dS) exactly 0
      const float lse = qrow < nq ? lse2[stat_idx] : INFINITY;
      const float dlt = qrow < nq ? delta[stat_idx] : 0.f;
      for (int j = 0; j < n_chunks; ++j, ++c) {
        mbar_wait(&gb[SD_FULL], c & 1);
        tc_fence_after();
        PLB_TRACE(10, c);
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tg + half * 32, sv);
        tmem_ld_32x32(tg + 64 + half * 32, dv);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gb[SD_FREE]);
        PLB_TRACE(11, c);
        const int valid = nk - j * BC - half * 32;  // keys >= kv_len are masked: P = exp2(-inf) = 0 (last chunk only)
        if (valid < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= valid) sv[i] = 0xff800000u;
        }
        uint32_t pk[16];
        const float nlse = turn_acquire<128>(grp, quad, c, steps1, turn_addr) - lse;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float pr = ex2_fast(fmaf(__uint_as_float(sv[cc * 8 + i]), p.scale_log2, nlse));
            e[i] = pr * (__uint_as_float(dv[cc * 8 + i]) - dlt);
          }
          pk[cc * 4 + 0] = pack_bf16x2(e[0], e[1]); pk[cc * 4 + 1] = pack_bf16x2(e[2], e[3]);
          pk[cc * 4 + 2] = pack_bf16x2(e[4], e[5]); pk[cc * 4 + 3] = pack_bf16x2(e[6], e[7]);
        }
        turn_release<128>(grp, quad, c, steps0, steps1, turn_addr, pk[1] ^ pk[7] ^ pk[8] ^ pk[15]);
        PLB_TRACE(14, c);
        if (c >= 1) mbar_wait(&gb[DQ_DONE], (c - 1) & 1);  // the previous dQ MMA has consumed the dS columns
        PLB_TRACE(15, c);
        // dS goes to tensor memory (the A operand of dQ += dS K is read from there): no shared-memory round trip
        tc_fence_after();
        tmem_st_32x16(tg + 192 + half * 16, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gb[DS_READY]);
        PLB_TRACE(16, c);
      }
      // ---- tile epilogue: dQ * scale -> bf16 -> staging (dS buffer) -> global ----
      mbar_wait(&gb[DQ_DONE], (c - 1) & 1);
      tc_fence_after();
      PLB_TRACE(20, t);
      {
        uint32_t o[32];
        tmem_ld_32x32(tg + 128 + half * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[cc * 8 + 0]) * p.scale, __uint_as_float(o[cc * 8 + 1]) * p.scale);
          v.y = pack_bf16x2(__uint_as_float(o[cc * 8 + 2]) * p.scale, __uint_as_float(o[cc * 8 + 3]) * p.scale);
          v.z = pack_bf16x2(__uint_as_float(o[cc * 8 + 4]) * p.scale, __uint_as_float(o[cc * 8 + 5]) * p.scale);
          v.w = pack_bf16x2(__uint_as_float(o[cc * 8 + 6]) * p.scale, __uint_as_float(o[cc * 8 + 7]) * p.scale);
          sts128(sDS_row + (((half * 4 + cc) ^ (row & 7)) << 4), v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&gb[DQ_FREE]);
      group_bar_sync256(grp);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = gtid + i * 256;
        const int r = idx >> 3, cc = idx & 7;
        if (t * QT + r < nq) {
          const uint4 v = lds128(sDS + r * 128 + ((cc ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(dqkv + static_cast<size_t>(row_base + t * QT + r) * ld_dqkv + h * D + cc * 8) = v;
        }
      }
      if (dbias != nullptr) tile_colsum_atomic<16>(sDS, min(QT, nq - t * QT), gtid, dbias + h * D);
      group_bar_sync256(grp);
      PLB_TRACE(21, t);
    }
  }
  tc_fence_before();
  __syncthreads();
  PLB_TRACE(99, 0);
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ dK/dV kernel
// One key tile at a time per CTA, shared by BOTH softmax groups: the 64-query chunks of the tile alternate between the
// groups (chunk C -> group C & 1, TMEM slot C & 1), and every chunk's dV += P^T dO / dK += dS^T Q accumulates into the
// single dK / dV accumulator pair.  That frees enough tensor memory to keep P^T and dS^T there as packed bf16 — the A
// operand of the accumulating MMAs is read from TMEM, so the softmax threads never store to shared memory in the loop
// (no fence.proxy.async, no swizzled stores) and those MMAs only fetch their 2 KB B operand.  Q and dO of the whole
// sequence stay resident (loaded once instead of once per key tile), K_t / V_t are double-buffered and the buffer of a
// finished tile doubles as the staging area of its epilogue.
//   TMEM columns: slot s: S^T [s*192, +64), dP^T [+64, +64), P^T bf16 pairs [+128, +32), dS^T bf16 pairs [+160, +32);
//                 dK [384, +64), dV [448, +64).
//   Warps: 0 TMA, 1 MMA issuer of the S^T / dP^T stream, 2 MMA issuer of the accumulating stream, 3 idle,
//          4..11 softmax group 0, 12..19 softmax group 1 (two threads per key row, 32 query columns each).
struct DkvSmem {
  static constexpr int kQ = 0;                                // Q of the sequence: 8 chunks of [64 x 64] (8 KB each)
  static constexpr int kDO = kQ + MAX_BOXES * TILE16K;        // dO of the sequence, same layout
  static constexpr int kKV = kDO + MAX_BOXES * TILE16K;       // 2 buffers x (K_t 16 KB | V_t 16 KB); epilogue staging of tile t
  static constexpr int kStat = kKV + 2 * 2 * TILE16K;         // float [2 (lse, delta)][512]
  static constexpr int kBar = kStat + 2 * MAX_BOXES * 128 * 4;
  // qdo_full[8], kv_full[2], kv_free[2], st_full[2], st_free[2], pds_ready[2], pds_free[2], acc_full, acc_free
  static constexpr int kNumBars = 2 * MAX_BOXES + 2 + 2 + 2 + 2 + 2 + 2 + 2;
  static constexpr int kTurn = kBar + kNumBars * 8 + 16;      // 17 words for the exponential-phase turn-taking
  static constexpr int kAlloc = kTurn + 80;
};
static_assert(DkvSmem::kAlloc <= 232448, "dK/dV kernel exceeds the 227 KB shared-memory limit");

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                        const __grid_constant__ CUtensorMap tm_dctx64, const AttnTcParams p,
                        const float* __restrict__ lse2, const float* __restrict__ delta,
                        __nv_bfloat16* __restrict__ dqkv, int ld_dqkv, float* __restrict__ dbias) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DkvSmem::kBar);
  constexpr int NQC = 2 * MAX_BOXES;  // at most 8 query chunks
  uint64_t* qdo_full = bars;                 // [8]
  uint64_t* kv_full = bars + NQC;            // [2]
  uint64_t* kv_free = kv_full + 2;           // [2]
  uint64_t* st_full = kv_free + 2;           // [2]
  uint64_t* st_free = st_full + 2;           // [2]
  uint64_t* pds_ready = st_free + 2;         // [2]
  uint64_t* pds_free = pds_ready + 2;        // [2]
  uint64_t* acc_full = pds_free + 2;
  uint64_t* acc_free = acc_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + DkvSmem::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PLB_TRACE_INIT(2);
  PLB_TRACE(0, 0);
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int row_base = p.seq_start[b], nq = p.q_len[b], nk = p.kv_len[b];
  const int H = p.heads * D;
  const int n_qc = (nq + BC - 1) / BC;            // query chunks per key tile
  const int n_kt = (nk + 127) / 128;              // key tiles that hold at least one valid key
  const uint32_t total = static_cast<uint32_t>(n_kt) * n_qc;  // chunks of the CTA, numbered C = t * n_qc + j

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv128);
    tma_prefetch_desc(&tm_qkv64);
    tma_prefetch_desc(&tm_dctx64);
    for (int i = 0; i < NQC + 4; ++i) mbar_init(&bars[i], 1);   // qdo_full, kv_full, kv_free
    for (int s = 0; s < 2; ++s) {
      mbar_init(&st_full[s], 1);
      mbar_init(&st_free[s], 8);      // one arrive per softmax warp of the group
      mbar_init(&pds_ready[s], 8);
      mbar_init(&pds_free[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 16);          // every softmax warp reads a part of the accumulators
    *reinterpret_cast<volatile float*>(smem + DkvSmem::kTurn) = 0.f;
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  {
    // log-sum-exp and delta of every query row of the sequence (rows beyond it: lse = +inf -> P = 0)
    float* stat_all = reinterpret_cast<float*>(smem + DkvSmem::kStat);
    const size_t base = static_cast<size_t>(h) * p.total_rows + row_base;
    for (int q = threadIdx.x; q < MAX_BOXES * 128; q += BWD_THREADS) {
      stat_all[q] = q < nq ? lse2[base + q] : INFINITY;
      stat_all[MAX_BOXES * 128 + q] = q < nq ? delta[base + q] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BWD_REGS_CONTROL));
    if (warp == 0 && elect_one_sync()) {
      // ====================================== TMA producer ======================================
      auto load_kv = [&](int t) {
        uint8_t* buf = smem + DkvSmem::kKV + (t & 1) * 2 * TILE16K;
        mbar_arrive_expect_tx(&kv_full[t & 1], 2 * TILE16K);
        tma_load_2d(buf, &tm_qkv128, &kv_full[t & 1], H + h * D, row_base + t * 128);
        tma_load_2d(buf + TILE16K, &tm_qkv128, &kv_full[t & 1], 2 * H + h * D, row_base + t * 128);
      };
      auto load_qdo = [&](int j) {
        mbar_arrive_expect_tx(&qdo_full[j], TILE16K);
        tma_load_2d(smem + DkvSmem::kQ + j * CHUNK8K, &tm_qkv64, &qdo_full[j], h * D, row_base + j * BC);
        tma_load_2d(smem + DkvSmem::kDO + j * CHUNK8K, &tm_dctx64, &qdo_full[j], h * D, row_base + j * BC);
      };
      if (total > 0) {
        load_qdo(0);
        load_kv(0);
        for (int j = 1; j < n_qc; ++j) load_qdo(j);
        if (n_kt > 1) load_kv(1);
        for (int t = 2; t < n_kt; ++t) {
          mbar_wait(&kv_free[t & 1], ((t >> 1) - 1) & 1);  // the epilogue of tile t-2 has left the buffer
          load_kv(t);
        }
      }
    } else if (warp == 1 && elect_one_sync()) {
      // ====================================== MMA issuer: S^T = K_t Q_j^T, dP^T = V_t dO_j^T ======================================
      constexpr uint32_t idesc_st = umma_idesc_bf16(128, BC, 0, 0);
      const uint64_t dQk = umma_smem_desc(smem_u32(smem + DkvSmem::kQ), 16, 1024);
      const uint64_t dDOk = umma_smem_desc(smem_u32(smem + DkvSmem::kDO), 16, 1024);
      const uint64_t dKV = umma_smem_desc(smem_u32(smem + DkvSmem::kKV), 16, 1024);
      uint32_t C = 0;
      for (int t = 0; t < n_kt; ++t) {
        for (int j = 0; j < n_qc; ++j, ++C) {
          const uint32_t s = C & 1, u = C >> 1;
          if (t == 0) mbar_wait(&qdo_full[j], 0);
          if (j == 0) mbar_wait(&kv_full[t & 1], (t >> 1) & 1);
          if (u >= 1) mbar_wait(&st_free[s], (u - 1) & 1);
          tc_fence_after();
          const uint64_t kb = dKV + static_cast<uint64_t>((t & 1) * ((2 * TILE16K) >> 4)), vb = kb + (TILE16K >> 4);
          const uint64_t qb = dQk + static_cast<uint64_t>(j * (CHUNK8K >> 4)), dob = dDOk + static_cast<uint64_t>(j * (CHUNK8K >> 4));
          const uint32_t ts = tmem_base + s * 192;
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) umma_bf16(ts, kb + 2 * kk, qb + 2 * kk, idesc_st, kk > 0);
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) umma_bf16(ts + 64, vb + 2 * kk, dob + 2 * kk, idesc_st, kk > 0);
          umma_commit(&st_full[s]);
          PLB_TRACE(1, C);
        }
      }
    } else if (warp == 2 && elect_one_sync()) {
      // ====================================== MMA issuer: dV += P^T dO_j, dK += dS^T Q_j (A operands in TMEM) ======================================
      constexpr uint32_t idesc_acc = umma_idesc_bf16(128, D, 0, 1);  // B MN-major
      const uint64_t dQm = umma_smem_desc(smem_u32(smem + DkvSmem::kQ), TILE16K, 1024);
      const uint64_t dDOm = umma_smem_desc(smem_u32(smem + DkvSmem::kDO), TILE16K, 1024);
      uint32_t C = 0;
      for (int t = 0; t < n_kt; ++t) {
        for (int j = 0; j < n_qc; ++j, ++C) {
          const uint32_t s = C & 1, u = C >> 1;
          mbar_wait(&pds_ready[s], u & 1);
          if (j == 0 && t >= 1) mbar_wait(acc_free, (t - 1) & 1);  // the epilogue of the previous tile has read dK / dV
          tc_fence_after();
          const uint64_t qb = dQm + static_cast<uint64_t>(j * (CHUNK8K >> 4)), dob = dDOm + static_cast<uint64_t>(j * (CHUNK8K >> 4));
          const uint32_t ts = tmem_base + s * 192;
#pragma unroll
          for (int kk = 0; kk < BC / 16; ++kk) umma_bf16_tmem_a(tmem_base + 448, ts + 128 + 8 * kk, dob + 128 * kk, idesc_acc, (j > 0 || kk > 0));
#pragma unroll
          for (int kk = 0; kk < BC / 16; ++kk) umma_bf16_tmem_a(tmem_base + 384, ts + 160 + 8 * kk, qb + 128 * kk, idesc_acc, (j > 0 || kk > 0));
          umma_commit(&pds_free[s]);
          if (j == n_qc - 1) umma_commit(acc_full);
          PLB_TRACE(2, C);
        }
      }
    }
    __syncwarp();
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BWD_REGS_SOFTMAX));
    // ====================================== P^T / dS^T producers + dK/dV epilogue ======================================
    const int grp = (warp - 4) >> 3;
    const int half = ((warp - 4) >> 2) & 1;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int stid = threadIdx.x - 128;  // 0..511 over both groups
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t ts = tmem_base + lane_addr + grp * 192;
    const uint32_t stat_addr = smem_u32(smem + DkvSmem::kStat);
    const uint32_t turn_addr = smem_u32(smem + DkvSmem::kTurn);
    const uint32_t steps0 = (total + 1) >> 1, steps1 = total >> 1;  // chunks per group

    // key rows in [128 * n_kt, nq): padded keys inside the allocated sequence -> their gradients are exactly zero
    for (int t = n_kt; t * 128 < nq; ++t) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = stid + i * 512;
        const int r = t * 128 + (idx >> 3), cc = idx & 7;
        if (r < nq) {
          __nv_bfloat16* dst = dqkv + static_cast<size_t>(row_base + r) * ld_dqkv + h * D + cc * 8;
          *reinterpret_cast<uint4*>(dst + H) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(dst + 2 * H) = make_uint4(0, 0, 0, 0);
        }
      }
    }

    uint32_t c = 0;  // chunks this group has processed (its step counter; global chunk C = 2 c + grp)
    for (int t = 0; t < n_kt; ++t) {
      const bool key_valid = (t * 128 + row) < nk;
      // this group's chunks of tile t: global C in [t * n_qc, (t + 1) * n_qc) with C & 1 == grp
      const uint32_t C_end = static_cast<uint32_t>(t + 1) * n_qc;
      for (uint32_t C = 2 * c + grp; C < C_end; C += 2, ++c) {
        const int j = static_cast<int>(C - static_cast<uint32_t>(t) * n_qc);
        mbar_wait(&st_full[grp], c & 1);
        tc_fence_after();
        PLB_TRACE(10, c);
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(ts + half * 32, sv);
        tmem_ld_32x32(ts + 64 + half * 32, dv);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_free[grp]);
        if (!key_valid) {  // key rows >= kv_len (last tile only): P = exp2(-inf) = 0, hence dS = 0
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[i] = 0xff800000u;
        }
        PLB_TRACE(11, c);
        const uint32_t st = stat_addr + (j * BC + half * 32) * 4;
        uint32_t pkp[16], pkd[16];
        const float tz = turn_acquire<128>(grp, quad, c, steps1, turn_addr);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4 l0 = lds128_ro(st + cc * 32), l1 = lds128_ro(st + cc * 32 + 16);
          const float4 d0 = lds128_ro(st + MAX_BOXES * 512 + cc * 32), d1 = lds128_ro(st + MAX_BOXES * 512 + cc * 32 + 16);
          const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
          const float dl[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
          float pr[8], ds[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pr[i] = ex2_fast(fmaf(__uint_as_float(sv[cc * 8 + i]), p.scale_log2, tz - ls[i]));
            ds[i] = pr[i] * (__uint_as_float(dv[cc * 8 + i]) - dl[i]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            pkp[cc * 4 + i] = pack_bf16x2(pr[2 * i], pr[2 * i + 1]);
            pkd[cc * 4 + i] = pack_bf16x2(ds[2 * i], ds[2 * i + 1]);
          }
        }
        turn_release<128>(grp, quad, c, steps0, steps1, turn_addr, pkd[1] ^ pkd[7] ^ pkd[8] ^ pkd[15]);
        PLB_TRACE(14, c);
        if (c >= 1) mbar_wait(&pds_free[grp], (c - 1) & 1);  // the previous accumulating MMAs of this slot have read P^T / dS^T
        PLB_TRACE(15, c);
        tc_fence_after();
        tmem_st_32x16(ts + 128 + half * 16, pkp);
        tmem_st_32x16(ts + 160 + half * 16, pkd);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_ready[grp]);
        PLB_TRACE(16, c);
      }
      // ---- tile epilogue (all 16 warps): group 0 drains dK * scale, group 1 drains dV -> bf16 -> staging -> global ----
      uint8_t* stage = smem + DkvSmem::kKV + (t & 1) * 2 * TILE16K;  // K_t / V_t of this tile are dead by now
      const uint32_t stage_mine = smem_u32(stage) + grp * TILE16K;    // dK staging | dV staging
      mbar_wait(acc_full, t & 1);
      tc_fence_after();
      PLB_TRACE(20, t);
      {
        const float mul = grp == 0 ? p.scale : 1.f;
        uint32_t o[32];
        tmem_ld_32x32(tmem_base + lane_addr + 384 + grp * 64 + half * 32, o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free);  // accumulators are in registers: the next tile may start accumulating
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[cc * 8 + 0]) * mul, __uint_as_float(o[cc * 8 + 1]) * mul);
          v.y = pack_bf16x2(__uint_as_float(o[cc * 8 + 2]) * mul, __uint_as_float(o[cc * 8 + 3]) * mul);
          v.z = pack_bf16x2(__uint_as_float(o[cc * 8 + 4]) * mul, __uint_as_float(o[cc * 8 + 5]) * mul);
          v.w = pack_bf16x2(__uint_as_float(o[cc * 8 + 6]) * mul, __uint_as_float(o[cc * 8 + 7]) * mul);
          sts128(stage_mine + row * 128 + (((half * 4 + cc) ^ (row & 7)) << 4), v);
        }
      }
      asm volatile("bar.sync 11, 512;" ::: "memory");  // both staging tiles complete
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = stid + i * 512;
        const int r = idx >> 3, cc = idx & 7;
        const int key = t * 128 + r;
        if (key < nq) {  // rows in [nk, nq) were masked above: their accumulators are exact zeros
          __nv_bfloat16* dst = dqkv + static_cast<size_t>(row_base + key) * ld_dqkv + h * D + cc * 8;
          *reinterpret_cast<uint4*>(dst + H) = lds128(smem_u32(stage) + r * 128 + ((cc ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(dst + 2 * H) = lds128(smem_u32(stage) + TILE16K + r * 128 + ((cc ^ (r & 7)) << 4));
        }
      }
      if (dbias != nullptr) {
        // (the key-bias gradient is mathematically zero; it is accumulated anyway so the tensor holds what autograd would)
        tile_colsum_atomic<8>(smem_u32(stage), min(128, nk - t * 128), stid, dbias + H + h * D);
        tile_colsum_atomic<8>(smem_u32(stage) + TILE16K, min(128, nk - t * 128), stid, dbias + 2 * H + h * D);
      }
      fence_proxy_async_smem();  // generic reads / writes of the staging area are ordered before the TMA load that reuses it
      asm volatile("bar.sync 11, 512;" ::: "memory");
      if (stid == 0) mbar_arrive(&kv_free[t & 1]);
      PLB_TRACE(21, t);
    }
  }
  tc_fence_before();
  __syncthreads();
  PLB_TRACE(99, 0);
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// delta = rowsum(dO o O) per (head, row) for callers that did not fuse it into the O-projection DGRAD (shared with
// attn_bwd_fused.cu)
int launch_attn_delta(const void* dctx, const void* ctx, int ld_ctx, int total_rows, int heads, float* delta,
                      cudaStream_t stream) {
  const int dgrid = (total_rows + 7) / 8 < sm_count() * 8 ? (total_rows + 7) / 8 : sm_count() * 8;
  attn_delta_kernel<<<dgrid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dctx), static_cast<const __nv_bfloat16*>(ctx),
                                               ld_ctx, total_rows, heads, total_rows, delta);
  return check_launch("plb_attn_bwd(delta)");
}
}  // namespace plb

using namespace plb;

extern "C" int plb_attn_bwd_tc(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                               const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                               const void* ctx, const void* dctx, int ld_ctx, const float* lse2, float* delta,
                               void* dqkv, int ld_dqkv, float* dbias_qkv_accum, plb_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLB_REQUIRE(qkv && seq_start && q_len && kv_len && dctx && lse2 && delta && dqkv, PLB_ERR_ARG, "attn_bwd_tc: null pointer");
  PLB_REQUIRE(head_dim == D, PLB_ERR_UNSUPPORTED, "attn_bwd_tc: head_dim %d not supported (only 64)", head_dim);
  PLB_REQUIRE(B > 0 && heads > 0 && total_rows > 0, PLB_ERR_ARG, "attn_bwd_tc: bad sizes");
  PLB_REQUIRE(max_q_len > 0 && max_q_len <= MAX_BOXES * 128, PLB_ERR_UNSUPPORTED,
              "attn_bwd_tc: sequences longer than %d rows are not supported (got %d)", MAX_BOXES * 128, max_q_len);
  PLB_REQUIRE(ld_qkv % 8 == 0 && ld_ctx % 8 == 0 && ld_dqkv % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(ctx) % 16) == 0 && (reinterpret_cast<uintptr_t>(dctx) % 16) == 0 &&  /* NULL ctx passes */
                  (reinterpret_cast<uintptr_t>(dqkv) % 16) == 0,
              PLB_ERR_ARG, "attn_bwd_tc: operands must be 16-byte aligned with ld %% 8 == 0");
  CUtensorMap tm_qkv128, tm_qkv64, tm_dctx128, tm_dctx64;
  if (!tensor_map_bf16(&tm_qkv128, qkv, total_rows, 3 * heads * D, ld_qkv, 128, 64) ||
      !tensor_map_bf16(&tm_qkv64, qkv, total_rows, 3 * heads * D, ld_qkv, 64, 64) ||
      !tensor_map_bf16(&tm_dctx128, dctx, total_rows, heads * D, ld_ctx, 128, 64) ||
      !tensor_map_bf16(&tm_dctx64, dctx, total_rows, heads * D, ld_ctx, 64, 64))
    return PLB_ERR_CUDA;
  static bool configured = false;
  if (!configured) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::kAlloc);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::kAlloc);
    PLB_REQUIRE(e1 == cudaSuccess && e2 == cudaSuccess, PLB_ERR_CUDA, "attn_bwd_tc: cudaFuncSetAttribute failed");
    configured = true;
  }
  AttnTcParams p{seq_start, q_len, kv_len, heads, total_rows, 1.4426950408889634f / sqrtf((float)head_dim),
                 1.f / sqrtf((float)head_dim)};
  int rc = PLB_OK;
  if (ctx != nullptr) {  // otherwise delta was produced by the O-projection DGRAD (PLB_EPI_ROWDOT)
    rc = launch_attn_delta(dctx, ctx, ld_ctx, total_rows, heads, delta, stream);
    if (rc) return rc;
  }
  attn_bwd_dq_tc_kernel<<<B * heads, BWD_THREADS, DqSmem::kAlloc, stream>>>(tm_qkv128, tm_dctx128, p, lse2, delta,
                                                                            static_cast<__nv_bfloat16*>(dqkv), ld_dqkv,
                                                                            dbias_qkv_accum);
  rc = check_launch("plb_attn_bwd_tc(dq)");
  if (rc) return rc;
  attn_bwd_dkv_tc_kernel<<<B * heads, BWD_THREADS, DkvSmem::kAlloc, stream>>>(tm_qkv128, tm_qkv64, tm_dctx64, p, lse2, delta,
                                                                              static_cast<__nv_bfloat16*>(dqkv), ld_dqkv,
                                                                              dbias_qkv_accum);
  return check_launch("plb_attn_bwd_tc(dkv)");
}

#ifdef PLB_ATTN_TRACE
// Trace build only: copies the record table (zero = unused slot) to the host and clears it.
extern "C" int plb_attn_trace_dump(unsigned long long* host, int cap) {
  if (cap < static_cast<int>(plb::TRACE_CAP)) return -1;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host, plb::g_trace, sizeof(plb::g_trace));
  void* dev = nullptr;
  cudaGetSymbolAddress(&dev, plb::g_trace);
  cudaMemset(dev, 0, sizeof(plb::g_trace));
  cudaDeviceSynchronize();
  return static_cast<int>(plb::TRACE_CAP);
}
#endif
