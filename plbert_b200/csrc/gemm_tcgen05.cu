// plbert_b200 — bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into a 128-byte-swizzled shared-memory ring.
//
// One persistent, warp-specialised kernel template serves the three GEMM shapes of an nn.Linear
// (HF modeling_albert.py:180-182,198-200,239-241,276; R/model.py:15-16):
//
//   FWD   (A K-major,  B K-major ):  C[M,N]  = A[M,K] W[N,K]^T
//   DGRAD (A K-major,  B MN-major):  C[M,N]  = A[M,K] W[K,N]
//   WGRAD (A MN-major, B MN-major):  C[M,N] += A[K,M]^T B[K,N]   (fp32 atomics, split-K)
//
// CTA = 11 warps: warp 0 TMA producer, warp 1 TMEM owner + single-thread MMA issuer, warps 2..9
// epilogue (one accumulator row per thread, tcgen05.ld 32x32b; the two warps that share a TMEM lane
// quadrant take 32 columns each of every 64-column chunk), warp 10 store warp (TMA stores of staged
// chunks + aux-operand prefetches).  Tile 128 x 256 x 64, fp32 accumulators double-buffered in TMEM
// (2 x 256 columns) so the epilogue of unit i overlaps the main loop of unit i+1.  Epilogue output goes
// registers -> swizzled smem -> TMA store; a residual / pre-activation operand comes in by TMA into the
// same staging buffer and is updated in place.
#include "common.cuh"

#include <cstdlib>

namespace plb {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * BK * 2;       // 16 KB
constexpr int B_TILE_BYTES = BN * BK * 2;       // 32 KB (one CTA) or 16 KB per CTA of a pair
constexpr int EPI_COLS = 64;                    // columns per epilogue chunk (128 B of bf16)
constexpr int EPI_BUF_BYTES = BM * EPI_COLS * 2;  // 16 KB
constexpr int MN_BOX_BYTES = 64 * 64 * 2;       // one MN-major TMA box: 64 K-rows x 64 elements
// Epilogue warps: EPI_GROUPS groups of four warps (one per TMEM lane quadrant).  With two groups the warps that share a
// quadrant (and an SM sub-partition scheduler) split every 64-column chunk into 32-column halves, so tcgen05.ld / LDS /
// MUFU latencies of one warp hide behind the other (the single-group epilogue left the scheduler idle 61 % of the cycles,
// profiles/r1e_gemm_ncu_full_details.txt).  -DPLB_EPI_GROUPS=1 rebuilds the single-group kernel for A/B runs.
#ifndef PLB_EPI_GROUPS
#define PLB_EPI_GROUPS 2
#endif
constexpr int EPI_GROUPS = PLB_EPI_GROUPS;
static_assert(EPI_GROUPS == 1 || EPI_GROUPS == 2, "one or two epilogue warp groups");
constexpr int EPI_THREADS = 128 * EPI_GROUPS;
// A dedicated store warp (one elected thread) issues the TMA stores of finished chunks and the aux-in prefetches and hands
// the staging slots back through mbarriers, so the epilogue warps never meet at a CTA barrier per chunk and never wait
// behind the leader's bulk-group waits (ncu: 13-22 % of the fused-epilogue kernels' samples sat on that barrier; the
// GELU' DGRAD gained 7 %, the O projection 4 % — profiles/r1f_gemm_selftest_perf.txt).
constexpr int STORE_WARP = 2 + EPI_THREADS / 32;  // warp index of the store warp
constexpr int NUM_THREADS = 64 + EPI_THREADS + 32;
constexpr int GRP_COLS = EPI_COLS / EPI_GROUPS;   // columns of a chunk one thread handles
constexpr int GRP_PIECES = GRP_COLS / 8;          // 16-byte pieces of a staged row one thread handles
constexpr int TMEM_COLS = 512;

// Timeline instrumentation of ONE epilogue warp (CTA 0, warp 2): compiled only into the trace library that
// profiles/gemm_trace.py builds with -DPLB_GEMM_TRACE; the product library contains none of it.
#ifdef PLB_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[8192];
#define PLB_GT(tag)                                                                                   \
  do {                                                                                                \
    if (blockIdx.x == 0 && warp == 2 && lane == 0 && gt_n < 8192)                                      \
      g_gemm_trace[gt_n++] = (static_cast<unsigned long long>(tag) << 56) | (clock64() & 0xffffffffffffffULL); \
  } while (0)
#else
#define PLB_GT(tag) do {} while (0)
#endif

enum Epi : int { kNone = 0, kResidual = 1, kGelu = 2, kGeluBwd = 3, kAtomicF32 = 4, kRowDot = 5 };

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits;
  int m_fast;         // 1: consecutive units walk down M (the B tile stays hot in L2), 0: along N (the A tile stays hot)
  int k_iters;        // ceil(K / BK)
  int k_per_split;    // k iterations per split
  int num_units;      // m_tiles * n_tiles * splits (m_tiles counts 128*kCta-row tiles)
  const float* bias;  // [N] or null
  float* c_f32;       // WGRAD output
  int c_multimem;     // WGRAD: c_f32 is a MULTICAST address (NVSwitch): every add lands in all peers' copies (multimem.red)
  int ldc;
  float* rowdot;      // kRowDot: fp32 [N / 64][rowdot_ld] accumulators of sum_n C[m,n] * aux_in[m,n] per 64-column block
  int rowdot_ld;
  float* colsum;      // kGeluBwd (optional): fp32 [N], column sums of the stored C are ADDED (the producing Linear's bias gradient)
};

template <int kStages, int kBufs, int kCta>
struct SmemLayout {
  static constexpr int kStageBytes = A_TILE_BYTES + B_TILE_BYTES / kCta;  // bytes THIS CTA stages per k-iteration
  static constexpr int kStageOff = 0;
  static constexpr int kEpiOff = kStages * kStageBytes;
  static constexpr int kBarOff = kEpiOff + kBufs * EPI_BUF_BYTES;
  // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], epi_in[kBufs], tmem ptr
  // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], epi_in[kBufs], chunk_ready[kBufs], slot_free[kBufs]
  static constexpr int kNumBars = 2 * kStages + 4 + 3 * kBufs;
  static constexpr int kBiasOff = kBarOff + kNumBars * 8 + 16;   // fp32 bias of the current tile: BN floats
  static constexpr int kBytes = kBiasOff + BN * 4;
  static constexpr int kAlloc = kBytes + 1024;  // slack for manual 1024 B alignment
};

// Walks the units unit0, unit0 + stride, ... of a persistent CTA without per-unit integer divisions: the div/mod chain
// of the straightforward index -> (m, n, split) mapping cost ~470 clk per tile on the epilogue's critical path
// (profiles/r1f_gemm_epilogue_timeline.txt).  The position is kept as (fast tile, slow tile, split) digits and advanced
// by the pre-divided stride with carries.  unit = (split * slow_size + slow) * fast_size + fast; the fast digit is the
// tile dimension consecutive units walk along (GemmParams::m_fast).
struct UnitWalker {
  int u, stride;
  int m_blk, n_blk, k_begin, k_end;
  int f, s, sp, F, Sn, df, ds, dsp;
  __device__ __forceinline__ UnitWalker(const GemmParams& p, int unit0, int stride_) : u(unit0), stride(stride_) {
    F = p.m_fast ? p.m_tiles : p.n_tiles;
    Sn = p.m_fast ? p.n_tiles : p.m_tiles;
    int t = unit0 / F;
    f = unit0 - t * F;
    sp = t / Sn;
    s = t - sp * Sn;
    t = stride_ / F;
    df = stride_ - t * F;
    dsp = t / Sn;
    ds = t - dsp * Sn;
    set(p);
  }
  __device__ __forceinline__ void set(const GemmParams& p) {
    m_blk = p.m_fast ? f : s;
    n_blk = p.m_fast ? s : f;
    k_begin = sp * p.k_per_split;
    k_end = min(k_begin + p.k_per_split, p.k_iters);
  }
  __device__ __forceinline__ bool valid(const GemmParams& p) const { return u < p.num_units; }
  __device__ __forceinline__ void next(const GemmParams& p) {
    u += stride;
    f += df;
    int carry = f >= F ? 1 : 0;
    f -= carry ? F : 0;
    s += ds + carry;
    carry = s >= Sn ? 1 : 0;
    s -= carry ? Sn : 0;
    sp += dsp + carry;
    set(p);
  }
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

template <bool kAMn, bool kBMn, int kEpi, int kStages, int kBufs, int kCta>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
            const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_aux_in,
            const __grid_constant__ CUtensorMap tm_aux_out, const GemmParams p) {
  using L = SmemLayout<kStages, kBufs, kCta>;
  constexpr int STAGE_BYTES = L::kStageBytes;
  constexpr int B_ROWS = BN / kCta;  // rows (FWD) / columns (MN-major) of the B tile this CTA loads
  // CTA pair (kCta == 2): the two CTAs of a cluster own rows [0,128) and [128,256) of a 256-row tile and each stages half
  // of the B tile; one tcgen05.mma.cta_group::2 issued by the leader (rank 0) drives both tensor cores.
  const uint32_t rank = (kCta == 2) ? cluster_ctarank() : 0;
  const int unit0 = blockIdx.x / kCta, unit_stride = gridDim.x / kCta;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint64_t* epi_in = bars + 2 * kStages + 4;
  uint64_t* chunk_ready = epi_in + kBufs;   // epilogue warps -> store warp: the chunk is staged (count = epilogue warps)
  uint64_t* slot_free = chunk_ready + kBufs;  // store warp -> epilogue warps: the slot's store has been read out
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (kEpi != kAtomicF32) tma_prefetch_desc(&tm_c);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], kCta);  // leader: own arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4 * EPI_GROUPS * kCta);  // one arrive per epilogue warp of every CTA of the pair
    }
    for (int i = 0; i < kBufs; ++i) {
      mbar_init(&epi_in[i], 1);
      mbar_init(&chunk_ready[i], EPI_THREADS / 32);
      mbar_init(&slot_free[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kCta == 2) tmem_alloc_2cta(tmem_ptr_smem, TMEM_COLS); else tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  }
  tc_fence_before();
  if (kCta == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_grid_sync();   // everything above touched only this CTA's shared / tensor memory; operands and outputs are global

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one_sync()) {  // elect.sync (not lane == 0): ptxas then emits the TMA / MMA instructions without waterfall loops
      int stage = 0;
      uint32_t phase = 0;
      for (UnitWalker w(p, unit0, unit_stride); w.valid(p); w.next(p)) {
        const int k_begin = w.k_begin, k_end = w.k_end;
        const int m0 = (w.m_blk * kCta + rank) * BM, n0 = w.n_blk * BN + rank * B_ROWS;
        for (int kb = k_begin; kb < k_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          if (kCta == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            if (!kAMn) {
              tma_load_2d(sa, &tm_a, &full_bar[stage], kb * BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d(sa + j * MN_BOX_BYTES, &tm_a, &full_bar[stage], m0 + j * 64, kb * BK);
            }
            if (!kBMn) {
              tma_load_2d(sb, &tm_b, &full_bar[stage], kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j)
                tma_load_2d(sb + j * MN_BOX_BYTES, &tm_b, &full_bar[stage], n0 + j * 64, kb * BK);
            }
          } else {
            // both CTAs load their halves into their own shared memory; every byte completes on the LEADER's barrier
            const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], kCta * STAGE_BYTES);
            else mbar_arrive_cluster(lead_bar);
            if (!kAMn) {
              tma_load_2d_2sm(sa, &tm_a, lead_bar, kb * BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(sa + j * MN_BOX_BYTES, &tm_a, lead_bar, m0 + j * 64, kb * BK);
            }
            if (!kBMn) {
              tma_load_2d_2sm(sb, &tm_b, lead_bar, kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j) tma_load_2d_2sm(sb + j * MN_BOX_BYTES, &tm_b, lead_bar, n0 + j * 64, kb * BK);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (rank == 0 && elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * kCta, BN, kAMn ? 1 : 0, kBMn ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (UnitWalker w(p, unit0, unit_stride); w.valid(p); w.next(p)) {
        const int k_begin = w.k_begin, k_end = w.k_end;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = k_begin; kb < k_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_TILE_BYTES;
          // descriptor of the stage once; the start-address field counts 16-byte units, so a K step is a plain add
          const uint64_t da0 = kAMn ? umma_smem_desc(sa, MN_BOX_BYTES, 1024) : umma_smem_desc(sa, 16, 1024);
          const uint64_t db0 = kBMn ? umma_smem_desc(sb, MN_BOX_BYTES, 1024) : umma_smem_desc(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = da0 + static_cast<uint64_t>(k * ((kAMn ? UMMA_K * 128 : UMMA_K * 2) >> 4));
            const uint64_t db = db0 + static_cast<uint64_t>(k * ((kBMn ? UMMA_K * 128 : UMMA_K * 2) >> 4));
            if (kCta == 2) umma_bf16_2cta(tmem_d, da, db, idesc, (kb > k_begin || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, da, db, idesc, (kb > k_begin || k > 0) ? 1u : 0u);
          }
          // smem slot reusable (in BOTH CTAs) once these MMAs have read it
          if (kCta == 2) umma_commit_2cta(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (kCta == 2) umma_commit_2cta(&tmem_full[acc], 3); else umma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == STORE_WARP) {
    // =========================== store warp ===========================
    // Walks the same (unit, chunk) sequence as the epilogue warps: waits until a chunk is staged, stores it, and once
    // the store has been read out of shared memory recycles the slot (aux-in prefetch of chunk seq + kSlots, or a
    // slot_free arrive when there is no aux operand).
    constexpr bool kHasAuxIn = (kEpi == kResidual || kEpi == kGeluBwd || kEpi == kRowDot);
    constexpr bool kDual = (kEpi == kGelu);
    constexpr int kSlots = kDual ? kBufs / 2 : kBufs;
    if (kEpi == kGeluBwd && p.colsum != nullptr) {
      // Same walk with all 32 lanes: besides the elected thread's TMA traffic, the warp adds the column sums of every staged
      // [128 x 64] chunk to p.colsum (d_u's column sums = the FFN1 bias gradient; the separate 62 us pass over [T, 2048] per
      // layer is gone).  Lane l owns the column pair 2l, 2l+1 and reads it straight from the swizzled staging buffer while the
      // TMA store reads the same bytes; the sums run over the ROUNDED values, i.e. what the weight-gradient GEMM reads.
      uint8_t* epi_smem = smem + L::kEpiOff;
      UnitWalker la(p, unit0, unit_stride);
      int la_c = 0;
      uint32_t la_seq = 0;
      auto la_issue = [&]() {  // aux load of the next not-yet-issued chunk, if any (state kept by every lane, issued by one)
        while (la.valid(p)) {
          const int m_blk = la.m_blk;
          const int n0 = la.n_blk * BN + la_c * EPI_COLS;
          if (la_c < BN / EPI_COLS && n0 < p.N) {
            const int slot = la_seq % kSlots;
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&epi_in[slot], EPI_BUF_BYTES);
              tma_load_2d(epi_smem + slot * EPI_BUF_BYTES, &tm_aux_in, &epi_in[slot], n0, (m_blk * kCta + rank) * BM);
            }
            ++la_seq;
            ++la_c;
            return;
          }
          la.next(p);
          la_c = 0;
        }
      };
      for (int i = 0; i < kSlots; ++i) la_issue();
      uint32_t seq = 0;
      for (UnitWalker w(p, unit0, unit_stride); w.valid(p); w.next(p)) {
        const int m0 = (w.m_blk * kCta + rank) * BM, n0 = w.n_blk * BN;
        const int rows = min(BM, p.M - m0);   // rows of the tile inside the matrix (<= 0: the pair's second half is past the end)
        for (int c = 0; c < BN / EPI_COLS; ++c) {
          const int nc = n0 + c * EPI_COLS;
          if (nc >= p.N) break;
          const int slot = seq % kSlots;
          uint8_t* buf = epi_smem + slot * EPI_BUF_BYTES;
          mbar_wait(&chunk_ready[slot], (seq / kSlots) & 1);
          if (elect_one_sync()) {
            tma_store_2d(&tm_c, buf, nc, m0);
            bulk_commit();
          }
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          const uint32_t col_addr = smem_u32(buf) + (lane & 3) * 4;
          int r = 0;
#pragma unroll 4
          for (; r + 1 < rows; r += 2) {
            uint32_t va, vb;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(va) : "r"(col_addr + r * 128 + (((lane >> 2) ^ (r & 7)) << 4)));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(vb) : "r"(col_addr + (r + 1) * 128 + (((lane >> 2) ^ ((r + 1) & 7)) << 4)));
            s0 += bf16_lo(va); s1 += bf16_hi(va);
            s2 += bf16_lo(vb); s3 += bf16_hi(vb);
          }
          if (r < rows) {
            uint32_t va;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(va) : "r"(col_addr + r * 128 + (((lane >> 2) ^ (r & 7)) << 4)));
            s0 += bf16_lo(va); s1 += bf16_hi(va);
          }
          if (rows > 0) {
            const int n = nc + 2 * lane;
            if (n < p.N) atomicAdd(p.colsum + n, s0 + s2);
            if (n + 1 < p.N) atomicAdd(p.colsum + n + 1, s1 + s3);
          }
          if (elect_one_sync()) bulk_wait_read<0>();   // the slot has left shared memory
          __syncwarp();                                // ... and every lane has read its column pair
          la_issue();                                  // chunk seq + kSlots goes into it; epi_in[slot] announces it
          ++seq;
        }
      }
      if (elect_one_sync()) bulk_wait<0>();  // all stores globally visible before exit
    } else if (kEpi != kAtomicF32 && elect_one_sync()) {
      uint8_t* epi_smem = smem + L::kEpiOff;
      UnitWalker la(p, unit0, unit_stride);
      int la_c = 0;
      uint32_t la_seq = 0;
      auto la_issue = [&]() {  // aux load of the next not-yet-issued chunk, if any
        while (la.valid(p)) {
          const int m_blk = la.m_blk;
          const int n0 = la.n_blk * BN + la_c * EPI_COLS;
          if (la_c < BN / EPI_COLS && n0 < p.N) {
            const int slot = la_seq % kSlots;
            mbar_arrive_expect_tx(&epi_in[slot], EPI_BUF_BYTES);
            tma_load_2d(epi_smem + slot * EPI_BUF_BYTES, &tm_aux_in, &epi_in[slot], n0, (m_blk * kCta + rank) * BM);
            ++la_seq;
            ++la_c;
            return;
          }
          la.next(p);
          la_c = 0;
        }
      };
      if (kHasAuxIn)
        for (int i = 0; i < kSlots; ++i) la_issue();
      uint32_t seq = 0;
      for (UnitWalker w(p, unit0, unit_stride); w.valid(p); w.next(p)) {
        const int m0 = (w.m_blk * kCta + rank) * BM, n0 = w.n_blk * BN;
        for (int c = 0; c < BN / EPI_COLS; ++c) {
          const int nc = n0 + c * EPI_COLS;
          if (nc >= p.N) break;
          const int slot = seq % kSlots;
          uint8_t* buf = epi_smem + (kDual ? 2 * slot : slot) * EPI_BUF_BYTES;
          mbar_wait(&chunk_ready[slot], (seq / kSlots) & 1);
          tma_store_2d(&tm_c, buf, nc, m0);
          if (kDual) tma_store_2d(&tm_aux_out, buf + EPI_BUF_BYTES, nc, m0);
          bulk_commit();
          bulk_wait_read<0>();                       // the slot has left shared memory
          if (kHasAuxIn) la_issue();                 // chunk seq + kSlots goes into it; epi_in[slot] announces it
          else mbar_arrive(&slot_free[slot]);
          ++seq;
        }
      }
      bulk_wait<0>();  // all stores globally visible before exit
    }
    __syncwarp();
  } else {
    // =========================== epilogue (warps 2..) ===========================
    const int quad = warp & 3;             // TMEM lane quadrant this warp may read
    const int grp = (warp - 2) >> 2;       // which column half of every chunk this warp handles (0 with one group)
    const int row = quad * 32 + lane;      // accumulator row owned by this thread
    uint8_t* epi_smem = smem + L::kEpiOff;
    float* s_bias = reinterpret_cast<float*>(smem + L::kBiasOff);
    const uint32_t s_bias_addr = smem_u32(s_bias);
    const int epi_tid = threadIdx.x - 64;
    constexpr bool kHasAuxIn = (kEpi == kResidual || kEpi == kGeluBwd || kEpi == kRowDot);
    constexpr bool kDual = (kEpi == kGelu);
    constexpr int kSlots = kDual ? kBufs / 2 : kBufs;
    static_assert(kEpi == kAtomicF32 || kSlots >= 2, "need at least two staging slots");

    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t seq = 0;  // running chunk sequence number (staging slot = seq % kSlots)
#ifdef PLB_GEMM_TRACE
    int gt_n = 0;
#endif

    const uint32_t lead_tmem_empty0 = (kCta == 2) ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0;
    for (UnitWalker w(p, unit0, unit_stride); w.valid(p); w.next(p)) {
      const int m0 = (w.m_blk * kCta + rank) * BM, n0 = w.n_blk * BN;
      PLB_GT(1);  // tile start
      if (kEpi != kAtomicF32 && p.bias != nullptr) {
        // stage this tile's bias in shared memory: the per-element __ldg it replaces serialised the epilogue
        // (the compiler cannot hoist global loads across the generic staging stores).  The epilogue warps no longer meet
        // at a barrier per chunk, so a fast warp must not overwrite the bias while a slow one still reads the previous
        // tile's: one barrier before the writes, one after.
        epi_bar_sync();
#pragma unroll
        for (int i = 0; i < BN / EPI_THREADS; ++i) {
          const int n = n0 + epi_tid + i * EPI_THREADS;
          s_bias[epi_tid + i * EPI_THREADS] = (n < p.N) ? __ldg(p.bias + n) : 0.f;
        }
        epi_bar_sync();
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      PLB_GT(2);  // accumulator ready
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;

      if (kEpi == kAtomicF32) {
        const int m = m0 + row;
#pragma unroll 1
        for (int c = grp; c < BN / 32; c += EPI_GROUPS) {
          if (n0 + c * 32 >= p.N) break;
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (m < p.M) {
            float* dst = p.c_f32 + static_cast<size_t>(m) * p.ldc + n0 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const int n = n0 + c * 32 + i;
              if (p.c_multimem) {
                // gradient all-reduce fused into the WGRAD epilogue: the switch adds the four values into EVERY rank's copy
                // of the arena (the host checked N % 4 == 0 and the 16-byte alignment for this mode)
                asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i),
                             "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                             "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                             : "memory");
              } else if (n + 3 < p.N && ((reinterpret_cast<uintptr_t>(dst + i) & 15) == 0)) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i),
                             "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                             "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                             : "memory");
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (n + j < p.N) atomicAdd(dst + i + j, __uint_as_float(v[i + j]));
              }
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / EPI_COLS; ++c) {
          const int nc = n0 + c * EPI_COLS;
          if (nc >= p.N) break;
          const int slot = seq % kSlots;
          uint8_t* buf = epi_smem + (kDual ? 2 * slot : slot) * EPI_BUF_BYTES;
          const uint32_t my_row = smem_u32(buf) + row * 128;

          uint32_t v0[32], v1[32];
          tmem_ld_32x32(taddr + c * EPI_COLS + grp * GRP_COLS, v0);
          if constexpr (EPI_GROUPS == 1) tmem_ld_32x32(taddr + c * EPI_COLS + 32, v1);
          if (kHasAuxIn) mbar_wait(&epi_in[slot], (seq / kSlots) & 1);
          else mbar_wait(&slot_free[slot], ((seq / kSlots) & 1) ^ 1);  // the previous use of the slot has been stored
          PLB_GT(3);  // slot / aux operand ready
          tmem_ld_wait();
          PLB_GT(4);  // accumulator chunk in registers
          f32x2 dot2 = f2_splat(0.f);  // kRowDot: this thread's part of sum_n C[m,n] * aux_in[m,n] over the chunk

#pragma unroll
          for (int jl = 0; jl < GRP_PIECES; ++jl) {  // this thread's 16-byte pieces of its 128-byte staged row
            const int j = grp * GRP_PIECES + jl;
            // four fp32 pairs (FFMA2-class packed math: the epilogue is issue-bound)
            f32x2 f[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = (jl & 3) * 8 + 2 * i;
              f[i] = (EPI_GROUPS == 2 || jl < 4) ? f2_pack(__uint_as_float(v0[e]), __uint_as_float(v0[e + 1]))
                                                 : f2_pack(__uint_as_float(v1[e]), __uint_as_float(v1[e + 1]));
            }
            if (p.bias != nullptr) {
              const float4 b0 = lds128_ro(s_bias_addr + (c * EPI_COLS + j * 8) * 4);
              const float4 b1 = lds128_ro(s_bias_addr + (c * EPI_COLS + j * 8 + 4) * 4);
              f[0] = f2_add(f[0], f2_pack(b0.x, b0.y)); f[1] = f2_add(f[1], f2_pack(b0.z, b0.w));
              f[2] = f2_add(f[2], f2_pack(b1.x, b1.y)); f[3] = f2_add(f[3], f2_pack(b1.z, b1.w));
            }
            const uint32_t piece = my_row + ((j ^ (row & 7)) << 4);
            uint32_t aw[4] = {0u, 0u, 0u, 0u};
            if (kHasAuxIn) {
              const uint4 a = lds128(piece);
              aw[0] = a.x; aw[1] = a.y; aw[2] = a.z; aw[3] = a.w;
              if (kEpi != kRowDot) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const f32x2 x = f2_from_bf16x2(aw[i]);
                  if (kEpi == kResidual) f[i] = f2_add(f[i], x);
                  else f[i] = f2_mul(f[i], x);  // kGeluBwd: acc * gelu'(u), the derivative saved by the forward epilogue
                }
              }
            }
            if (kDual) {
              // gelu'(u) goes to the second buffer of the pair (all the backward pass needs of u: its DGRAD epilogue becomes
              // one multiply, and this epilogue is bound by its two stores, not by its math), gelu(u) to the first.  Both are
              // evaluated on the bf16-rounded u, the value a stored pre-activation would have had.
              f32x2 gp[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) gelu_new_fwd_grad_x2(f2_from_bf16x2(f2_to_bf16x2(f[i])), f[i], gp[i]);
              uint4 pu;
              pu.x = f2_to_bf16x2(gp[0]); pu.y = f2_to_bf16x2(gp[1]);
              pu.z = f2_to_bf16x2(gp[2]); pu.w = f2_to_bf16x2(gp[3]);
              sts128(my_row + EPI_BUF_BYTES + ((j ^ (row & 7)) << 4), pu);
            }
            uint4 o;
            o.x = f2_to_bf16x2(f[0]); o.y = f2_to_bf16x2(f[1]);
            o.z = f2_to_bf16x2(f[2]); o.w = f2_to_bf16x2(f[3]);
            if (kEpi == kRowDot) {  // the dot product runs over the ROUNDED outputs: what the consumer of C will read
              dot2 = f2_fma(f2_from_bf16x2(o.x), f2_from_bf16x2(aw[0]), dot2);
              dot2 = f2_fma(f2_from_bf16x2(o.y), f2_from_bf16x2(aw[1]), dot2);
              dot2 = f2_fma(f2_from_bf16x2(o.z), f2_from_bf16x2(aw[2]), dot2);
              dot2 = f2_fma(f2_from_bf16x2(o.w), f2_from_bf16x2(aw[3]), dot2);
            }
            sts128(piece, o);
          }
          if (kEpi == kRowDot) {
            float lo, hi;
            f2_unpack(dot2, lo, hi);
            const int m = m0 + row;
            // out-of-range columns of the chunk contribute exact zeros (TMA zero-fills aux_in and B)
            if (m < p.M) atomicAdd(p.rowdot + static_cast<size_t>(nc / EPI_COLS) * p.rowdot_ld + m, lo + hi);
          }
          PLB_GT(5);  // math done, chunk staged
          fence_proxy_async_smem();
          PLB_GT(6);  // proxy fence done
          __syncwarp();
          if (lane == 0) mbar_arrive(&chunk_ready[slot]);  // one arrival per epilogue warp: the store warp takes it from here
          ++seq;
        }
      }
      // accumulator drained -> hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCta == 2) mbar_arrive_cluster(lead_tmem_empty0 + acc * 8);  // the leader's MMA warp waits for both CTAs
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      PLB_GT(7);  // tile handed back
    }
  }

  tc_fence_before();
  if (kCta == 2) cluster_sync_all(); else __syncthreads();  // the peer may still read this CTA's smem / signal its barriers
  if (warp == 1) {
    if (kCta == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <bool kAMn, bool kBMn, int kEpi, int kStages, int kBufs, int kCta>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tin,
           const CUtensorMap& tout, const GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<kStages, kBufs, kCta>;
  static_assert(L::kAlloc <= 232448, "GEMM stage ring exceeds the 227 KB shared-memory limit");
  auto kern = gemm_kernel<kAMn, kBMn, kEpi, kStages, kBufs, kCta>;
  static bool configured = false;  // benign race: the attribute call is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc);
    PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "gemm: cudaFuncSetAttribute(smem=%d): %s", L::kAlloc,
                cudaGetErrorString(e));
    configured = true;
  }
  const int max_clusters = sm_count() / kCta;
  const int clusters = p.num_units < max_clusters ? p.num_units : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * kCta);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::kAlloc;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = launch_overlap_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, tin, tout, p);
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "plb_gemm_bf16: launch failed: %s", cudaGetErrorString(e));
  return check_launch("plb_gemm_bf16");
}

}  // namespace

}  // namespace plb

using namespace plb;

// CTA pairs (cta_group::2) halve the B-operand shared-memory traffic per SM; a single CTA is kept for problems with
// fewer than two 128-row tiles.
static int pick_ctas(int M) { return M > BM ? 2 : 1; }

static int gemm_dispatch(int op, int epilogue, int M, int N, int K, const void* A, int lda, const void* B,
                         int ldb, void* C, int ldc, const float* bias, const void* aux_in, int ld_aux_in,
                         void* aux_out, int ld_aux_out, plb_stream_t stream_, int c_multimem) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLB_REQUIRE(M > 0 && N > 0 && K > 0, PLB_ERR_ARG, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  PLB_REQUIRE(A && B && C, PLB_ERR_ARG, "gemm: null operand");
  PLB_REQUIRE(op >= PLB_GEMM_FWD && op <= PLB_GEMM_WGRAD, PLB_ERR_ARG, "gemm: bad op %d", op);
  PLB_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0, PLB_ERR_ARG, "gemm: lda/ldb must be multiples of 8 (got %d, %d)",
              lda, ldb);
  PLB_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0, PLB_ERR_ARG, "gemm: A/B must be 16-byte aligned");

  const int ctas = pick_ctas(M);
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + BM * ctas - 1) / (BM * ctas);
  p.n_tiles = (N + BN - 1) / BN;
  // Units run concurrently in launch order, so the operand that consecutive units SHARE is read from HBM once and the other
  // one is streamed once per tile row/column.  Stream the smaller one: for the layer GEMMs (M = tokens >> N) that is the
  // weight matrix (walk along N); for the 64 k-row token head on the masked rows (N*K = 98 MB of weights vs 17 MB of
  // activations) walking along N would pull the whole weight matrix through L2 once per 256 activation rows.
  p.m_fast = (op != PLB_GEMM_WGRAD && static_cast<long long>(N) > static_cast<long long>(M)) ? 1 : 0;
  p.k_iters = (K + BK - 1) / BK;
  p.bias = bias;
  p.splits = 1;

  CUtensorMap ta, tb, tc, tin, tout;
  bool ok;
  if (op == PLB_GEMM_FWD) {
    ok = tensor_map_bf16(&ta, A, M, K, lda, BM, 64) && tensor_map_bf16(&tb, B, N, K, ldb, BN / ctas, 64);
  } else if (op == PLB_GEMM_DGRAD) {
    ok = tensor_map_bf16(&ta, A, M, K, lda, BM, 64) && tensor_map_bf16(&tb, B, K, N, ldb, 64, 64);
  } else {
    ok = tensor_map_bf16(&ta, A, K, M, lda, 64, 64) && tensor_map_bf16(&tb, B, K, N, ldb, 64, 64);
  }
  if (!ok) return PLB_ERR_CUDA;

#define PLB_LAUNCH(AMN, BMN, EPI, S1, B1, S2, B2)                                                      \
  return ctas == 2 ? launch<AMN, BMN, EPI, S2, B2, 2>(ta, tb, tc, tin, tout, p, stream)                \
                   : launch<AMN, BMN, EPI, S1, B1, 1>(ta, tb, tc, tin, tout, p, stream)

  if (op == PLB_GEMM_WGRAD) {
    PLB_REQUIRE(epilogue == PLB_EPI_NONE && bias == nullptr, PLB_ERR_ARG, "gemm: WGRAD takes no epilogue/bias");
    PLB_REQUIRE(ldc >= N, PLB_ERR_ARG, "gemm: ldc < N");
    // Split K so that the units fill whole waves of the persistent grid: with P workers (CTAs or CTA pairs) the GEMM takes
    // ceil(units / P) rounds of one unit each, a unit costing its k-iterations plus a fixed fill / epilogue overhead
    // (~8 iterations).  Pick the split count that minimises rounds x unit cost; every unit keeps at least 8 k-iterations.
    // (The old rule "about two units per SM" left the last round 20-30 % full for the four weight shapes of the layer.)
    const int tiles = p.m_tiles * p.n_tiles;
    const int workers = sm_count() / ctas;
    const int max_splits = p.k_iters / 8 > 0 ? p.k_iters / 8 : 1;
    long best_cost = -1;
    int best_kps = p.k_iters;
    for (int sp = 1; sp <= max_splits && sp <= 4 * workers; ++sp) {
      const int kps = (p.k_iters + sp - 1) / sp;
      const int eff_splits = (p.k_iters + kps - 1) / kps;
      const long rounds = (static_cast<long>(tiles) * eff_splits + workers - 1) / workers;
      const long cost = rounds * (kps + 8);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_kps = kps;
      }
    }
    p.k_per_split = best_kps;
    p.splits = (p.k_iters + p.k_per_split - 1) / p.k_per_split;
    p.num_units = tiles * p.splits;
    p.c_f32 = static_cast<float*>(C);
    p.ldc = ldc;
    p.c_multimem = c_multimem;
    PLB_REQUIRE(!c_multimem || ((N % 4) == 0 && (ldc % 4) == 0 && ((uintptr_t)C % 16) == 0), PLB_ERR_ARG,
                "gemm: the multicast WGRAD needs N %% 4 == 0, ldc %% 4 == 0 and a 16-byte aligned C");
    tc = tin = tout = ta;
    PLB_LAUNCH(true, true, kAtomicF32, 4, 2, 6, 2);
  }

  PLB_REQUIRE((ldc % 8) == 0 && ((uintptr_t)C % 16) == 0, PLB_ERR_ARG, "gemm: C must be 16-byte aligned, ldc %% 8 == 0");
  p.k_per_split = p.k_iters;
  p.num_units = p.m_tiles * p.n_tiles;
  if (!tensor_map_bf16(&tc, C, M, N, ldc, BM, EPI_COLS)) return PLB_ERR_CUDA;
  tin = tout = tc;
  if (epilogue == PLB_EPI_ROWDOT) {
    PLB_REQUIRE(op == PLB_GEMM_DGRAD && aux_out && ld_aux_out >= M && (N % EPI_COLS) == 0 && bias == nullptr, PLB_ERR_ARG,
                "gemm: ROWDOT needs op DGRAD, an fp32 aux_out with ld_aux_out >= M, N %% 64 == 0 and no bias");
    p.rowdot = static_cast<float*>(aux_out);
    p.rowdot_ld = ld_aux_out;
  }
  if (epilogue == PLB_EPI_GELU_BWD && aux_out != nullptr) {
    PLB_REQUIRE(((uintptr_t)aux_out % 4) == 0, PLB_ERR_ARG, "gemm: GELU_BWD column-sum accumulator must be an fp32 array");
    p.colsum = static_cast<float*>(aux_out);
  }
  if (epilogue == PLB_EPI_RESIDUAL || epilogue == PLB_EPI_GELU_BWD || epilogue == PLB_EPI_ROWDOT) {
    PLB_REQUIRE(aux_in && (ld_aux_in % 8) == 0 && ((uintptr_t)aux_in % 16) == 0, PLB_ERR_ARG,
                "gemm: epilogue %d needs an aligned aux_in", epilogue);
    if (!tensor_map_bf16(&tin, aux_in, M, N, ld_aux_in, BM, EPI_COLS)) return PLB_ERR_CUDA;
  }
  if (epilogue == PLB_EPI_GELU) {
    PLB_REQUIRE(aux_out && (ld_aux_out % 8) == 0 && ((uintptr_t)aux_out % 16) == 0, PLB_ERR_ARG,
                "gemm: GELU epilogue needs an aligned aux_out");
    if (!tensor_map_bf16(&tout, aux_out, M, N, ld_aux_out, BM, EPI_COLS)) return PLB_ERR_CUDA;
  }

  if (op == PLB_GEMM_FWD) {
    switch (epilogue) {
      case PLB_EPI_NONE:     PLB_LAUNCH(false, false, kNone, 4, 2, 6, 2);
      case PLB_EPI_RESIDUAL: PLB_LAUNCH(false, false, kResidual, 3, 4, 5, 4);
#ifndef PLB_GELU_S2
#define PLB_GELU_S2 5
#define PLB_GELU_B2 4
#endif
      case PLB_EPI_GELU:     PLB_LAUNCH(false, false, kGelu, 3, 4, PLB_GELU_S2, PLB_GELU_B2);
      default: break;
    }
  } else {
    switch (epilogue) {
      case PLB_EPI_NONE:     PLB_LAUNCH(false, true, kNone, 4, 2, 6, 2);
      case PLB_EPI_RESIDUAL: PLB_LAUNCH(false, true, kResidual, 3, 4, 5, 4);
      case PLB_EPI_GELU_BWD: PLB_LAUNCH(false, true, kGeluBwd, 3, 4, 5, 4);
      case PLB_EPI_ROWDOT:   PLB_LAUNCH(false, true, kRowDot, 3, 4, 5, 4);
      default: break;
    }
  }
#undef PLB_LAUNCH
  set_error("gemm: op %d does not support epilogue %d", op, epilogue);
  return PLB_ERR_UNSUPPORTED;
}

extern "C" int plb_gemm_bf16(int op, int epilogue, int M, int N, int K, const void* A, int lda, const void* B,
                             int ldb, void* C, int ldc, const float* bias, const void* aux_in, int ld_aux_in,
                             void* aux_out, int ld_aux_out, plb_stream_t stream) {
  return gemm_dispatch(op, epilogue, M, N, K, A, lda, B, ldb, C, ldc, bias, aux_in, ld_aux_in, aux_out, ld_aux_out, stream, 0);
}

extern "C" int plb_gemm_wgrad_allreduce(int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                                        float* C_multicast, int ldc, plb_stream_t stream) {
  PLB_REQUIRE(C_multicast != nullptr, PLB_ERR_ARG, "gemm_wgrad_allreduce: null multicast address");
  return gemm_dispatch(PLB_GEMM_WGRAD, PLB_EPI_NONE, M, N, K, A, lda, B, ldb, C_multicast, ldc, nullptr, nullptr, 0, nullptr, 0,
                       stream, 1);
}

#ifdef PLB_GEMM_TRACE
// trace library only: copies the timeline of the last launches (tag << 56 | clock64) to the host
extern "C" int plb_debug_gemm_trace(unsigned long long* host_out, int n) {
  if (n > 8192) n = 8192;
  cudaError_t e = cudaMemcpyFromSymbol(host_out, plb::g_gemm_trace, sizeof(unsigned long long) * n);
  return e == cudaSuccess ? 0 : -1;
}
extern "C" int plb_debug_gemm_trace_clear() {
  static unsigned long long zeros[8192];
  cudaError_t e = cudaMemcpyToSymbol(plb::g_gemm_trace, zeros, sizeof(zeros));
  return e == cudaSuccess ? 0 : -1;
}
#endif
