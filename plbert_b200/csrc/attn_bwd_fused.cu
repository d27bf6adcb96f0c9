// plbert_b200 — attention backward as ONE kernel on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators).
//
// Replaces autograd of HF modeling_albert.py:184-197 (SDPA + key-padding mask); math of SURVEY.md appendix A:
//     P = exp2(S * k - lse2),  dP = dO V^T,  dS = P o (dP - delta),  dQ = s dS K,  dK = s dS^T Q,  dV = P^T dO.
// The two-kernel backward (attn_tc.cu: a dQ kernel that owns query tiles and a dK/dV kernel that owns key tiles) computes
// S, dP and the exponentials TWICE (14 L^2*64 MMA units, 2 L^2 exp).  Here every [128 keys x 128 queries] block is
// visited once (10 units, 1 L^2 exp):
//     S^T  = K_t Q_j^T          tcgen05.mma M128 N128 K64  (smem x smem)      -> TMEM [  0,128)
//     dP^T = V_t dO_j^T         tcgen05.mma M128 N128 K64  (smem x smem)      -> TMEM [128,256)
//     P^T, dS^T                 8 softmax warps (32 key rows x 64 query columns each, matrix-fragment register layout);
//                               one exp2 per element
//        P^T  -> TMEM [256,320) packed bf16 (A operand of the dV MMA)
//        dS^T -> shared memory [128 keys x 128 q] bf16, 128-byte swizzled, DOUBLE-buffered (read by two MMAs under two views)
//     dV  += P^T  dO_j          M128 N64 K128, A from TMEM, B = dO_j MN-major            -> TMEM [384,448)   (issued first)
//     dQ_j = dS_j K_t           M128 N64 K128, A = dS^T read MN-major (i.e. transposed), B = K_t MN-major -> TMEM [448,512)
//     dK  += dS^T Q_j           M128 N64 K128, A = dS^T K-major, B = Q_j MN-major        -> TMEM [320,384)
// dK / dV accumulate in tensor memory over the query tiles of a key tile.  dQ_j gets one partial per KEY tile (<= 4): the
// partials are summed in an fp32 scratch of the CTA (128 KB, L2-resident: store for the first tile, red.add for the middle
// ones, read + add + convert for the last), because 512 queries x 64 fp32 accumulators do not fit beside S^T / dP^T.
//
// One persistent CTA per SM walks the (sequence, head) items.  Shared memory (224 KB of tiles): Q of the item resident
// (4 x 16 KB), dO_j streamed per block through 3 slots, K_t double-buffered, V_t single (its only reader, the dP^T MMA, runs
// early in a block), dS^T 2 x 32 KB.  16 warps:
//   warp 0   TMA producer of Q / K / V         warp 1   MMA issuer A: the S^T / dP^T stream
//   warp 2   MMA issuer B: dV / dQ / dK        warp 3   per-block stager: dO_j (TMA) and (-lse2, -delta) of the query tile
//   warps 4..11   softmax (P^T, dS^T)          warps 12..15   epilogue (dQ partials, dK / dV tiles, bias column sums)
// All roles walk the same block sequence n = 0, 1, ... (item-major, then key tile, then query tile); every mbarrier phase
// is a function of n (or of the running key-tile / slot-use counters), so the pipeline keeps running across items.
//
// What bounds it (profiles/r2_attn_bwd_fused.md): not one pipe — tensor 28 %, MUFU 22 %, issue 36 %, and the shared-memory /
// L1 data pipe, which carries the MMA operand reads (30 % of its wavefronts) AND every LDS / STS / global store of the
// roles (47 %), at 77 %.  Hence the design rules of this file: no staging of outputs through shared memory, statistics
// fetched in the fragment layout (a lane loads only its own column pairs), dO streamed instead of kept resident.
#include "common.cuh"
#include "rng.cuh"

namespace plb {
namespace {

constexpr int D = 64;
constexpr int TILE16K = 128 * 64 * 2;       // one [128 x 64] bf16 swizzled tile
constexpr int MAX_TILES = 4;                // sequences up to 512 rows
constexpr int FB_THREADS = 512;
// setmaxnreg moves registers inside the launch allocation (512 x 128 = 65536): 128*56 + 256*168 + 128*120 = 65536
constexpr int FB_REGS_CONTROL = 56;
constexpr int FB_REGS_SOFTMAX = 168;
constexpr int FB_REGS_EPI = 120;
constexpr int SCRATCH_FLOATS = MAX_TILES * 128 * D;   // fp32 dQ partial sums of one CTA: 512 queries x 64

struct FbSmem {
  static constexpr int kQ = 0;                               // 4 slots [128 q x 64]: Q of the item, resident
  static constexpr int kDO = kQ + MAX_TILES * TILE16K;       // 3 slots: dO_j of blocks n, n+1, n+2 (streamed per block)
  static constexpr int kK = kDO + 3 * TILE16K;               // 2 slots [128 keys x 64]
  static constexpr int kV = kK + 2 * TILE16K;                // 1 slot: V_t is only read by the dP^T MMAs, early in every block
  static constexpr int kDS = kV + TILE16K;                   // dS^T: 2 buffers x 2 sub-tiles [128 keys x 64 q]
  static constexpr int kStat = kDS + 4 * TILE16K;            // 2 buffers x (128 x -lse2 | 128 x -delta) fp32
  static constexpr int kBar = kStat + 2 * 256 * 4;
  // q_full[4], q_free[4], k_full[2], k_free[2], v_full, v_free, do_full[3], do_free[3], pipeline, stat_full[2], stat_free[2]
  static constexpr int kNumBars = 4 + 4 + 2 + 2 + 2 + 3 + 3 + 11 + 4;
  static constexpr int kAlloc = kBar + kNumBars * 8 + 16;
};
static_assert(FbSmem::kAlloc <= 232448, "fused attention backward exceeds the 227 KB shared-memory limit");

struct FbParams {
  const int32_t* seq_start;
  const int32_t* q_len;
  const int32_t* kv_len;
  int heads;
  int total_rows;
  float scale_log2;  // log2(e) / sqrt(d)
  float scale;       // 1 / sqrt(d)
};

struct FbItem {
  int h, row_base, nq, nk, n_kt, n_qt;
};
__device__ __forceinline__ FbItem fb_item(const FbParams& p, int item) {
  FbItem it;
  const int b = item / p.heads;
  it.h = item % p.heads;
  it.row_base = p.seq_start[b];
  it.nq = p.q_len[b];
  it.nk = p.kv_len[b];
  // (lengths are >= 1 by contract; the clamps only keep the roles' bookkeeping consistent on bad input)
  it.n_kt = min(MAX_TILES, max(1, (it.nk + 127) >> 7));
  it.n_qt = min(MAX_TILES, max(1, (it.nq + 127) >> 7));
  return it;
}

// per-slot use counters of the four Q / dO slots, one byte each (only the parity matters; 256 is even)
__device__ __forceinline__ uint32_t slot_count(uint32_t packed, int s) { return (packed >> (8 * s)) & 0xffu; }

__device__ __forceinline__ void st_global_f4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_global_f4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 32 bytes per lane (one full L2 sector): the bf16 output rows are 4608 bytes apart, so a warp's store touches 32 lines either
// way; a 16-byte store leaves every sector half written until its partner arrives
__device__ __forceinline__ void st_global_b256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
// L2-only load: the scratch lines are modified by red operations, which bypass L1
__device__ __forceinline__ float4 ld_global_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// Column sums of a [32 rows (lanes) x 16 columns] register tile, added to dst[0..16) with atomics: reduce-scatter
// butterfly (8 + 4 + 2 + 1 exchanges halve the live values) and one plain exchange; even lanes end up owning a column.
// (Staging the output tiles through a dead shared-memory slot for coalesced stores was measured SLOWER, 679 vs 635 us:
// the kernel is bound by the shared-memory / L1 data pipe — LSU 55 % + tensor-core operand reads 29 % of its wavefronts —
// and staging adds wavefronts, profiles/r2_attn_bwd_fused.md.)
__device__ __forceinline__ void colsum16_atomic(const float (&v)[16], float* dst, int lane) {
  float w8[8], w4[4], w2[2];
  bool up = (lane & 16) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
    w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  up = (lane & 8) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = up ? w8[i] : w8[i + 4], keep = up ? w8[i + 4] : w8[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  up = (lane & 4) != 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = up ? w4[i] : w4[i + 2], keep = up ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  up = (lane & 2) != 0;
  const float send = up ? w2[0] : w2[1], keep = up ? w2[1] : w2[0];
  float w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
  const int col = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
  if ((lane & 1) == 0) atomicAdd(dst + col, w1);
}

// Optional timeline instrumentation (profiles/attn_trace.py builds a separate library with -DPLB_ATTN_TRACE): lane 0 of
// every warp of one CTA appends (clock, warp, code, index) records.  Compiled out of the product library.
#ifdef PLB_ATTN_TRACE
#ifndef PLB_TRACE_BLOCK
#define PLB_TRACE_BLOCK 0
#endif
constexpr unsigned FB_TRACE_SLOTS = 1024;
constexpr unsigned FB_TRACE_CAP = 16 * FB_TRACE_SLOTS;
__device__ unsigned long long g_fb_trace[FB_TRACE_CAP];
#define FB_TRACE_INIT() unsigned trace_k = 0; const unsigned trace_w = threadIdx.x >> 5
#define FB_TRACE(code, idx)                                                                                             \
  do {                                                                                                                  \
    if (blockIdx.x == PLB_TRACE_BLOCK && (threadIdx.x & 31) == 0 && trace_k < FB_TRACE_SLOTS)                           \
      g_fb_trace[trace_w * FB_TRACE_SLOTS + trace_k++] = (static_cast<unsigned long long>(clock64()) << 24) |           \
                                                         (trace_w << 16) | ((code) << 8) | ((idx) & 255);               \
  } while (0)
#else
#define FB_TRACE_INIT() do { } while (0)
#define FB_TRACE(code, idx) do { } while (0)
#endif

// pipeline barriers (phase = block counter n unless noted)
enum { ST_FULL = 0, ST_FREE, PDS_READY, PV_FREE, DQ_FULL, ACC_DONE, DQ_FREE, DV_FREE, DK_FREE /* phase = key-tile counter */,
       DS_FREE0, DS_FREE1 /* one per dS^T buffer: phase = n >> 1 */, NUM_PIPE_BARS };

// multiplier (0 or 1 / (1 - p)) of attention probability (item, query q, key k): the forward kernel's mask, regenerated
__device__ __forceinline__ float drop_mult(const DropoutParams& d, uint32_t item, uint32_t q, uint32_t k) {
  const uint32_t x = rng32(d.seed_lo, d.seed_hi, d.stream, attn_dropout_ctr(item, q, k));
  return ((k & 1u) ? (x >> 16) : (x & 0xffffu)) < d.thresh16 ? d.scale : 0.f;
}

// kDrop: attention-probability dropout (HF :134 / :194).  With O = (P o m) V the chain rule gives dV = (P o m)^T dO,
// dP = (dO V^T) o m and dS = P o (dP - delta) with the UNdropped P and delta = rowsum(dO o O) as before.
template <bool kDrop>
__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_dctx,
                      const FbParams p, int n_items, const float* __restrict__ lse2, const float* __restrict__ delta,
                      __nv_bfloat16* __restrict__ dqkv, int ld_dqkv, float* __restrict__ dbias,
                      float* __restrict__ scratch_all, const DropoutParams drop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();  // 128-byte swizzle needs 1024-byte aligned tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FbSmem::kBar);
  uint64_t* q_full = bars;               // [4]  Q_j of the item has landed in slot s
  uint64_t* q_free = bars + 4;           // [4]  every MMA of the item that reads slot s has completed
  uint64_t* k_full = bars + 8;           // [2]  K_t in slot T & 1
  uint64_t* k_free = bars + 10;          // [2]
  uint64_t* v_full = bars + 12;          //      V_t (single slot)
  uint64_t* v_free = bars + 13;          //      the dP^T MMAs of the tile's last block have completed
  uint64_t* do_full = bars + 14;         // [3]  dO_j of block n in slot n % 3
  uint64_t* do_free = bars + 17;         // [3]  dP^T (issuer A) and dV (issuer B) of the block have read it
  uint64_t* pipe = bars + 20;            // [NUM_PIPE_BARS]
  uint64_t* stat_full = pipe + NUM_PIPE_BARS;   // [2]
  uint64_t* stat_free = stat_full + 2;          // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + FbSmem::kNumBars);
  static_assert(20 + NUM_PIPE_BARS + 4 == FbSmem::kNumBars, "barrier layout");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.heads * D;
  FB_TRACE_INIT();
  FB_TRACE(0, 0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_dctx);
    for (int i = 0; i < 17; ++i) mbar_init(&bars[i], 1);   // q_full, q_free, k_full, k_free, v_full, v_free, do_full
    for (int i = 0; i < 3; ++i) mbar_init(&do_free[i], 2);  // issuer A (dP^T) and issuer B (dV)
    mbar_init(&pipe[ST_FULL], 1);
    mbar_init(&pipe[ST_FREE], 8);       // one arrive per softmax warp
    mbar_init(&pipe[PDS_READY], 8);
    mbar_init(&pipe[PV_FREE], 1);
    mbar_init(&pipe[DQ_FULL], 1);
    mbar_init(&pipe[ACC_DONE], 1);
    mbar_init(&pipe[DQ_FREE], 4);       // one arrive per epilogue warp
    mbar_init(&pipe[DV_FREE], 4);
    mbar_init(&pipe[DK_FREE], 4);
    mbar_init(&pipe[DS_FREE0], 1);
    mbar_init(&pipe[DS_FREE1], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&stat_full[i], 1);
      mbar_init(&stat_free[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_grid_sync();   // the prologue above is CTA-local; everything below reads / writes global memory
  // TMEM columns: S^T [0,128)  dP^T [128,256)  P^T bf16 pairs [256,320)  dK [320,384)  dV [384,448)  dQ [448,512)

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FB_REGS_CONTROL));
    if (warp == 0) {
      // ====================================== TMA producer ======================================
      if (elect_one_sync()) {
        uint32_t T = 0, qbase = 0, quse = 0;
        FbItem nxt = blockIdx.x < n_items ? fb_item(p, blockIdx.x) : FbItem{};
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
          const FbItem it = nxt;
          if (item + static_cast<int>(gridDim.x) < n_items) nxt = fb_item(p, item + gridDim.x);
          auto load_k = [&](int t) {
            const uint32_t Tt = T + t, s = Tt & 1;
            if (Tt >= 2) mbar_wait(&k_free[s], ((Tt >> 1) - 1) & 1);   // tile Tt - 2 has left the slot
            mbar_arrive_expect_tx(&k_full[s], TILE16K);
            tma_load_2d(smem + FbSmem::kK + s * TILE16K, &tm_qkv, &k_full[s], H + it.h * D, it.row_base + t * 128);
          };
          auto load_v = [&](int t) {
            const uint32_t Tt = T + t;
            if (Tt >= 1) mbar_wait(v_free, (Tt - 1) & 1);               // the previous tile's last dP^T has run
            mbar_arrive_expect_tx(v_full, TILE16K);
            tma_load_2d(smem + FbSmem::kV, &tm_qkv, v_full, 2 * H + it.h * D, it.row_base + t * 128);
          };
          auto load_q = [&](int j) {
            const int s = (qbase + j) & 3;
            const uint32_t uses = slot_count(quse, s);
            if (uses >= 1) mbar_wait(&q_free[s], (uses - 1) & 1);
            quse = (quse & ~(0xffu << (8 * s))) | (((uses + 1) & 0xffu) << (8 * s));
            mbar_arrive_expect_tx(&q_full[s], TILE16K);
            tma_load_2d(smem + FbSmem::kQ + s * TILE16K, &tm_qkv, &q_full[s], it.h * D, it.row_base + j * 128);
          };
          // In the order of first use; every wait is for MMAs that need nothing this thread issues later (dO comes from
          // warp 3).  Q slots rotate, so short items prefetch across item borders.
          load_k(0);
          load_v(0);
          for (int j = 0; j < it.n_qt; ++j) load_q(j);
          for (int t = 1; t < it.n_kt; ++t) {
            load_k(t);
            load_v(t);
          }
          T += it.n_kt;
          qbase += it.n_qt;
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ====================================== MMA issuer A: S^T = K_t Q_j^T, dP^T = V_t dO_j^T ======================================
      if (elect_one_sync()) {
        constexpr uint32_t idesc_st = umma_idesc_bf16(128, 128, 0, 0);
        const uint64_t dQk = umma_smem_desc(smem_u32(smem + FbSmem::kQ), 16, 1024);
        const uint64_t dDOk = umma_smem_desc(smem_u32(smem + FbSmem::kDO), 16, 1024);
        const uint64_t dKk = umma_smem_desc(smem_u32(smem + FbSmem::kK), 16, 1024);
        const uint64_t dVk = umma_smem_desc(smem_u32(smem + FbSmem::kV), 16, 1024);
        uint32_t n = 0, T = 0, qbase = 0, quse = 0, dslot = 0, dphase = 0;   // dO ring: slot n % 3, phase (n / 3) & 1
        FbItem nxt = blockIdx.x < n_items ? fb_item(p, blockIdx.x) : FbItem{};
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
          const FbItem it = nxt;
          if (item + static_cast<int>(gridDim.x) < n_items) nxt = fb_item(p, item + gridDim.x);
          for (int t = 0; t < it.n_kt; ++t, ++T) {
            const uint64_t kv_off = static_cast<uint64_t>((T & 1) * (TILE16K >> 4));
            for (int j = 0; j < it.n_qt; ++j, ++n) {
              const int qs = (qbase + j) & 3;
              if (t == 0) mbar_wait(&q_full[qs], slot_count(quse, qs) & 1);
              if (j == 0) {
                mbar_wait(&k_full[T & 1], (T >> 1) & 1);
                mbar_wait(v_full, T & 1);
              }
              mbar_wait(&do_full[dslot], dphase);
              FB_TRACE(7, n);
              if (n >= 1) mbar_wait(&pipe[ST_FREE], (n - 1) & 1);   // every softmax warp has pulled block n-1 out of TMEM
              tc_fence_after();
              FB_TRACE(8, n);
              const uint64_t q_off = static_cast<uint64_t>(qs * (TILE16K >> 4)), do_off = static_cast<uint64_t>(dslot * (TILE16K >> 4));
#pragma unroll
              for (int kk = 0; kk < D / 16; ++kk)
                umma_bf16(tmem_base, dKk + kv_off + 2 * kk, dQk + q_off + 2 * kk, idesc_st, kk > 0);
#pragma unroll
              for (int kk = 0; kk < D / 16; ++kk)
                umma_bf16(tmem_base + 128, dVk + 2 * kk, dDOk + do_off + 2 * kk, idesc_st, kk > 0);
              umma_commit(&pipe[ST_FULL]);
              umma_commit(&do_free[dslot]);   // (one of two arrivals: the dV MMAs of issuer B read the slot too)
              if (j == it.n_qt - 1) umma_commit(v_free);   // last reader of V_t
              if (++dslot == 3) { dslot = 0; dphase ^= 1; }
              FB_TRACE(1, n);
            }
          }
          for (int j = 0; j < it.n_qt; ++j) quse += 1u << (8 * ((qbase + j) & 3));   // (only the parity of the low 7 bits
          quse &= 0x7f7f7f7fu;                                                       //  of each byte is read)
          qbase += it.n_qt;
        }
      }
      __syncwarp();
    } else if (warp == 2) {
      // ====================================== MMA issuer B: dQ_j = dS_j K_t, dK += dS^T Q_j, dV += P^T dO_j ======================================
      if (elect_one_sync()) {
        constexpr uint32_t idesc_acc = umma_idesc_bf16(128, D, 0, 1);   // A K-major (smem) or TMEM, B MN-major
        constexpr uint32_t idesc_dq = umma_idesc_bf16(128, D, 1, 1);    // A MN-major (dS^T read transposed), B MN-major
        const uint64_t dQm = umma_smem_desc(smem_u32(smem + FbSmem::kQ), TILE16K, 1024);
        const uint64_t dDOm = umma_smem_desc(smem_u32(smem + FbSmem::kDO), TILE16K, 1024);
        const uint64_t dKm = umma_smem_desc(smem_u32(smem + FbSmem::kK), TILE16K, 1024);
        const uint64_t dDSk = umma_smem_desc(smem_u32(smem + FbSmem::kDS), 16, 1024);        // [keys x q]: q is K
        const uint64_t dDSm = umma_smem_desc(smem_u32(smem + FbSmem::kDS), TILE16K, 1024);   // [q x keys]: q is M, 2 x 64
        uint32_t n = 0, T = 0, qbase = 0, dslot = 0;
        FbItem nxt = blockIdx.x < n_items ? fb_item(p, blockIdx.x) : FbItem{};
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
          const FbItem it = nxt;
          if (item + static_cast<int>(gridDim.x) < n_items) nxt = fb_item(p, item + gridDim.x);
          for (int t = 0; t < it.n_kt; ++t, ++T) {
            const uint64_t kv_off = static_cast<uint64_t>((T & 1) * (TILE16K >> 4));
            for (int j = 0; j < it.n_qt; ++j, ++n) {
              const int qs = (qbase + j) & 3;
              const uint64_t q_off = static_cast<uint64_t>(qs * (TILE16K >> 4));
              const uint64_t do_off = static_cast<uint64_t>(dslot * (TILE16K >> 4));
              const uint64_t ds_off = static_cast<uint64_t>((n & 1) * ((2 * TILE16K) >> 4));
              mbar_wait(&pipe[PDS_READY], n & 1);                    // P^T in tensor memory, dS^T in shared memory
              FB_TRACE(2, n);
              if (j == 0 && T >= 1) mbar_wait(&pipe[DV_FREE], (T - 1) & 1);   // new key tile: the previous dV is out of TMEM
              tc_fence_after();
              // (rolled loops: 24 unrolled MMAs keep 48 descriptors live and spill the issuer's registers)
              // dV first: P^T is single-buffered in tensor memory, so it is released as early as possible
#pragma unroll 1
              for (int kk = 0; kk < 8; ++kk)
                umma_bf16_tmem_a(tmem_base + 384, tmem_base + 256 + 8 * kk, dDOm + do_off + 128 * kk, idesc_acc, (j > 0 || kk > 0));
              umma_commit(&pipe[PV_FREE]);
              umma_commit(&do_free[dslot]);
              if (++dslot == 3) dslot = 0;
              FB_TRACE(5, n);
              if (n >= 1) mbar_wait(&pipe[DQ_FREE], (n - 1) & 1);    // the epilogue has pulled the previous dQ partial out of TMEM
              tc_fence_after();
              FB_TRACE(3, n);
#pragma unroll 1
              for (int kk = 0; kk < 8; ++kk)
                umma_bf16(tmem_base + 448, dDSm + ds_off + 128 * kk, dKm + kv_off + 128 * kk, idesc_dq, kk > 0);
              umma_commit(&pipe[DQ_FULL]);                           // the epilogue drains dQ under the dK MMAs
              FB_TRACE(6, n);
              if (j == 0 && T >= 1) {   // ... and the previous dK (the epilogue converts dV first, then dK: by now both are out)
                mbar_wait(&pipe[DK_FREE], (T - 1) & 1);
                tc_fence_after();
              }
#pragma unroll 1
              for (int kk = 0; kk < 8; ++kk)
                umma_bf16(tmem_base + 320, dDSk + ds_off + static_cast<uint64_t>((kk >> 2) * (TILE16K >> 4) + 2 * (kk & 3)),
                          dQm + q_off + 128 * kk, idesc_acc, (j > 0 || kk > 0));
              umma_commit(&pipe[DS_FREE0 + (n & 1)]);                // this dS^T buffer may be rewritten (block n + 2)
              umma_commit(&pipe[ACC_DONE]);
              FB_TRACE(4, n);
              if (t == it.n_kt - 1) umma_commit(&q_free[qs]);        // last reader of Q_j in this item
              if (j == it.n_qt - 1) umma_commit(&k_free[T & 1]);     // last reader of K_t
            }
          }
          qbase += it.n_qt;
        }
      }
      __syncwarp();
    } else {
      // ====================================== warp 3: per-block staging — dO_j (TMA, 3 slots) and (-lse2, -delta) of the query tile ======================================
      float* stat = reinterpret_cast<float*>(smem + FbSmem::kStat);
      uint32_t n = 0, dslot = 0, dphase = 0;
      FbItem nxt = blockIdx.x < n_items ? fb_item(p, blockIdx.x) : FbItem{};
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const FbItem it = nxt;
        if (item + static_cast<int>(gridDim.x) < n_items) nxt = fb_item(p, item + gridDim.x);
        const size_t base = static_cast<size_t>(it.h) * p.total_rows + it.row_base;
        for (int t = 0; t < it.n_kt; ++t) {
          for (int j = 0; j < it.n_qt; ++j, ++n) {
            if (lane == 0) {
              if (n >= 3) mbar_wait(&do_free[dslot], dphase ^ 1);   // block n - 3 has read the slot
              FB_TRACE(30, n);
              mbar_arrive_expect_tx(&do_full[dslot], TILE16K);
              tma_load_2d(smem + FbSmem::kDO + dslot * TILE16K, &tm_dctx, &do_full[dslot], it.h * D, it.row_base + j * 128);
            }
            if (++dslot == 3) { dslot = 0; dphase ^= 1; }
            __syncwarp();
            const uint32_t b = n & 1;
            float nl[4], nd[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int q = j * 128 + i * 32 + lane;
              // query rows beyond the sequence hold another sequence's data: lse = +inf makes their P (and dS) exactly 0
              nl[i] = q < it.nq ? -lse2[base + q] : -INFINITY;
              nd[i] = q < it.nq ? -delta[base + q] : 0.f;
            }
            if (n >= 2) mbar_wait(&stat_free[b], ((n >> 1) - 1) & 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              stat[b * 256 + i * 32 + lane] = nl[i];
              stat[b * 256 + 128 + i * 32 + lane] = nd[i];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&stat_full[b]);
          }
        }
      }
    }
  } else if (warp < 12) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FB_REGS_SOFTMAX));
    // ====================================== softmax warps: P^T and dS^T of every block ======================================
    // A warp owns 32 key rows (its TMEM lane quadrant = warp & 3) x 64 of the block's 128 query columns (ch = (warp - 4) >> 2).
    // S^T / dP^T are pulled out of tensor memory in the MATRIX-FRAGMENT shape (16x256b: thread t holds rows t/4 and t/4 + 8,
    // column pairs 2(t%4) + 8r), not one row per thread: a lane then needs the (-lse2, -delta) of ITS column pairs only —
    // 16 eight-byte loads per block instead of 32 sixteen-byte broadcasts at four shared-memory wavefronts each (the kernel is
    // bound by the shared-memory / L1 data pipe, where those broadcasts were 21 % of all wavefronts).  P^T goes back to tensor
    // memory with the matching 16x128b shape, dS^T to shared memory with stmatrix.  No cross-column reduction exists in the
    // backward pass, so the layout is free.  Packed fp32 math.
    const int quad = warp & 3, ch = (warp - 4) >> 2;
    const uint32_t t_quad = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    // stmatrix row address of this lane inside a dS^T sub-tile: matrix m = lane >> 3 -> (row block m & 1, 16-byte chunk m >> 1)
    const int st_row = quad * 32 + ((lane >> 3) & 1) * 8 + (lane & 7);   // + 16 * half
    const uint32_t sDS_lane = smem_u32(smem + FbSmem::kDS) + ch * TILE16K + st_row * 128;
    const uint32_t stat_lane = smem_u32(smem + FbSmem::kStat) + (ch * 64 + 2 * (lane & 3)) * 4;
    const f32x2 sl2 = f2_splat(p.scale_log2);
    uint32_t n = 0;
    FbItem nxt = blockIdx.x < n_items ? fb_item(p, blockIdx.x) : FbItem{};
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const FbItem it = nxt;
      if (item + static_cast<int>(gridDim.x) < n_items) nxt = fb_item(p, item + gridDim.x);
      for (int t = 0; t < it.n_kt; ++t) {
        // key rows >= kv_len (last tile only): P = 0 and dS = 0.  Their S^T / dP^T rows are finite (rows of a neighbouring
        // sequence or TMA zero fill), so the packed results are simply cleared.  Rows of this thread: r0 + {0, 8, 16, 24}.
        const bool partial = (t + 1) * 128 > it.nk;
        const int r0 = t * 128 + quad * 32 + (lane >> 2);
        uint32_t keep[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) keep[i] = (r0 + 8 * i) < it.nk ? 0xffffffffu : 0u;
        for (int j = 0; j < it.n_qt; ++j, ++n) {
          const uint32_t b = n & 1;
          mbar_wait(&stat_full[b], (n >> 1) & 1);
          mbar_wait(&pipe[ST_FULL], n & 1);
          tc_fence_after();
          FB_TRACE(10, n);
          // rounds: (32-column chunk c32, lane half h) = (0,0) (0,1) (1,0) (1,1); results pk[c32][h][8]: word 2r + e = repeat r
          // (columns 8r + 2(t%4), +1 of the chunk), row +0 (e = 0) or +8 (e = 1) — the register order of 16x128b and stmatrix
          uint32_t pkp[4][8], pkd[4][8];
          uint32_t sv[2][16], dv[2][16];
          tmem_ld_16x256b_x4(t_quad + ch * 64, sv[0]);
          tmem_ld_16x256b_x4(t_quad + 128 + ch * 64, dv[0]);
          f32x2 nl[4], nd[4];
#pragma unroll
          for (int rd = 0; rd < 4; ++rd) {
            const int c32 = rd >> 1, h = rd & 1;
            tmem_ld_wait();   // round rd has landed; round rd + 1 is loaded under this round's math
            if (rd < 3) {
              const uint32_t ta = t_quad + (static_cast<uint32_t>(((rd + 1) & 1) * 16) << 16) + ch * 64 + ((rd + 1) >> 1) * 32;
              tmem_ld_16x256b_x4(ta, sv[(rd + 1) & 1]);
              tmem_ld_16x256b_x4(ta + 128, dv[(rd + 1) & 1]);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&pipe[ST_FREE]);   // S^T / dP^T of block n+1 may overwrite the buffers
              FB_TRACE(11, n);
            }
            if (h == 0) {   // the statistics of the chunk's column pairs serve both lane halves
              const uint32_t st = stat_lane + b * 1024 + c32 * 128;
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const float2 l = lds64_f2(st + r * 32), d = lds64_f2(st + 512 + r * 32);
                nl[r] = f2_pack(l.x, l.y);
                nd[r] = f2_pack(d.x, d.y);
              }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = 4 * r + 2 * e;
                const f32x2 x = f2_fma(f2_pack(__uint_as_float(sv[rd & 1][i]), __uint_as_float(sv[rd & 1][i + 1])), sl2, nl[r]);
                float x0, x1;
                f2_unpack(x, x0, x1);
                const float p0 = ex2_fast(x0), p1 = ex2_fast(x1);
                const f32x2 dp2 = f2_pack(__uint_as_float(dv[rd & 1][i]), __uint_as_float(dv[rd & 1][i + 1]));
                if (kDrop) {
                  const uint32_t q0 = static_cast<uint32_t>(j * 128 + ch * 64 + c32 * 32 + 8 * r + 2 * (lane & 3));
                  const uint32_t kr = static_cast<uint32_t>(r0 + 16 * h + 8 * e);
                  const float m0 = drop_mult(drop, static_cast<uint32_t>(item), q0, kr);
                  const float m1 = drop_mult(drop, static_cast<uint32_t>(item), q0 + 1, kr);
                  const f32x2 tt = f2_fma(dp2, f2_pack(m0, m1), nd[r]);
                  pkp[rd][2 * r + e] = pack_bf16x2(p0 * m0, p1 * m1);
                  pkd[rd][2 * r + e] = f2_to_bf16x2(f2_mul(f2_pack(p0, p1), tt));
                } else {
                  const f32x2 tt = f2_add(dp2, nd[r]);
                  pkp[rd][2 * r + e] = pack_bf16x2(p0, p1);
                  pkd[rd][2 * r + e] = f2_to_bf16x2(f2_mul(f2_pack(p0, p1), tt));
                }
              }
            }
            if (partial) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                pkp[rd][i] &= keep[2 * h + (i & 1)];
                pkd[rd][i] &= keep[2 * h + (i & 1)];
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&stat_free[b]);
          FB_TRACE(12, n);
          // dS^T is double-buffered (block n - 2 used this buffer): its stmatrix stores and the proxy fence run first, so that
          // the wait for P^T's release (the dV MMAs of block n - 1, issued first by issuer B) hides under them
          if (n >= 2) mbar_wait(&pipe[DS_FREE0 + b], ((n >> 1) - 1) & 1);
          FB_TRACE(15, n);
#pragma unroll
          for (int rd = 0; rd < 4; ++rd) {
            const int c32 = rd >> 1, h = rd & 1;
            // lane -> row (st_row + 16 h) of matrix m = lane >> 3: chunk 4 c32 + 2 k + (m >> 1) for the k-th instruction
            const uint32_t base = sDS_lane + b * (2 * TILE16K) + h * (16 * 128);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int chunk16 = 4 * c32 + 2 * k;
              stmatrix_x4(base + (((chunk16 + (lane >> 4)) ^ (lane & 7)) << 4), pkd[rd][4 * k], pkd[rd][4 * k + 1], pkd[rd][4 * k + 2],
                          pkd[rd][4 * k + 3]);
            }
          }
          fence_proxy_async_smem();   // generic-proxy stores of dS^T -> visible to the tensor core's async-proxy reads
          FB_TRACE(16, n);
          if (n >= 1) mbar_wait(&pipe[PV_FREE], (n - 1) & 1);
          tc_fence_after();
          FB_TRACE(13, n);
#pragma unroll
          for (int rd = 0; rd < 4; ++rd)   // packed columns: 32 per 64-column half, 16 per chunk
            tmem_st_16x128b_x4(t_quad + (static_cast<uint32_t>((rd & 1) * 16) << 16) + 256 + ch * 32 + (rd >> 1) * 16, pkp[rd]);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pipe[PDS_READY]);
          FB_TRACE(14, n);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FB_REGS_EPI));
    // ====================================== epilogue warps: dQ partial sums, dK / dV tiles, bias column sums ======================================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    // scratch layout [16 column groups of 4][512 queries][4] fp32: a warp's 16-byte accesses are contiguous
    float* scratch = scratch_all + static_cast<size_t>(blockIdx.x) * SCRATCH_FLOATS;
    uint32_t n = 0;
    FbItem nxt = blockIdx.x < n_items ? fb_item(p, blockIdx.x) : FbItem{};
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const FbItem it = nxt;
      if (item + static_cast<int>(gridDim.x) < n_items) nxt = fb_item(p, item + gridDim.x);
      __nv_bfloat16* out_head = dqkv + static_cast<size_t>(it.row_base) * ld_dqkv + it.h * D;
      // key rows in [128 * n_kt, nq): padded keys inside the allocated sequence -> their gradients are exactly zero
      for (int t = it.n_kt; t * 128 < it.nq; ++t) {
        const int key = t * 128 + row;
        if (key < it.nq) {
          __nv_bfloat16* dst = out_head + static_cast<size_t>(key) * ld_dqkv;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<uint4*>(dst + H + c * 8) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(dst + 2 * H + c * 8) = make_uint4(0, 0, 0, 0);
          }
        }
      }
      for (int t = 0; t < it.n_kt; ++t) {
        const bool first = t == 0, last = t == it.n_kt - 1;
        for (int j = 0; j < it.n_qt; ++j, ++n) {
          // ---- dQ partial of (t, j) ----
          const int q = j * 128 + row;
          float* sp = scratch + static_cast<size_t>(q) * 4;   // column group g is 512 * 4 floats further
          if (!last) {
            mbar_wait(&pipe[DQ_FULL], n & 1);
            tc_fence_after();
            FB_TRACE(20, n);
            uint32_t a[64];
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) tmem_ld_32x16(t_row + 448 + pc * 16, *reinterpret_cast<uint32_t(*)[16]>(&a[pc * 16]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pipe[DQ_FREE]);
            FB_TRACE(21, n);
#pragma unroll
            for (int g = 0; g < 16; ++g) {
              float* dst = sp + g * 2048;
              if (first) st_global_f4(dst, __uint_as_float(a[g * 4]), __uint_as_float(a[g * 4 + 1]), __uint_as_float(a[g * 4 + 2]), __uint_as_float(a[g * 4 + 3]));
              else red_global_f4(dst, __uint_as_float(a[g * 4]), __uint_as_float(a[g * 4 + 1]), __uint_as_float(a[g * 4 + 2]), __uint_as_float(a[g * 4 + 3]));
            }
          } else {
            // last key tile: the partial sums of the earlier tiles come back from L2 while the MMAs of this block run
            float v[64];
            if (!first) {
#pragma unroll
              for (int g = 0; g < 16; ++g) {
                const float4 s4 = ld_global_cg_f4(sp + g * 2048);
                v[g * 4] = s4.x; v[g * 4 + 1] = s4.y; v[g * 4 + 2] = s4.z; v[g * 4 + 3] = s4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 64; ++i) v[i] = 0.f;
            }
            mbar_wait(&pipe[DQ_FULL], n & 1);
            tc_fence_after();
            FB_TRACE(20, n);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t a[32];
              tmem_ld_32x16(t_row + 448 + hf * 32, *reinterpret_cast<uint32_t(*)[16]>(&a[0]));
              tmem_ld_32x16(t_row + 448 + hf * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&a[16]));
              tmem_ld_wait();
              if (hf == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pipe[DQ_FREE]);
                FB_TRACE(21, n);
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) v[hf * 32 + i] += __uint_as_float(a[i]);
            }
            const bool q_ok = q < it.nq;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
              uint32_t w[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(v[pc * 16 + 2 * i] * p.scale, v[pc * 16 + 2 * i + 1] * p.scale);
              if (q_ok) st_global_b256(out_head + static_cast<size_t>(q) * ld_dqkv + pc * 16, w);
              if (dbias != nullptr) {   // column sums of the ROUNDED values: exactly what the weight-gradient GEMM reads
                float cs[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  cs[2 * i] = q_ok ? bf16_lo(w[i]) : 0.f;
                  cs[2 * i + 1] = q_ok ? bf16_hi(w[i]) : 0.f;
                }
                colsum16_atomic(cs, dbias + it.h * D + pc * 16, lane);
              }
            }
          }
          FB_TRACE(22, n);
          // ---- dK / dV of key tile t, complete after its last query tile ----
          if (j == it.n_qt - 1) {
            mbar_wait(&pipe[ACC_DONE], n & 1);
            tc_fence_after();
            {
              // one accumulator at a time: pulling both out before converting (earlier release) was measured slower, 710 vs 670 us
              const int key = t * 128 + row;
              const bool k_ok = key < it.nq;
              __nv_bfloat16* dst_row = out_head + static_cast<size_t>(key) * ld_dqkv;
#pragma unroll
              for (int mm = 0; mm < 2; ++mm) {   // dV first (the next tile's MMAs need its accumulator first), then dK (x scale)
                const int m = 1 - mm;
                uint32_t a[64];
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) tmem_ld_32x16(t_row + 320 + m * 64 + pc * 16, *reinterpret_cast<uint32_t(*)[16]>(&a[pc * 16]));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pipe[m == 1 ? DV_FREE : DK_FREE]);   // the accumulator is in registers
                if (m == 0) FB_TRACE(23, n);
                const float mul = m == 0 ? p.scale : 1.f;
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) {
                  uint32_t w[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(__uint_as_float(a[pc * 16 + 2 * i]) * mul, __uint_as_float(a[pc * 16 + 2 * i + 1]) * mul);
                  const int col = (m + 1) * H + pc * 16;
                  if (k_ok) st_global_b256(dst_row + col, w);
                  if (dbias != nullptr) {
                    float cs[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      cs[2 * i] = k_ok ? bf16_lo(w[i]) : 0.f;
                      cs[2 * i + 1] = k_ok ? bf16_hi(w[i]) : 0.f;
                    }
                    colsum16_atomic(cs, dbias + col + it.h * D, lane);
                  }
                }
              }
            }
            FB_TRACE(24, n);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace plb

using namespace plb;

extern "C" int plb_attn_bwd_fused_workspace_bytes(int B, int heads) {
  const long long items = static_cast<long long>(B) * heads;
  const long long ctas = items < sm_count() ? items : sm_count();
  return static_cast<int>(ctas * SCRATCH_FLOATS * static_cast<long long>(sizeof(float)));
}

static int attn_bwd_fused_launch(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                                 const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                                 const void* ctx, const void* dctx, int ld_ctx, const float* lse2, float* delta,
                                 void* dqkv, int ld_dqkv, float* dbias_qkv_accum, void* workspace,
                                 long long workspace_bytes, const DropoutParams& drop, plb_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLB_REQUIRE(qkv && seq_start && q_len && kv_len && dctx && lse2 && delta && dqkv && workspace, PLB_ERR_ARG,
              "attn_bwd_fused: null pointer");
  PLB_REQUIRE(head_dim == D, PLB_ERR_UNSUPPORTED, "attn_bwd_fused: head_dim %d not supported (only 64)", head_dim);
  PLB_REQUIRE(B > 0 && heads > 0 && total_rows > 0, PLB_ERR_ARG, "attn_bwd_fused: bad sizes");
  PLB_REQUIRE(max_q_len > 0 && max_q_len <= MAX_TILES * 128, PLB_ERR_UNSUPPORTED,
              "attn_bwd_fused: sequences longer than %d rows are not supported (got %d)", MAX_TILES * 128, max_q_len);
  PLB_REQUIRE(ld_qkv % 8 == 0 && ld_ctx % 8 == 0 && ld_dqkv % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(ctx) % 16) == 0 && (reinterpret_cast<uintptr_t>(dctx) % 16) == 0 &&  /* NULL ctx passes */
                  (reinterpret_cast<uintptr_t>(dqkv) % 32) == 0 && ld_dqkv % 16 == 0 && (reinterpret_cast<uintptr_t>(workspace) % 16) == 0,
              PLB_ERR_ARG, "attn_bwd_fused: operands must be 16-byte aligned with ld %% 8 == 0 (dqkv: 32 bytes, ld %% 16 == 0)");
  PLB_REQUIRE(workspace_bytes >= plb_attn_bwd_fused_workspace_bytes(B, heads), PLB_ERR_ARG,
              "attn_bwd_fused: workspace of %lld bytes is too small (plb_attn_bwd_fused_workspace_bytes: %d)", workspace_bytes,
              plb_attn_bwd_fused_workspace_bytes(B, heads));
  PLB_REQUIRE(drop.thresh16 == 0 || static_cast<long long>(B) * heads <= 32768, PLB_ERR_UNSUPPORTED,
              "attn_bwd_fused: dropout counters cover at most 32768 (sequence, head) items");
  CUtensorMap tm_qkv, tm_dctx;
  if (!tensor_map_bf16(&tm_qkv, qkv, total_rows, 3 * heads * D, ld_qkv, 128, 64) ||
      !tensor_map_bf16(&tm_dctx, dctx, total_rows, heads * D, ld_ctx, 128, 64))
    return PLB_ERR_CUDA;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FbSmem::kAlloc);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FbSmem::kAlloc);
    PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "attn_bwd_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  if (ctx != nullptr) {  // otherwise delta was produced by the O-projection DGRAD (PLB_EPI_ROWDOT)
    const int rc = launch_attn_delta(dctx, ctx, ld_ctx, total_rows, heads, delta, stream);
    if (rc) return rc;
  }
  FbParams p{seq_start, q_len, kv_len, heads, total_rows, 1.4426950408889634f / sqrtf((float)head_dim),
             1.f / sqrtf((float)head_dim)};
  const int n_items = B * heads, sms = sm_count();
  const int grid = n_items < sms ? n_items : sms;
  cudaError_t le = launch_kernel(drop.thresh16 != 0 ? attn_bwd_fused_kernel<true> : attn_bwd_fused_kernel<false>, dim3(grid),
                                 dim3(FB_THREADS), FbSmem::kAlloc, stream, tm_qkv, tm_dctx, p, n_items, lse2, delta,
                                 static_cast<__nv_bfloat16*>(dqkv), ld_dqkv, dbias_qkv_accum, static_cast<float*>(workspace), drop);
  PLB_REQUIRE(le == cudaSuccess, PLB_ERR_CUDA, "plb_attn_bwd_fused: launch failed: %s", cudaGetErrorString(le));
  return check_launch("plb_attn_bwd_fused");
}

extern "C" int plb_attn_bwd_fused(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                                  const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                                  const void* ctx, const void* dctx, int ld_ctx, const float* lse2, float* delta,
                                  void* dqkv, int ld_dqkv, float* dbias_qkv_accum, void* workspace,
                                  long long workspace_bytes, plb_stream_t stream) {
  return attn_bwd_fused_launch(qkv, ld_qkv, seq_start, q_len, kv_len, B, heads, head_dim, max_q_len, total_rows, ctx, dctx, ld_ctx,
                               lse2, delta, dqkv, ld_dqkv, dbias_qkv_accum, workspace, workspace_bytes, make_dropout(0.f, 0, 0),
                               stream);
}

extern "C" int plb_attn_bwd_fused_dropout(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                                          const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                                          const void* ctx, const void* dctx, int ld_ctx, const float* lse2, float* delta,
                                          void* dqkv, int ld_dqkv, float* dbias_qkv_accum, void* workspace,
                                          long long workspace_bytes, float p_drop, unsigned long long seed,
                                          unsigned int rng_stream, plb_stream_t stream) {
  PLB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, PLB_ERR_ARG, "attn_bwd_fused_dropout: p = %f outside [0, 1)", p_drop);
  return attn_bwd_fused_launch(qkv, ld_qkv, seq_start, q_len, kv_len, B, heads, head_dim, max_q_len, total_rows, ctx, dctx, ld_ctx,
                               lse2, delta, dqkv, ld_dqkv, dbias_qkv_accum, workspace, workspace_bytes,
                               make_dropout(p_drop, seed, rng_stream), stream);
}

#ifdef PLB_ATTN_TRACE
// Trace build only: copies the record table (zero = unused slot) to the host and clears it.
extern "C" int plb_attn_fb_trace_dump(unsigned long long* host, int cap) {
  if (cap < static_cast<int>(plb::FB_TRACE_CAP)) return -1;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host, plb::g_fb_trace, sizeof(plb::g_fb_trace));
  void* dev = nullptr;
  cudaGetSymbolAddress(&dev, plb::g_fb_trace);
  cudaMemset(dev, 0, sizeof(plb::g_fb_trace));
  cudaDeviceSynchronize();
  return static_cast<int>(plb::FB_TRACE_CAP);
}
#endif
