// plbert_b200 — fused padded/varlen multi-head attention, forward and backward, head_dim 64.
//
// Replaces HF modeling_albert.py:184-197 (SDPA / eager attention with the [B,1,L,L] key-padding mask built by
// masking_utils.py:495-520).  The mask is never materialised: each sequence carries (first row, number of query
// rows, number of valid keys) and keys >= kv_len are excluded inside the kernel; S and P live in registers only.
// Saved for backward: LSE in the log2 domain, one float per (head, row).
//
// Round-1 implementation: flash-style kernels on the warp-level tensor-core path (mma.sync m16n8k16 bf16,
// ldmatrix from XOR-swizzled shared memory, cp.async double buffering).  Backward is two kernels so that no
// atomics are needed: one owns query tiles (dQ, and the row term delta = rowsum(dO o O)), one owns key tiles
// (dK, dV).  Formulas: SURVEY.md appendix A.
#include "common.cuh"

namespace plb {
namespace {

constexpr int D = 64;          // head dim
constexpr int TILE = 64;       // rows per tile (queries or keys)
constexpr int NTHREADS = 128;  // 4 warps x 16 rows
constexpr int TILE_BYTES = TILE * D * 2;

struct AttnParams {
  const __nv_bfloat16* qkv;  // [T, 3*H]: q | k | v, each H = heads*64 wide
  int ld_qkv;
  const int32_t* seq_start;  // [B] first row of the sequence
  const int32_t* q_len;      // [B] rows to compute (padded layout: L; packed layout: length)
  const int32_t* kv_len;     // [B] valid keys (length)
  int heads;
  int total_rows;            // T (LSE / delta row stride)
  float scale_log2;          // log2(e) / sqrt(d)
  float scale;               // 1 / sqrt(d)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}
// 64 x 64 bf16 tile: rows [row0, row0+64) of a matrix with leading dim ld, 64 columns from col0.
// Rows >= row_end are zero-filled (their addresses are clamped, never dereferenced out of range).
__device__ __forceinline__ void load_tile(uint32_t smem, const __nv_bfloat16* base, int ld, int row0, int row_end,
                                          int col0) {
#pragma unroll
  for (int i = 0; i < (TILE * 8) / NTHREADS; ++i) {
    const int idx = threadIdx.x + i * NTHREADS;
    const int r = idx >> 3, c = idx & 7;
    const bool ok = (row0 + r) < row_end;
    const __nv_bfloat16* src = base + static_cast<size_t>(ok ? row0 + r : row0) * ld + col0 + c * 8;
    cp_async16(tile_addr(smem, r, c), src, ok);
  }
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A-operand fragments of a [16 rows x 64] slab of a swizzled tile: 4 k-steps x 4 registers
__device__ __forceinline__ void load_a_frags(uint32_t tile, int row0, uint32_t (&f)[4][4]) {
  const int lane = threadIdx.x & 31;
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldsm_x4(tile_addr(tile, r, kk * 2 + (lane >> 4)), f[kk][0], f[kk][1], f[kk][2], f[kk][3]);
}
// acc[16 x 64] += A[16 x 64(k)] * Bt, with B given as a [64(n) x 64(k)] row-major tile  (acc = A * tile^T)
__device__ __forceinline__ void mma_a_bT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(tile_addr(tile, np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
      mma16816(acc[2 * np], a[kk], b0, b1);
      mma16816(acc[2 * np + 1], a[kk], b2, b3);
    }
  }
}
// acc[16 x 64(n)] += P[16 x 64(k)] * tile, with tile [64(k) x 64(n)] row-major and P given as fp32 accumulator
// fragments (converted to bf16 A fragments on the fly)
__device__ __forceinline__ void mma_p_b(float (&acc)[8][4], const float (&p)[8][4], uint32_t tile) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile_addr(tile, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma16816(acc[2 * dp], a, b0, b1);
      mma16816(acc[2 * dp + 1], a, b2, b3);
    }
  }
}
__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
}
// write a warp's [16 x 64] fp32 accumulator slab as bf16 into a swizzled tile (rows row0..row0+15)
__device__ __forceinline__ void store_acc_tile(uint8_t* tile, int row0, const float (&acc)[8][4], float mul) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r0 = row0 + g, r1 = r0 + 8;
    *reinterpret_cast<uint32_t*>(tile + r0 * 128 + ((j ^ (r0 & 7)) << 4) + t * 4) = pack_bf16x2(acc[j][0] * mul, acc[j][1] * mul);
    *reinterpret_cast<uint32_t*>(tile + r1 * 128 + ((j ^ (r1 & 7)) << 4) + t * 4) = pack_bf16x2(acc[j][2] * mul, acc[j][3] * mul);
  }
}
// coalesced copy of a swizzled 64 x 64 tile to global rows [row0, row_end), 64 columns from col0;
// rows in [row_end, zero_end) are written as zeros
__device__ __forceinline__ void store_tile(const uint8_t* tile, __nv_bfloat16* base, int ld, int row0, int row_end,
                                           int zero_end, int col0) {
#pragma unroll
  for (int i = 0; i < (TILE * 8) / NTHREADS; ++i) {
    const int idx = threadIdx.x + i * NTHREADS;
    const int r = idx >> 3, c = idx & 7;
    const int row = row0 + r;
    if (row < zero_end) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row < row_end) v = *reinterpret_cast<const uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
      *reinterpret_cast<uint4*>(base + static_cast<size_t>(row) * ld + col0 + c * 8) = v;
    }
  }
}

// ================================================================================================
// forward
// ================================================================================================
__global__ void __launch_bounds__(NTHREADS) attn_fwd_kernel(AttnParams p, __nv_bfloat16* ctx, int ld_ctx, float* lse2) {
  __shared__ __align__(1024) uint8_t smem[TILE_BYTES * 5];  // Q | K0 V0 | K1 V1
  const int b = blockIdx.z, h = blockIdx.y;
  const int row_base = p.seq_start[b], nq = p.q_len[b], nk = p.kv_len[b];
  const int q0 = blockIdx.x * TILE;
  if (q0 >= nq) return;
  const int H = p.heads * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sKV = sQ + TILE_BYTES;
  const __nv_bfloat16* qkv = p.qkv;

  load_tile(sQ, qkv, p.ld_qkv, row_base + q0, row_base + nq, h * D);
  const int n_kt = (nk + TILE - 1) / TILE;
  load_tile(sKV, qkv, p.ld_qkv, row_base, row_base + nk, H + h * D);
  load_tile(sKV + TILE_BYTES, qkv, p.ld_qkv, row_base, row_base + nk, 2 * H + h * D);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
  zero_acc(o);
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int kt = 0; kt < n_kt; ++kt) {
    const uint32_t sK = sKV + (kt & 1) * 2 * TILE_BYTES, sV = sK + TILE_BYTES;
    if (kt + 1 < n_kt) {
      const uint32_t nK = sKV + ((kt + 1) & 1) * 2 * TILE_BYTES;
      load_tile(nK, qkv, p.ld_qkv, row_base + (kt + 1) * TILE, row_base + nk, H + h * D);
      load_tile(nK + TILE_BYTES, qkv, p.ld_qkv, row_base + (kt + 1) * TILE, row_base + nk, 2 * H + h * D);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) load_a_frags(sQ, warp * 16, qf);

    float s[8][4];
    zero_acc(s);
    mma_a_bT(s, qf, sK);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kt * TILE + j * 8 + 2 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool valid = (key + (i & 1)) < nk;
        s[j][i] = valid ? s[j][i] * p.scale_log2 : -INFINITY;
        mx[i >> 1] = fmaxf(mx[i >> 1], s[j][i]);
      }
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);  // finite: every processed tile has >= 1 valid key
      alpha[r] = exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[j][i] = exp2f(s[j][i] - m_run[i >> 1]);
        rs[i >> 1] += s[j][i];
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * alpha[r] + rs[r];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
      o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
    }
    mma_p_b(o, s, sV);
    __syncthreads();  // everyone is done with this K/V buffer before it is refilled
  }

  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    o[j][0] *= inv0; o[j][1] *= inv0;
    o[j][2] *= inv1; o[j][3] *= inv1;
  }
  // the Q tile is dead: reuse it as the output staging buffer
  store_acc_tile(smem, warp * 16, o, 1.f);
  if (t == 0) {
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    float* out = lse2 + static_cast<size_t>(h) * p.total_rows + row_base;
    if (r0 < nq) out[r0] = m_run[0] + log2f(l_run[0]);
    if (r1 < nq) out[r1] = m_run[1] + log2f(l_run[1]);
  }
  __syncthreads();
  store_tile(smem, ctx, ld_ctx, row_base + q0, row_base + nq, row_base + nq, h * D);
}

// ================================================================================================
// backward, query-tile owner: delta = rowsum(dO o O) and dQ = scale * dS K
// ================================================================================================
__global__ void __launch_bounds__(NTHREADS)
attn_bwd_dq_kernel(AttnParams p, const __nv_bfloat16* dctx, const __nv_bfloat16* ctx, int ld_ctx, const float* lse2,
                   float* delta, __nv_bfloat16* dqkv, int ld_dqkv) {
  extern __shared__ __align__(1024) uint8_t smem[];  // Q | dO | K0 V0 | K1 V1 | delta   (O staged in K0 first)
  float* s_delta = reinterpret_cast<float*>(smem + 6 * TILE_BYTES);
  const int b = blockIdx.z, h = blockIdx.y;
  const int row_base = p.seq_start[b], nq = p.q_len[b], nk = p.kv_len[b];
  const int q0 = blockIdx.x * TILE;
  if (q0 >= nq) return;
  const int H = p.heads * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint32_t sQ = smem_u32(smem), sdO = sQ + TILE_BYTES, sKV = sdO + TILE_BYTES;

  load_tile(sQ, p.qkv, p.ld_qkv, row_base + q0, row_base + nq, h * D);
  load_tile(sdO, dctx, ld_ctx, row_base + q0, row_base + nq, h * D);
  load_tile(sKV, ctx, ld_ctx, row_base + q0, row_base + nq, h * D);  // O, only for delta
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  {
    // delta: two threads per row, 32 elements each
    const int r = threadIdx.x >> 1, half = threadIdx.x & 1;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int chunk = half * 4 + c;
      const uint4 a = *reinterpret_cast<const uint4*>(smem + TILE_BYTES + r * 128 + ((chunk ^ (r & 7)) << 4));
      const uint4 o = *reinterpret_cast<const uint4*>(smem + 2 * TILE_BYTES + r * 128 + ((chunk ^ (r & 7)) << 4));
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) acc += bf16_lo(aw[i]) * bf16_lo(ow[i]) + bf16_hi(aw[i]) * bf16_hi(ow[i]);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0) {
      s_delta[r] = acc;
      if (q0 + r < nq) delta[static_cast<size_t>(h) * p.total_rows + row_base + q0 + r] = acc;
    }
  }
  uint32_t qf[4][4], dof[4][4];
  load_a_frags(sQ, warp * 16, qf);
  load_a_frags(sdO, warp * 16, dof);
  __syncthreads();  // O tile consumed -> K0 may be overwritten

  const int n_kt = (nk + TILE - 1) / TILE;
  load_tile(sKV, p.qkv, p.ld_qkv, row_base, row_base + nk, H + h * D);
  load_tile(sKV + TILE_BYTES, p.qkv, p.ld_qkv, row_base, row_base + nk, 2 * H + h * D);
  cp_async_commit();

  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const float* lse_row = lse2 + static_cast<size_t>(h) * p.total_rows + row_base;
  const float lse[2] = {r0 < nq ? lse_row[r0] : 0.f, r1 < nq ? lse_row[r1] : 0.f};
  const float dl[2] = {s_delta[warp * 16 + g], s_delta[warp * 16 + g + 8]};
  float dq[8][4];
  zero_acc(dq);

  for (int kt = 0; kt < n_kt; ++kt) {
    const uint32_t sK = sKV + (kt & 1) * 2 * TILE_BYTES, sV = sK + TILE_BYTES;
    if (kt + 1 < n_kt) {
      const uint32_t nK = sKV + ((kt + 1) & 1) * 2 * TILE_BYTES;
      load_tile(nK, p.qkv, p.ld_qkv, row_base + (kt + 1) * TILE, row_base + nk, H + h * D);
      load_tile(nK + TILE_BYTES, p.qkv, p.ld_qkv, row_base + (kt + 1) * TILE, row_base + nk, 2 * H + h * D);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_a_bT(s, qf, sK);
    mma_a_bT(dp, dof, sV);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kt * TILE + j * 8 + 2 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool valid = (key + (i & 1)) < nk;
        const float pr = valid ? exp2f(s[j][i] * p.scale_log2 - lse[i >> 1]) : 0.f;
        s[j][i] = pr * (dp[j][i] - dl[i >> 1]);  // dS
      }
    }
    mma_p_b(dq, s, sK);
    __syncthreads();
  }
  store_acc_tile(smem, warp * 16, dq, p.scale);  // Q tile is dead (fragments are in registers)
  __syncthreads();
  store_tile(smem, dqkv, ld_dqkv, row_base + q0, row_base + nq, row_base + nq, h * D);
}

// ================================================================================================
// backward, key-tile owner: dV = P^T dO, dK = scale * dS^T Q
// ================================================================================================
__global__ void __launch_bounds__(NTHREADS)
attn_bwd_dkv_kernel(AttnParams p, const __nv_bfloat16* dctx, int ld_ctx, const float* lse2, const float* delta,
                    __nv_bfloat16* dqkv, int ld_dqkv) {
  extern __shared__ __align__(1024) uint8_t smem[];  // K | V | Q0 dO0 | Q1 dO1 | lse[2][64] | delta[2][64]
  float (*s_lse)[TILE] = reinterpret_cast<float (*)[TILE]>(smem + 6 * TILE_BYTES);
  float (*s_dl)[TILE] = reinterpret_cast<float (*)[TILE]>(smem + 6 * TILE_BYTES + 2 * TILE * 4);
  const int b = blockIdx.z, h = blockIdx.y;
  const int row_base = p.seq_start[b], nq = p.q_len[b], nk = p.kv_len[b];
  const int k0 = blockIdx.x * TILE;
  if (k0 >= nq) return;  // rows beyond the allocated sequence do not exist
  const int H = p.heads * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  if (k0 >= nk) {
    // padded key rows inside the allocated sequence: their gradients are exactly zero
    store_tile(smem, dqkv, ld_dqkv, row_base + k0, row_base + k0, row_base + nq, H + h * D);
    store_tile(smem, dqkv, ld_dqkv, row_base + k0, row_base + k0, row_base + nq, 2 * H + h * D);
    return;
  }
  const uint32_t sK = smem_u32(smem), sV = sK + TILE_BYTES, sQdO = sV + TILE_BYTES;
  const float* lse_row = lse2 + static_cast<size_t>(h) * p.total_rows + row_base;
  const float* dl_row = delta + static_cast<size_t>(h) * p.total_rows + row_base;

  auto load_q_tile = [&](int qt, int buf) {
    const uint32_t dst = sQdO + buf * 2 * TILE_BYTES;
    load_tile(dst, p.qkv, p.ld_qkv, row_base + qt * TILE, row_base + nq, h * D);
    load_tile(dst + TILE_BYTES, dctx, ld_ctx, row_base + qt * TILE, row_base + nq, h * D);
    if (threadIdx.x < TILE) {
      const int r = qt * TILE + threadIdx.x;
      s_lse[buf][threadIdx.x] = r < nq ? lse_row[r] : INFINITY;  // +inf => P = 0 for rows that do not exist
      s_dl[buf][threadIdx.x] = r < nq ? dl_row[r] : 0.f;
    }
  };
  load_tile(sK, p.qkv, p.ld_qkv, row_base + k0, row_base + nk, H + h * D);
  load_tile(sV, p.qkv, p.ld_qkv, row_base + k0, row_base + nk, 2 * H + h * D);
  load_q_tile(0, 0);
  cp_async_commit();

  uint32_t kf[4][4], vf[4][4];
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  const int n_qt = (nq + TILE - 1) / TILE;
  const int key0 = k0 + warp * 16 + g, key1 = key0 + 8;

  for (int qt = 0; qt < n_qt; ++qt) {
    const int buf = qt & 1;
    const uint32_t sQ = sQdO + buf * 2 * TILE_BYTES, sdO = sQ + TILE_BYTES;
    if (qt + 1 < n_qt) {
      load_q_tile(qt + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (qt == 0) {
      load_a_frags(sK, warp * 16, kf);
      load_a_frags(sV, warp * 16, vf);
    }
    float st[8][4], dpt[8][4];  // S^T and dP^T: rows = keys of this warp, columns = queries of the tile
    zero_acc(st);
    zero_acc(dpt);
    mma_a_bT(st, kf, sQ);
    mma_a_bT(dpt, vf, sdO);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qc = j * 8 + 2 * t + (i & 1);
        const bool valid = ((i >> 1) ? key1 : key0) < nk;
        const float pr = valid ? exp2f(st[j][i] * p.scale_log2 - s_lse[buf][qc]) : 0.f;
        st[j][i] = pr;                                  // P^T
        dpt[j][i] = pr * (dpt[j][i] - s_dl[buf][qc]);   // dS^T
      }
    }
    mma_p_b(dv, st, sdO);
    mma_p_b(dk, dpt, sQ);
    __syncthreads();
  }
  // K and V tiles are dead (fragments in registers): stage dK in the K tile, dV in the V tile
  store_acc_tile(smem, warp * 16, dk, p.scale);
  store_acc_tile(smem + TILE_BYTES, warp * 16, dv, 1.f);
  __syncthreads();
  store_tile(smem, dqkv, ld_dqkv, row_base + k0, row_base + nk, row_base + nq, H + h * D);
  store_tile(smem + TILE_BYTES, dqkv, ld_dqkv, row_base + k0, row_base + nk, row_base + nq, 2 * H + h * D);
}

int check_args(const void* qkv, int ld_qkv, const int32_t* a, const int32_t* b, const int32_t* c, int B, int heads,
               int head_dim, int max_q_len, int total_rows) {
  PLB_REQUIRE(qkv && a && b && c, PLB_ERR_ARG, "attn: null pointer");
  PLB_REQUIRE(head_dim == D, PLB_ERR_UNSUPPORTED, "attn: head_dim %d not supported (only 64)", head_dim);
  PLB_REQUIRE(B > 0 && heads > 0 && max_q_len > 0 && total_rows > 0, PLB_ERR_ARG, "attn: bad sizes");
  PLB_REQUIRE(heads <= 65535 && B <= 65535, PLB_ERR_ARG, "attn: grid too large");
  PLB_REQUIRE(ld_qkv % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) % 16) == 0, PLB_ERR_ARG,
              "attn: qkv must be 16-byte aligned with ld %% 8 == 0");
  return PLB_OK;
}

}  // namespace
}  // namespace plb

using namespace plb;

extern "C" int plb_attn_fwd(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                            const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                            void* ctx, int ld_ctx, float* lse2, plb_stream_t stream) {
  int rc = check_args(qkv, ld_qkv, seq_start, q_len, kv_len, B, heads, head_dim, max_q_len, total_rows);
  if (rc) return rc;
  PLB_REQUIRE(ctx && lse2 && ld_ctx % 8 == 0 && (reinterpret_cast<uintptr_t>(ctx) % 16) == 0, PLB_ERR_ARG,
              "attn_fwd: bad ctx/lse");
  AttnParams p{static_cast<const __nv_bfloat16*>(qkv), ld_qkv, seq_start, q_len, kv_len, heads, total_rows,
               1.4426950408889634f / sqrtf((float)head_dim), 1.f / sqrtf((float)head_dim)};
  dim3 grid((max_q_len + TILE - 1) / TILE, heads, B);
  attn_fwd_kernel<<<grid, NTHREADS, 0, static_cast<cudaStream_t>(stream)>>>(p, static_cast<__nv_bfloat16*>(ctx), ld_ctx,
                                                                            lse2);
  return check_launch("plb_attn_fwd");
}

extern "C" int plb_attn_bwd(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                            const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                            const void* ctx, const void* dctx, int ld_ctx, const float* lse2, float* delta,
                            void* dqkv, int ld_dqkv, plb_stream_t stream) {
  int rc = check_args(qkv, ld_qkv, seq_start, q_len, kv_len, B, heads, head_dim, max_q_len, total_rows);
  if (rc) return rc;
  PLB_REQUIRE(ctx && dctx && lse2 && delta && dqkv, PLB_ERR_ARG, "attn_bwd: null pointer");
  PLB_REQUIRE(ld_ctx % 8 == 0 && ld_dqkv % 8 == 0 && (reinterpret_cast<uintptr_t>(ctx) % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(dctx) % 16) == 0 && (reinterpret_cast<uintptr_t>(dqkv) % 16) == 0,
              PLB_ERR_ARG, "attn_bwd: operands must be 16-byte aligned with ld %% 8 == 0");
  AttnParams p{static_cast<const __nv_bfloat16*>(qkv), ld_qkv, seq_start, q_len, kv_len, heads, total_rows,
               1.4426950408889634f / sqrtf((float)head_dim), 1.f / sqrtf((float)head_dim)};
  dim3 grid((max_q_len + TILE - 1) / TILE, heads, B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  constexpr int kBwdSmem = 6 * TILE_BYTES + 4 * TILE * 4;
  static bool configured = false;
  if (!configured) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    PLB_REQUIRE(e1 == cudaSuccess && e2 == cudaSuccess, PLB_ERR_CUDA, "attn_bwd: cudaFuncSetAttribute failed");
    configured = true;
  }
  attn_bwd_dq_kernel<<<grid, NTHREADS, kBwdSmem, s>>>(p, static_cast<const __nv_bfloat16*>(dctx),
                                               static_cast<const __nv_bfloat16*>(ctx), ld_ctx, lse2, delta,
                                               static_cast<__nv_bfloat16*>(dqkv), ld_dqkv);
  rc = check_launch("plb_attn_bwd(dq)");
  if (rc) return rc;
  attn_bwd_dkv_kernel<<<grid, NTHREADS, kBwdSmem, s>>>(p, static_cast<const __nv_bfloat16*>(dctx), ld_ctx, lse2, delta,
                                                static_cast<__nv_bfloat16*>(dqkv), ld_dqkv);
  return check_launch("plb_attn_bwd(dkv)");
}
