// plbert_b200 — counter-based random bits for dropout (HF modeling_albert.py:109,194,201: nn.Dropout after the embedding
// LayerNorm, on the attention probabilities and on the attention output projection).  No state, no mask tensors: the
// backward pass regenerates the forward mask from (seed, stream, element counter).  32 bits per call, used as two 16-bit
// draws for a PAIR of adjacent elements; an element is kept iff its draw < round((1 - p) * 65536) and then scaled by
// 1 / (1 - p) like torch.  oracle/dropout_rng.py is the numpy replica the parity tests use.
#pragma once
#include <stdint.h>

namespace plb {

struct DropoutParams {
  uint32_t seed_lo, seed_hi;
  uint32_t stream;      // which mask: (optimizer step, layer application, site) folded by the host
  uint32_t thresh16;    // keep iff draw < thresh16; 0 disables dropout (the draw is never evaluated)
  float scale;          // 1 / (1 - p)
};

__host__ __device__ __forceinline__ uint32_t rng32(uint32_t seed_lo, uint32_t seed_hi, uint32_t stream, uint32_t ctr) {
  uint32_t x = ctr * 0x9E3779B1u + seed_lo;
  x ^= x >> 15;
  x *= 0x85EBCA6Bu;
  x ^= stream * 0xC2B2AE35u + seed_hi;
  x ^= x >> 13;
  x *= 0xC2B2AE3Du;
  x ^= x >> 16;
  return x;
}

// multipliers (0 or scale) of the two elements of pair `ctr`
__device__ __forceinline__ void dropout_pair(const DropoutParams& d, uint32_t ctr, float& m0, float& m1) {
  const uint32_t x = rng32(d.seed_lo, d.seed_hi, d.stream, ctr);
  m0 = (x & 0xffffu) < d.thresh16 ? d.scale : 0.f;
  m1 = (x >> 16) < d.thresh16 ? d.scale : 0.f;
}

inline DropoutParams make_dropout(float p, unsigned long long seed, unsigned int stream) {
  DropoutParams d{};
  d.seed_lo = static_cast<uint32_t>(seed);
  d.seed_hi = static_cast<uint32_t>(seed >> 32);
  d.stream = stream;
  if (p > 0.f) {
    const double keep = 1.0 - static_cast<double>(p);
    d.thresh16 = static_cast<uint32_t>(keep * 65536.0 + 0.5);
    d.scale = static_cast<float>(1.0 / keep);
  } else {
    d.thresh16 = 0;
    d.scale = 1.f;
  }
  return d;
}

// Attention probabilities: the pair (k, k + 1), k even, of query row q of item (sequence * heads + head)
__host__ __device__ __forceinline__ uint32_t attn_dropout_ctr(uint32_t item, uint32_t q, uint32_t k) {
  return (item * 512u + q) * 256u + (k >> 1);
}

}  // namespace plb
