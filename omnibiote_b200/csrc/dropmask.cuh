// Attention-dropout keep mask shared by every attention kernel of the library.
//
// The Bernoulli(1-p) keep decisions of one (batch, head) are a bit matrix keep[b,h,i,kw] (uint32, kw = j / 32):
// ONE standalone kernel (dropmask.cu) draws it per layer and micro-batch; the forward, dQ and dK/dV kernels (and the
// generic CUDA-core kernels) only read bits. Before, each of the three tensor-core kernels re-hashed every score
// element (8-10 integer instructions per element, 47 % of the forward's issued instructions:
// profiles/r01_attn_v5_fwd.source.txt).
//
// Bit layout inside a word: key e (0..31) of the group sits at bit 8*(e&3) + 7 - (e>>2), i.e. after `w << (e>>2)` the
// keep bits of keys 4s..4s+3 are the sign bits of bytes 0..3. A PRMT in sign-replicate mode then expands two of them
// into the 0x0000/0xFFFF halves of an AND mask for a packed bf16x2 pair: 1.25 instructions per element.
#pragma once
#include <stdint.h>

namespace obt {

__host__ __device__ constexpr int keep_bit_pos(int e) { return 8 * (e & 3) + 7 - (e >> 2); }

// number of 32-key words per query row
__host__ __device__ constexpr int keep_words(int T) { return (T + 31) / 32; }

// AND mask for the packed bf16x2 pair of keys (e, e+1), e even, from the word pre-shifted by (e >> 2):
// byte (e&3) -> low half, byte (e&3)+1 -> high half, sign-replicated.
__device__ __forceinline__ uint32_t keep_pair_mask(uint32_t shifted, int e) {
  uint32_t m;
  const uint32_t sel = (e & 2) ? 0xBBAAu : 0x9988u;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(shifted), "r"(0u), "r"(sel));
  return m;
}

}  // namespace obt
