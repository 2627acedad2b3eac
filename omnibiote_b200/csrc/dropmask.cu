// Generator of the attention-dropout keep mask (see dropmask.cuh): the dropout of
// F.scaled_dot_product_attention(..., dropout_p) (training/model.py:118,134) as a bit matrix keep[B,H,T,ceil(T/32)].
//
// One thread draws whole 32-key words: twelve 32-bit hash words r_11..r_0 are read as the bit planes of thirty-two
// 12-bit uniforms u_e, and the comparison u_e < round(p * 2^12) is evaluated for all 32 keys at once with bitwise
// logic (a bit-serial magnitude comparator, MSB first): ~3.5 integer instructions per element at full issue rate,
// none of them inside the attention kernels. The drop probability is p rounded to 1/4096 (0.1 -> 0.10010); the
// 1/(1-p) scaling uses the caller's p. The stream is counter based: (seed, offset, flat word index).
#include "common.cuh"
#include "dropmask.cuh"
#include "ptx.cuh"

namespace obt {

constexpr int KEEP_PLANES = 12;  // bits of the per-element uniform

__device__ __forceinline__ uint32_t keep_word_draw(uint32_t k0, uint32_t k1, unsigned long long widx, uint32_t thr) {
  const uint32_t c0 = static_cast<uint32_t>(widx), c1 = static_cast<uint32_t>(widx >> 32);
  // two keyed 32-bit murmur3 finalisers of the word index
  uint32_t a = (c0 * 0x9E3779B1u) ^ k0 ^ (c1 * 0x85EBCA77u);
  uint32_t b = (c0 * 0xC2B2AE3Du) + k1 + (c1 * 0x27D4EB2Fu);
  a ^= a >> 16; a *= 0x85EBCA6Bu; a ^= a >> 13; a *= 0xC2B2AE35u; a ^= a >> 16;
  b ^= b >> 16; b *= 0x7FEB352Du; b ^= b >> 15; b *= 0x846CA68Bu; b ^= b >> 16;
  uint32_t lt = 0u, eq = 0xffffffffu;
#pragma unroll
  for (int k = KEEP_PLANES - 1; k >= 0; --k) {
    // bit plane k: a two-multiply mix of (a, b, k)
    uint32_t x = a * 0x2C1B3C6Du + static_cast<uint32_t>(k + 1) * 0x9E3779B9u;
    x ^= x >> 15;
    x = x * 0x297A2D39u + b;
    x ^= x >> 13;
    if ((thr >> k) & 1u) {  // uniform branch
      lt |= eq & ~x;
      eq &= x;
    } else {
      eq &= ~x;
    }
  }
  return ~lt;  // keep <=> u >= thr
}

// With an interval mask (row_lo / row_hi, int32 [B,T]) the words of a row that lie entirely outside its visible interval
// are not drawn but stored as all-ones: every consumer multiplies those bits with an exactly-zero probability, and with
// packed documents they are ~60 % of the matrix. Fully-masked rows (lo >= hi) attend to every key and are drawn
// completely. The 32 lanes of a warp take the SAME 4-word group (128 keys) of 32 consecutive query rows, whose
// intervals nearly coincide, so a skipped word is skipped by the whole warp (with a linear mapping the lanes of a row
// look at different key ranges and some lane always draws). Grid (warp groups of one (b, h), B * H): no 64-bit index
// arithmetic. The kernel is bound by the integer ALU pipe (LOP3 / SHF / SEL at half the issue rate,
// profiles/r02s_keepmask.details.txt); 4 words per lane instead of 8 halves the work of the last wave.
__global__ void attn_keep_mask_rows_kernel(uint32_t* __restrict__ keep, uint32_t thr, uint32_t k0, uint32_t k1,
                                           const int* __restrict__ row_lo, const int* __restrict__ row_hi, int H, int T,
                                           int nw) {
  const int groups = nw >> 2;
  const int wl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // warp within this (b, h)
  const int rblk = wl / groups;                                        // warp-uniform
  const int wr = (wl - rblk * groups) * 4;       // first word of this lane's group within its row
  const int i = rblk * 32 + (threadIdx.x & 31);  // query row
  if (i >= T) return;
  const int bh = blockIdx.y, b = bh / H;
  const long long w0 = (static_cast<long long>(bh) * T + i) * nw + wr;
  const int lo = row_lo[b * T + i], hi = row_hi[b * T + i];
  const bool dead = lo >= hi;
  uint32_t out[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int k = (wr + u) * 32;
    const bool draw = dead || (k < hi && k + 32 > lo);
    out[u] = draw ? keep_word_draw(k0, k1, static_cast<unsigned long long>(w0 + u), thr) : 0xffffffffu;
  }
  *reinterpret_cast<uint4*>(keep + w0) = make_uint4(out[0], out[1], out[2], out[3]);  // nw % 4 == 0: 16-byte aligned
}

// Linear mapping (no interval mask, or a row length that is not a multiple of 128 keys): grid.x covers the B*H*T*nw
// words, 8 consecutive words per thread.
__global__ void attn_keep_mask_kernel(uint32_t* __restrict__ keep, long long n_words, uint32_t thr, uint32_t k0,
                                      uint32_t k1, const int* __restrict__ row_lo, const int* __restrict__ row_hi,
                                      int H, int T, int nw) {
  uint32_t out[8];
  const long long w0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (w0 >= n_words) return;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    bool draw = true;
    if (row_lo != nullptr && w0 + u < n_words) {  // odd shapes only: a division per word is fine
      const long long row = (w0 + u) / nw;
      const int w = static_cast<int>((w0 + u) - row * nw);
      const long long b = row / (static_cast<long long>(H) * T);
      const int i = static_cast<int>(row % T);
      const int lo = row_lo[b * T + i], hi = row_hi[b * T + i];
      draw = (lo >= hi) || (w * 32 < hi && w * 32 + 32 > lo);
    }
    out[u] = draw ? keep_word_draw(k0, k1, static_cast<unsigned long long>(w0 + u), thr) : 0xffffffffu;
  }
  if (w0 + 8 <= n_words && (reinterpret_cast<uintptr_t>(keep + w0) & 31) == 0) {
    st_global_256(keep + w0, out);  // one full 32-byte sector per lane
  } else {
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (w0 + u < n_words) keep[w0 + u] = out[u];
  }
}

}  // namespace obt

using namespace obt;

extern "C" int obt_attn_keep_mask(unsigned int* keep, int B, int H, int T, float drop_p, unsigned long long seed,
                                  unsigned long long offset, const int* row_lo, const int* row_hi,
                                  cudaStream_t stream) {
  OBT_REQUIRE(keep != nullptr, "obt_attn_keep_mask: null pointer");
  OBT_REQUIRE(B > 0 && H > 0 && T > 0, "obt_attn_keep_mask: empty problem");
  OBT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "obt_attn_keep_mask: dropout p=%f", drop_p);
  const long long n_words = static_cast<long long>(B) * H * T * keep_words(T);
  uint32_t thr = static_cast<uint32_t>(drop_p * static_cast<float>(1 << KEEP_PLANES) + 0.5f);
  if (thr > (1u << KEEP_PLANES) - 1u) thr = (1u << KEEP_PLANES) - 1u;
  const unsigned long long key = seed ^ (offset * 0xD1B54A32D192ED03ull);
  const int threads = 256;
  const int nw = keep_words(T);
  const int* lo = (row_lo != nullptr && row_hi != nullptr) ? row_lo : nullptr;
  if (lo != nullptr && nw % 4 == 0 && T % 32 == 0 && static_cast<long long>(B) * H < 65536 &&
      (reinterpret_cast<uintptr_t>(keep) & 15) == 0) {
    // per (b, h): (T / 32) row blocks x (nw / 4) word groups, one warp each; 8 warps per block
    const int warps = (T / 32) * (nw / 4);
    attn_keep_mask_rows_kernel<<<dim3((warps + 7) / 8, B * H), threads, 0, stream>>>(
        keep, thr, static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), lo, row_hi, H, T, nw);
  } else {
    const long long per_block = static_cast<long long>(threads) * 8;
    attn_keep_mask_kernel<<<static_cast<unsigned>((n_words + per_block - 1) / per_block), threads, 0, stream>>>(
        keep, n_words, thr, static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), lo, row_hi, H, T, nw);
  }
  return check_launch("attn_keep_mask");
}
