// Shared pieces of the tcgen05 attention kernels (forward: attn_tc.cu, backward: attn_tc_bwd.cu), head_dim = 128.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace obt {

constexpr int ATT_D = 128;      // head dim
constexpr int ATT_BM = 128;     // query rows per tile
constexpr int ATT_BN = 128;     // keys per tile
constexpr uint32_t ATT_TILE_BYTES = 128 * 128 * 2;  // one [128 x 128] bf16 operand tile = two 16 KB swizzle sub-tiles
constexpr float LOG2E = 1.4426950408889634f;

struct AttnTcParams {
  int B, H, T;
  float scale;
  // mask: dense additive bias, or per-row visible key interval, or neither
  const __nv_bfloat16* mask;
  long long msb, msh, msq;
  const int* row_lo;
  const int* row_hi;
  // forward outputs / backward inputs
  __nv_bfloat16* y;
  long long ldy;
  float* lse;  // [B,H,T,2] (row max, log exp-sum) in natural-log units of the scaled+biased scores
  float drop_p;
  unsigned long long seed, offset;
  // backward
  const float* delta;  // [B,H,T] = rowsum(dO * O)
  __nv_bfloat16* dq;   // gradients are written into the fused dqkv buffer [M, 3C] (pitch ldd)
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  long long ldd;
};

// byte offset of the 16-byte chunk holding elements [c, c+8) of row r inside a [128 x 128] bf16 tile stored as two
// [128 x 64] K-major sub-tiles with the 128B swizzle (what TMA SWIZZLE_128B produces and UMMA descriptors expect)
__device__ __forceinline__ uint32_t sw128_chunk_off(int r, int c) {
  const int sub = c >> 6;
  const int chunk = (c & 63) >> 3;
  return static_cast<uint32_t>(sub * 16384 + r * 128 + ((chunk ^ (r & 7)) << 4));
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Issue the 8 UMMA (k = 16 each) of one 128x128x128 product.
//   a_addr: K-major A tile (two 64-wide sub-tiles); b_addr: B tile, K-major (kBMN = false) or MN-major (kBMN = true:
//   the same TMA-written bytes read as [K rows][128 B of N], second 64 columns of N 16 KB further).
template <bool kBMN>
__device__ __forceinline__ void issue_128x128x128(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, bool accumulate) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, kBMN);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint64_t a_desc = make_smem_desc_sw128(a_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024);
    const uint64_t b_desc = kBMN ? make_smem_desc_sw128(b_addr + kk * 2048, 16384, 1024)
                                 : make_smem_desc_sw128(b_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024);
    umma_bf16_ss<1>(d_tmem, a_desc, b_desc, idesc, (accumulate || kk > 0) ? 1u : 0u);
  }
}

// dropout keep * 1/(1-p) factors for 4 consecutive keys starting at flat element index e0 (e0 % 4 == 0)
__device__ __forceinline__ void keep4(const AttnTcParams& p, unsigned long long e0, float (&ks)[4]) {
  const uint4 rnd = rand4x32(p.seed, e0 >> 2, p.offset);
  const float s = 1.0f / (1.0f - p.drop_p);
  ks[0] = ((rnd.x >> 8) * (1.0f / 16777216.0f) >= p.drop_p) ? s : 0.f;
  ks[1] = ((rnd.y >> 8) * (1.0f / 16777216.0f) >= p.drop_p) ? s : 0.f;
  ks[2] = ((rnd.z >> 8) * (1.0f / 16777216.0f) >= p.drop_p) ? s : 0.f;
  ks[3] = ((rnd.w >> 8) * (1.0f / 16777216.0f) >= p.drop_p) ? s : 0.f;
}

// named barrier among the 256 compute threads (warps 2..9)
__device__ __forceinline__ void compute_bar_sync256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

constexpr int ATT_COMPUTE_WARPS = 8;
constexpr int ATT_THREADS_BWD = 64 + 32 * ATT_COMPUTE_WARPS;

}  // namespace obt
