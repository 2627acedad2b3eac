// Shared pieces of the tcgen05 attention kernels (forward: attn_tc.cu, backward: attn_tc_bwd.cu), head_dim = 128.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "dropmask.cuh"

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
  // optional tile metadata of the interval mask (obt_attn_tile_meta): saves every CTA its own scan of the intervals
  const int* qmeta;            // [B, ceil(T/128), 4] = {min lo, max hi, any fully-masked row, 0} per 128-query tile
  const unsigned int* kmeta;   // [B, ceil(T/128), 4] = relevance bits of the 64-query sub-tiles per 128-key tile
  // forward outputs / backward inputs
  __nv_bfloat16* y;
  long long ldy;
  float* lse;  // [B,H,T,2] (row max, log exp-sum) in natural-log units of the scaled+biased scores
  float drop_p;
  const uint32_t* keep;  // dropout keep bits [B,H,T,nw] (dropmask.cuh); required when drop_p > 0
  int nw;                // words per query row = ceil(T / 32)
  // backward
  const float* delta;  // [B,H,T] = rowsum(dO * O)
  __nv_bfloat16* dq;   // gradients are written into the fused dqkv buffer [M, 3C] (pitch ldd)
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  long long ldd;
  // optional dS hand-over (round 2): the dK/dV kernel also stores its dS^T tiles (bf16 [B*H][T keys][ds_pitch queries],
  // ds_pitch = T rounded up to 64) and a score-free dQ kernel (attn_tc_dq2_kernel) accumulates dQ = dS K from them
  __nv_bfloat16* ds_out;
  long long ds_pitch;
  // optional adjoint of the rotary embedding on dq / dk (fp32 tables [>= T, 64]; sin == nullptr: cosine scaling)
  const float* rope_cos;
  const float* rope_sin;
};

// adjoint rotary on 32 consecutive d-columns (16 pairs) of one gradient row, values already scaled (fp32); cs / sn:
// the 16 table entries of those pairs (4 x float4), already in registers
__device__ __forceinline__ void rope_adjoint32(float (&f)[32], const float4* cs, const float4* sn, bool has_sin) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float c[4] = {cs[g].x, cs[g].y, cs[g].z, cs[g].w};
    if (has_sin) {
      const float s4[4] = {sn[g].x, sn[g].y, sn[g].z, sn[g].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = f[g * 8 + 2 * j], b = f[g * 8 + 2 * j + 1];
        f[g * 8 + 2 * j] = a * c[j] + b * s4[j];
        f[g * 8 + 2 * j + 1] = b * c[j] - a * s4[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[g * 8 + 2 * j] *= c[j];
        f[g * 8 + 2 * j + 1] *= c[j];
      }
    }
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Measured and rejected (round 2, profiles/r02e_attn_probe_poly_{on,off}.txt): a quarter of the exponentials on the
// FMA / ALU pipes (Cody-Waite range reduction + degree-4 polynomial, relative error 5.6e-5) instead of the MUFU, the
// FlashAttention-4 trick: forward 130.7 -> 139.9 us, dQ 212.3 -> 216.0 us, dK/dV 259.2 -> 267.1 us. The ~9 extra issue
// slots per converted element cost more than the MUFU slot they free: none of the three kernels is MUFU-bound.

// D[128 x 64] = A[128 x 128(d)] * B[64 x 128(d)]^T, both K-major; a tile = two 64-column sub-tiles `a_sub` /
// `b_sub` bytes apart.
__device__ __forceinline__ void issue_scores_128x64(uint32_t d_tmem, uint32_t a_addr, uint32_t a_sub, uint32_t b_addr,
                                                    uint32_t b_sub) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint64_t a_desc = make_smem_desc_sw128(a_addr + (kk >> 2) * a_sub + (kk & 3) * 32, 0, 1024);
    const uint64_t b_desc = make_smem_desc_sw128(b_addr + (kk >> 2) * b_sub + (kk & 3) * 32, 0, 1024);
    umma_bf16_ss<1>(d_tmem, a_desc, b_desc, idesc, kk > 0 ? 1u : 0u);
  }
}

// D[128 x 128(d)] += A[128 x 64] * B[64 x 128(d)]: A = bf16 [128 x 64] held in TMEM (32 columns, two keys per 32-bit
// column, lane = row); B is read MN-major from a TMA-written [64 rows x 128] tile whose two 64-column halves are `b_lbo`
// bytes apart.
__device__ __forceinline__ void issue_pv_ts_128x128x64(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, uint32_t b_lbo,
                                                       bool accumulate) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, true);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint64_t b_desc = make_smem_desc_sw128(b_addr + kk * 2048, b_lbo, 1024);
    umma_bf16_ts(d_tmem, a_tmem + kk * 8, b_desc, idesc, (accumulate || kk > 0) ? 1u : 0u);
  }
}

// D[128 x 64] = A * B^T with A = bf16 [128 x 128(d)] held in TMEM (64 columns, two d per column, lane = row) and
// B = K-major [64 x 128(d)] smem tile (two 64-column sub-tiles `b_sub` bytes apart).
__device__ __forceinline__ void issue_scores_ts_128x64(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, uint32_t b_sub) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint64_t b_desc = make_smem_desc_sw128(b_addr + (kk >> 2) * b_sub + (kk & 3) * 32, 0, 1024);
    umma_bf16_ts(d_tmem, a_tmem + kk * 8, b_desc, idesc, kk > 0 ? 1u : 0u);
  }
}

// D[128 x 128(d)] += A[128 x 64] * B[64 x 128(d)] with A in TMEM as two 16-column pieces (reduction indices 0..31
// at a_lo, 32..63 at a_hi: each compute thread wrote its piece over the score columns it had just read) and B read
// MN-major from a TMA-written [64 rows x 128] tile whose two 64-column halves are `b_lbo` bytes apart.
__device__ __forceinline__ void issue_grad_ts_128x128x64(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_addr,
                                                         uint32_t b_lbo, bool accumulate) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, true);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint64_t b_desc = make_smem_desc_sw128(b_addr + kk * 2048, b_lbo, 1024);
    umma_bf16_ts(d_tmem, (kk < 2 ? a_lo : a_hi) + (kk & 1) * 8, b_desc, idesc, (accumulate || kk > 0) ? 1u : 0u);
  }
}

// bit k set <=> position base + k lies in [lo, hi), k = 0..31
__device__ __forceinline__ uint32_t interval_bits32(int lo, int hi, int base) {
  const int a = min(max(lo - base, 0), 32), b = min(max(hi - base, 0), 32);
  if (b <= a) return 0u;
  const uint32_t upto_b = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
  return upto_b & ~((1u << a) - 1u);  // a < 32 here
}

constexpr int ATT_COMPUTE_WARPS = 8;
// Backward kernels: 12 warps. Warpgroup 0 = {TMA producer, MMA issuer, 2 idle}, warpgroups 1-2 = the 8 compute warps.
// Three warps share each scheduler's 16 K registers (168 per thread at launch); with the roles aligned to warpgroups,
// warpgroup 0 returns registers (setmaxnreg.dec) and the compute warps take them (224 each): the 10-warp layout
// spilled in the compute loops and reloaded the spills on the critical path (profiles/r01_attn_v7_bwd.source.txt).
constexpr int ATT_BWD_THREADS = 384;
constexpr int ATT_BWD_FIRST_COMPUTE_WARP = 4;
// Measured and rejected (round 2, profiles/r02c_attn_bwd_w16.details.txt): 16 compute warps with 16 score columns each
// (four threads per TMEM lane; 96 registers at launch, setmaxnreg to 104 because it only redistributes the CTA's own
// 640 x 96 registers) raised the IPC (1.13 -> 1.38, 1.45 -> 1.94) but not the speed: dQ 211.8 -> 220.8 us, dK/dV
// 267.6 -> 308.5 us. The sub-tile loop is bound by the dependency chain dS(s) -> [dQ += dS K ; S(s+2), dP(s+2)] ->
// compute(s+2) through the two TMEM score buffers, and ~45 % of a CTA's life is fixed prologue / epilogue latency.

}  // namespace obt
