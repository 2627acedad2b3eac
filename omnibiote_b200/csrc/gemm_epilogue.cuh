// GEMM epilogue: TMEM -> registers -> (swizzled smem transpose) -> coalesced, fused global stores.
//
// tcgen05.ld hands each thread one accumulator ROW, so storing straight from those registers makes every warp-wide
// store touch 32 different 128-byte lines (the first ncu capture showed L1TEX at 83 % and the tensor pipe at 41 %).
// Instead each epilogue warp stages its 32 rows x 64 columns (bf16, 128 B per row) in a 4 KB XOR-swizzled smem
// buffer and reads it back with 8 lanes per row, so every global access of the fused epilogue (residual / U loads,
// D / U stores) is a full 128-byte line per 8 lanes. All epilogues start from rb(acc), which is exactly what the
// staged bf16 value is.
#pragma once
#include "ptx.cuh"

namespace obt {

enum : int {
  EPI_PLAIN = 0,     // D = rb(acc)
  EPI_RESID = 1,     // D = rb(float(aux_in) + float(rb(acc)))            (residual add / gradient accumulation)
  EPI_GELU = 2,      // aux_out = U = rb(acc); D = rb(gelu(U))             (model.py:23-25,163-165)
  EPI_GELU_BWD = 3,  // D = rb(float(rb(acc)) * gelu'(float(aux_in)))      (aux_in = U saved by EPI_GELU)
  EPI_PARTIAL = 4,   // split-K: fp32 partial tile -> workspace[split]
  EPI_RESID_DROPOUT = 5,  // D = rb(float(aux_in) + float(rb(rb(acc) * keep/(1-p))))   (resid_dropout, model.py:151,167)
};

struct GemmParams {
  int M, N, K;
  int num_m, num_n, splits, kb_per_split, num_kb;
  int epi, gelu_mode, vec_ok;
  __nv_bfloat16* D;
  long long ldd;
  const __nv_bfloat16* aux_in;
  long long ld_aux_in;
  __nv_bfloat16* aux_out;
  long long ld_aux_out;
  float* partial;
  float drop_p;
  unsigned long long seed, offset;
};

constexpr int GEMM_BK = 64;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BM_CTA = 128;
constexpr int GEMM_GROUP_M = 16;
constexpr uint32_t GEMM_STAGE_BYTES_PER_WARP = 32 * 128;  // epilogue staging: 32 rows x 128 B
constexpr int GEMM_EPI_WARPS = 8;                          // two per TMEM lane quadrant
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;     // warp 0 = TMA, warp 1 = MMA, warps 2.. = epilogue

// 1 + erf(t) and exp(-t^2) in ~14 instructions (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7, no cancellation on the
// negative side). libdevice erff costs ~40 instructions and made the GELU epilogues 5-8x longer than the main loop
// (profiles/r01_launches_v1.txt); the result is rounded to bf16 (2^-9 relative) right after.
__device__ __forceinline__ void one_plus_erf(float t, float& cdf2, float& e) {
  const float a = fabsf(t);
  const float k = __frcp_rn(fmaf(0.3275911f, a, 1.0f));
  e = __expf(-a * a);
  float poly = fmaf(1.061405429f, k, -1.453152027f);
  poly = fmaf(poly, k, 1.421413741f);
  poly = fmaf(poly, k, -0.284496736f);
  poly = fmaf(poly, k, 0.254829592f);
  const float pe = poly * k * e;        // = 1 - erf(|t|)
  cdf2 = (t >= 0.f) ? 2.0f - pe : pe;   // = 1 + erf(t)
}

__device__ __forceinline__ float gelu_ref(float x, int mode) {
  // reference: x * 0.5 * (1.0 + erf(x / 1.41421))  -- the constant is 1.41421, not sqrt(2) (model.py:25)
  if (mode == 0) {
    float cdf2, e;
    one_plus_erf(x * (1.0f / 1.41421f), cdf2, e);
    return x * 0.5f * cdf2;
  }
  // per-primitive bf16 rounding (un-fused TorchScript / CPU eager execution of the same expression)
  float a = rb(x * 0.5f);
  float b = rb(x / 1.41421f);
  float c = rb(erff(b));
  float d = rb(1.0f + c);
  return a * d;  // caller rounds
}

__device__ __forceinline__ float gelu_grad_ref(float x) {
  const float inv = 1.0f / 1.41421f;
  float cdf2, e;
  one_plus_erf(x * inv, cdf2, e);
  // d/dx [x * 0.5 * (1 + erf(x/c))] = 0.5 (1 + erf(x/c)) + x * 0.5 * 2/sqrt(pi) * exp(-(x/c)^2) / c
  return 0.5f * cdf2 + x * (0.5f * 1.1283791670955126f * inv) * e;
}

__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

__device__ __forceinline__ void load8(const __nv_bfloat16* src, bool vec, int nvalid, float (&f)[8]) {
  if (vec) {
    unpack8f(*reinterpret_cast<const uint4*>(src), f);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = (e < nvalid) ? __bfloat162float(src[e]) : 0.f;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* dst, bool vec, int nvalid, const float (&f)[8]) {
  if (vec) {
    *reinterpret_cast<uint4*>(dst) = pack8f(f);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (e < nvalid) dst[e] = __float2bfloat16_rn(f[e]);
  }
}

// One 8-element (16-byte) segment of an output row: v holds rb(acc) for columns [gcol, gcol+8) of row grow.
__device__ __forceinline__ void epilogue_segment(const GemmParams& p, float (&v)[8], long long grow, int gcol) {
  const int nvalid = p.N - gcol;  // > 0 by construction
  const bool vec = p.vec_ok && nvalid >= 8;
  if (p.epi == EPI_RESID || p.epi == EPI_GELU_BWD || p.epi == EPI_RESID_DROPOUT) {
    float a[8];
    load8(p.aux_in + grow * p.ld_aux_in + gcol, vec, nvalid, a);
    if (p.epi == EPI_RESID) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = a[e] + v[e];
    } else if (p.epi == EPI_GELU_BWD) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = v[e] * gelu_grad_ref(a[e]);
    } else {
      // resid dropout: one Philox call per 4 consecutive columns, keyed by the flat element index row*N + col
      const float scale = 1.0f / (1.0f - p.drop_p);
      const unsigned long long base = (static_cast<unsigned long long>(grow) * p.N + gcol) >> 2;
#pragma unroll
      for (int j4 = 0; j4 < 2; ++j4) {
        const uint4 rnd = rand4x32(p.seed, base + j4, p.offset);
        const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float u01 = (rr[e] >> 8) * (1.0f / 16777216.0f);
          const float d = (u01 >= p.drop_p) ? rb(v[4 * j4 + e] * scale) : 0.f;
          v[4 * j4 + e] = a[4 * j4 + e] + d;
        }
      }
    }
  } else if (p.epi == EPI_GELU) {
    store8(p.aux_out + grow * p.ld_aux_out + gcol, vec, nvalid, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = gelu_ref(v[e], p.gelu_mode);
  }
  store8(p.D + grow * p.ldd + gcol, vec, nvalid, v);
}

// Stage one thread-row of 8 x 16-byte chunks (chunk k of row `lane` goes to slot k ^ (lane & 7)).
__device__ __forceinline__ void stage_write_bf16(uint8_t* stage, int lane, const uint32_t (&r0)[32], const uint32_t (&r1)[32]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 w = make_uint4(pack_bf16x2(__uint_as_float(r0[8 * k + 0]), __uint_as_float(r0[8 * k + 1])),
                               pack_bf16x2(__uint_as_float(r0[8 * k + 2]), __uint_as_float(r0[8 * k + 3])),
                               pack_bf16x2(__uint_as_float(r0[8 * k + 4]), __uint_as_float(r0[8 * k + 5])),
                               pack_bf16x2(__uint_as_float(r0[8 * k + 6]), __uint_as_float(r0[8 * k + 7])));
    *reinterpret_cast<uint4*>(stage + lane * 128 + ((k ^ (lane & 7)) << 4)) = w;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 w = make_uint4(pack_bf16x2(__uint_as_float(r1[8 * k + 0]), __uint_as_float(r1[8 * k + 1])),
                               pack_bf16x2(__uint_as_float(r1[8 * k + 2]), __uint_as_float(r1[8 * k + 3])),
                               pack_bf16x2(__uint_as_float(r1[8 * k + 4]), __uint_as_float(r1[8 * k + 5])),
                               pack_bf16x2(__uint_as_float(r1[8 * k + 6]), __uint_as_float(r1[8 * k + 7])));
    *reinterpret_cast<uint4*>(stage + lane * 128 + (((k + 4) ^ (lane & 7)) << 4)) = w;
  }
}

__device__ __forceinline__ void stage_write_f32(uint8_t* stage, int lane, const uint32_t (&r)[32]) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    *reinterpret_cast<uint4*>(stage + lane * 128 + ((k ^ (lane & 7)) << 4)) =
        make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
}

// split-K partial: 32 fp32 columns starting at gcol0, rows row_base .. row_base+31 (already staged)
__device__ __forceinline__ void partial_readback(const GemmParams& p, const uint8_t* stage, int lane, long long row_base,
                                                 int gcol0, int split) {
  const int rsub = lane >> 3, seg = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rl = it * 4 + rsub;
    const long long grow = row_base + rl;
    const int gcol = gcol0 + seg * 4;
    const uint4 w = *reinterpret_cast<const uint4*>(stage + rl * 128 + ((seg ^ (rl & 7)) << 4));
    if (grow < p.M && gcol < p.N) {
      float* dst = p.partial + (static_cast<size_t>(split) * p.M + grow) * p.N + gcol;
      if ((p.N & 3) == 0) {
        *reinterpret_cast<uint4*>(dst) = w;
      } else {
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (gcol + e < p.N) dst[e] = __uint_as_float(ww[e]);
      }
    }
  }
}

// Whole 128 x 256 accumulator slice of one epilogue warp: rows row_base..row_base+31 (TMEM lanes of this warp's
// quadrant), columns n0..n0+255.
__device__ __forceinline__ void epilogue_warp_tile(const GemmParams& p, uint32_t taddr, uint8_t* stage, int lane,
                                                   long long row_base, int n0, int split, int c_begin, int c_end) {
  const int rsub = lane >> 3, seg = lane & 7;
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const int col_base = n0 + c * 64;
    if (col_base >= p.N) break;  // warp-uniform
    uint32_t r0[32], r1[32];
    __syncwarp();
    tmem_ld_32x32(taddr + c * 64, r0);
    tmem_ld_32x32(taddr + c * 64 + 32, r1);
    tmem_ld_wait();
    if (p.epi == EPI_PARTIAL) {
      stage_write_f32(stage, lane, r0);
      __syncwarp();
      partial_readback(p, stage, lane, row_base, col_base, split);
      __syncwarp();
      stage_write_f32(stage, lane, r1);
      __syncwarp();
      partial_readback(p, stage, lane, row_base, col_base + 32, split);
      __syncwarp();
      continue;
    }
    stage_write_bf16(stage, lane, r0, r1);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rl = it * 4 + rsub;
      const long long grow = row_base + rl;
      const int gcol = col_base + seg * 8;
      const uint4 w = *reinterpret_cast<const uint4*>(stage + rl * 128 + ((seg ^ (rl & 7)) << 4));
      if (grow < p.M && gcol < p.N) {
        float v[8];
        unpack8f(w, v);
        epilogue_segment(p, v, grow, gcol);
      }
    }
    __syncwarp();  // the staging buffer is reused by the next 64-column chunk
  }
}

}  // namespace obt
