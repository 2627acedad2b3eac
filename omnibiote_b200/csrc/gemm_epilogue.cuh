// GEMM epilogue: TMEM -> registers -> (swizzled smem transpose) -> coalesced, fused global stores.
//
// tcgen05.ld hands each thread one accumulator ROW, so storing straight from those registers makes every warp-wide
// store touch 32 different 128-byte lines (first ncu capture: L1TEX 83 %, tensor pipe 41 %, profiles/r01_gemm_v0_*).
// Instead each epilogue warp stages 32 rows x 64 columns (bf16, 128 B per row) in a 4 KB XOR-swizzled smem buffer
// and reads it back with 8 lanes per row, so every global access of the fused epilogue (residual / U loads, D / U
// stores) is a full 128-byte line per 8 lanes. All epilogues start from rb(acc), which is exactly the staged bf16.
//
// The epilogue kind is a template parameter of the inner loops (one switch per tile, none per element) and the aux
// loads of a 64-column chunk are issued back to back before any math: the second capture
// (profiles/r01_gemm_v2_gelu_epilogue.details.txt) showed the runtime-switched version issue-bound at ~56
// instructions per element and the residual variant latency-bound on dependent load->store pairs.
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
  EPI_GELU_EAGER = 6,     // internal: EPI_GELU with one bf16 rounding per primitive (gelu_mode = 1)
  EPI_ROPE = 7,           // D = rb(rotary(rb(acc))) on columns < rope_cols (q | k of the fused c_attn output):
                          // apply_rotary_emb (model.py:39-50,108) fused into the c_attn GEMM
  EPI_GELU_DG = 9,        // aux_out = rb(gelu'(rb(acc))); D = rb(gelu(rb(acc))): the forward stores the activation's
                          // DERIVATIVE instead of the pre-activation (one erf evaluation serves both), so that
  EPI_MUL = 10,           // D = rb(rb(acc) * aux_in) is all the backward's epilogue has to do (the GELU' epilogue
                          // was the slowest GEMM of the step: 291 us vs 185 us for the same FLOPs)
  EPI_DELTA = 11,         // D = rb(acc) and delta[b, h, t] = sum over the 128 columns of head h of D * aux_in: the
                          // attention backward's per-row dO . O (aux_in = the forward's y), emitted by the GEMM that
                          // PRODUCES dO = d_a Wo (model.py:151) instead of a separate pass over dO and O
  EPI_ROWMASK = 8,        // D = row_mask[row] ? rb(acc) : 0 with aux_in = uint8 [M]: the MLM head's logits of rows
                          // outside the loss mask are never read (their loss weight and gradient are exactly zero,
                          // train_encoder.py:301-305), so the zeros d loss / d logits needs there are stored right away
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
  uint32_t drop_thr;  // ceil(drop_p * 2^24): keep <=> (r >> 8) >= drop_thr
  unsigned long long seed, offset;
  // EPI_ROPE: fp32 tables [rope_T rows, rope_d / 2]; rope_sin == nullptr -> cosine scaling (real bf16 freqs_cis)
  const float* rope_cos;
  const float* rope_sin;
  int rope_T, rope_d, rope_cols;
  int rope_T_mask, rope_d_mask;  // T - 1 / d - 1 when they are powers of two (no integer division in the epilogue), else -1
  // EPI_DELTA: fp32 [B, H, T] with H = N / 128, rows = b * delta_T + t
  float* delta;
  int delta_T;
};

// offset of the 4 table entries of columns gcol..gcol+7 at row grow
__device__ __forceinline__ long long rope_table_off(const GemmParams& p, long long grow, int gcol) {
  const int t = p.rope_T_mask >= 0 ? static_cast<int>(grow) & p.rope_T_mask : static_cast<int>(grow % p.rope_T);
  const int c = p.rope_d_mask >= 0 ? gcol & p.rope_d_mask : gcol % p.rope_d;
  return static_cast<long long>(t) * (p.rope_d >> 1) + (c >> 1);
}

constexpr int GEMM_BK = 64;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BM_CTA = 128;
constexpr int GEMM_GROUP_M = 16;
constexpr uint32_t GEMM_STAGE_BYTES_PER_WARP = 32 * 128;  // epilogue staging: 32 rows x 128 B
constexpr int GEMM_EPI_WARPS = 8;                          // two per TMEM lane quadrant
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;     // warp 0 = TMA, warp 1 = MMA, warps 2.. = epilogue

// 1 + erf(t) and exp(-t^2) in ~15 instructions (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7, no cancellation on the
// negative side); the result is rounded to bf16 (2^-9 relative) right after. libdevice erff costs ~40.
__device__ __forceinline__ void one_plus_erf(float t, float& cdf2, float& e) {
  const float a = fabsf(t);
  float k;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(k) : "f"(fmaf(0.3275911f, a, 1.0f)));  // 1 MUFU, rel. error 2^-23
  e = fast_ex2(a * a * -1.4426950408889634f);
  float poly = fmaf(1.061405429f, k, -1.453152027f);
  poly = fmaf(poly, k, 1.421413741f);
  poly = fmaf(poly, k, -0.284496736f);
  poly = fmaf(poly, k, 0.254829592f);
  const float pe = poly * k * e;        // = 1 - erf(|t|)
  cdf2 = (t >= 0.f) ? 2.0f - pe : pe;   // = 1 + erf(t)
}

// reference: x * 0.5 * (1.0 + erf(x / 1.41421))  -- the constant is 1.41421, not sqrt(2) (model.py:25)
__device__ __forceinline__ float gelu_fused(float x) {
  float cdf2, e;
  one_plus_erf(x * (1.0f / 1.41421f), cdf2, e);
  return (0.5f * x) * cdf2;
}

// same expression with one bf16 rounding per primitive (un-fused TorchScript / CPU eager execution)
__device__ __noinline__ float gelu_eager(float x) {
  float a = rb(x * 0.5f);
  float b = rb(x / 1.41421f);
  float c = rb(erff(b));
  float d = rb(1.0f + c);
  return a * d;  // caller rounds
}

// gelu(x) and gelu'(x) from ONE erf / exp evaluation
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& dg) {
  const float inv = 1.0f / 1.41421f;
  float cdf2, e;
  one_plus_erf(x * inv, cdf2, e);
  g = (0.5f * x) * cdf2;
  dg = fmaf(x * (0.5f * 1.1283791670955126f * inv), e, 0.5f * cdf2);
}

// Two elements at a time with packed fp32x2 instructions: the polynomial, the squares and the products of the pair
// cost one instruction instead of two (the GELU forward epilogue is instruction-issue bound: c_fc ran at 225 us
// against 181 us for mlp.c_proj with the same FLOPs). Same formula and selection as one_plus_erf / gelu_and_grad.
__device__ __forceinline__ void gelu_and_grad_x2(float x0, float x1, float& g0, float& g1, float& d0, float& d1) {
  const float inv = 1.0f / 1.41421f;
  const uint64_t x = f2_pack(x0, x1);
  const uint64_t t = f2_mul(x, f2_splat(inv));
  float t0, t1;
  f2_unpack(t, t0, t1);
  const uint64_t a = f2_pack(fabsf(t0), fabsf(t1));
  float den0, den1;
  f2_unpack(f2_fma(f2_splat(0.3275911f), a, f2_splat(1.0f)), den0, den1);
  float k0, k1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(k0) : "f"(den0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(k1) : "f"(den1));
  float q0, q1;
  f2_unpack(f2_mul(f2_mul(a, a), f2_splat(-1.4426950408889634f)), q0, q1);
  const float e0 = fast_ex2(q0), e1 = fast_ex2(q1);
  const uint64_t k = f2_pack(k0, k1), e = f2_pack(e0, e1);
  uint64_t poly = f2_fma(f2_splat(1.061405429f), k, f2_splat(-1.453152027f));
  poly = f2_fma(poly, k, f2_splat(1.421413741f));
  poly = f2_fma(poly, k, f2_splat(-0.284496736f));
  poly = f2_fma(poly, k, f2_splat(0.254829592f));
  const uint64_t pe = f2_mul(f2_mul(poly, k), e);                      // 1 - erf(|t|)
  float pe0, pe1, cp0, cp1;
  f2_unpack(pe, pe0, pe1);
  f2_unpack(f2_fma(pe, f2_splat(-1.0f), f2_splat(2.0f)), cp0, cp1);    // 2 - pe
  const uint64_t cdf2 = f2_pack(t0 >= 0.f ? cp0 : pe0, t1 >= 0.f ? cp1 : pe1);  // 1 + erf(t), no cancellation
  f2_unpack(f2_mul(f2_mul(x, f2_splat(0.5f)), cdf2), g0, g1);
  f2_unpack(f2_fma(f2_mul(x, f2_splat(0.5f * 1.1283791670955126f * inv)), e, f2_mul(cdf2, f2_splat(0.5f))), d0, d1);
}

__device__ __forceinline__ float gelu_grad_ref(float x) {
  const float inv = 1.0f / 1.41421f;
  float cdf2, e;
  one_plus_erf(x * inv, cdf2, e);
  // d/dx [x * 0.5 * (1 + erf(x/c))] = 0.5 (1 + erf(x/c)) + x * 0.5 * 2/sqrt(pi) * exp(-(x/c)^2) / c
  return fmaf(x * (0.5f * 1.1283791670955126f * inv), e, 0.5f * cdf2);
}

__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// rotary of 4 adjacent (even, odd) pairs: cs/sn are the table entries of the pairs; sn == nullptr semantics are
// expressed by has_sin = false (cosine scaling). inverse = adjoint rotation (backward).
__device__ __forceinline__ void rope8(float (&v)[8], const float4& cs, const float4& sn, bool has_sin, bool inverse) {
  const float c[4] = {cs.x, cs.y, cs.z, cs.w};
  const float s4[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (has_sin) {
      const float sj = inverse ? -s4[j] : s4[j];
      const float a = v[2 * j], b = v[2 * j + 1];
      v[2 * j] = a * c[j] - b * sj;
      v[2 * j + 1] = a * sj + b * c[j];
    } else {
      v[2 * j] *= c[j];
      v[2 * j + 1] *= c[j];
    }
  }
}

template <int EPI>
struct EpiTraits {
  static constexpr bool kAuxIn =
      (EPI == EPI_RESID || EPI == EPI_GELU_BWD || EPI == EPI_RESID_DROPOUT || EPI == EPI_MUL || EPI == EPI_DELTA);
  static constexpr bool kAuxOut = (EPI == EPI_GELU || EPI == EPI_GELU_EAGER || EPI == EPI_GELU_DG);
};

// v: rb(acc) for 8 consecutive columns; a: aux_in values (when the epilogue has one). Returns D in v, U in u.
// eidx: flat element index row * N + col of v[0] (the dropout stream is keyed by it).
template <int EPI>
__device__ __forceinline__ void epilogue_math(const GemmParams& p, float (&v)[8], const float (&a)[8], float (&u)[8],
                                              unsigned long long eidx) {
  if constexpr (EPI == EPI_RESID) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = a[e] + v[e];
  } else if constexpr (EPI == EPI_GELU_BWD) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = v[e] * gelu_grad_ref(a[e]);
  } else if constexpr (EPI == EPI_RESID_DROPOUT) {
    // one RNG call per 4 consecutive columns, keyed by the flat element index row*N + col
    const float scale = 1.0f / (1.0f - p.drop_p);
    const unsigned long long base = eidx >> 2;
#pragma unroll
    for (int j4 = 0; j4 < 2; ++j4) {
      const uint4 rnd = rand4x32(p.seed, base + j4, p.offset);
      const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // keep <=> (r >> 8) * 2^-24 >= p <=> (r >> 8) >= ceil(p * 2^24): exact on both sides, one integer compare
        const float d = ((rr[e] >> 8) >= p.drop_thr) ? rb(v[4 * j4 + e] * scale) : 0.f;
        v[4 * j4 + e] = a[4 * j4 + e] + d;
      }
    }
  } else if constexpr (EPI == EPI_GELU) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      u[e] = v[e];
      v[e] = gelu_fused(v[e]);
    }
  } else if constexpr (EPI == EPI_GELU_DG) {
#pragma unroll
    for (int e = 0; e < 8; e += 2) gelu_and_grad_x2(v[e], v[e + 1], v[e], v[e + 1], u[e], u[e + 1]);
  } else if constexpr (EPI == EPI_MUL) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = v[e] * a[e];
  } else if constexpr (EPI == EPI_GELU_EAGER) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      u[e] = v[e];
      v[e] = gelu_eager(v[e]);
    }
  }
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

// Rare path: ragged N / unaligned pointers. Element-wise with bounds checks, runtime-dispatched, kept out of line.
template <int EPI>
__device__ __noinline__ void epilogue_segment_slow(const GemmParams& p, uint4 w, long long grow, int gcol) {
  float v[8], a[8], u[8];
  unpack8f(w, v);
  const int nvalid = p.N - gcol;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    a[e] = 0.f;
    if constexpr (EpiTraits<EPI>::kAuxIn)
      if (e < nvalid) a[e] = __bfloat162float(p.aux_in[grow * p.ld_aux_in + gcol + e]);
  }
  if constexpr (EPI == EPI_ROPE) {  // rope_cols, rope_d are multiples of 8: a chunk is either all rotary or none
    if (gcol < p.rope_cols) {
      const long long off = rope_table_off(p, grow, gcol);
      const float4 cs = *reinterpret_cast<const float4*>(p.rope_cos + off);
      float4 sn = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.rope_sin != nullptr) sn = *reinterpret_cast<const float4*>(p.rope_sin + off);
      rope8(v, cs, sn, p.rope_sin != nullptr, false);
    }
  }
  if constexpr (EPI == EPI_ROWMASK) {
    if (reinterpret_cast<const unsigned char*>(p.aux_in)[grow] == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
  }
  epilogue_math<EPI>(p, v, a, u, static_cast<unsigned long long>(grow) * p.N + gcol);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e < nvalid) {
      if constexpr (EpiTraits<EPI>::kAuxOut) p.aux_out[grow * p.ld_aux_out + gcol + e] = __float2bfloat16_rn(u[e]);
      p.D[grow * p.ldd + gcol + e] = __float2bfloat16_rn(v[e]);
    }
  }
}

// Interior fast path of one staged 32-row x 64-column chunk: every row and column is inside the matrix and 16-byte
// aligned (warp-uniform test in the caller), so the 8 row-iterations carry NO branches and the compiler interleaves
// the four iterations of a group (with the per-iteration bounds branches of the general path it emitted them one
// after the other, each a dependent chain: `stall_wait` was the top stall of every second-operand epilogue,
// profiles/r02n_gemm_epi5.source.txt). Row pointers and the dropout stream index advance by constants.
template <int EPI>
__device__ __forceinline__ void epilogue_rows_interior(const GemmParams& p, const uint8_t* stage, int lane,
                                                       long long row_base, int col_base, const uint4 (&axs)[8],
                                                       float (&dacc0)[4], float (&dacc1)[4]) {
  const int rsub = lane >> 3, seg = lane & 7;
  const int gcol = col_base + seg * 8;
  const long long row0 = row_base + rsub;
  __nv_bfloat16* drow = p.D + row0 * p.ldd + gcol;
  __nv_bfloat16* urow = nullptr;
  if constexpr (EpiTraits<EPI>::kAuxOut) urow = p.aux_out + row0 * p.ld_aux_out + gcol;
  unsigned long long eidx = static_cast<unsigned long long>(row0) * p.N + gcol;
  const unsigned long long estep = 4ull * p.N;                         // 4 rows further per iteration
  const bool rope_chunk = (EPI == EPI_ROPE) && col_base < p.rope_cols;  // chunk-uniform: rope_cols % 64 == 0
  const unsigned char* rmask = reinterpret_cast<const unsigned char*>(p.aux_in);
#pragma unroll 1
  for (int hf = 0; hf < 2; ++hf) {
    uint4 w[4];
    float4 rc[4], rs[4];
    unsigned char rm[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rl = (hf * 4 + it) * 4 + rsub;
      w[it] = *reinterpret_cast<const uint4*>(stage + rl * 128 + ((seg ^ (rl & 7)) << 4));
      if constexpr (EPI == EPI_ROPE) {
        rc[it] = make_float4(1.f, 1.f, 1.f, 1.f);
        rs[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rope_chunk) {
          const long long off = rope_table_off(p, row_base + rl, gcol);
          rc[it] = *reinterpret_cast<const float4*>(p.rope_cos + off);
          if (p.rope_sin != nullptr) rs[it] = *reinterpret_cast<const float4*>(p.rope_sin + off);
        }
      }
      if constexpr (EPI == EPI_ROWMASK) rm[it] = rmask[row_base + rl];
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      float v[8], a[8], u[8];
      unpack8f(w[it], v);
      if constexpr (EpiTraits<EPI>::kAuxIn) unpack8f(hf == 0 ? axs[it] : axs[4 + it], a);
      if constexpr (EPI == EPI_ROPE) {
        if (rope_chunk) rope8(v, rc[it], rs[it], p.rope_sin != nullptr, false);
      }
      if constexpr (EPI == EPI_ROWMASK) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = rm[it] != 0 ? v[e] : 0.f;
      }
      if constexpr (EPI == EPI_DELTA) {
        float d8 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) d8 = fmaf(v[e], a[e], d8);
        if (hf == 0) dacc0[it] += d8; else dacc1[it] += d8;
      }
      epilogue_math<EPI>(p, v, a, u, eidx);
      eidx += estep;
      if constexpr (EpiTraits<EPI>::kAuxOut) {
        *reinterpret_cast<uint4*>(urow) = pack8f(u);
        urow += 4 * p.ld_aux_out;
      }
      *reinterpret_cast<uint4*>(drow) = pack8f(v);
      drow += 4 * p.ldd;
    }
  }
}

// 64-column chunks [c_begin, c_end) of this warp's 32 accumulator rows.
template <int EPI>
__device__ __forceinline__ void epilogue_chunks(const GemmParams& p, uint32_t taddr, uint8_t* stage, int lane,
                                                long long row_base, int n0, int split, int c_begin, int c_end) {
  const int rsub = lane >> 3, seg = lane & 7;
  // EPI_DELTA: per-row partial dot products of this lane's 8 columns, summed over the warp's two 64-column chunks
  // (= one 128-wide head); rows (hf * 4 + it) * 4 + rsub
  float dacc0[4] = {0.f, 0.f, 0.f, 0.f}, dacc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const int col_base = n0 + c * 64;
    if (col_base >= p.N) break;  // warp-uniform
    // second operand of all 8 row-iterations FIRST: 8 independent 16-byte loads in flight per lane, issued before the
    // accumulator read-out so that their (L2) latency overlaps the tcgen05.ld and the smem staging. They are the
    // epilogue's only long-latency accesses; issued after the staging, four at a time, they made every K = 1024 GEMM
    // with a second operand epilogue-bound (profiles/r02e_bench_n1.json: 610 / 700 TFLOP/s against ~1590 for the head).
    const int gcol_a = col_base + seg * 8;
    const bool vec_a = p.vec_ok && (p.N - gcol_a >= 8);
    uint4 axs[8];
    if constexpr (EpiTraits<EPI>::kAuxIn) {
#pragma unroll
      for (int it8 = 0; it8 < 8; ++it8) {
        const long long grow = row_base + it8 * 4 + rsub;
        axs[it8] = make_uint4(0, 0, 0, 0);
        if (vec_a && grow < p.M) axs[it8] = *reinterpret_cast<const uint4*>(p.aux_in + grow * p.ld_aux_in + gcol_a);
      }
    }
    uint32_t r0[32], r1[32];
    __syncwarp();
    tmem_ld_32x32(taddr + c * 64, r0);
    tmem_ld_32x32(taddr + c * 64 + 32, r1);
    tmem_ld_wait();
    if constexpr (EPI == EPI_PARTIAL) {
      stage_write_f32(stage, lane, r0);
      __syncwarp();
      partial_readback(p, stage, lane, row_base, col_base, split);
      __syncwarp();
      stage_write_f32(stage, lane, r1);
      __syncwarp();
      partial_readback(p, stage, lane, row_base, col_base + 32, split);
      __syncwarp();
    } else {
      stage_write_bf16(stage, lane, r0, r1);
      __syncwarp();
      const int gcol = col_base + seg * 8;
      // warp-uniform: the whole 32 x 64 chunk inside the matrix, 16-byte accesses allowed
      bool interior = p.vec_ok && (col_base + 64 <= p.N) && (row_base + 32 <= p.M);
      if constexpr (EPI == EPI_ROPE) interior = interior && ((p.rope_cols & 63) == 0);
      if (interior) {
        epilogue_rows_interior<EPI>(p, stage, lane, row_base, col_base, axs, dacc0, dacc1);
      } else {
        // matrix edges (ragged M / N, unaligned pointers): one row-iteration at a time with bounds checks
        const bool col_ok = gcol < p.N;
        const bool vec = p.vec_ok && (p.N - gcol >= 8);
#pragma unroll 1
        for (int it8 = 0; it8 < 8; ++it8) {
          const int rl = it8 * 4 + rsub;
          const long long grow = row_base + rl;
          const uint4 w = *reinterpret_cast<const uint4*>(stage + rl * 128 + ((seg ^ (rl & 7)) << 4));
          if (grow < p.M && col_ok) {
            if (vec) {
              float v[8], a[8], u[8];
              unpack8f(w, v);
              if constexpr (EpiTraits<EPI>::kAuxIn) {
                uint4 ax = axs[0];  // register array: select without dynamic indexing
#pragma unroll
                for (int k = 1; k < 8; ++k) ax = it8 == k ? axs[k] : ax;
                unpack8f(ax, a);
              }
              if constexpr (EPI == EPI_ROPE) {
                if (gcol < p.rope_cols) {
                  const long long off = rope_table_off(p, grow, gcol);
                  const float4 rc = *reinterpret_cast<const float4*>(p.rope_cos + off);
                  float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (p.rope_sin != nullptr) rs = *reinterpret_cast<const float4*>(p.rope_sin + off);
                  rope8(v, rc, rs, p.rope_sin != nullptr, false);
                }
              }
              if constexpr (EPI == EPI_ROWMASK) {
                if (reinterpret_cast<const unsigned char*>(p.aux_in)[grow] == 0) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = 0.f;
                }
              }
              if constexpr (EPI == EPI_DELTA) {
                float d8 = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) d8 = fmaf(v[e], a[e], d8);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (it8 == k) dacc0[k] += d8;
                  if (it8 == 4 + k) dacc1[k] += d8;
                }
              }
              epilogue_math<EPI>(p, v, a, u, static_cast<unsigned long long>(grow) * p.N + gcol);
              if constexpr (EpiTraits<EPI>::kAuxOut)
                *reinterpret_cast<uint4*>(p.aux_out + grow * p.ld_aux_out + gcol) = pack8f(u);
              *reinterpret_cast<uint4*>(p.D + grow * p.ldd + gcol) = pack8f(v);
            } else {
              epilogue_segment_slow<EPI>(p, w, grow, gcol);
            }
          }
        }
      }
      __syncwarp();  // the staging buffer is reused by the next 64-column chunk
    }
  }
  if constexpr (EPI == EPI_DELTA) {
    // the 8 lanes of a row group hold the partial sums of 8 columns each: butterfly over seg, lane seg == 0 stores
    const int head = (n0 + c_begin * 64) >> 7;
    const int n_heads = p.N >> 7;
    if (n0 + c_begin * 64 < p.N) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float t = k < 4 ? dacc0[k & 3] : dacc1[k & 3];
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        const long long grow = row_base + k * 4 + rsub;
        if (seg == 0 && grow < p.M) {
          const long long b = grow / p.delta_T, tt = grow - b * p.delta_T;
          p.delta[(b * n_heads + head) * p.delta_T + tt] = t;
        }
      }
    }
  }
}

// L2 prefetch of the aux_in slice an epilogue warp will read for a tile: rows row_base + lane (32 rows), 128 columns from
// column col0 (256 bytes = two 128-byte lines per row). Issued one tile AHEAD: the in-step per-shape table
// (profiles/r02e_bench_n1.json, roofline.by_shape) showed every GEMM whose epilogue reads a second operand (residual,
// GELU derivative, y for delta) far below the plain ones - 610 / 700 TFLOP/s at K = 1024 against ~1590 for the head -
// because those reads are cold HBM accesses issued only after the tile's main loop, four 16-byte loads per thread at
// a time: latency-bound. With the lines already in L2 the same loads cost a third of the latency.
__device__ __forceinline__ void epilogue_prefetch_aux(const GemmParams& p, long long row, int col0) {
  const bool has_aux = (p.epi == EPI_RESID || p.epi == EPI_GELU_BWD || p.epi == EPI_RESID_DROPOUT || p.epi == EPI_MUL ||
                        p.epi == EPI_DELTA);
  if (!has_aux || row >= p.M || col0 >= p.N) return;
  const char* a = reinterpret_cast<const char*>(p.aux_in + row * p.ld_aux_in + col0);
  asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
  if (col0 + 64 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 128));
}

// Whole accumulator slice of one epilogue warp: rows row_base..row_base+31 (TMEM lanes of this warp's quadrant),
// 64-column chunks [c_begin, c_end) of the tile starting at column n0. One switch per tile.
__device__ __forceinline__ void epilogue_warp_tile(const GemmParams& p, uint32_t taddr, uint8_t* stage, int lane,
                                                   long long row_base, int n0, int split, int c_begin, int c_end) {
  switch (p.epi) {
    case EPI_PLAIN: epilogue_chunks<EPI_PLAIN>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    case EPI_RESID: epilogue_chunks<EPI_RESID>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    case EPI_GELU: epilogue_chunks<EPI_GELU>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    case EPI_GELU_BWD: epilogue_chunks<EPI_GELU_BWD>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    case EPI_PARTIAL: epilogue_chunks<EPI_PARTIAL>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    case EPI_RESID_DROPOUT:
      epilogue_chunks<EPI_RESID_DROPOUT>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end);
      break;
    case EPI_ROPE: epilogue_chunks<EPI_ROPE>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    case EPI_ROWMASK: epilogue_chunks<EPI_ROWMASK>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    case EPI_GELU_DG: epilogue_chunks<EPI_GELU_DG>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    case EPI_MUL: epilogue_chunks<EPI_MUL>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    case EPI_DELTA: epilogue_chunks<EPI_DELTA>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
    default: epilogue_chunks<EPI_GELU_EAGER>(p, taddr, stage, lane, row_base, n0, split, c_begin, c_end); break;
  }
}

}  // namespace obt
