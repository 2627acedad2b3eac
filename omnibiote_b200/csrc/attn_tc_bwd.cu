// Attention backward on tcgen05 tensor cores (head_dim = 128): the adjoint of attn_tc.cu.
//
//   delta_i = dO_i . O_i                                   (attn_delta_kernel, memory-bound pre-pass)
//   dQ kernel : one CTA per 128-query tile, loops over 64-key sub-tiles:  S = Q K^T, dP = dO V^T (TMEM),
//               dS = P o (keep o dP - delta) -> bf16 written back over the S columns,   dQ += dS K   (TMEM)
//   dKV kernel: one CTA per 128-key tile, loops over 64-query sub-tiles: S^T = K Q^T, dP^T = V dO^T (TMEM),
//               P^T, dS^T -> bf16 written back over the score columns,   dV += P^T dO,   dK += dS^T Q   (TMEM)
// P is recomputed from the forward's (row max, log exp-sum) pair. Splitting dQ from dK/dV recomputes S and dP once
// more (7 instead of 5 tile products) but needs no atomics and keeps every accumulator in TMEM.
//
// Common structure: the score tiles are 64 wide so that S/dP fit TWICE in TMEM next to the gradient accumulators
// (2 x 128 + 128 columns for dQ + 128 for Q and dO, 2 x 128 + 256 for dV/dK = all 512). The MMA warp issues the score
// products of sub-tile s+1 before the gradient products of sub-tile s, so the 8 compute warps (two threads per TMEM
// lane, 32 columns each) have a finished score tile to work on while the tensor pipe is busy. Every A operand of a
// gradient product (dS, P^T, dS^T) - and for dQ also Q and dO - lives in TMEM; K / V / Q / dO tiles written by TMA
// serve as K-major B operands of the score products and, read through an MN-major descriptor, as the
// [reduction x 128] B operands of the gradient products: no transposes are materialised, no operand is staged by
// the compute warps in shared memory. Scalings (softmax scale, 1/(1-p)) are folded into the epilogues, which also
// apply the rotary adjoint to dQ / dK. 12 warps: warpgroup 0 = TMA producer + MMA issuer (+ 2 idle), warpgroups
// 1-2 = compute, with register reallocation between them (attn_tc_common.cuh).
#include <stdlib.h>
#include "attn_tc_common.cuh"

namespace obt {

// ---------------------------------------------------------------------------------------------
// delta[b,h,i] = sum_e dO[row, h*128+e] * O[row, h*128+e]; one warp per token row, loops over heads
// ---------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ y,
                                  long long ldy, float* __restrict__ delta, int B, int H, int T) {
  const int warps = blockDim.x >> 5;
  const long long row = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * T) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(row / T), i = static_cast<int>(row % T);
  for (int h = 0; h < H; ++h) {
    const uint2 a = *reinterpret_cast<const uint2*>(dy + row * lddy + h * ATT_D + lane * 4);
    const uint2 c = *reinterpret_cast<const uint2*>(y + row * ldy + h * ATT_D + lane * 4);
    float s = bf16_lo(a.x) * bf16_lo(c.x) + bf16_hi(a.x) * bf16_hi(c.x) + bf16_lo(a.y) * bf16_lo(c.y) +
              bf16_hi(a.y) * bf16_hi(c.y);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) delta[(static_cast<long long>(b) * H + h) * T + i] = s;
  }
}

// =============================================================================================
// dQ kernel (v6). profiles/r01_attn_v6_bwd.source.txt showed v5 bound by shared-memory operand fetch: every 64-key
// sub-tile re-read the Q and dO tiles (A operands, 32 KB each) from smem: 128 KB per sub-tile = 1024 cycles at
// 128 B/clk against 768 cycles of MMA. Now Q, dO AND dS are TMEM A operands:
//   TMEM  [0,128) / [128,256)  score buffers: S (64 columns) | dP (64 columns), ping-pong
//         [256,384)            dQ accumulator
//         [384,448) [448,512)  Q and dO as bf16 (two d per column), written once by the compute threads straight
//                              from global memory (no TMA, no smem tile)
//   dS is written back as bf16 over the S columns its thread has just read and consumed from there by dQ += dS K.
// Shared memory only holds K/V tiles (3 stages): 48 KB of operand fetch per sub-tile.
// =============================================================================================
constexpr int ATT_DQ_KV_STAGES = 3;

struct AttnDqSmem {
  static constexpr uint32_t K_OFF = 0;                                           // stages of 128 keys
  static constexpr uint32_t V_OFF = K_OFF + ATT_DQ_KV_STAGES * ATT_TILE_BYTES;
  static constexpr uint32_t BAR_OFF = V_OFF + ATT_DQ_KV_STAGES * ATT_TILE_BYTES;
  static constexpr uint32_t BYTES = BAR_OFF + 256 + 1024;
};

template <bool kDrop>
__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
attn_tc_dq_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __nv_bfloat16* __restrict__ qkv, long long ld,
                  const __nv_bfloat16* __restrict__ dy, long long lddy, const AttnTcParams p, int C) {
  // 1024-byte alignment (128B-swizzle atoms) is requested from the toolchain, so every smem address below is a
  // link-time constant instead of a live register (the run-time round-up cost registers / spill reloads in the loops)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* sK = smem + AttnDqSmem::K_OFF;
  uint8_t* sV = smem + AttnDqSmem::V_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnDqSmem::BAR_OFF);
  uint64_t* qdo_ready = bars + 0;  // Q / dO are in TMEM (8 compute warps)
  uint64_t* k_full = bars + 1;     // [3]
  uint64_t* v_full = bars + 4;     // [3]
  uint64_t* kv_empty = bars + 7;   // [3]
  uint64_t* sdp_full = bars + 10;  // [2]
  uint64_t* ds_full = bars + 12;   // [2]
  uint64_t* dq_done = bars + 14;   // last dQ MMA completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  int* s_range = reinterpret_cast<int*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;

  // tile metadata precomputed once per micro-batch (obt_attn_tile_meta) instead of a per-CTA scan of the intervals
  int4 qm = make_int4(T, 0, 0, 0);
  if (p.qmeta != nullptr)
    qm = *reinterpret_cast<const int4*>(p.qmeta + (static_cast<long long>(b) * gridDim.x + blockIdx.x) * 4);

  // Compute threads (warps 4..11: two per query row, d columns [64*hh, +64)) issue the global loads of their Q and dO
  // half-rows FIRST: ~1 us of latency that now overlaps the barrier / TMEM set-up and the interval scan instead of
  // heading the critical path (17 % of the stall samples of v6, profiles/r01_attn_v8_dq.source.txt).
  uint4 qv[8], dov[8];
  {
    const int r_ = (warp & 3) * 32 + lane, hh_ = (warp - ATT_BWD_FIRST_COMPUTE_WARP) >> 2;
    const bool ok_ = warp >= ATT_BWD_FIRST_COMPUTE_WARP && t0 + r_ < T;
    const long long row_ = static_cast<long long>(b) * T + t0 + r_;
    const uint4* src = reinterpret_cast<const uint4*>(qkv + row_ * ld + h * ATT_D + hh_ * 64);
    const uint4* src2 = reinterpret_cast<const uint4*>(dy + row_ * lddy + h * ATT_D + hh_ * 64);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      qv[g] = ok_ ? src[g] : make_uint4(0, 0, 0, 0);
      dov[g] = ok_ ? src2[g] : make_uint4(0, 0, 0, 0);
    }
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(qdo_ready, ATT_COMPUTE_WARPS);
    for (int i = 0; i < ATT_DQ_KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&ds_full[i], ATT_COMPUTE_WARPS);
    }
    mbar_init(dq_done, 1);
    fence_barrier_init();
    s_range[0] = T;
    s_range[1] = 0;
    s_range[2] = 0;
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  __syncthreads();
  if (p.row_lo != nullptr && p.qmeta == nullptr && threadIdx.x < ATT_BM && t0 + static_cast<int>(threadIdx.x) < T) {
    const int lo = p.row_lo[static_cast<long long>(b) * T + t0 + threadIdx.x];
    const int hi = p.row_hi[static_cast<long long>(b) * T + t0 + threadIdx.x];
    if (lo >= hi) {
      atomicExch(&s_range[2], 1);
    } else {
      atomicMin(&s_range[0], lo);
      atomicMax(&s_range[1], hi);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // shuffled from lane 0 so that the compiler knows the value is warp-uniform: tcgen05 operands then go through
  // uniform registers directly instead of a per-lane R2UR waterfall loop around every MMA
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(tmem_slot), 0);
  // first / one-past-last 128-key tile; evaluated inside each role branch (values live across the role split would
  // be spilled for the register-poor branch and reloaded in the compute loop)
  auto tile_range = [&](int& jb_out, int& je_out) {
    jb_out = 0;
    je_out = (T + ATT_BN - 1) / ATT_BN;
    if (p.row_lo != nullptr) {
      const int r_lo = p.qmeta ? qm.x : s_range[0], r_hi = p.qmeta ? qm.y : s_range[1];
      const int r_dead = p.qmeta ? qm.z : s_range[2];
      if (r_dead == 0 && r_hi > r_lo) {
        jb_out = r_lo / ATT_BN;
        je_out = (r_hi + ATT_BN - 1) / ATT_BN;
      }
    }
  };
  const int row0 = b * T;
  const int kcol = C + h * ATT_D, vcol = 2 * C + h * ATT_D;
  constexpr uint32_t TM_DQ = 256, TM_Q = 384, TM_DO = 448;

  if (warp < ATT_BWD_FIRST_COMPUTE_WARP) {
   reg_dealloc<56>();
   int jb, je;
   tile_range(jb, je);
   const int n_tiles = je - jb;   // 128-key tiles
   const int n_sub = 2 * n_tiles;  // 64-key sub-tiles
   if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int jj = 0; jj < n_tiles; ++jj) {
        const int krow = row0 + (jb + jj) * ATT_BN;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
        tma_load_2d(&tm_qkv, &k_full[st], sK + st * ATT_TILE_BYTES, kcol, krow);
        tma_load_2d(&tm_qkv, &k_full[st], sK + st * ATT_TILE_BYTES + 16384, kcol + 64, krow);
        mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_2d(&tm_qkv, &v_full[st], sV + st * ATT_TILE_BYTES, vcol, krow);
        tma_load_2d(&tm_qkv, &v_full[st], sV + st * ATT_TILE_BYTES + 16384, vcol + 64, krow);
        if (++st == ATT_DQ_KV_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // The whole warp runs the warp-uniform control flow and waits; one elected lane issues (operands stay in
    // uniform registers: no per-lane R2UR waterfall loop around every MMA).
    const int n_sub_u = __shfl_sync(0xffffffffu, n_sub, 0);
    const bool leader = elect_one();
    mbar_wait(qdo_ready, 0);
    // scores of sub-tile s: S -> buffer (s&1) columns [0,64), dP -> columns [64,128)
    auto issue_scores = [&](int s) {
      const int jj = s >> 1, hsub = s & 1, st = jj % ATT_DQ_KV_STAGES;
      if (hsub == 0) {
        const uint32_t ph = (jj / ATT_DQ_KV_STAGES) & 1;
        mbar_wait(&k_full[st], ph);
        mbar_wait(&v_full[st], ph);
      }
      tc_fence_after();
      if (leader) {
        const uint32_t k_addr = smem_u32(sK + st * ATT_TILE_BYTES) + hsub * 8192;
        const uint32_t v_addr = smem_u32(sV + st * ATT_TILE_BYTES) + hsub * 8192;
        const uint32_t d = tmem_base + (s & 1) * 128;
        issue_scores_ts_128x64(d, tmem_base + TM_Q, k_addr, 16384);        // S  = Q K^T
        issue_scores_ts_128x64(d + 64, tmem_base + TM_DO, v_addr, 16384);  // dP = dO V^T
        umma_commit(&sdp_full[s & 1]);
      }
      __syncwarp();
    };
    issue_scores(0);
    for (int s = 0; s < n_sub_u; ++s) {
      // S(s+1) overwrites the buffer dQ(s-1) read its dS from: MMAs execute in issue order
      if (s + 1 < n_sub_u) issue_scores(s + 1);
      const int jj = s >> 1, hsub = s & 1, st = jj % ATT_DQ_KV_STAGES;
      mbar_wait(&ds_full[s & 1], (s >> 1) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t k_rows = smem_u32(sK + st * ATT_TILE_BYTES) + hsub * 8192;  // key rows 64*hsub .. of the tile
        const uint32_t dsb = tmem_base + (s & 1) * 128;
        issue_grad_ts_128x128x64(tmem_base + TM_DQ, dsb, dsb + 32, k_rows, 16384, s > 0);  // dQ += dS K
        if (hsub == 1) umma_commit(&kv_empty[st]);
        if (s == n_sub_u - 1) umma_commit(dq_done);
      }
      __syncwarp();
    }
   }
  } else {
    reg_alloc<224>();
    int jb, je;
    tile_range(jb, je);
    const int n_sub = 2 * (je - jb);  // 64-key sub-tiles
    const int q = warp & 3;
    const int hh = (warp - ATT_BWD_FIRST_COMPUTE_WARP) >> 2;  // two threads per query row: columns [32*hh, +32)
    const int r = q * 32 + lane;
    const int i = t0 + r;
    const bool row_ok = i < T;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // ---- Q and dO rows (loaded at kernel entry) -> TMEM: 32 packed words of each
    {
      uint32_t w[32];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        w[g * 4 + 0] = qv[g].x; w[g * 4 + 1] = qv[g].y; w[g * 4 + 2] = qv[g].z; w[g * 4 + 3] = qv[g].w;
      }
      __syncwarp();
      tmem_st_32x32(lane_addr + TM_Q + hh * 32, w);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        w[g * 4 + 0] = dov[g].x; w[g * 4 + 1] = dov[g].y; w[g * 4 + 2] = dov[g].z; w[g * 4 + 3] = dov[g].w;
      }
      tmem_st_32x32(lane_addr + TM_DO + hh * 32, w);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qdo_ready);
    }
    int lo = 0, hi = T;
    float row_scale = p.scale;  // natural-log units here
    if (p.row_lo != nullptr && row_ok) {
      lo = p.row_lo[static_cast<long long>(b) * T + i];
      hi = p.row_hi[static_cast<long long>(b) * T + i];
      if (lo >= hi) { lo = 0; hi = T; row_scale = 0.f; }
    }
    const long long bh = static_cast<long long>(b) * p.H + h;
    float off_nat = 0.f, ls2 = 0.f, dl = 0.f;  // row max (natural units), log-sum in log2 units, delta
    if (row_ok) {
      off_nat = p.lse[2 * (bh * T + i)];
      ls2 = p.lse[2 * (bh * T + i) + 1] * LOG2E;
      dl = p.delta[bh * T + i];
    }
    const float neg = off_nat * LOG2E + ls2;  // interval / no-mask path: the max is a plain score, safe to fold
    const float sc2 = row_scale * LOG2E;
    const __nv_bfloat16* mrow =
        (p.mask != nullptr && row_ok) ? p.mask + b * p.msb + h * p.msh + static_cast<long long>(i) * p.msq : nullptr;
    // dropout: dP~ = keep o dP / (1-p). The 1/(1-p) is folded: dS' = P o (keep o dP - delta (1-p)) here and
    // dQ = scale / (1-p) * dS' K in the epilogue.
    const float inv_keep = kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const float dlp = kDrop ? dl * (1.0f - p.drop_p) : dl;
    const uint32_t* keep_row = nullptr;
    if (kDrop && row_ok) keep_row = p.keep + (bh * T + i) * p.nw;

    // keep words one sub-tile ahead (they come from HBM: written a whole forward pass earlier)
    auto load_kw = [&](int s) -> uint32_t {
      const int w = ((jb + (s >> 1)) * ATT_BN + (s & 1) * 64 + hh * 32) >> 5;
      return (kDrop && keep_row != nullptr && s < n_sub && w < p.nw) ? keep_row[w] : 0xffffffffu;
    };
    uint32_t kw_next = load_kw(0);
    for (int s = 0; s < n_sub; ++s) {
      const int bsel = s & 1;
      const int j0 = (jb + (s >> 1)) * ATT_BN + (s & 1) * 64 + hh * 32;  // first key of this thread's 32 columns
      const uint32_t kw = kw_next;
      kw_next = load_kw(s + 1);
      mbar_wait(&sdp_full[bsel], (s >> 1) & 1);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + bsel * 128 + hh * 32, sv);
      tmem_ld_32x32(lane_addr + bsel * 128 + 64 + hh * 32, dv);
      tmem_ld_wait();
      float ds[32];  // first P, then dS
      bool none_visible = false;
      if (p.mask != nullptr) {  // dense additive bias (kernel-uniform branch)
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = j0 + e;
          const bool vis = (j < T) && (mrow != nullptr);
          const float bias = vis ? __bfloat162float(mrow[j]) : 0.f;
          const float sp = __fadd_rn(__fmul_rn(__uint_as_float(sv[e]), p.scale), bias);
          ds[e] = vis ? fast_exp2((sp - off_nat) * LOG2E - ls2) : 0.f;
        }
      } else if (__all_sync(0xffffffffu, row_ok && j0 >= lo && j0 + 32 <= hi)) {
        // interior of a document for every row of the warp: no per-element interval tests
        const float nneg = -neg;
#pragma unroll
        for (int e = 0; e < 32; ++e) ds[e] = fast_exp2(fmaf(__uint_as_float(sv[e]), sc2, nneg));
      } else {
        // the 32 keys straddle an interval end for some row: per-row visibility bits, one bit test per element
        const uint32_t vm = row_ok ? interval_bits32(lo, hi, j0) : 0u;
        if (__all_sync(0xffffffffu, vm == 0u)) {
          none_visible = true;  // no row of the warp sees any of these keys: dS = 0, no exponentials
        } else {
          const float nneg = -neg;
#pragma unroll
          for (int e = 0; e < 32; ++e)
            ds[e] = (vm & (1u << e)) ? fast_exp2(fmaf(__uint_as_float(sv[e]), sc2, nneg)) : 0.f;
        }
      }
      if (none_visible) {
#pragma unroll
        for (int e = 0; e < 32; ++e) ds[e] = 0.f;
      } else if (kDrop) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float dpk = (kw & (1u << keep_bit_pos(e))) ? __uint_as_float(dv[e]) : 0.f;
          ds[e] *= dpk - dlp;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) ds[e] *= __uint_as_float(dv[e]) - dlp;
      }
      // bf16 dS over the first 16 of the S columns this thread has just read (keys 32*hh.. of the sub-tile)
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] = pack_bf16x2(ds[2 * e], ds[2 * e + 1]);
      __syncwarp();
      tmem_st_32x16(lane_addr + bsel * 128 + hh * 32, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[bsel]);
    }
    // dQ epilogue: this thread stores columns [64*hh, 64*hh+64) of its row; its rotary table entries are fetched
    // before the wait for the last MMA
    const bool do_rope = row_ok && p.rope_cos != nullptr;
    float4 rcs[8], rsn[8];
    if (do_rope) {
      const long long toff = static_cast<long long>(i) * (ATT_D / 2) + hh * 32;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        rcs[g] = reinterpret_cast<const float4*>(p.rope_cos + toff)[g];
        rsn[g] = p.rope_sin ? reinterpret_cast<const float4*>(p.rope_sin + toff)[g] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    mbar_wait(dq_done, 0);
    tc_fence_after();
    __nv_bfloat16* drow = p.dq + (static_cast<long long>(row0) + i) * p.ldd + h * ATT_D;
    const float oscale = row_scale * inv_keep;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {  // unrolled: the prefetched table entries stay in registers
      const int c = hh * 2 + cc;
      uint32_t o[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + TM_DQ + c * 32, o);
      tmem_ld_wait();
      if (row_ok) {
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(o[e]) * oscale;
        if (do_rope) {  // adjoint of the rotary embedding: dqkv is the gradient of c_attn's raw output
#pragma unroll
          for (int e = 0; e < 32; ++e) f[e] = rb(f[e]);
          rope_adjoint32(f, &rcs[cc * 4], &rsn[cc * 4], p.rope_sin != nullptr);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          reinterpret_cast<uint4*>(drow + c * 32)[g] =
              make_uint4(pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]), pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]),
                         pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// =============================================================================================
// dK / dV kernel (v6). Same finding as for dQ (operand fetch from shared memory bound the tensor pipe, and the
// single P^T / dS^T smem tile serialised sub-tiles): P^T and dS^T are now written as bf16 INTO the score columns
// their thread has just read and consumed from TMEM by dV += P^T dO and dK += dS^T Q. The per-query parameters are
// staged per warp (each warp only needs the 32 queries of its column half), so the 8 compute warps no longer meet
// at a block barrier every sub-tile (14 % of all stall samples in v5).
//   TMEM  [0,128) / [128,256)  S^T (64 columns) | dP^T (64 columns), ping-pong;  [256,384) dV;  [384,512) dK
// =============================================================================================
constexpr uint32_t ATT_SUB_BYTES = 64 * 128 * 2;  // one [64 rows x 128] bf16 tile = two 8 KB swizzle sub-tiles
constexpr int ATT_QDO_STAGES = 4;

// per warp and buffer: 32 x float2 {-(max+lsum)*log2e, delta*(1-p)} + 32 x int2 {lo,hi} + 32 x float2 {max, lsum*log2e}
// + 32 x float live + 32 keep words + 32 visibility words (which of the warp's 32 keys each query sees)
constexpr uint32_t ATT_WPAR_BYTES = 1152;

struct AttnDkvSmem {
  static constexpr uint32_t K_OFF = 0;
  static constexpr uint32_t V_OFF = K_OFF + ATT_TILE_BYTES;
  static constexpr uint32_t Q_OFF = V_OFF + ATT_TILE_BYTES;
  static constexpr uint32_t DO_OFF = Q_OFF + ATT_QDO_STAGES * ATT_SUB_BYTES;
  static constexpr uint32_t PAR_OFF = DO_OFF + ATT_QDO_STAGES * ATT_SUB_BYTES;  // [8 warps][2 buffers]
  static constexpr uint32_t BAR_OFF = PAR_OFF + ATT_COMPUTE_WARPS * 2 * ATT_WPAR_BYTES;
  static constexpr uint32_t BYTES = BAR_OFF + 256 + 1024;
};

template <bool kDrop>
__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
attn_tc_dkv_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_q64,
                   const __grid_constant__ CUtensorMap tm_dy64, const AttnTcParams p, int C) {
  // 1024-byte alignment (128B-swizzle atoms) is requested from the toolchain, so every smem address below is a
  // link-time constant instead of a live register (the run-time round-up cost registers / spill reloads in the loops)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* sK = smem + AttnDkvSmem::K_OFF;
  uint8_t* sV = smem + AttnDkvSmem::V_OFF;
  uint8_t* sQ = smem + AttnDkvSmem::Q_OFF;
  uint8_t* sDO = smem + AttnDkvSmem::DO_OFF;
  uint8_t* sPar = smem + AttnDkvSmem::PAR_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnDkvSmem::BAR_OFF);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;    // [4]
  uint64_t* qdo_free = bars + 5;    // [4] dV/dK MMAs that read Q/dO stage s completed
  uint64_t* sdp_full = bars + 9;    // [2]
  uint64_t* pds_full = bars + 11;   // [2]
  uint64_t* grads_done = bars + 13; // all dV/dK MMAs completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  unsigned int* s_rel = reinterpret_cast<unsigned int*>(bars + 15);  // relevance bits of the 64-query sub-tiles (<=128)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * ATT_BN;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;
  const int nq = (T + 63) / 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_q64);
    tma_prefetch_desc(&tm_dy64);
    mbar_init(kv_full, 1);
    for (int i = 0; i < ATT_QDO_STAGES; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&pds_full[i], ATT_COMPUTE_WARPS);
    }
    mbar_init(grads_done, 1);
    fence_barrier_init();
    s_rel[0] = s_rel[1] = s_rel[2] = s_rel[3] = 0;
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  __syncthreads();
  const int row0 = b * T;
  const int qcol = h * ATT_D, kcol = C + h * ATT_D, vcol = 2 * C + h * ATT_D;
  if (warp == 0 && lane == 0) {  // K / V do not depend on the relevance scan below: get them in flight first
    mbar_expect_tx(kv_full, 2 * ATT_TILE_BYTES);
    tma_load_2d(&tm_qkv, kv_full, sK, kcol, row0 + j0);
    tma_load_2d(&tm_qkv, kv_full, sK + 16384, kcol + 64, row0 + j0);
    tma_load_2d(&tm_qkv, kv_full, sV, vcol, row0 + j0);
    tma_load_2d(&tm_qkv, kv_full, sV + 16384, vcol + 64, row0 + j0);
  }
  // which 64-query sub-tiles can see this key tile at all: precomputed once per micro-batch (obt_attn_tile_meta) or,
  // without it, scanned here
  if (p.row_lo != nullptr && p.kmeta != nullptr) {
    if (threadIdx.x < 4)
      s_rel[threadIdx.x] = p.kmeta[(static_cast<long long>(b) * gridDim.x + blockIdx.x) * 4 + threadIdx.x];
  } else if (p.row_lo != nullptr) {
    // four rows per thread and pass: all loads in flight before the first dependent atomic
    for (int i0 = threadIdx.x; i0 < T; i0 += 4 * blockDim.x) {
      int lo[4], hi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        lo[u] = 1; hi[u] = 1;  // beyond T: an empty, irrelevant interval
        if (i < T) {
          lo[u] = p.row_lo[static_cast<long long>(b) * T + i];
          hi[u] = p.row_hi[static_cast<long long>(b) * T + i];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        const bool rel = i < T && ((lo[u] >= hi[u]) || (lo[u] < j0 + ATT_BN && hi[u] > j0));
        if (rel) atomicOr(&s_rel[(i >> 6) >> 5], 1u << ((i >> 6) & 31));
      }
    }
  } else if (threadIdx.x < 4) {
    s_rel[threadIdx.x] = 0xffffffffu;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // shuffled from lane 0 so that the compiler knows the value is warp-uniform: tcgen05 operands then go through
  // uniform registers directly instead of a per-lane R2UR waterfall loop around every MMA
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(tmem_slot), 0);
  // read from shared memory at every use: held in registers these words were live across the role split and got
  // spilled, with the reloads on the compute loop's critical path
  auto relevant = [&](int it) -> bool { return (s_rel[it >> 5] >> (it & 31)) & 1u; };

  if (warp < ATT_BWD_FIRST_COMPUTE_WARP) {
   reg_dealloc<56>();
   if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < nq; ++it) {
        if (!relevant(it)) continue;
        mbar_wait(&qdo_free[st], ph ^ 1);
        mbar_expect_tx(&qdo_full[st], 2 * ATT_SUB_BYTES);
        const int qrow = row0 + it * 64;
        tma_load_2d(&tm_q64, &qdo_full[st], sQ + st * ATT_SUB_BYTES, qcol, qrow);
        tma_load_2d(&tm_q64, &qdo_full[st], sQ + st * ATT_SUB_BYTES + 8192, qcol + 64, qrow);
        tma_load_2d(&tm_dy64, &qdo_full[st], sDO + st * ATT_SUB_BYTES, qcol, qrow);
        tma_load_2d(&tm_dy64, &qdo_full[st], sDO + st * ATT_SUB_BYTES + 8192, qcol + 64, qrow);
        if (++st == ATT_QDO_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp in the uniform control flow, one elected issuer (see the dQ kernel)
    const bool leader = elect_one();
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    int n_total = 0;
    for (int it = 0; it < nq; ++it) n_total += relevant(it) ? 1 : 0;
    n_total = __shfl_sync(0xffffffffu, n_total, 0);
    mbar_wait(kv_full, 0);
    auto issue_scores = [&](int n) {  // Q/dO stage n % 4, TMEM score buffer n & 1
      const int st = n % ATT_QDO_STAGES, tb = n & 1;
      mbar_wait(&qdo_full[st], (n / ATT_QDO_STAGES) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t d = tmem_base + tb * 128;
        issue_scores_128x64(d, k_addr, 16384, smem_u32(sQ + st * ATT_SUB_BYTES), 8192);        // S^T  = K Q^T
        issue_scores_128x64(d + 64, v_addr, 16384, smem_u32(sDO + st * ATT_SUB_BYTES), 8192);  // dP^T = V dO^T
        umma_commit(&sdp_full[tb]);
      }
      __syncwarp();
    };
    if (n_total > 0) issue_scores(0);
    for (int n = 0; n < n_total; ++n) {
      // the scores of n+1 overwrite the buffer the gradient products of n-1 read from: MMAs execute in issue order
      if (n + 1 < n_total) issue_scores(n + 1);
      const int st = n % ATT_QDO_STAGES, tb = n & 1;
      mbar_wait(&pds_full[tb], (n >> 1) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t buf = tmem_base + tb * 128;
        issue_grad_ts_128x128x64(tmem_base + 256, buf, buf + 32, smem_u32(sDO + st * ATT_SUB_BYTES), 8192, n > 0);      // dV += P^T dO
        issue_grad_ts_128x128x64(tmem_base + 384, buf + 64, buf + 96, smem_u32(sQ + st * ATT_SUB_BYTES), 8192, n > 0);  // dK += dS^T Q
        umma_commit(&qdo_free[st]);
        if (n == n_total - 1) umma_commit(grads_done);
      }
      __syncwarp();
    }
   }
  } else {
    reg_alloc<224>();
    const int q = warp & 3;
    const int hh = (warp - ATT_BWD_FIRST_COMPUTE_WARP) >> 2;  // two threads per key row: query columns [32*hh, +32)
    const int r = q * 32 + lane;     // key row within the tile
    const int j = j0 + r;
    const bool key_ok = j < T;
    const int kq0 = j0 + q * 32;     // first key of this warp
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const long long bh = static_cast<long long>(b) * p.H + h;
    const float inv_keep = kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const float keep_frac = kDrop ? 1.0f - p.drop_p : 1.0f;
    const float sc2 = p.scale * LOG2E;
    const uint32_t mybit = 1u << keep_bit_pos(lane);  // this key's bit inside the keep word of its 32-key group
    uint8_t* wpar = sPar + (warp - ATT_BWD_FIRST_COMPUTE_WARP) * 2 * ATT_WPAR_BYTES;
    // Per-query parameters of THIS warp's 32 query columns (lane = query), software-pipelined: the global loads for
    // the next relevant sub-tile are issued before the math of the current one and staged in the warp's other smem
    // buffer afterwards; the math reads them back as warp-wide broadcasts.
    // load_params only ISSUES the global loads (raw values; no arithmetic or branches on loaded data, which would
    // stall the in-order warp right there); finish_params turns them into the staged form after the math.
    struct QParams { int lo, hi; float off, ls2, dl, live; uint32_t kw, vm; };
    auto load_params = [&](int it) -> QParams {
      QParams z;
      z.lo = 0; z.hi = 0; z.off = 0.f; z.ls2 = 0.f; z.dl = 0.f; z.live = 1.f;  // query beyond T: contributes nothing
      z.kw = 0xffffffffu;
      z.vm = 0u;
      const int i = it * 64 + hh * 32 + lane;
      if (i < T) {
        z.hi = T;
        if (p.row_lo != nullptr) {
          z.lo = p.row_lo[static_cast<long long>(b) * T + i];
          z.hi = p.row_hi[static_cast<long long>(b) * T + i];
        }
        const float2 ml = *reinterpret_cast<const float2*>(p.lse + 2 * (bh * T + i));
        z.off = ml.x;
        z.ls2 = ml.y;
        z.dl = p.delta[bh * T + i];
        if (kDrop && (kq0 >> 5) < p.nw) z.kw = p.keep[(bh * T + i) * p.nw + (kq0 >> 5)];
      }
      return z;
    };
    auto finish_params = [&](QParams& z, int it) {
      const int i = it * 64 + hh * 32 + lane;
      if (i < T) {
        if (z.lo >= z.hi) { z.lo = 0; z.hi = T; z.live = 0.f; }  // fully-masked row: uniform P, no dS
        z.vm = interval_bits32(z.lo, z.hi, kq0);  // hi <= T: keys beyond the sequence are never visible
      }
      z.ls2 *= LOG2E;
      z.dl *= keep_frac;
    };
    auto store_params = [&](int buf, const QParams& z) {
      uint8_t* base = wpar + buf * ATT_WPAR_BYTES;
      reinterpret_cast<float2*>(base)[lane] = make_float2(-(z.off * LOG2E + z.ls2), z.dl);
      reinterpret_cast<int2*>(base + 256)[lane] = make_int2(z.lo, z.hi);
      reinterpret_cast<float2*>(base + 512)[lane] = make_float2(z.off, z.ls2);
      reinterpret_cast<float*>(base + 768)[lane] = z.live;
      reinterpret_cast<uint32_t*>(base + 896)[lane] = z.kw;
      reinterpret_cast<uint32_t*>(base + 1024)[lane] = z.vm;
    };
    auto next_relevant = [&](int it) -> int {
      ++it;
      while (it < nq && !relevant(it)) ++it;
      return it;
    };

    // dS hand-over to the score-free dQ kernel: that kernel accumulates over 128-query tiles, so a 64-query sub-tile this
    // CTA skips while its sibling is processed must read as zeros
    __nv_bfloat16* ds_row = nullptr;
    if (p.ds_out != nullptr && key_ok) ds_row = p.ds_out + (bh * T + j) * p.ds_pitch + hh * 32;
    if (ds_row != nullptr) {
      for (int z = 0; z < nq; ++z)
        if (!relevant(z) && relevant(z ^ 1)) {
          const uint32_t zero8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          st_global_256(ds_row + z * 64, zero8);
          st_global_256(ds_row + z * 64 + 16, zero8);
        }
    }

    // two sub-tiles of look-ahead: the keep words come from HBM (written a whole forward pass earlier) and one
    // sub-tile of math (~1 us) did not cover that latency (11 % of the stall samples sat on the first use)
    int n = 0;
    int it = next_relevant(-1);
    int nx = it < nq ? next_relevant(it) : nq;
    QParams cur = {}, zn = {};
    if (it < nq) {
      cur = load_params(it);
      if (nx < nq) zn = load_params(nx);
      finish_params(cur, it);
      store_params(0, cur);
    }
    __syncwarp();
    while (it < nq) {
      const int st = n & 1;
      const int i0 = it * 64 + hh * 32;
      const int nx2 = nx < nq ? next_relevant(nx) : nq;
      QParams zn2 = {};
      if (nx2 < nq) zn2 = load_params(nx2);  // in flight during the math of this AND the next sub-tile
      const uint8_t* base = wpar + st * ATT_WPAR_BYTES;
      const float4* nd4 = reinterpret_cast<const float4*>(base);            // two queries per float4
      const int2* c_lh = reinterpret_cast<const int2*>(base + 256);
      const float2* c_x = reinterpret_cast<const float2*>(base + 512);
      const float* c_live = reinterpret_cast<const float*>(base + 768);
      const uint4* kp4 = reinterpret_cast<const uint4*>(base + 896);        // four queries per uint4
      const uint4* vp4 = reinterpret_cast<const uint4*>(base + 1024);
      const uint32_t lanebit = 1u << lane;
      // every key of this warp visible to every (live) query of its 32 columns: one vote
      const bool interior = (p.mask == nullptr) && (kq0 + 32 <= T) &&
                            __all_sync(0xffffffffu, cur.live != 0.f && cur.lo <= kq0 && cur.hi >= kq0 + 32);
      mbar_wait(&sdp_full[st], (n >> 1) & 1);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + st * 128 + hh * 32, sv);
      tmem_ld_32x32(lane_addr + st * 128 + 64 + hh * 32, dv);
      tmem_ld_wait();
      uint32_t ptw[16], dsw[16];  // bf16 pairs of P^T and dS^T
      if (interior) {
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          uint4 kk = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (kDrop) kk = kp4[e4];
          const uint32_t kws[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float4 nd = nd4[e4 * 2 + h2];
            const int e = e4 * 4 + h2 * 2;
            const float pr0 = fast_exp2(fmaf(__uint_as_float(sv[e]), sc2, nd.x));
            const float pr1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), sc2, nd.z));
            const bool kb0 = !kDrop || (kws[h2 * 2] & mybit), kb1 = !kDrop || (kws[h2 * 2 + 1] & mybit);
            ptw[e >> 1] = pack_bf16x2(kb0 ? pr0 : 0.f, kb1 ? pr1 : 0.f);
            dsw[e >> 1] = pack_bf16x2(pr0 * ((kb0 ? __uint_as_float(dv[e]) : 0.f) - nd.y),
                                      pr1 * ((kb1 ? __uint_as_float(dv[e + 1]) : 0.f) - nd.w));
          }
        }
      } else if (p.mask == nullptr && __all_sync(0xffffffffu, cur.vm == 0u && cur.live != 0.f)) {
        // no query of the chunk sees any key of this warp: P^T = dS^T = 0, no exponentials
#pragma unroll
        for (int e = 0; e < 16; ++e) ptw[e] = dsw[e] = 0u;
      } else if (p.mask == nullptr && __all_sync(0xffffffffu, cur.live != 0.f)) {
        // an interval end crosses the 32 x 32 block: per-query visibility words, one bit test per element
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          uint4 kk = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (kDrop) kk = kp4[e4];
          const uint4 vv = vp4[e4];
          const uint32_t kws[4] = {kk.x, kk.y, kk.z, kk.w};
          const uint32_t vms[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float4 nd = nd4[e4 * 2 + h2];
            const int e = e4 * 4 + h2 * 2;
            const float pr0 = (vms[h2 * 2] & lanebit) ? fast_exp2(fmaf(__uint_as_float(sv[e]), sc2, nd.x)) : 0.f;
            const float pr1 = (vms[h2 * 2 + 1] & lanebit) ? fast_exp2(fmaf(__uint_as_float(sv[e + 1]), sc2, nd.z)) : 0.f;
            const bool kb0 = !kDrop || (kws[h2 * 2] & mybit), kb1 = !kDrop || (kws[h2 * 2 + 1] & mybit);
            ptw[e >> 1] = pack_bf16x2(kb0 ? pr0 : 0.f, kb1 ? pr1 : 0.f);
            dsw[e >> 1] = pack_bf16x2(pr0 * ((kb0 ? __uint_as_float(dv[e]) : 0.f) - nd.y),
                                      pr1 * ((kb1 ? __uint_as_float(dv[e + 1]) : 0.f) - nd.w));
          }
        }
      } else if (p.mask == nullptr) {
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          uint4 kk = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (kDrop) kk = kp4[e4];
          const uint32_t kws[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            float prs[2], dss[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int e = e4 * 4 + h2 * 2 + u;
              const float2 nd = reinterpret_cast<const float2*>(nd4)[e];
              const int2 lh = c_lh[e];
              const float live = c_live[e];
              const bool vis = key_ok && j >= lh.x && j < lh.y;
              const float pr = vis ? fast_exp2(fmaf(__uint_as_float(sv[e]) * live, sc2, nd.x)) : 0.f;
              const bool kb = !kDrop || (kws[h2 * 2 + u] & mybit);
              prs[u] = kb ? pr : 0.f;
              dss[u] = pr * ((kb ? __uint_as_float(dv[e]) : 0.f) - nd.y) * live;
            }
            ptw[e4 * 2 + h2] = pack_bf16x2(prs[0], prs[1]);
            dsw[e4 * 2 + h2] = pack_bf16x2(dss[0], dss[1]);
          }
        }
      } else {  // dense additive bias
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          uint4 kk = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (kDrop) kk = kp4[e4];
          const uint32_t kws[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            float prs[2], dss[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int e = e4 * 4 + h2 * 2 + u;
              const int i = i0 + e;
              const float2 nd = reinterpret_cast<const float2*>(nd4)[e];
              const float2 cx = c_x[e];
              const bool vis = key_ok && i < T;
              const float bias =
                  vis ? __bfloat162float(p.mask[b * p.msb + h * p.msh + static_cast<long long>(i) * p.msq + j]) : 0.f;
              const float sp = __fadd_rn(__fmul_rn(__uint_as_float(sv[e]), p.scale), bias);
              const float pr = vis ? fast_exp2((sp - cx.x) * LOG2E - cx.y) : 0.f;
              const bool kb = !kDrop || (kws[h2 * 2 + u] & mybit);
              prs[u] = kb ? pr : 0.f;
              dss[u] = pr * ((kb ? __uint_as_float(dv[e]) : 0.f) - nd.y);
            }
            ptw[e4 * 2 + h2] = pack_bf16x2(prs[0], prs[1]);
            dsw[e4 * 2 + h2] = pack_bf16x2(dss[0], dss[1]);
          }
        }
      }
      if (ds_row != nullptr) {
        // dS^T[key j][queries i0 .. i0+31] for the score-free dQ kernel: the lanes of a warp write 32 different rows, so
        // two 256-bit stores (one full sector each) instead of four 128-bit ones halve the LSU transactions - with the
        // latter the stores cost this kernel 75 us (264 -> 339 us, profiles/r02f_attn_probe_launches_ds*.txt)
        const uint32_t lo8[8] = {dsw[0], dsw[1], dsw[2], dsw[3], dsw[4], dsw[5], dsw[6], dsw[7]};
        const uint32_t hi8[8] = {dsw[8], dsw[9], dsw[10], dsw[11], dsw[12], dsw[13], dsw[14], dsw[15]};
        st_global_256(ds_row + it * 64, lo8);
        st_global_256(ds_row + it * 64 + 16, hi8);
      }
      // P^T over the first 16 of the S^T columns this thread has read, dS^T over the first 16 of its dP^T columns
      __syncwarp();
      tmem_st_32x16(lane_addr + st * 128 + hh * 32, ptw);
      tmem_st_32x16(lane_addr + st * 128 + 64 + hh * 32, dsw);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[st]);
      // parameters of the next sub-tile into the warp's other buffer (last read during sub-tile n-1)
      if (nx < nq) {
        finish_params(zn, nx);
        store_params(st ^ 1, zn);
      }
      cur = zn;
      zn = zn2;
      __syncwarp();
      it = nx;
      nx = nx2;
      ++n;
    }
    // epilogue: hh = 0 stores dV (TMEM columns 256..383) / (1-p), hh = 1 stores dK (384..511) * scale / (1-p)
    __nv_bfloat16* dvrow = p.dv + (static_cast<long long>(row0) + j) * p.ldd + h * ATT_D;
    __nv_bfloat16* dkrow = p.dk + (static_cast<long long>(row0) + j) * p.ldd + h * ATT_D;
    const float oscale = hh == 0 ? inv_keep : inv_keep * p.scale;
    // rotary table row of this key (dK only), fetched before the wait for the last MMAs
    const bool do_rope = hh == 1 && key_ok && p.rope_cos != nullptr;
    float4 rcs[16], rsn[16];
    if (do_rope) {
      const long long toff = static_cast<long long>(j) * (ATT_D / 2);
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        rcs[g] = reinterpret_cast<const float4*>(p.rope_cos + toff)[g];
        rsn[g] = p.rope_sin ? reinterpret_cast<const float4*>(p.rope_sin + toff)[g] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (n > 0) {
      mbar_wait(grads_done, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {  // unrolled: the prefetched table entries stay in registers
      const int c = hh * 4 + cc;
      uint32_t o[32];
      __syncwarp();
      if (n > 0) {
        tmem_ld_32x32(lane_addr + 256 + c * 32, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = 0u;
      }
      if (key_ok) {
        __nv_bfloat16* dstp = (c < 4 ? dvrow : dkrow) + (c & 3) * 32;
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(o[e]) * oscale;
        if (do_rope) {  // dK: adjoint of the rotary embedding at key position j
#pragma unroll
          for (int e = 0; e < 32; ++e) f[e] = rb(f[e]);
          rope_adjoint32(f, &rcs[cc * 4], &rsn[cc * 4], p.rope_sin != nullptr);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          reinterpret_cast<uint4*>(dstp)[g] =
              make_uint4(pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]), pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]),
                         pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 512);
  }
}


// =============================================================================================
// Score-free dQ kernel (round 2). The dK/dV kernel computes every dS^T tile anyway (as bf16, for dK += dS^T Q); when it
// also stores them (AttnTcParams::ds_out), dQ = scale / (1-p) * dS K needs neither the score products S = Q K^T and
// dP = dO V^T nor the exponentials a second time: per 64 keys ONE 128 x 128 x 64 product instead of three, no MUFU, no
// per-row parameters, 128 TMEM columns and 96 KB of shared memory, so two CTAs share an SM and hide each other's
// prologue / epilogue. The price is the round trip of the visited dS tiles through L2 / HBM (2 bytes per score).
//   A = dS[128 queries x 64 keys]: the stored dS^T tile [64 keys][128 queries] read as an MN-major operand (two
//       64-query halves 8 KB apart); B = K[64 keys x 128 d] read MN-major like in the dQ += dS K product above.
// A 128-key tile is visited iff the dK/dV kernel wrote it: relevance bit of either 64-query half of this query tile.
// =============================================================================================
constexpr int ATT_DQ2_STAGES = 3;
constexpr int ATT_DQ2_THREADS = 64 + 128;  // TMA warp, MMA warp, 4 epilogue warps (one thread per query row)

struct AttnDq2Smem {
  static constexpr uint32_t A_OFF = 0;                                      // dS^T sub-tiles [64 keys x 128 queries]
  static constexpr uint32_t B_OFF = A_OFF + ATT_DQ2_STAGES * ATT_SUB_BYTES;  // K sub-tiles [64 keys x 128 d]
  static constexpr uint32_t BAR_OFF = B_OFF + ATT_DQ2_STAGES * ATT_SUB_BYTES;
  static constexpr uint32_t BYTES = BAR_OFF + 256 + 1024;
};

__global__ void __launch_bounds__(ATT_DQ2_THREADS, 2)
attn_tc_dq2_kernel(const __grid_constant__ CUtensorMap tm_ds, const __grid_constant__ CUtensorMap tm_k64,
                   const AttnTcParams p, int C) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* sA = smem + AttnDq2Smem::A_OFF;
  uint8_t* sB = smem + AttnDq2Smem::B_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnDq2Smem::BAR_OFF);
  uint64_t* full = bars + 0;    // [3]
  uint64_t* empty = bars + 3;   // [3]
  uint64_t* done = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tq = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;
  const int nT = gridDim.x;
  const int t0 = tq * ATT_BM;
  // which 128-key tiles the dK/dV kernel visited for either 64-query half of this tile (bit kt)
  uint32_t visit = 0;
  for (int kt = 0; kt < nT; ++kt) {
    bool v = true;
    if (p.kmeta != nullptr) {
      const unsigned int* km = p.kmeta + (static_cast<long long>(b) * nT + kt) * 4;
      const int s0 = 2 * tq, s1 = 2 * tq + 1;
      v = ((km[s0 >> 5] >> (s0 & 31)) & 1u) || ((km[s1 >> 5] >> (s1 & 31)) & 1u);
    }
    if (v) visit |= 1u << kt;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_ds);
    tma_prefetch_desc(&tm_k64);
    for (int i = 0; i < ATT_DQ2_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(tmem_slot), 0);
  const int n_sub = 2 * __popc(visit);  // 64-key sub-tiles
  const int row0 = b * T;
  const int bh = b * p.H + h;
  const int kcol = C + h * ATT_D;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int kt = 0; kt < nT; ++kt) {
        if (!((visit >> kt) & 1u)) continue;
#pragma unroll 1
        for (int hs = 0; hs < 2; ++hs) {
          const int key0 = kt * ATT_BN + hs * 64;
          mbar_wait(&empty[st], ph ^ 1);
          mbar_expect_tx(&full[st], 2 * ATT_SUB_BYTES);
          // dS^T rows key0.. (zero-filled beyond T keys), queries t0..t0+127 as two 64-wide halves
          tma_load_3d(&tm_ds, &full[st], sA + st * ATT_SUB_BYTES, t0, key0, bh);
          tma_load_3d(&tm_ds, &full[st], sA + st * ATT_SUB_BYTES + 8192, t0 + 64, key0, bh);
          tma_load_2d(&tm_k64, &full[st], sB + st * ATT_SUB_BYTES, kcol, row0 + key0);
          tma_load_2d(&tm_k64, &full[st], sB + st * ATT_SUB_BYTES + 8192, kcol + 64, row0 + key0);
          if (++st == ATT_DQ2_STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const int n_sub_u = __shfl_sync(0xffffffffu, n_sub, 0);
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, true, true);
    for (int s = 0; s < n_sub_u; ++s) {
      const int st = s % ATT_DQ2_STAGES;
      mbar_wait(&full[st], (s / ATT_DQ2_STAGES) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t a_addr = smem_u32(sA + st * ATT_SUB_BYTES), b_addr = smem_u32(sB + st * ATT_SUB_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t a_desc = make_smem_desc_sw128(a_addr + kk * 2048, 8192, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(b_addr + kk * 2048, 8192, 1024);
          umma_bf16_ss<1>(tmem_base, a_desc, b_desc, idesc, (s > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty[st]);
        if (s == n_sub_u - 1) umma_commit(done);
      }
      __syncwarp();
    }
  } else {
    // epilogue: one thread per query row; dq = scale / (1-p) * acc, rotary adjoint, bf16
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int i = t0 + r;
    const bool row_ok = i < T;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float oscale = p.scale * (p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f);
    const bool do_rope = row_ok && p.rope_cos != nullptr;
    __nv_bfloat16* drow = p.dq + (static_cast<long long>(row0) + i) * p.ldd + h * ATT_D;
    if (n_sub > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      __syncwarp();
      if (n_sub > 0) {
        tmem_ld_32x32(lane_addr + c * 32, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = 0u;
      }
      if (row_ok) {
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(o[e]) * oscale;
        if (do_rope) {
          const long long toff = static_cast<long long>(i) * (ATT_D / 2) + c * 16;
          float4 rcs[4], rsn[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            rcs[g] = reinterpret_cast<const float4*>(p.rope_cos + toff)[g];
            rsn[g] = p.rope_sin ? reinterpret_cast<const float4*>(p.rope_sin + toff)[g] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) f[e] = rb(f[e]);
          rope_adjoint32(f, rcs, rsn, p.rope_sin != nullptr);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          reinterpret_cast<uint4*>(drow + c * 32)[g] =
              make_uint4(pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]), pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]),
                         pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 128);
  }
}

}  // namespace obt

using namespace obt;

// Measured and rejected (round 2): persistent forms of both kernels (one CTA per SM fetching (tile, head, batch) items
// from a device counter, the TMA producer running ahead across items, Q / dO of the next item loaded before the
// accumulator read-out). Bit-identical results, but dQ 211.4 -> 198.4 us and dK/dV 257.9 -> 322.9 us stand-alone and
// no gain inside the step (profiles/r02d_attn_probe_launches_*.txt, r02d_attn_bwd_persist.details.txt): the per-item
// latencies that remain (first parameter loads, first score product, accumulator read-out) are on the compute warps'
// critical path in either form, and the sub-tile loop itself (MUFU 512 cycles + ~600 cycles of FP32 / pack / TMEM
// traffic per 64 keys on 8 warps, against 768 cycles of MMA) sets the pace.

extern "C" int obt_attn_tc_bwd(const void* qkv, long long ld, const void* mask, long long msb, long long msh,
                               long long msq, const int* row_lo, const int* row_hi, const void* y, long long ldy,
                               const void* dy, long long lddy, const float* lse, float* delta, int delta_ready,
                               void* dqkv, long long ldd, int B, int H, int T, int d, float scale, float drop_p,
                               const unsigned int* keep, const float* rope_cos, const float* rope_sin,
                               const int* qmeta, const unsigned int* kmeta, void* ds_scratch, cudaStream_t stream) {
  OBT_REQUIRE(qkv && y && dy && lse && delta && dqkv, "obt_attn_tc_bwd: null pointer");
  OBT_REQUIRE((reinterpret_cast<uintptr_t>(rope_cos) & 15) == 0 && (reinterpret_cast<uintptr_t>(rope_sin) & 15) == 0,
              "obt_attn_tc_bwd: rotary tables must be 16-byte aligned");
  OBT_REQUIRE(d == ATT_D, "obt_attn_tc_bwd: head_dim=%d, the tensor-core kernel is specialised for 128", d);
  OBT_REQUIRE(B > 0 && H > 0 && T > 0, "obt_attn_tc_bwd: empty problem");
  OBT_REQUIRE(T <= 128 * 64, "obt_attn_tc_bwd: T=%d exceeds %d", T, 128 * 64);
  OBT_REQUIRE(ld % 8 == 0 && ldy % 8 == 0 && lddy % 8 == 0 && ldd % 8 == 0, "obt_attn_tc_bwd: pitches must be multiples of 8");
  OBT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "obt_attn_tc_bwd: dropout p=%f", drop_p);
  OBT_REQUIRE(drop_p == 0.f || keep != nullptr, "obt_attn_tc_bwd: dropout needs the keep mask (obt_attn_keep_mask)");
  const int C = H * d;
  const long long M = static_cast<long long>(B) * T;
  OBT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0,
              "obt_attn_tc_bwd: qkv and dy must be 16-byte aligned");
  CUtensorMap tm_qkv, tm_q64, tm_dy64;
  int rc = get_tensor_map_2d(&tm_qkv, qkv, static_cast<uint64_t>(3) * C, static_cast<uint64_t>(M),
                             static_cast<uint64_t>(ld), 64, 128);
  if (rc) return rc;
  rc = get_tensor_map_2d(&tm_q64, qkv, static_cast<uint64_t>(3) * C, static_cast<uint64_t>(M), static_cast<uint64_t>(ld),
                         64, 64);
  if (rc) return rc;
  rc = get_tensor_map_2d(&tm_dy64, dy, static_cast<uint64_t>(C), static_cast<uint64_t>(M), static_cast<uint64_t>(lddy), 64,
                         64);
  if (rc) return rc;
  if (!delta_ready) {  // otherwise the GEMM that produced dy already emitted delta (epilogue 11)
    const int warps = 8;
    attn_delta_kernel<<<static_cast<unsigned>((M + warps - 1) / warps), warps * 32, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(dy), lddy, static_cast<const __nv_bfloat16*>(y), ldy, delta, B, H, T);
    rc = check_launch("attn_delta");
    if (rc) return rc;
  }
  AttnTcParams p = {};
  p.B = B; p.H = H; p.T = T;
  p.scale = scale;
  p.mask = static_cast<const __nv_bfloat16*>(mask);
  p.msb = msb; p.msh = msh; p.msq = msq;
  p.row_lo = mask ? nullptr : row_lo;
  p.row_hi = mask ? nullptr : row_hi;
  p.qmeta = (p.row_lo != nullptr) ? qmeta : nullptr;
  p.kmeta = (p.row_lo != nullptr) ? kmeta : nullptr;
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.drop_p = drop_p;
  p.keep = keep;
  p.nw = keep_words(T);
  p.rope_cos = rope_cos;
  p.rope_sin = rope_cos ? rope_sin : nullptr;
  p.dq = static_cast<__nv_bfloat16*>(dqkv);
  p.dk = p.dq + C;
  p.dv = p.dq + 2 * C;
  p.ldd = ldd;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaSuccess;
    auto set = [&](auto kern, size_t bytes) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    };
    set(attn_tc_dq_kernel<false>, AttnDqSmem::BYTES);
    set(attn_tc_dq_kernel<true>, AttnDqSmem::BYTES);
    set(attn_tc_dkv_kernel<false>, AttnDkvSmem::BYTES);
    set(attn_tc_dkv_kernel<true>, AttnDkvSmem::BYTES);
    set(attn_tc_dq2_kernel, AttnDq2Smem::BYTES);
    if (e != cudaSuccess) {
      set_last_error("obt_attn_tc_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return OBT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((T + ATT_BM - 1) / ATT_BM, H, B);
  const bool drop = drop_p > 0.f;
  // dS hand-over (ds_scratch given): the dK/dV kernel stores its dS^T tiles, a score-free kernel turns them into dQ
  const long long ds_pitch = (static_cast<long long>(T) + 63) / 64 * 64;
  // the score-free kernel must visit exactly the tiles the dK/dV kernel wrote: it needs the same relevance bits (tile
  // metadata) whenever an interval mask is in use, and keeps its visit set in one 32-bit word (T <= 4096)
  if (ds_scratch != nullptr && ((p.row_lo != nullptr && p.kmeta == nullptr) || T > 32 * ATT_BN)) ds_scratch = nullptr;
  if (ds_scratch != nullptr) {
    OBT_REQUIRE((reinterpret_cast<uintptr_t>(ds_scratch) & 127) == 0, "obt_attn_tc_bwd: ds_scratch must be 128-byte aligned");
    p.ds_out = static_cast<__nv_bfloat16*>(ds_scratch);
    p.ds_pitch = ds_pitch;
  } else {
    auto qk = static_cast<const __nv_bfloat16*>(qkv);
    auto dyp = static_cast<const __nv_bfloat16*>(dy);
    if (drop) attn_tc_dq_kernel<true><<<grid, ATT_BWD_THREADS, AttnDqSmem::BYTES, stream>>>(tm_qkv, qk, ld, dyp, lddy, p, C);
    else attn_tc_dq_kernel<false><<<grid, ATT_BWD_THREADS, AttnDqSmem::BYTES, stream>>>(tm_qkv, qk, ld, dyp, lddy, p, C);
    rc = check_launch("attn_tc_dq");
    if (rc) return rc;
  }
  if (drop) attn_tc_dkv_kernel<true><<<grid, ATT_BWD_THREADS, AttnDkvSmem::BYTES, stream>>>(tm_qkv, tm_q64, tm_dy64, p, C);
  else attn_tc_dkv_kernel<false><<<grid, ATT_BWD_THREADS, AttnDkvSmem::BYTES, stream>>>(tm_qkv, tm_q64, tm_dy64, p, C);
  if (ds_scratch != nullptr) {
    rc = check_launch("attn_tc_dkv");
    if (rc) return rc;
    CUtensorMap tm_ds;
    // [B*H][T keys][ds_pitch queries]: queries contiguous; keys beyond T read as zeros (per (batch, head) slab)
    rc = get_tensor_map_3d(&tm_ds, ds_scratch, static_cast<uint64_t>(ds_pitch), static_cast<uint64_t>(T),
                           static_cast<uint64_t>(B) * H, static_cast<uint64_t>(ds_pitch),
                           static_cast<uint64_t>(ds_pitch) * T, 64, 64, 1);
    if (rc) return rc;
    attn_tc_dq2_kernel<<<grid, ATT_DQ2_THREADS, AttnDq2Smem::BYTES, stream>>>(tm_ds, tm_q64, p, C);
    return check_launch("attn_tc_dq2");
  }
  return check_launch("attn_tc_dkv");
}
