// Attention backward on tcgen05 tensor cores (head_dim = 128): the adjoint of attn_tc.cu.
//
//   delta_i = dO_i . O_i                                   (attn_delta_kernel, memory-bound pre-pass)
//   dQ kernel : one CTA per 128-query tile, loops over 64-key sub-tiles:  S = Q K^T, dP = dO V^T (TMEM),
//               dS = P o (dP - delta) * scale -> bf16 smem tile,   dQ += dS K   (accumulated in TMEM)
//   dKV kernel: one CTA per 128-key tile, loops over 64-query sub-tiles: S^T = K Q^T, dP^T = V dO^T (TMEM),
//               P^T, dS^T -> bf16 smem tiles,   dV += P^T dO,   dK += dS^T Q   (accumulated in TMEM)
// P is recomputed from the forward's (row max, log exp-sum) pair. Splitting dQ from dK/dV recomputes S and dP once
// more (7 instead of 5 tile products) but needs no atomics and keeps every accumulator in TMEM.
//
// Pipelining (v3): the score tiles are 64 wide so that S/dP fit TWICE in TMEM next to the gradient accumulators
// (2 x 128 + 128 columns for dQ, 2 x 128 + 256 for dV/dK = all 512). The MMA thread issues the score products of
// sub-tile s+1 before the gradient products of sub-tile s, so the 8 compute warps (two threads per TMEM lane, 32
// columns each) always have a finished score tile to work on while the tensor pipe is busy (v2 alternated between
// the two and measured 552 / 789 us per layer, profiles/r01_launches_v3.txt).
// The Q / dO / K / V smem tiles written by TMA serve both as K-major operands (scores) and, read through an
// MN-major descriptor, as the [reduction x 128] operands of the gradient products: no transposes are materialised.
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
// dQ kernel
// =============================================================================================
struct AttnDqSmem {
  static constexpr uint32_t Q_OFF = 0;
  static constexpr uint32_t DO_OFF = Q_OFF + ATT_TILE_BYTES;
  static constexpr uint32_t K_OFF = DO_OFF + ATT_TILE_BYTES;      // 2 stages of 128 keys
  static constexpr uint32_t V_OFF = K_OFF + 2 * ATT_TILE_BYTES;   // 2 stages
  static constexpr uint32_t DS_OFF = V_OFF + 2 * ATT_TILE_BYTES;  // 2 buffers of [128 x 64] (16 KB each)
  static constexpr uint32_t BAR_OFF = DS_OFF + ATT_TILE_BYTES;
  static constexpr uint32_t BYTES = BAR_OFF + 256 + 1024;
};

template <bool kDrop>
__global__ void __launch_bounds__(ATT_THREADS_BWD, 1)
attn_tc_dq_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_dy,
                  const AttnTcParams p, int C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + AttnDqSmem::Q_OFF;
  uint8_t* sDO = smem + AttnDqSmem::DO_OFF;
  uint8_t* sK = smem + AttnDqSmem::K_OFF;
  uint8_t* sV = smem + AttnDqSmem::V_OFF;
  uint8_t* sDS = smem + AttnDqSmem::DS_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnDqSmem::BAR_OFF);
  uint64_t* qdo_full = bars + 0;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* sdp_full = bars + 7;   // [2]
  uint64_t* ds_full = bars + 9;    // [2]
  uint64_t* buf_free = bars + 11;  // [2]  dQ MMA of the sub-tile that used score/dS buffer b has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  int* s_range = reinterpret_cast<int*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_dy);
    mbar_init(qdo_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&sdp_full[i], 1);
      mbar_init(&ds_full[i], ATT_COMPUTE_WARPS);
      mbar_init(&buf_free[i], 1);
    }
    fence_barrier_init();
    s_range[0] = T;
    s_range[1] = 0;
    s_range[2] = 0;
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  __syncthreads();
  if (p.row_lo != nullptr && threadIdx.x < ATT_BM && t0 + static_cast<int>(threadIdx.x) < T) {
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
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  int jb = 0, je = (T + ATT_BN - 1) / ATT_BN;
  if (p.row_lo != nullptr && s_range[2] == 0 && s_range[1] > s_range[0]) {
    jb = s_range[0] / ATT_BN;
    je = (s_range[1] + ATT_BN - 1) / ATT_BN;
  }
  const int n_tiles = je - jb;   // 128-key tiles
  const int n_sub = 2 * n_tiles;  // 64-key sub-tiles
  const int row0 = b * T;
  const int qcol = h * ATT_D, kcol = C + h * ATT_D, vcol = 2 * C + h * ATT_D;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qdo_full, 2 * ATT_TILE_BYTES);
      tma_load_2d(&tm_qkv, qdo_full, sQ, qcol, row0 + t0);
      tma_load_2d(&tm_qkv, qdo_full, sQ + 16384, qcol + 64, row0 + t0);
      tma_load_2d(&tm_dy, qdo_full, sDO, qcol, row0 + t0);
      tma_load_2d(&tm_dy, qdo_full, sDO + 16384, qcol + 64, row0 + t0);
      for (int jj = 0; jj < n_tiles; ++jj) {
        const int st = jj & 1;
        const uint32_t par = (jj >> 1) & 1;
        const int krow = row0 + (jb + jj) * ATT_BN;
        mbar_wait(&kv_empty[st], par ^ 1);
        mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
        tma_load_2d(&tm_qkv, &k_full[st], sK + st * ATT_TILE_BYTES, kcol, krow);
        tma_load_2d(&tm_qkv, &k_full[st], sK + st * ATT_TILE_BYTES + 16384, kcol + 64, krow);
        mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_2d(&tm_qkv, &v_full[st], sV + st * ATT_TILE_BYTES, vcol, krow);
        tma_load_2d(&tm_qkv, &v_full[st], sV + st * ATT_TILE_BYTES + 16384, vcol + 64, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), ds_addr = smem_u32(sDS);
      mbar_wait(qdo_full, 0);
      // scores of sub-tile s: S -> buffer (s&1) columns [0,64), dP -> columns [64,128)
      auto issue_scores = [&](int s) {
        const int jj = s >> 1, hsub = s & 1, st = jj & 1;
        if (hsub == 0) {
          mbar_wait(&k_full[st], (jj >> 1) & 1);
          mbar_wait(&v_full[st], (jj >> 1) & 1);
        }
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * ATT_TILE_BYTES) + hsub * 8192;
        const uint32_t v_addr = smem_u32(sV + st * ATT_TILE_BYTES) + hsub * 8192;
        const uint32_t d = tmem_base + (s & 1) * 128;
        issue_scores_128x64(d, q_addr, 16384, k_addr, 16384);
        issue_scores_128x64(d + 64, do_addr, 16384, v_addr, 16384);
        umma_commit(&sdp_full[s & 1]);
      };
      issue_scores(0);
      for (int s = 0; s < n_sub; ++s) {
        if (s + 1 < n_sub) issue_scores(s + 1);  // buffer (s+1)&1 was drained before dQ(s-1) was issued
        const int jj = s >> 1, hsub = s & 1, st = jj & 1;
        mbar_wait(&ds_full[s & 1], (s >> 1) & 1);
        tc_fence_after();
        const uint32_t k_rows = smem_u32(sK + st * ATT_TILE_BYTES) + hsub * 8192;  // key rows 64*hsub .. of the tile
        issue_grad_128x128x64(tmem_base + 256, ds_addr + (s & 1) * 16384, k_rows, 16384, s > 0);  // dQ += dS K
        umma_commit(&buf_free[s & 1]);
        if (hsub == 1) umma_commit(&kv_empty[st]);
      }
    }
  } else {
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;  // two threads per query row: columns [32*hh, +32) of every 64-key sub-tile
    const int r = q * 32 + lane;
    const int i = t0 + r;
    const bool row_ok = i < T;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
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

    for (int s = 0; s < n_sub; ++s) {
      const int bsel = s & 1;
      const int j0 = (jb + (s >> 1)) * ATT_BN + (s & 1) * 64 + hh * 32;  // first key of this thread's 32 columns
      uint32_t kw = 0xffffffffu;
      if (kDrop && keep_row != nullptr && (j0 >> 5) < p.nw) kw = keep_row[j0 >> 5];
      mbar_wait(&sdp_full[bsel], (s >> 1) & 1);
      tc_fence_after();
      if (s >= 2) mbar_wait(&buf_free[bsel], ((s >> 1) - 1) & 1);  // dS buffer consumed by dQ of sub-tile s-2
      uint32_t sv[32], dv[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + bsel * 128 + hh * 32, sv);
      tmem_ld_32x32(lane_addr + bsel * 128 + 64 + hh * 32, dv);
      tmem_ld_wait();
      float ds[32];  // first P, then dS
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
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = j0 + e;
          ds[e] = (j >= lo && j < hi && row_ok) ? fast_exp2(__uint_as_float(sv[e]) * sc2 - neg) : 0.f;
        }
      }
      if (kDrop) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float dpk = (kw & (1u << keep_bit_pos(e))) ? __uint_as_float(dv[e]) : 0.f;
          ds[e] *= dpk - dlp;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) ds[e] *= __uint_as_float(dv[e]) - dlp;
      }
      uint8_t* dsb = sDS + bsel * 16384;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 w = make_uint4(pack_bf16x2(ds[g * 8 + 0], ds[g * 8 + 1]), pack_bf16x2(ds[g * 8 + 2], ds[g * 8 + 3]),
                                   pack_bf16x2(ds[g * 8 + 4], ds[g * 8 + 5]), pack_bf16x2(ds[g * 8 + 6], ds[g * 8 + 7]));
        *reinterpret_cast<uint4*>(dsb + sw128_row64_off(r, hh * 32 + g * 8)) = w;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[bsel]);
    }
    // dQ epilogue: this thread stores columns [64*hh, 64*hh+64) of its row
    const int last = n_sub - 1;
    mbar_wait(&buf_free[last & 1], (last >> 1) & 1);
    tc_fence_after();
    __nv_bfloat16* drow = p.dq + (static_cast<long long>(row0) + i) * p.ldd + h * ATT_D;
    const float oscale = row_scale * inv_keep;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = hh * 2 + cc;
      uint32_t o[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + 256 + c * 32, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          reinterpret_cast<uint4*>(drow + c * 32)[g] = make_uint4(
              pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * oscale, __uint_as_float(o[g * 8 + 1]) * oscale),
              pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * oscale, __uint_as_float(o[g * 8 + 3]) * oscale),
              pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * oscale, __uint_as_float(o[g * 8 + 5]) * oscale),
              pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * oscale, __uint_as_float(o[g * 8 + 7]) * oscale));
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
// dK / dV kernel
// =============================================================================================
constexpr uint32_t ATT_SUB_BYTES = 64 * 128 * 2;  // one [64 rows x 128] bf16 tile = two 8 KB swizzle sub-tiles
// per 64-query sub-tile: 64 x float2 {-(max+lsum)*log2e, delta*(1-p)} + 64 x int2 {lo,hi} + 64 x float2 {max, lsum*log2e}
// + 64 x float live + 4 x 64 keep words + 2 x int4 chunk ranges, rounded up
constexpr uint32_t ATT_COL_STRIDE = 3072;
constexpr int ATT_QDO_STAGES = 3;

struct AttnDkvSmem {
  static constexpr uint32_t K_OFF = 0;
  static constexpr uint32_t V_OFF = K_OFF + ATT_TILE_BYTES;
  // 3 TMA stages of 64 queries (16 KB each for Q and for dO): with 2 stages the ~1.5 us TMA round trip was exposed
  // every other sub-tile (profiles/r01_attn_bwd_v4.details.txt: IPC 0.93, every pipe < 25 %)
  static constexpr uint32_t Q_OFF = V_OFF + ATT_TILE_BYTES;
  static constexpr uint32_t DO_OFF = Q_OFF + ATT_QDO_STAGES * ATT_SUB_BYTES;
  static constexpr uint32_t PT_OFF = DO_OFF + ATT_QDO_STAGES * ATT_SUB_BYTES;  // [128 keys x 64 q] bf16, single buffer
  static constexpr uint32_t DST_OFF = PT_OFF + 16384;
  // per-query (column) parameters, double buffered (layout: see ATT_COL_STRIDE users in the kernel)
  static constexpr uint32_t COL_OFF = DST_OFF + 16384;
  static constexpr uint32_t BAR_OFF = COL_OFF + 2 * ATT_COL_STRIDE;
  static constexpr uint32_t BYTES = BAR_OFF + 256 + 1024;
};

template <bool kDrop>
__global__ void __launch_bounds__(ATT_THREADS_BWD, 1)
attn_tc_dkv_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_q64,
                   const __grid_constant__ CUtensorMap tm_dy64, const AttnTcParams p, int C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + AttnDkvSmem::K_OFF;
  uint8_t* sV = smem + AttnDkvSmem::V_OFF;
  uint8_t* sQ = smem + AttnDkvSmem::Q_OFF;
  uint8_t* sDO = smem + AttnDkvSmem::DO_OFF;
  uint8_t* sPT = smem + AttnDkvSmem::PT_OFF;
  uint8_t* sDST = smem + AttnDkvSmem::DST_OFF;
  uint8_t* sCol = smem + AttnDkvSmem::COL_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnDkvSmem::BAR_OFF);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [3]
  uint64_t* qdo_free = bars + 4;   // [3] dV/dK MMAs that read Q/dO stage s completed
  uint64_t* sdp_full = bars + 7;   // [2]
  uint64_t* pds_full = bars + 9;   // [2]
  uint64_t* grad_done = bars + 11; // dV/dK MMAs of a sub-tile completed: P^T / dS^T smem may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  unsigned int* s_rel = reinterpret_cast<unsigned int*>(bars + 13);  // relevance bits of the 64-query sub-tiles (<=128)

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
    mbar_init(grad_done, 1);
    fence_barrier_init();
    s_rel[0] = s_rel[1] = s_rel[2] = s_rel[3] = 0;
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  __syncthreads();
  // which 64-query sub-tiles can see this key tile at all
  if (p.row_lo != nullptr) {
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
      const int lo = p.row_lo[static_cast<long long>(b) * T + i], hi = p.row_hi[static_cast<long long>(b) * T + i];
      const bool rel = (lo >= hi) || (lo < j0 + ATT_BN && hi > j0);
      if (rel) atomicOr(&s_rel[(i >> 6) >> 5], 1u << ((i >> 6) & 31));
    }
  } else if (threadIdx.x < 4) {
    s_rel[threadIdx.x] = 0xffffffffu;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const unsigned int rel0 = s_rel[0], rel1 = s_rel[1], rel2 = s_rel[2], rel3 = s_rel[3];
  auto relevant = [&](int it) -> bool {
    const unsigned int w = (it >> 5) == 0 ? rel0 : (it >> 5) == 1 ? rel1 : (it >> 5) == 2 ? rel2 : rel3;
    return (w >> (it & 31)) & 1u;
  };
  const int row0 = b * T;
  const int qcol = h * ATT_D, kcol = C + h * ATT_D, vcol = 2 * C + h * ATT_D;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * ATT_TILE_BYTES);
      tma_load_2d(&tm_qkv, kv_full, sK, kcol, row0 + j0);
      tma_load_2d(&tm_qkv, kv_full, sK + 16384, kcol + 64, row0 + j0);
      tma_load_2d(&tm_qkv, kv_full, sV, vcol, row0 + j0);
      tma_load_2d(&tm_qkv, kv_full, sV + 16384, vcol + 64, row0 + j0);
      int n = 0, st = 0;
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
        ++n;
        if (++st == ATT_QDO_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      int n_total = 0;
      for (int it = 0; it < nq; ++it) n_total += relevant(it) ? 1 : 0;
      mbar_wait(kv_full, 0);
      auto issue_scores = [&](int n) {  // Q/dO stage n % 3, TMEM score buffer n & 1
        const int st = n % ATT_QDO_STAGES, tb = n & 1;
        mbar_wait(&qdo_full[st], (n / ATT_QDO_STAGES) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + tb * 128;
        issue_scores_128x64(d, k_addr, 16384, smem_u32(sQ + st * ATT_SUB_BYTES), 8192);        // S^T  = K Q^T
        issue_scores_128x64(d + 64, v_addr, 16384, smem_u32(sDO + st * ATT_SUB_BYTES), 8192);  // dP^T = V dO^T
        umma_commit(&sdp_full[tb]);
      };
      if (n_total > 0) issue_scores(0);
      for (int n = 0; n < n_total; ++n) {
        if (n + 1 < n_total) issue_scores(n + 1);
        const int st = n % ATT_QDO_STAGES, tb = n & 1;
        mbar_wait(&pds_full[tb], (n >> 1) & 1);
        tc_fence_after();
        issue_grad_128x128x64(tmem_base + 256, smem_u32(sPT), smem_u32(sDO + st * ATT_SUB_BYTES), 8192, n > 0);   // dV += P^T dO
        issue_grad_128x128x64(tmem_base + 384, smem_u32(sDST), smem_u32(sQ + st * ATT_SUB_BYTES), 8192, n > 0);  // dK += dS^T Q
        umma_commit(&qdo_free[st]);
        umma_commit(grad_done);
      }
    }
  } else {
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;  // two threads per key row: query columns [32*hh, +32) of every 64-query sub-tile
    const int r = q * 32 + lane;     // key row within the tile
    const int j = j0 + r;
    const bool key_ok = j < T;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const long long bh = static_cast<long long>(b) * p.H + h;
    const float inv_keep = kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const float keep_frac = kDrop ? 1.0f - p.drop_p : 1.0f;
    const float sc2 = p.scale * LOG2E;
    const uint32_t mybit = 1u << keep_bit_pos(lane);  // this key's bit inside the keep word of its 32-key group
    const int ct = threadIdx.x - 64;  // 0..255 among the compute threads
    // Per-query parameters are software-pipelined: the global loads for the NEXT relevant sub-tile are issued before
    // the math of the current one and written to the other smem buffer afterwards, so their latency never sits on
    // the critical path.
    struct QParams { int lo, hi; float off, ls2, dl, live; uint32_t kw; };
    auto load_params = [&](int it) -> QParams {
      QParams z;
      z.lo = 0; z.hi = 0; z.off = 0.f; z.ls2 = 0.f; z.dl = 0.f; z.live = 1.f;  // query beyond T: contributes nothing
      z.kw = 0xffffffffu;
      if (ct < 64) {
        const int i = it * 64 + ct;
        if (i < T) {
          z.lo = 0; z.hi = T;
          if (p.row_lo != nullptr) {
            z.lo = p.row_lo[static_cast<long long>(b) * T + i];
            z.hi = p.row_hi[static_cast<long long>(b) * T + i];
            if (z.lo >= z.hi) { z.lo = 0; z.hi = T; z.live = 0.f; }  // fully-masked row: uniform P, no dS
          }
          z.off = p.lse[2 * (bh * T + i)];
          z.ls2 = p.lse[2 * (bh * T + i) + 1] * LOG2E;
          z.dl = p.delta[bh * T + i] * keep_frac;
        }
      }
      if (kDrop) {  // thread ct: query ct >> 2, key word ct & 3 of this 128-key tile
        const int i = it * 64 + (ct >> 2);
        const int w = (j0 >> 5) + (ct & 3);
        if (i < T && w < p.nw) z.kw = p.keep[(bh * T + i) * p.nw + w];
      }
      return z;
    };
    auto store_params = [&](int buf, const QParams& z) {
      uint8_t* base = sCol + buf * ATT_COL_STRIDE;
      float2* c_nd = reinterpret_cast<float2*>(base);          // {-(max + lsum) * log2e, delta * (1-p)}
      int2* c_lh = reinterpret_cast<int2*>(base + 512);        // {lo, hi}
      float2* c_x = reinterpret_cast<float2*>(base + 1024);    // {max (natural), lsum * log2e} (dense path)
      float* c_live = reinterpret_cast<float*>(base + 1536);   // 0 for fully-masked queries
      uint32_t* c_keep = reinterpret_cast<uint32_t*>(base + 1792);  // [4 key words][64 queries]
      int4* c_rng = reinterpret_cast<int4*>(base + 2816);      // per 32-query chunk: {max lo, min hi}
      if (ct < 64) {
        c_nd[ct] = make_float2(-(z.off * LOG2E + z.ls2), z.dl);
        c_lh[ct] = make_int2(z.lo, z.hi);
        c_x[ct] = make_float2(z.off, z.ls2);
        c_live[ct] = z.live;
        // threads 0..63 are exactly two warps, one per 32-query chunk: interval common to the whole chunk
        int mlo = (z.live == 0.f) ? 0x7fffffff : z.lo, mhi = z.hi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mlo = max(mlo, __shfl_xor_sync(0xffffffffu, mlo, o));
          mhi = min(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
        }
        if ((ct & 31) == 0) c_rng[ct >> 5] = make_int4(mlo, mhi, 0, 0);
      }
      if (kDrop) c_keep[(ct & 3) * 64 + (ct >> 2)] = z.kw;
    };
    auto next_relevant = [&](int it) -> int {
      ++it;
      while (it < nq && !relevant(it)) ++it;
      return it;
    };

    int n = 0;
    int it = next_relevant(-1);
    if (it < nq) {
      const QParams z0 = load_params(it);
      store_params(0, z0);
    }
    compute_bar_sync256();
    while (it < nq) {
      const int st = n & 1;
      const int i0 = it * 64;
      const int nx = next_relevant(it);
      QParams zn;
      if (nx < nq) zn = load_params(nx);  // in flight during the math below
      const uint8_t* base = sCol + st * ATT_COL_STRIDE;
      const float4* nd4 = reinterpret_cast<const float4*>(base) + hh * 16;               // two queries per float4
      const int2* c_lh = reinterpret_cast<const int2*>(base + 512) + hh * 32;
      const float2* c_x = reinterpret_cast<const float2*>(base + 1024) + hh * 32;
      const float* c_live = reinterpret_cast<const float*>(base + 1536) + hh * 32;
      const uint4* kp4 = reinterpret_cast<const uint4*>(base + 1792 + (q * 64 + hh * 32) * 4);  // four queries per uint4
      const int4 rng = reinterpret_cast<const int4*>(base + 2816)[hh];
      mbar_wait(&sdp_full[st], (n >> 1) & 1);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + st * 128 + hh * 32, sv);
      tmem_ld_32x32(lane_addr + st * 128 + 64 + hh * 32, dv);
      tmem_ld_wait();
      float pt[32], dst[32];
      // every key of this warp visible to every (live) query of the chunk: warp-uniform test
      const bool interior = (p.mask == nullptr) && (j0 + q * 32 >= rng.x) && (j0 + q * 32 + 32 <= rng.y) &&
                            (j0 + q * 32 + 32 <= T);
      if (interior) {
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          uint4 kk = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (kDrop) kk = kp4[e4];
          const uint32_t kws[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float4 nd = nd4[e4 * 2 + h2];
            const float nneg[2] = {nd.x, nd.z}, dlp[2] = {nd.y, nd.w};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int e = e4 * 4 + h2 * 2 + u;
              const float pr = fast_exp2(fmaf(__uint_as_float(sv[e]), sc2, nneg[u]));
              const bool kb = !kDrop || (kws[h2 * 2 + u] & mybit);
              pt[e] = kb ? pr : 0.f;
              dst[e] = pr * ((kb ? __uint_as_float(dv[e]) : 0.f) - dlp[u]);
            }
          }
        }
      } else if (p.mask == nullptr) {
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          uint4 kk = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (kDrop) kk = kp4[e4];
          const uint32_t kws[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e4 * 4 + u;
            const float2 nd = reinterpret_cast<const float2*>(nd4)[e];
            const int2 lh = c_lh[e];
            const float live = c_live[e];
            const bool vis = key_ok && j >= lh.x && j < lh.y;
            const float pr = vis ? fast_exp2(fmaf(__uint_as_float(sv[e]) * live, sc2, nd.x)) : 0.f;
            const bool kb = !kDrop || (kws[u] & mybit);
            pt[e] = kb ? pr : 0.f;
            dst[e] = pr * ((kb ? __uint_as_float(dv[e]) : 0.f) - nd.y) * live;
          }
        }
      } else {  // dense additive bias
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          uint4 kk = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (kDrop) kk = kp4[e4];
          const uint32_t kws[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e4 * 4 + u;
            const int i = i0 + hh * 32 + e;
            const float2 nd = reinterpret_cast<const float2*>(nd4)[e];
            const float2 cx = c_x[e];
            const bool vis = key_ok && i < T;
            const float bias =
                vis ? __bfloat162float(p.mask[b * p.msb + h * p.msh + static_cast<long long>(i) * p.msq + j]) : 0.f;
            const float sp = __fadd_rn(__fmul_rn(__uint_as_float(sv[e]), p.scale), bias);
            const float pr = vis ? fast_exp2((sp - cx.x) * LOG2E - cx.y) : 0.f;
            const bool kb = !kDrop || (kws[u] & mybit);
            pt[e] = kb ? pr : 0.f;
            dst[e] = pr * ((kb ? __uint_as_float(dv[e]) : 0.f) - nd.y);
          }
        }
      }
      // the single P^T / dS^T smem tile is free once the dV/dK MMAs of the previous sub-tile have completed
      if (n >= 1) mbar_wait(grad_done, (n - 1) & 1);
      uint8_t* ptb = sPT;
      uint8_t* dsb = sDST;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t o = sw128_row64_off(r, hh * 32 + g * 8);
        *reinterpret_cast<uint4*>(ptb + o) =
            make_uint4(pack_bf16x2(pt[g * 8 + 0], pt[g * 8 + 1]), pack_bf16x2(pt[g * 8 + 2], pt[g * 8 + 3]),
                       pack_bf16x2(pt[g * 8 + 4], pt[g * 8 + 5]), pack_bf16x2(pt[g * 8 + 6], pt[g * 8 + 7]));
        *reinterpret_cast<uint4*>(dsb + o) =
            make_uint4(pack_bf16x2(dst[g * 8 + 0], dst[g * 8 + 1]), pack_bf16x2(dst[g * 8 + 2], dst[g * 8 + 3]),
                       pack_bf16x2(dst[g * 8 + 4], dst[g * 8 + 5]), pack_bf16x2(dst[g * 8 + 6], dst[g * 8 + 7]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[st]);
      // parameters of the next sub-tile into the other buffer (last read during sub-tile n-1, before the barrier
      // that ended that iteration)
      if (nx < nq) store_params(st ^ 1, zn);
      compute_bar_sync256();
      it = nx;
      ++n;
    }
    // epilogue: hh = 0 stores dV (TMEM columns 256..383) / (1-p), hh = 1 stores dK (384..511) * scale / (1-p)
    __nv_bfloat16* dvrow = p.dv + (static_cast<long long>(row0) + j) * p.ldd + h * ATT_D;
    __nv_bfloat16* dkrow = p.dk + (static_cast<long long>(row0) + j) * p.ldd + h * ATT_D;
    const float oscale = hh == 0 ? inv_keep : inv_keep * p.scale;
    if (n > 0) {
      mbar_wait(grad_done, (n - 1) & 1);
      tc_fence_after();
    }
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
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
#pragma unroll
        for (int g = 0; g < 4; ++g)
          reinterpret_cast<uint4*>(dstp)[g] = make_uint4(
              pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * oscale, __uint_as_float(o[g * 8 + 1]) * oscale),
              pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * oscale, __uint_as_float(o[g * 8 + 3]) * oscale),
              pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * oscale, __uint_as_float(o[g * 8 + 5]) * oscale),
              pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * oscale, __uint_as_float(o[g * 8 + 7]) * oscale));
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

}  // namespace obt

using namespace obt;

extern "C" int obt_attn_tc_bwd(const void* qkv, long long ld, const void* mask, long long msb, long long msh,
                               long long msq, const int* row_lo, const int* row_hi, const void* y, long long ldy,
                               const void* dy, long long lddy, const float* lse, float* delta, void* dqkv, long long ldd,
                               int B, int H, int T, int d, float scale, float drop_p, const unsigned int* keep,
                               cudaStream_t stream) {
  OBT_REQUIRE(qkv && y && dy && lse && delta && dqkv, "obt_attn_tc_bwd: null pointer");
  OBT_REQUIRE(d == ATT_D, "obt_attn_tc_bwd: head_dim=%d, the tensor-core kernel is specialised for 128", d);
  OBT_REQUIRE(B > 0 && H > 0 && T > 0, "obt_attn_tc_bwd: empty problem");
  OBT_REQUIRE(T <= 128 * 64, "obt_attn_tc_bwd: T=%d exceeds %d", T, 128 * 64);
  OBT_REQUIRE(ld % 8 == 0 && ldy % 8 == 0 && lddy % 8 == 0 && ldd % 8 == 0, "obt_attn_tc_bwd: pitches must be multiples of 8");
  OBT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "obt_attn_tc_bwd: dropout p=%f", drop_p);
  OBT_REQUIRE(drop_p == 0.f || keep != nullptr, "obt_attn_tc_bwd: dropout needs the keep mask (obt_attn_keep_mask)");
  const int C = H * d;
  const long long M = static_cast<long long>(B) * T;
  CUtensorMap tm_qkv, tm_dy, tm_q64, tm_dy64;
  int rc = get_tensor_map_2d(&tm_qkv, qkv, static_cast<uint64_t>(3) * C, static_cast<uint64_t>(M),
                             static_cast<uint64_t>(ld), 64, 128);
  if (rc) return rc;
  rc = get_tensor_map_2d(&tm_dy, dy, static_cast<uint64_t>(C), static_cast<uint64_t>(M), static_cast<uint64_t>(lddy), 64,
                         128);
  if (rc) return rc;
  rc = get_tensor_map_2d(&tm_q64, qkv, static_cast<uint64_t>(3) * C, static_cast<uint64_t>(M), static_cast<uint64_t>(ld),
                         64, 64);
  if (rc) return rc;
  rc = get_tensor_map_2d(&tm_dy64, dy, static_cast<uint64_t>(C), static_cast<uint64_t>(M), static_cast<uint64_t>(lddy), 64,
                         64);
  if (rc) return rc;
  {
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
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.drop_p = drop_p;
  p.keep = keep;
  p.nw = keep_words(T);
  p.dq = static_cast<__nv_bfloat16*>(dqkv);
  p.dk = p.dq + C;
  p.dv = p.dq + 2 * C;
  p.ldd = ldd;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_tc_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnDqSmem::BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(attn_tc_dkv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnDkvSmem::BYTES);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(attn_tc_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnDqSmem::BYTES);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(attn_tc_dkv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnDkvSmem::BYTES);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      set_last_error("obt_attn_tc_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      return OBT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((T + ATT_BM - 1) / ATT_BM, H, B);
  if (drop_p > 0.f)
    attn_tc_dq_kernel<true><<<grid, ATT_THREADS_BWD, AttnDqSmem::BYTES, stream>>>(tm_qkv, tm_dy, p, C);
  else
    attn_tc_dq_kernel<false><<<grid, ATT_THREADS_BWD, AttnDqSmem::BYTES, stream>>>(tm_qkv, tm_dy, p, C);
  rc = check_launch("attn_tc_dq");
  if (rc) return rc;
  if (drop_p > 0.f)
    attn_tc_dkv_kernel<true><<<grid, ATT_THREADS_BWD, AttnDkvSmem::BYTES, stream>>>(tm_qkv, tm_q64, tm_dy64, p, C);
  else
    attn_tc_dkv_kernel<false><<<grid, ATT_THREADS_BWD, AttnDkvSmem::BYTES, stream>>>(tm_qkv, tm_q64, tm_dy64, p, C);
  return check_launch("attn_tc_dkv");
}
