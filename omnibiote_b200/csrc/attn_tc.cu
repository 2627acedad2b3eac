// Attention forward on tcgen05 tensor cores (head_dim = 128), flash-style: S = Q K^T and O += P V are
// tcgen05.mma tiles with S/O accumulators in TMEM, K/V tiles streamed by TMA, online softmax in registers.
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask, dropout_p, scale=8/n_embd) and the head merge
// (training/model.py:111-148): q,k,v are read in place from the fused qkv buffer [M, 3C] (no transposes), y is
// written head-major into [M, C].
//
// One CTA per (128-query tile, head, batch); 10 warps:
//   warp 0     TMA producer: Q once, then K_j / V_j tiles through 2-stage rings
//   warp 1     MMA issuer  : S_{j+1} = Q K_{j+1}^T is issued before O += P_j V_j so it overlaps softmax(j)
//   warps 2-9  softmax     : TWO threads per query row (TMEM lane): warps 2-5 own key columns 0..63 of every S tile
//                            and O columns 0..63, warps 6-9 the other halves; the pair exchanges its half-row maxima
//                            through smem (one 256-thread named barrier per tile). S is read with tcgen05.ld, P is
//                            written as bf16 into a 128B-swizzled smem tile (A operand of the PV MMA), O is rescaled
//                            in TMEM only when the running max moved.
// TMEM: S ping-pong (2 x 128 columns) + O (128 columns).
// Mask modes: none | per-row visible key interval [lo,hi) (document / padding masks; KV tiles outside the union of
// the tile's intervals are skipped) | dense additive bf16 bias (arbitrary masks). Fully-masked rows (finite -1e9 on
// every key) attend uniformly to all T keys like the reference.
#include "attn_tc_common.cuh"

namespace obt {

struct AttnFwdSmem {
  static constexpr uint32_t Q_OFF = 0;
  static constexpr uint32_t K_OFF = Q_OFF + ATT_TILE_BYTES;          // 2 stages
  static constexpr uint32_t V_OFF = K_OFF + 2 * ATT_TILE_BYTES;      // 2 stages
  static constexpr uint32_t P_OFF = V_OFF + 2 * ATT_TILE_BYTES;
  static constexpr uint32_t X_OFF = P_OFF + ATT_TILE_BYTES;          // half-row exchange: float [2 buffers][2 halves][128]
  static constexpr uint32_t BAR_OFF = X_OFF + 2 * 2 * 128 * 4;
  static constexpr uint32_t BYTES = BAR_OFF + 256 + 1024;
};

constexpr int ATT_THREADS = 64 + 256;  // TMA warp, MMA warp, 8 compute warps

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnTcParams p, int C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + AttnFwdSmem::Q_OFF;
  uint8_t* sK = smem + AttnFwdSmem::K_OFF;
  uint8_t* sV = smem + AttnFwdSmem::V_OFF;
  uint8_t* sP = smem + AttnFwdSmem::P_OFF;
  float* sX = reinterpret_cast<float*>(smem + AttnFwdSmem::X_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnFwdSmem::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* pv_done = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  int* s_range = reinterpret_cast<int*>(bars + 17);  // [0] = min lo, [1] = max hi, [2] = any fully-masked row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
    }
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    fence_barrier_init();
    s_range[0] = T;
    s_range[1] = 0;
    s_range[2] = 0;
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  __syncthreads();
  // KV tile range covering the union of this tile's visible intervals
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
  const int n_tiles = je - jb;

  const int row0 = b * T;                 // first token row of this sequence in the qkv buffer
  const int qcol = h * ATT_D, kcol = C + h * ATT_D, vcol = 2 * C + h * ATT_D;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(&tm_qkv, q_full, sQ, qcol, row0 + t0);
      tma_load_2d(&tm_qkv, q_full, sQ + 16384, qcol + 64, row0 + t0);
      for (int jj = 0; jj < n_tiles; ++jj) {
        const int st = jj & 1;
        const uint32_t par = (jj >> 1) & 1;
        const int krow = row0 + (jb + jj) * ATT_BN;
        mbar_wait(&k_empty[st], par ^ 1);
        mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
        tma_load_2d(&tm_qkv, &k_full[st], sK + st * ATT_TILE_BYTES, kcol, krow);
        tma_load_2d(&tm_qkv, &k_full[st], sK + st * ATT_TILE_BYTES + 16384, kcol + 64, krow);
        mbar_wait(&v_empty[st], par ^ 1);
        mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_2d(&tm_qkv, &v_full[st], sV + st * ATT_TILE_BYTES, vcol, krow);
        tma_load_2d(&tm_qkv, &v_full[st], sV + st * ATT_TILE_BYTES + 16384, vcol + 64, krow);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      const uint32_t o_tmem = tmem_base + 256;
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_128x128x128<false>(tmem_base, q_addr, smem_u32(sK), false);
      umma_commit(&s_full[0]);
      umma_commit(&k_empty[0]);
      for (int jj = 0; jj < n_tiles; ++jj) {
        if (jj + 1 < n_tiles) {
          const int st = (jj + 1) & 1;
          const uint32_t par = ((jj + 1) >> 1) & 1;
          mbar_wait(&k_full[st], par);
          mbar_wait(&s_empty[st], par ^ 1);
          tc_fence_after();
          issue_128x128x128<false>(tmem_base + st * 128, q_addr, smem_u32(sK + st * ATT_TILE_BYTES), false);
          umma_commit(&s_full[st]);
          umma_commit(&k_empty[st]);
        }
        const int st = jj & 1;
        mbar_wait(p_full, jj & 1);
        mbar_wait(&v_full[st], (jj >> 1) & 1);
        tc_fence_after();
        issue_128x128x128<true>(o_tmem, p_addr, smem_u32(sV + st * ATT_TILE_BYTES), jj > 0);
        umma_commit(pv_done);
        umma_commit(&v_empty[st]);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (two threads per query row) =====================
    const int q = warp & 3;              // TMEM lane quadrant
    const int half = (warp - 2) >> 2;    // which 64 key columns of S / 64 columns of O this thread owns
    const int r = q * 32 + lane;         // row within the tile == TMEM lane
    const int i = t0 + r;                // query position
    const bool row_ok = i < T;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int lo = 0, hi = T;
    float row_scale = p.scale * LOG2E;
    if (p.row_lo != nullptr && row_ok) {
      lo = p.row_lo[static_cast<long long>(b) * T + i];
      hi = p.row_hi[static_cast<long long>(b) * T + i];
      if (lo >= hi) {  // fully-masked row: every key carries the same finite bias -> uniform over all T keys
        lo = 0;
        hi = T;
        row_scale = 0.f;
      }
    }
    const __nv_bfloat16* mrow =
        (p.mask != nullptr && row_ok) ? p.mask + b * p.msb + h * p.msh + static_cast<long long>(i) * p.msq : nullptr;
    const bool use_drop = p.drop_p > 0.f;
    // keep <=> u01 >= p with u01 = (w >> 8) * 2^-24  <=>  (w >> 8) >= ceil(p * 2^24): same decision as every other
    // dropout site of the library, as one integer compare
    const uint32_t drop_thr = static_cast<uint32_t>(ceilf(p.drop_p * 16777216.0f));
    const float drop_scale = 1.0f / (1.0f - p.drop_p);
    float m_run = -INFINITY, l_run = 0.f;  // running max (log2 domain, whole row) and exp-sum (this half only)

    for (int jj = 0; jj < n_tiles; ++jj) {
      const int st = jj & 1;
      const int j0 = (jb + jj) * ATT_BN;
      mbar_wait(&s_full[st], (jj >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = lane_addr + st * 128;
      // ---- pass 1: max of the scaled + biased scores of this thread's 64 columns
      float tmax = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(s_addr + c * 32, v);
        tmem_ld_wait();
        if (p.mask != nullptr) {  // warp-uniform (rows beyond T have mrow == nullptr and read no bias)
          const uint4* mp = reinterpret_cast<const uint4*>((mrow ? mrow : p.mask) + j0 + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int jbase = j0 + c * 32 + g * 8;
            uint4 mu = make_uint4(0, 0, 0, 0);
            if (jbase + 8 <= T && mrow != nullptr) mu = mp[g];
            const uint32_t mw[4] = {mu.x, mu.y, mu.z, mu.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float bias = (e & 1) ? bf16_hi(mw[e >> 1]) : bf16_lo(mw[e >> 1]);
              float s = __fadd_rn(__fmul_rn(__uint_as_float(v[g * 8 + e]), p.scale), bias) * LOG2E;
              if (jbase + e >= T) s = -INFINITY;
              tmax = fmaxf(tmax, s);
            }
          }
        } else if (__all_sync(0xffffffffu, (j0 + c * 32 >= lo) && (j0 + c * 32 + 32 <= hi))) {
          // every key of this chunk is visible to every row of the warp (interior of a document): no per-element
          // interval tests; the (non-negative) scale is applied once to the maximum
          float mx = __uint_as_float(v[0]);
#pragma unroll
          for (int e = 1; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
          tmax = fmaxf(tmax, mx * row_scale);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int j = j0 + c * 32 + e;
            const float s = (j >= lo && j < hi) ? __uint_as_float(v[e]) * row_scale : -INFINITY;
            tmax = fmaxf(tmax, s);
          }
        }
      }
      // exchange the half-row maxima with the partner thread (same row, other column half)
      float* xbuf = sX + (jj & 1) * 256;
      xbuf[half * 128 + r] = tmax;
      compute_bar_sync256();
      tmax = fmaxf(tmax, xbuf[(half ^ 1) * 128 + r]);
      const float m_new = fmaxf(m_run, tmax);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(m_run - m_use);  // m_run = -inf -> 0
      // ---- wait for O += P_{j-1} V_{j-1}; rescale this thread's 64 O columns when the running max moved
      if (jj > 0) {
        mbar_wait(pv_done, (jj - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int c = half * 2 + cc;
            uint32_t o[32];
            __syncwarp();
            tmem_ld_32x32(lane_addr + 256 + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st_32x32(lane_addr + 256 + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      // ---- pass 2: P = exp2(s - m), partial row sum, bf16 P into swizzled smem
      float lsum = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(s_addr + c * 32, v);
        tmem_ld_wait();
        float pr[32];
        if (p.mask != nullptr) {  // warp-uniform (rows beyond T have mrow == nullptr and read no bias)
          const uint4* mp = reinterpret_cast<const uint4*>((mrow ? mrow : p.mask) + j0 + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int jbase = j0 + c * 32 + g * 8;
            uint4 mu = make_uint4(0, 0, 0, 0);
            if (jbase + 8 <= T && mrow != nullptr) mu = mp[g];
            const uint32_t mw[4] = {mu.x, mu.y, mu.z, mu.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float bias = (e & 1) ? bf16_hi(mw[e >> 1]) : bf16_lo(mw[e >> 1]);
              const float s = __fadd_rn(__fmul_rn(__uint_as_float(v[g * 8 + e]), p.scale), bias) * LOG2E;
              pr[g * 8 + e] = (jbase + e < T) ? fast_exp2(s - m_use) : 0.f;
            }
          }
        } else if (__all_sync(0xffffffffu, (j0 + c * 32 >= lo) && (j0 + c * 32 + 32 <= hi))) {
          const float nm = -m_use;
#pragma unroll
          for (int e = 0; e < 32; ++e) pr[e] = fast_exp2(fmaf(__uint_as_float(v[e]), row_scale, nm));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int j = j0 + c * 32 + e;
            pr[e] = (j >= lo && j < hi) ? fast_exp2(__uint_as_float(v[e]) * row_scale - m_use) : 0.f;
          }
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) lsum += pr[e];
        if (use_drop && row_ok) {
          // one RNG call per 4 consecutive keys; element index e = ((b*H+h)*T + i)*T + j as in the generic kernel
          // (T % 4 == 0 so that groups of 4 keys never straddle rows; enforced by the host wrapper)
          const unsigned long long e0 =
              ((static_cast<unsigned long long>(b) * p.H + h) * T + i) * T + static_cast<unsigned long long>(j0 + c * 32);
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const uint4 rnd = rand4x32(p.seed, (e0 >> 2) + g4, p.offset);
            pr[g4 * 4 + 0] = ((rnd.x >> 8) >= drop_thr) ? pr[g4 * 4 + 0] * drop_scale : 0.f;
            pr[g4 * 4 + 1] = ((rnd.y >> 8) >= drop_thr) ? pr[g4 * 4 + 1] * drop_scale : 0.f;
            pr[g4 * 4 + 2] = ((rnd.z >> 8) >= drop_thr) ? pr[g4 * 4 + 2] * drop_scale : 0.f;
            pr[g4 * 4 + 3] = ((rnd.w >> 8) >= drop_thr) ? pr[g4 * 4 + 3] * drop_scale : 0.f;
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 w = make_uint4(pack_bf16x2(pr[g * 8 + 0], pr[g * 8 + 1]), pack_bf16x2(pr[g * 8 + 2], pr[g * 8 + 3]),
                                     pack_bf16x2(pr[g * 8 + 4], pr[g * 8 + 5]), pack_bf16x2(pr[g * 8 + 6], pr[g * 8 + 7]));
          *reinterpret_cast<uint4*>(sP + sw128_chunk_off(r, c * 32 + g * 8)) = w;
        }
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      // publish P (generic-proxy smem writes -> async proxy) and release the S buffer
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(p_full);
        mbar_arrive(&s_empty[st]);
      }
    }
    // ---- epilogue: total row sum = both halves; O / l -> y (this thread's 64 columns), (max, log-sum) -> lse
    float* xbuf = sX + (n_tiles & 1) * 256;
    xbuf[half * 128 + r] = l_run;
    compute_bar_sync256();
    const float l_tot = l_run + xbuf[(half ^ 1) * 128 + r];
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_tot;
    __nv_bfloat16* yrow = p.y + (static_cast<long long>(row0) + i) * p.ldy + h * ATT_D;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;
      uint32_t o[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + 256 + c * 32, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[g * 8 + e]) * inv_l;
          reinterpret_cast<uint4*>(yrow + c * 32)[g] =
              make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
      }
    }
    if (row_ok && half == 0 && p.lse != nullptr) {
      float* l = p.lse + 2 * ((static_cast<long long>(b) * p.H + h) * T + i);
      l[0] = m_run / LOG2E;
      l[1] = logf(l_tot);
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

extern "C" int obt_attn_tc_fwd(const void* qkv, long long ld, const void* mask, long long msb, long long msh,
                               long long msq, const int* row_lo, const int* row_hi, void* y, long long ldy, float* lse,
                               int B, int H, int T, int d, float scale, float drop_p, unsigned long long seed,
                               unsigned long long offset, cudaStream_t stream) {
  OBT_REQUIRE(qkv && y && lse, "obt_attn_tc_fwd: null pointer");
  OBT_REQUIRE(d == ATT_D, "obt_attn_tc_fwd: head_dim=%d, the tensor-core kernel is specialised for 128", d);
  OBT_REQUIRE(B > 0 && H > 0 && T > 0, "obt_attn_tc_fwd: empty problem");
  OBT_REQUIRE(ld % 8 == 0 && ldy % 8 == 0, "obt_attn_tc_fwd: pitches must be multiples of 8");
  OBT_REQUIRE(mask == nullptr || (msq % 8 == 0 && msb % 8 == 0 && msh % 8 == 0 && T % 8 == 0 &&
                                   (reinterpret_cast<uintptr_t>(mask) & 15) == 0),
              "obt_attn_tc_fwd: dense mask needs 16-byte aligned rows (T %% 8 == 0)");
  OBT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "obt_attn_tc_fwd: dropout p=%f", drop_p);
  OBT_REQUIRE(drop_p == 0.f || T % 4 == 0, "obt_attn_tc_fwd: attention dropout needs T %% 4 == 0 (T=%d)", T);
  const int C = H * d;
  CUtensorMap tm;
  int rc = get_tensor_map_2d(&tm, qkv, static_cast<uint64_t>(3) * C, static_cast<uint64_t>(B) * T,
                             static_cast<uint64_t>(ld), 64, 128);
  if (rc) return rc;
  AttnTcParams p = {};
  p.B = B; p.H = H; p.T = T;
  p.scale = scale;
  p.mask = static_cast<const __nv_bfloat16*>(mask);
  p.msb = msb; p.msh = msh; p.msq = msq;
  p.row_lo = mask ? nullptr : row_lo;
  p.row_hi = mask ? nullptr : row_hi;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.ldy = ldy;
  p.lse = lse;
  p.drop_p = drop_p; p.seed = seed; p.offset = offset;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         AttnFwdSmem::BYTES);
    if (e != cudaSuccess) {
      set_last_error("obt_attn_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return OBT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((T + ATT_BM - 1) / ATT_BM, H, B);
  attn_tc_fwd_kernel<<<grid, ATT_THREADS, AttnFwdSmem::BYTES, stream>>>(tm, p, C);
  return check_launch("attn_tc_fwd");
}
