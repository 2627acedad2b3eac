// Attention forward on tcgen05 tensor cores (head_dim = 128), flash-style: S = Q K^T and O += P V are
// tcgen05.mma tiles with S/P/O in TMEM, K/V tiles streamed by TMA, online softmax in registers.
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask, dropout_p, scale=8/n_embd) and the head merge
// (training/model.py:111-148): q,k,v are read in place from the fused qkv buffer [M, 3C] (no transposes), y is
// written head-major into [M, C].
//
// v5 layout (profiles/r01_attn_v5_fwd.source.txt showed v4 issue-bound: 21 instructions per score element, half of
// them dropout hashing, one CTA per SM so the tensor pipe idled during every softmax):
//   * one CTA per (128-query tile, head, batch), 6 warps, ~100 KB smem and 256 TMEM columns -> TWO CTAs per SM, so
//     one CTA's softmax overlaps the other's MMAs;
//   * warp 0 = TMA producer (Q once, 64-key K/V tiles through 2-stage rings), warp 1 = MMA issuer, warps 2-5 =
//     softmax with ONE thread per query row (= TMEM lane): no cross-thread max exchange, no named barriers;
//   * S ping-pong in TMEM (2 x 64 columns): S_{j+1} = Q K_{j+1}^T is in flight while softmax works on S_j;
//   * P is written back as bf16 INTO the S columns (tcgen05.st) and consumed as the TMEM A operand of O += P V:
//     no P tile in shared memory, no generic->async proxy fence, no smem bandwidth for A;
//   * O (128 columns) is rescaled lazily: only when the running max grew by more than 2^8 (exact: the same
//     reference max is used for P and for the row sum), which after the first tiles almost never happens;
//   * dropout = AND of the packed bf16 P pairs with masks expanded from precomputed keep bits (dropmask.cuh).
// Mask modes: none | per-row visible key interval [lo,hi) (document / padding masks; KV tiles outside the union of
// the tile's intervals are skipped) | dense additive bf16 bias (arbitrary masks). Fully-masked rows (finite -1e9 on
// every key) attend uniformly to all T keys like the reference.
#include <stdlib.h>
#include "attn_tc_common.cuh"

namespace obt {

constexpr int FWD_BN = 64;                         // keys per tile
constexpr uint32_t FWD_KV_BYTES = FWD_BN * 128 * 2;  // one [64 x 128] bf16 tile = two 8 KB swizzle sub-tiles

struct AttnFwdSmem {
  static constexpr uint32_t Q_OFF = 0;
  static constexpr uint32_t K_OFF = Q_OFF + ATT_TILE_BYTES;     // 2 stages
  static constexpr uint32_t V_OFF = K_OFF + 2 * FWD_KV_BYTES;   // 2 stages
  static constexpr uint32_t BAR_OFF = V_OFF + 2 * FWD_KV_BYTES;
  static constexpr uint32_t BYTES = BAR_OFF + 256 + 1024;
};

constexpr int ATT_FWD_THREADS = 64 + 128;  // TMA warp, MMA warp, 4 softmax warps
constexpr float ATT_LAZY_LOG2 = 8.0f;      // rescale O only when the row max grew by more than 2^8

template <bool kDrop>
__global__ void __launch_bounds__(ATT_FWD_THREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                   const AttnTcParams p, int C) {
  // 1024-byte alignment (128B-swizzle atoms) is requested from the toolchain, so every smem address below is a
  // link-time constant instead of a live register (the run-time round-up cost registers / spill reloads in the loops)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* sQ = smem + AttnFwdSmem::Q_OFF;
  uint8_t* sK = smem + AttnFwdSmem::K_OFF;
  uint8_t* sV = smem + AttnFwdSmem::V_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnFwdSmem::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_full = bars + 11;   // [2]
  uint64_t* pv_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  int* s_range = reinterpret_cast<int*>(bars + 15);  // [0] = min lo, [1] = max hi, [2] = any fully-masked row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;
  // tile metadata precomputed once per micro-batch (obt_attn_tile_meta): one broadcast load issued before anything
  // else instead of 128 interval loads + shared-memory atomics between the two block barriers below
  int4 qm = make_int4(T, 0, 0, 0);
  if (p.qmeta != nullptr)
    qm = *reinterpret_cast<const int4*>(p.qmeta + (static_cast<long long>(b) * gridDim.x + blockIdx.x) * 4);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
    }
    mbar_init(pv_done, 1);
    fence_barrier_init();
    s_range[0] = T;
    s_range[1] = 0;
    s_range[2] = 0;
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 256);
  __syncthreads();
  // KV tile range covering the union of this tile's visible intervals
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
  int jb = 0, je = (T + FWD_BN - 1) / FWD_BN;
  if (p.row_lo != nullptr) {
    const int r_lo = p.qmeta ? qm.x : s_range[0], r_hi = p.qmeta ? qm.y : s_range[1];
    const int r_dead = p.qmeta ? qm.z : s_range[2];
    if (r_dead == 0 && r_hi > r_lo) {
      jb = r_lo / FWD_BN;
      je = (r_hi + FWD_BN - 1) / FWD_BN;
    }
  }
  const int n_tiles = je - jb;

  const int row0 = b * T;                 // first token row of this sequence in the qkv buffer
  const int qcol = h * ATT_D, kcol = C + h * ATT_D, vcol = 2 * C + h * ATT_D;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(&tm_q, q_full, sQ, qcol, row0 + t0);
      tma_load_2d(&tm_q, q_full, sQ + 16384, qcol + 64, row0 + t0);
      for (int jj = 0; jj < n_tiles; ++jj) {
        const int st = jj & 1;
        const uint32_t par = (jj >> 1) & 1;
        const int krow = row0 + (jb + jj) * FWD_BN;
        mbar_wait(&k_empty[st], par ^ 1);
        mbar_expect_tx(&k_full[st], FWD_KV_BYTES);
        tma_load_2d(&tm_kv, &k_full[st], sK + st * FWD_KV_BYTES, kcol, krow);
        tma_load_2d(&tm_kv, &k_full[st], sK + st * FWD_KV_BYTES + 8192, kcol + 64, krow);
        mbar_wait(&v_empty[st], par ^ 1);
        mbar_expect_tx(&v_full[st], FWD_KV_BYTES);
        tma_load_2d(&tm_kv, &v_full[st], sV + st * FWD_KV_BYTES, vcol, krow);
        tma_load_2d(&tm_kv, &v_full[st], sV + st * FWD_KV_BYTES + 8192, vcol + 64, krow);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (warp-uniform) control flow and waits; one elected lane issues. Keeping every operand
    // provably uniform lets the compiler feed tcgen05.mma from uniform registers: with `if (lane == 0)` around the
    // loop it wrapped EVERY MMA in a per-lane R2UR/ELECT waterfall loop (~13 instructions,
    // profiles/r01_attn_v7_bwd.source.txt).
    const int n_t = __shfl_sync(0xffffffffu, n_tiles, 0);
    const bool leader = elect_one();
    const uint32_t q_addr = smem_u32(sQ);
    const uint32_t o_tmem = tmem_base + 128;
    mbar_wait(q_full, 0);
    const int pro = n_t < 2 ? n_t : 2;
    for (int t = 0; t < pro; ++t) {  // S_0, S_1
      mbar_wait(&k_full[t], 0);
      tc_fence_after();
      if (leader) {
        issue_scores_128x64(tmem_base + t * 64, q_addr, 16384, smem_u32(sK + t * FWD_KV_BYTES), 8192);
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[t]);
      }
      __syncwarp();
    }
    for (int jj = 0; jj < n_t; ++jj) {
      const int st = jj & 1;
      const uint32_t par = (jj >> 1) & 1;
      mbar_wait(&p_full[st], par);
      mbar_wait(&v_full[st], par);
      tc_fence_after();
      if (leader) {
        issue_pv_ts_128x128x64(o_tmem, tmem_base + st * 64, smem_u32(sV + st * FWD_KV_BYTES), 8192, jj > 0);
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
      __syncwarp();
      if (jj + 2 < n_t) {  // S_{j+2} reuses the score buffer P_j was just read from (MMAs execute in order)
        mbar_wait(&k_full[st], par ^ 1);
        tc_fence_after();
        if (leader) {
          issue_scores_128x64(tmem_base + st * 64, q_addr, 16384, smem_u32(sK + st * FWD_KV_BYTES), 8192);
          umma_commit(&s_full[st]);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax / correction / epilogue: one thread per query row =====================
    const int q = warp & 3;              // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;         // row within the tile == TMEM lane
    const int i = t0 + r;                // query position
    const bool row_ok = i < T;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float sc2 = p.scale * LOG2E;
    int lo = 0, hi = T;
    bool dead = false;  // fully-masked row: every key carries the same finite bias -> uniform over all T keys
    if (p.row_lo != nullptr && row_ok) {
      lo = p.row_lo[static_cast<long long>(b) * T + i];
      hi = p.row_hi[static_cast<long long>(b) * T + i];
      if (lo >= hi) {
        lo = 0;
        hi = T;
        dead = true;
      }
    }
    const bool any_dead = __any_sync(0xffffffffu, dead);
    const float live01 = dead ? 0.f : 1.f;
    const __nv_bfloat16* mrow =
        (p.mask != nullptr && row_ok) ? p.mask + b * p.msb + h * p.msh + static_cast<long long>(i) * p.msq : nullptr;
    const uint32_t* keep_row = nullptr;
    if (kDrop && row_ok) keep_row = p.keep + ((static_cast<long long>(b) * p.H + h) * T + i) * p.nw;
    float m_ref = -INFINITY, l_run = 0.f;  // reference max (log2 domain) and exp-sum relative to it

    // keep words one tile ahead (they come from HBM: the mask generator ran before the c_attn GEMM)
    auto load_kw = [&](int jj, int u) -> uint32_t {
      const int w = (((jb + jj) * FWD_BN) >> 5) + u;
      return (kDrop && keep_row != nullptr && jj < n_tiles && w < p.nw) ? keep_row[w] : 0xffffffffu;
    };
    uint32_t kw0_next = load_kw(0, 0), kw1_next = load_kw(0, 1);
    for (int jj = 0; jj < n_tiles; ++jj) {
      const int st = jj & 1;
      const int j0 = (jb + jj) * FWD_BN;
      const uint32_t kw0 = kw0_next, kw1 = kw1_next;
      kw0_next = load_kw(jj + 1, 0);
      kw1_next = load_kw(jj + 1, 1);
      mbar_wait(&s_full[st], (jj >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = lane_addr + st * 64;
      float s[64];
      {
        uint32_t v0[32], v1[32];
        __syncwarp();
        tmem_ld_32x32(s_addr, v0);
        tmem_ld_32x32(s_addr + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          s[e] = __uint_as_float(v0[e]);
          s[32 + e] = __uint_as_float(v1[e]);
        }
      }
      float mul = sc2;
      if (p.mask != nullptr) {  // dense additive bias (kernel-uniform branch): s <- (s * scale + bias) * log2(e)
        mul = 1.0f;
        const uint4* mp = reinterpret_cast<const uint4*>((mrow ? mrow : p.mask) + j0);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int jbase = j0 + g * 8;
          uint4 mu = make_uint4(0, 0, 0, 0);
          if (jbase + 8 <= T && mrow != nullptr) mu = mp[g];
          const uint32_t mw[4] = {mu.x, mu.y, mu.z, mu.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float bias = (e & 1) ? bf16_hi(mw[e >> 1]) : bf16_lo(mw[e >> 1]);
            const float t = __fadd_rn(__fmul_rn(s[g * 8 + e], p.scale), bias) * LOG2E;
            s[g * 8 + e] = (jbase + e < T) ? t : -INFINITY;
          }
        }
      } else if (any_dead) {
#pragma unroll
        for (int e = 0; e < 64; ++e) {
          const int j = j0 + e;
          s[e] = (j >= lo && j < hi) ? s[e] * live01 : -INFINITY;
        }
      } else if (!__all_sync(0xffffffffu, j0 >= lo && j0 + FWD_BN <= hi)) {
        // tile straddles an interval end (or the end of the sequence) for some row of the warp: per-row visibility
        // bits, one bit test per element
        const uint32_t vm0 = interval_bits32(lo, hi, j0), vm1 = interval_bits32(lo, hi, j0 + 32);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          s[e] = (vm0 & (1u << e)) ? s[e] : -INFINITY;
          s[32 + e] = (vm1 & (1u << e)) ? s[32 + e] : -INFINITY;
        }
      }
      // ---- row max of the tile, lazy update of the reference max
      float tmax = fmaxf(s[0], s[1]);
#pragma unroll
      for (int e = 2; e < 64; e += 2) tmax = fmaxf(tmax, fmaxf(s[e], s[e + 1]));
      const float m_tile = tmax * mul;
      const bool grow = m_tile > m_ref + ATT_LAZY_LOG2;  // also true for the first finite tile (m_ref = -inf)
      float alpha = 1.0f;
      if (grow) {
        alpha = fast_exp2(m_ref - m_tile);  // m_ref = -inf -> 0
        m_ref = m_tile;
        l_run *= alpha;
      }
      if (jj > 0 && __any_sync(0xffffffffu, grow)) {
        // O += P_{j-1} V_{j-1} must have landed before this row of O is rescaled
        mbar_wait(pv_done, (jj - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t o[32];
          __syncwarp();
          tmem_ld_32x32(lane_addr + 128 + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          tmem_st_32x32(lane_addr + 128 + c * 32, o);
        }
        tmem_st_wait();
      }
      // ---- P = exp2(s * mul - m_ref), row sum, dropout, bf16 pack, back into the S columns
      const float nm = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float lsum0 = 0.f, lsum1 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int e = 0; e < 64; e += 2) {
        const float p0 = fast_exp2(fmaf(s[e], mul, nm));
        const float p1 = fast_exp2(fmaf(s[e + 1], mul, nm));
        lsum0 += p0;
        lsum1 += p1;
        pk[e >> 1] = pack_bf16x2(p0, p1);
      }
      l_run += lsum0 + lsum1;
      if (kDrop) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t w = g ? kw1 : kw0;
#pragma unroll
          for (int sft = 0; sft < 8; ++sft) {
            const uint32_t t = w << sft;
            pk[g * 16 + sft * 2 + 0] &= keep_pair_mask(t, 0);
            pk[g * 16 + sft * 2 + 1] &= keep_pair_mask(t, 2);
          }
        }
      }
      // Keep every warp within one phase of pv_done (parity waits cannot tell phases two apart): PV_{j-1} was
      // issued a whole softmax tile ago, so this wait is almost always already satisfied.
      if (jj > 0) mbar_wait(pv_done, (jj - 1) & 1);
      __syncwarp();
      tmem_st_32x32(s_addr, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
    }
    // ---- epilogue: O / l (and the dropout 1/(1-p)) -> y, (max, log-sum) -> lse
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = (kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f) / l_run;
    __nv_bfloat16* yrow = p.y + (static_cast<long long>(row0) + i) * p.ldy + h * ATT_D;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + 128 + c * 32, o);
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
    if (row_ok && p.lse != nullptr) {
      float* l = p.lse + 2 * ((static_cast<long long>(b) * p.H + h) * T + i);
      l[0] = m_ref / LOG2E;
      l[1] = logf(l_run);
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 256);
  }
}

}  // namespace obt

using namespace obt;

extern "C" int obt_attn_tc_fwd(const void* qkv, long long ld, const void* mask, long long msb, long long msh,
                               long long msq, const int* row_lo, const int* row_hi, void* y, long long ldy, float* lse,
                               int B, int H, int T, int d, float scale, float drop_p, const unsigned int* keep,
                               const int* qmeta, cudaStream_t stream) {
  OBT_REQUIRE(qkv && y && lse, "obt_attn_tc_fwd: null pointer");
  OBT_REQUIRE(d == ATT_D, "obt_attn_tc_fwd: head_dim=%d, the tensor-core kernel is specialised for 128", d);
  OBT_REQUIRE(B > 0 && H > 0 && T > 0, "obt_attn_tc_fwd: empty problem");
  OBT_REQUIRE(scale > 0.f, "obt_attn_tc_fwd: scale must be positive");
  OBT_REQUIRE(ld % 8 == 0 && ldy % 8 == 0, "obt_attn_tc_fwd: pitches must be multiples of 8");
  OBT_REQUIRE(mask == nullptr || (msq % 8 == 0 && msb % 8 == 0 && msh % 8 == 0 && T % 8 == 0 &&
                                   (reinterpret_cast<uintptr_t>(mask) & 15) == 0),
              "obt_attn_tc_fwd: dense mask needs 16-byte aligned rows (T %% 8 == 0)");
  OBT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "obt_attn_tc_fwd: dropout p=%f", drop_p);
  OBT_REQUIRE(drop_p == 0.f || keep != nullptr, "obt_attn_tc_fwd: dropout needs the keep mask (obt_attn_keep_mask)");
  const int C = H * d;
  CUtensorMap tm_q, tm_kv;
  int rc = get_tensor_map_2d(&tm_q, qkv, static_cast<uint64_t>(3) * C, static_cast<uint64_t>(B) * T,
                             static_cast<uint64_t>(ld), 64, 128);
  if (rc) return rc;
  rc = get_tensor_map_2d(&tm_kv, qkv, static_cast<uint64_t>(3) * C, static_cast<uint64_t>(B) * T,
                         static_cast<uint64_t>(ld), 64, 64);
  if (rc) return rc;
  AttnTcParams p = {};
  p.B = B; p.H = H; p.T = T;
  p.scale = scale;
  p.mask = static_cast<const __nv_bfloat16*>(mask);
  p.msb = msb; p.msh = msh; p.msq = msq;
  p.row_lo = mask ? nullptr : row_lo;
  p.row_hi = mask ? nullptr : row_hi;
  p.qmeta = (p.row_lo != nullptr) ? qmeta : nullptr;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.ldy = ldy;
  p.lse = lse;
  p.drop_p = drop_p;
  p.keep = keep;
  p.nw = keep_words(T);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          AttnFwdSmem::BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(attn_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          AttnFwdSmem::BYTES);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      set_last_error("obt_attn_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      return OBT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((T + ATT_BM - 1) / ATT_BM, H, B);
  if (drop_p > 0.f)
    attn_tc_fwd_kernel<true><<<grid, ATT_FWD_THREADS, AttnFwdSmem::BYTES, stream>>>(tm_q, tm_kv, p, C);
  else
    attn_tc_fwd_kernel<false><<<grid, ATT_FWD_THREADS, AttnFwdSmem::BYTES, stream>>>(tm_q, tm_kv, p, C);
  return check_launch("attn_tc_fwd");
}
