// Attention backward, PERSISTENT forms of the dQ and dK/dV kernels of attn_tc_bwd.cu (same math, same TMEM layout, same
// inner loops): one CTA per SM fetches work items (query tile / key tile of one (batch, head)) from a device-side
// counter until none is left.
//
// Why (profiles/r02c_attn_dq_w8.source.txt, r02c_attn_dkv_w8.source.txt): an item is short (15-19 us for ~10
// sub-tiles) and ~45 % of the one-CTA-per-item kernels' life is fixed latency in front of and behind the sub-tile loop
// that nothing overlaps because the kernels hold all 512 TMEM columns (one CTA per SM): CTA launch, barrier
// initialisation, TMEM allocation, two block barriers, the interval scan, the first TMA round trip, the global loads of
// Q / dO and of the per-row parameters, the accumulator read-out, TMEM release, exit. Here the set-up happens once per
// SM and the per-item latencies are hidden behind the previous item:
//   * the TMA producer runs ahead ACROSS items (the K/V ring of dQ and the Q/dO ring of dK/dV simply keep cycling;
//     dK/dV's resident K/V tile is re-loaded as soon as the last score product of the previous item has read it);
//   * the tile range / relevance bits come from the per-micro-batch tile metadata (obt_attn_tile_meta): one load;
//   * the compute warps issue the next item's Q / dO global loads BEFORE they drain the current item's accumulator;
//   * every mbarrier phase is derived from counters that run across items, so nothing is re-initialised.
// Work distribution: items are handed out by atomicAdd on sched[0] by the producer warp, which publishes
// {item, range / relevance} in a small shared-memory ring guarded by mbarriers (item costs differ by up to 8x with
// document masks: a static assignment would leave SMs idle). The last CTA to finish re-zeroes the counters.
#include "attn_tc_common.cuh"

namespace obt {

constexpr int ATT_P_SLOTS = 8;  // item ring; the producer is never more than 3 K/V stages (<= 3 items) ahead

// ---------------------------------------------------------------------------------------------------------------------
// persistent dQ
//   TMEM  [0,128) / [128,256) S (64 columns) | dP (64 columns), ping-pong; [256,384) dQ; [384,448) Q; [448,512) dO
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ATT_PDQ_KV_STAGES = 3;

struct AttnPDqSmem {
  static constexpr uint32_t K_OFF = 0;  // stages of 128 keys
  static constexpr uint32_t V_OFF = K_OFF + ATT_PDQ_KV_STAGES * ATT_TILE_BYTES;
  static constexpr uint32_t BAR_OFF = V_OFF + ATT_PDQ_KV_STAGES * ATT_TILE_BYTES;  // 32 barriers
  static constexpr uint32_t SLOT_OFF = BAR_OFF + 256;                              // ATT_P_SLOTS x int4
  static constexpr uint32_t BYTES = SLOT_OFF + ATT_P_SLOTS * 16 + 64 + 1024;
};

template <bool kDrop>
__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
attn_tc_dq_persist_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __nv_bfloat16* __restrict__ qkv,
                          long long ld, const __nv_bfloat16* __restrict__ dy, long long lddy, const AttnTcParams p, int C,
                          int* __restrict__ sched) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* sK = smem + AttnPDqSmem::K_OFF;
  uint8_t* sV = smem + AttnPDqSmem::V_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnPDqSmem::BAR_OFF);
  uint64_t* qdo_ready = bars + 0;  // Q / dO of the item are in TMEM (8 compute warps)
  uint64_t* k_full = bars + 1;     // [3]
  uint64_t* v_full = bars + 4;     // [3]
  uint64_t* kv_empty = bars + 7;   // [3]
  uint64_t* sdp_full = bars + 10;  // [2]
  uint64_t* ds_full = bars + 12;   // [2]
  uint64_t* dq_done = bars + 14;   // last dQ MMA of the item completed
  uint64_t* item_full = bars + 15; // [ATT_P_SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15 + ATT_P_SLOTS);
  volatile int4* slots = reinterpret_cast<volatile int4*>(smem + AttnPDqSmem::SLOT_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T;
  const int nT = (T + ATT_BM - 1) / ATT_BM;
  const int n_items = nT * p.H * p.B;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(qdo_ready, ATT_COMPUTE_WARPS);
    for (int i = 0; i < ATT_PDQ_KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&ds_full[i], ATT_COMPUTE_WARPS);
    }
    mbar_init(dq_done, 1);
    for (int i = 0; i < ATT_P_SLOTS; ++i) mbar_init(&item_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(tmem_slot), 0);
  constexpr uint32_t TM_DQ = 256, TM_Q = 384, TM_DO = 448;

  // item -> (query tile, head, batch); query tile fastest so that concurrently running CTAs share K / V in L2
  auto decode = [&](int item, int& tq, int& h, int& b) {
    tq = item % nT;
    const int bh = item / nT;
    h = bh % p.H;
    b = bh / p.H;
  };
  // consumers: wait for ring entry k and read it (item < 0: no more work)
  auto wait_item = [&](int k) -> int4 {
    mbar_wait(&item_full[k % ATT_P_SLOTS], (k / ATT_P_SLOTS) & 1);
    const volatile int4* sp = &slots[k % ATT_P_SLOTS];
    return make_int4(sp->x, sp->y, sp->z, sp->w);
  };

  if (warp < ATT_BWD_FIRST_COMPUTE_WARP) {
    reg_dealloc<56>();
    if (warp == 0) {
      // ===================== scheduler + TMA producer =====================
      if (lane == 0) {
        int st = 0;
        uint32_t ph = 0;
        int item = atomicAdd(&sched[0], 1);
        for (int k = 0;; ++k) {
          volatile int4* sp = &slots[k % ATT_P_SLOTS];
          if (item >= n_items) {
            sp->x = -1; sp->y = 0; sp->z = 0; sp->w = 0;
            mbar_arrive(&item_full[k % ATT_P_SLOTS]);
            break;
          }
          int tq, h, b;
          decode(item, tq, h, b);
          int jb = 0, je = nT;
          if (p.qmeta != nullptr) {
            const int4 qm = *reinterpret_cast<const int4*>(p.qmeta + (static_cast<long long>(b) * nT + tq) * 4);
            if (qm.z == 0 && qm.y > qm.x) {
              jb = qm.x / ATT_BN;
              je = (qm.y + ATT_BN - 1) / ATT_BN;
            }
          }
          sp->x = item; sp->y = jb; sp->z = je; sp->w = 0;
          mbar_arrive(&item_full[k % ATT_P_SLOTS]);
          const int next = atomicAdd(&sched[0], 1);  // in flight while this item's tiles are issued
          const int row0 = b * T;
          const int kcol = C + h * ATT_D, vcol = 2 * C + h * ATT_D;
          for (int jj = jb; jj < je; ++jj) {
            const int krow = row0 + jj * ATT_BN;
            mbar_wait(&kv_empty[st], ph ^ 1);
            mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
            tma_load_2d(&tm_qkv, &k_full[st], sK + st * ATT_TILE_BYTES, kcol, krow);
            tma_load_2d(&tm_qkv, &k_full[st], sK + st * ATT_TILE_BYTES + 16384, kcol + 64, krow);
            mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
            tma_load_2d(&tm_qkv, &v_full[st], sV + st * ATT_TILE_BYTES, vcol, krow);
            tma_load_2d(&tm_qkv, &v_full[st], sV + st * ATT_TILE_BYTES + 16384, vcol + 64, krow);
            if (++st == ATT_PDQ_KV_STAGES) { st = 0; ph ^= 1; }
          }
          item = next;
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      const bool leader = elect_one();
      int g_tile = 0;  // K/V tiles consumed so far (ring position), all items
      int g_sub = 0;   // 64-key sub-tiles issued so far (score buffer / phase), all items
      for (int k = 0;; ++k) {
        const int4 it = wait_item(k);
        const int item = __shfl_sync(0xffffffffu, it.x, 0);
        if (item < 0) break;
        const int n_tiles = __shfl_sync(0xffffffffu, it.z - it.y, 0);
        const int n_sub = 2 * n_tiles;
        mbar_wait(qdo_ready, k & 1);
        // scores of sub-tile s of this item: S -> buffer columns [0,64), dP -> [64,128)
        auto issue_scores = [&](int s) {
          const int gt = g_tile + (s >> 1), hsub = s & 1, st = gt % ATT_PDQ_KV_STAGES, gs = g_sub + s;
          if (hsub == 0) {
            const uint32_t ph = (gt / ATT_PDQ_KV_STAGES) & 1;
            mbar_wait(&k_full[st], ph);
            mbar_wait(&v_full[st], ph);
          }
          tc_fence_after();
          if (leader) {
            const uint32_t k_addr = smem_u32(sK + st * ATT_TILE_BYTES) + hsub * 8192;
            const uint32_t v_addr = smem_u32(sV + st * ATT_TILE_BYTES) + hsub * 8192;
            const uint32_t d = tmem_base + (gs & 1) * 128;
            issue_scores_ts_128x64(d, tmem_base + TM_Q, k_addr, 16384);        // S  = Q K^T
            issue_scores_ts_128x64(d + 64, tmem_base + TM_DO, v_addr, 16384);  // dP = dO V^T
            umma_commit(&sdp_full[gs & 1]);
          }
          __syncwarp();
        };
        issue_scores(0);
        for (int s = 0; s < n_sub; ++s) {
          if (s + 1 < n_sub) issue_scores(s + 1);
          const int gt = g_tile + (s >> 1), hsub = s & 1, st = gt % ATT_PDQ_KV_STAGES, gs = g_sub + s;
          mbar_wait(&ds_full[gs & 1], (gs >> 1) & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t k_rows = smem_u32(sK + st * ATT_TILE_BYTES) + hsub * 8192;
            const uint32_t dsb = tmem_base + (gs & 1) * 128;
            issue_grad_ts_128x128x64(tmem_base + TM_DQ, dsb, dsb + 32, k_rows, 16384, s > 0);  // dQ += dS K
            if (hsub == 1) umma_commit(&kv_empty[st]);
            if (s == n_sub - 1) umma_commit(dq_done);
          }
          __syncwarp();
        }
        g_tile += n_tiles;
        g_sub += n_sub;
      }
    }
  } else {
    // ===================== compute warps: two threads per query row, 32 score columns each =====================
    reg_alloc<224>();
    const int q = warp & 3;
    const int hh = (warp - ATT_BWD_FIRST_COMPUTE_WARP) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float inv_keep = kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    int g_sub = 0;
    uint4 qv[8], dov[8];
    // half rows of Q and dO of an item (d columns [64 hh, +64)): only ISSUES the loads
    auto load_qdo = [&](const int4& it) {
      int tq, h, b;
      decode(it.x, tq, h, b);
      const int i_ = tq * ATT_BM + r;
      const bool ok_ = i_ < T;
      const long long row_ = static_cast<long long>(b) * T + (ok_ ? i_ : 0);
      const uint4* src = reinterpret_cast<const uint4*>(qkv + row_ * ld + h * ATT_D + hh * 64);
      const uint4* src2 = reinterpret_cast<const uint4*>(dy + row_ * lddy + h * ATT_D + hh * 64);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        qv[g] = ok_ ? src[g] : make_uint4(0, 0, 0, 0);
        dov[g] = ok_ ? src2[g] : make_uint4(0, 0, 0, 0);
      }
    };
    int4 it = wait_item(0);
    if (it.x >= 0) load_qdo(it);
    for (int k = 0; it.x >= 0; ++k) {
      int tq, h, b;
      decode(it.x, tq, h, b);
      const int jb = it.y;
      const int n_sub = 2 * (it.z - it.y);
      const int t0 = tq * ATT_BM;
      const int i = t0 + r;
      const bool row_ok = i < T;
      const int row0 = b * T;
      {  // Q and dO (already in registers) -> TMEM: 32 packed words of each
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
      float row_scale = p.scale;
      if (p.row_lo != nullptr && row_ok) {
        lo = p.row_lo[static_cast<long long>(b) * T + i];
        hi = p.row_hi[static_cast<long long>(b) * T + i];
        if (lo >= hi) { lo = 0; hi = T; row_scale = 0.f; }
      }
      const long long bh = static_cast<long long>(b) * p.H + h;
      float off_nat = 0.f, ls2 = 0.f, dl = 0.f;
      if (row_ok) {
        off_nat = p.lse[2 * (bh * T + i)];
        ls2 = p.lse[2 * (bh * T + i) + 1] * LOG2E;
        dl = p.delta[bh * T + i];
      }
      const float neg = off_nat * LOG2E + ls2;
      const float sc2 = row_scale * LOG2E;
      const __nv_bfloat16* mrow =
          (p.mask != nullptr && row_ok) ? p.mask + b * p.msb + h * p.msh + static_cast<long long>(i) * p.msq : nullptr;
      const float dlp = kDrop ? dl * (1.0f - p.drop_p) : dl;
      const uint32_t* keep_row = nullptr;
      if (kDrop && row_ok) keep_row = p.keep + (bh * T + i) * p.nw;

      auto load_kw = [&](int s) -> uint32_t {
        const int w = ((jb + (s >> 1)) * ATT_BN + (s & 1) * 64 + hh * 32) >> 5;
        return (kDrop && keep_row != nullptr && s < n_sub && w < p.nw) ? keep_row[w] : 0xffffffffu;
      };
      uint32_t kw_next = load_kw(0);
      for (int s = 0; s < n_sub; ++s) {
        const int gs = g_sub + s;
        const int bsel = gs & 1;
        const int j0 = (jb + (s >> 1)) * ATT_BN + (s & 1) * 64 + hh * 32;  // first key of this thread's 32 columns
        const uint32_t kw = kw_next;
        kw_next = load_kw(s + 1);
        mbar_wait(&sdp_full[bsel], (gs >> 1) & 1);
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
          const float nneg = -neg;
#pragma unroll
          for (int e = 0; e < 32; ++e) ds[e] = fast_exp2(fmaf(__uint_as_float(sv[e]), sc2, nneg));
        } else {
          const uint32_t vm = row_ok ? interval_bits32(lo, hi, j0) : 0u;
          if (__all_sync(0xffffffffu, vm == 0u)) {
            none_visible = true;
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
      g_sub += n_sub;
      // next item: its Q / dO loads fly while this item's accumulator is drained
      const int4 nxt = wait_item(k + 1);
      if (nxt.x >= 0) load_qdo(nxt);
      // ---- dQ epilogue of this item: columns [64 hh, 64 hh + 64) of the row
      mbar_wait(dq_done, k & 1);
      tc_fence_after();
      __nv_bfloat16* drow = p.dq + (static_cast<long long>(row0) + i) * p.ldd + h * ATT_D;
      const float oscale = row_scale * inv_keep;
      const bool do_rope = row_ok && p.rope_cos != nullptr;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
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
      tc_fence_before();  // the accumulator reads are ordered before the next item's first dQ MMA (ds_full arrival)
      it = nxt;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 512);
  }
  // the last CTA to leave re-arms the work counter for the next launch on this stream
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&sched[1], 1) == static_cast<int>(gridDim.x) - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// persistent dK / dV
//   TMEM  [0,128) / [128,256)  S^T (64 columns) | dP^T (64 columns), ping-pong;  [256,384) dV;  [384,512) dK
//   The resident K / V tile is single-buffered: the producer re-loads it for the next item as soon as the LAST score
//   product of the current item (the only MMAs that read sK / sV) has completed (kv_free), i.e. one sub-tile of math,
//   the last gradient products and the whole accumulator read-out before the next item needs it.
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint32_t ATT_PSUB_BYTES = 64 * 128 * 2;  // one [64 rows x 128] bf16 tile = two 8 KB swizzle sub-tiles
constexpr int ATT_PQDO_STAGES = 4;
constexpr uint32_t ATT_PWPAR_BYTES = 1152;         // per warp and buffer (see attn_tc_bwd.cu)

struct AttnPDkvSmem {
  static constexpr uint32_t K_OFF = 0;
  static constexpr uint32_t V_OFF = K_OFF + ATT_TILE_BYTES;
  static constexpr uint32_t Q_OFF = V_OFF + ATT_TILE_BYTES;
  static constexpr uint32_t DO_OFF = Q_OFF + ATT_PQDO_STAGES * ATT_PSUB_BYTES;
  static constexpr uint32_t PAR_OFF = DO_OFF + ATT_PQDO_STAGES * ATT_PSUB_BYTES;  // [8 warps][2 buffers]
  static constexpr uint32_t BAR_OFF = PAR_OFF + ATT_COMPUTE_WARPS * 2 * ATT_PWPAR_BYTES;  // 32 barriers
  static constexpr uint32_t SLOT_OFF = BAR_OFF + 256;  // ATT_P_SLOTS x {int4 item, uint4 relevance bits}
  static constexpr uint32_t BYTES = SLOT_OFF + ATT_P_SLOTS * 32 + 64 + 1024;
};

template <bool kDrop>
__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
attn_tc_dkv_persist_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_q64,
                           const __grid_constant__ CUtensorMap tm_dy64, const AttnTcParams p, int C,
                           int* __restrict__ sched) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* sK = smem + AttnPDkvSmem::K_OFF;
  uint8_t* sV = smem + AttnPDkvSmem::V_OFF;
  uint8_t* sQ = smem + AttnPDkvSmem::Q_OFF;
  uint8_t* sDO = smem + AttnPDkvSmem::DO_OFF;
  uint8_t* sPar = smem + AttnPDkvSmem::PAR_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnPDkvSmem::BAR_OFF);
  uint64_t* kv_full = bars + 0;     // K / V tile of the item landed
  uint64_t* kv_free = bars + 1;     // every score product of the item has read sK / sV
  uint64_t* qdo_full = bars + 2;    // [4]
  uint64_t* qdo_free = bars + 6;    // [4] dV/dK MMAs that read Q/dO stage s completed
  uint64_t* sdp_full = bars + 10;   // [2]
  uint64_t* pds_full = bars + 12;   // [2]
  uint64_t* grads_done = bars + 14; // all dV/dK MMAs of the item completed
  uint64_t* item_full = bars + 15;  // [ATT_P_SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15 + ATT_P_SLOTS);
  volatile int* slot_item = reinterpret_cast<volatile int*>(smem + AttnPDkvSmem::SLOT_OFF);               // [slots]
  volatile uint32_t* slot_rel = reinterpret_cast<volatile uint32_t*>(smem + AttnPDkvSmem::SLOT_OFF + ATT_P_SLOTS * 4);  // [slots][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T;
  const int nq = (T + 63) / 64;
  const int nT = (T + ATT_BN - 1) / ATT_BN;
  const int n_items = nT * p.H * p.B;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_q64);
    tma_prefetch_desc(&tm_dy64);
    mbar_init(kv_full, 1);
    mbar_init(kv_free, 1);
    for (int i = 0; i < ATT_PQDO_STAGES; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&pds_full[i], ATT_COMPUTE_WARPS);
    }
    mbar_init(grads_done, 1);
    for (int i = 0; i < ATT_P_SLOTS; ++i) mbar_init(&item_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(tmem_slot), 0);

  // item -> (key tile, head, batch)
  auto decode = [&](int item, int& tk, int& h, int& b) {
    tk = item % nT;
    const int bh = item / nT;
    h = bh % p.H;
    b = bh / p.H;
  };
  auto wait_item = [&](int k) -> int {
    mbar_wait(&item_full[k % ATT_P_SLOTS], (k / ATT_P_SLOTS) & 1);
    return slot_item[k % ATT_P_SLOTS];
  };

  if (warp < ATT_BWD_FIRST_COMPUTE_WARP) {
    reg_dealloc<56>();
    if (warp == 0) {
      // ===================== scheduler + TMA producer =====================
      if (lane == 0) {
        int st = 0;
        uint32_t ph = 0;
        int item = atomicAdd(&sched[0], 1);
        for (int k = 0;; ++k) {
          const int sl = k % ATT_P_SLOTS;
          if (item >= n_items) {
            slot_item[sl] = -1;
            mbar_arrive(&item_full[sl]);
            break;
          }
          int tk, h, b;
          decode(item, tk, h, b);
          uint4 rel = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (p.kmeta != nullptr)
            rel = *reinterpret_cast<const uint4*>(p.kmeta + (static_cast<long long>(b) * nT + tk) * 4);
          slot_rel[sl * 4 + 0] = rel.x; slot_rel[sl * 4 + 1] = rel.y;
          slot_rel[sl * 4 + 2] = rel.z; slot_rel[sl * 4 + 3] = rel.w;
          slot_item[sl] = item;
          mbar_arrive(&item_full[sl]);
          const int next = atomicAdd(&sched[0], 1);  // in flight while this item's tiles are issued
          const int row0 = b * T, j0 = tk * ATT_BN;
          const int qcol = h * ATT_D, kcol = C + h * ATT_D, vcol = 2 * C + h * ATT_D;
          if (k > 0) mbar_wait(kv_free, (k - 1) & 1);  // the previous item's last score product has read sK / sV
          mbar_expect_tx(kv_full, 2 * ATT_TILE_BYTES);
          tma_load_2d(&tm_qkv, kv_full, sK, kcol, row0 + j0);
          tma_load_2d(&tm_qkv, kv_full, sK + 16384, kcol + 64, row0 + j0);
          tma_load_2d(&tm_qkv, kv_full, sV, vcol, row0 + j0);
          tma_load_2d(&tm_qkv, kv_full, sV + 16384, vcol + 64, row0 + j0);
          const uint32_t relw[4] = {rel.x, rel.y, rel.z, rel.w};
          for (int it = 0; it < nq; ++it) {
            if (!((relw[it >> 5] >> (it & 31)) & 1u)) continue;
            mbar_wait(&qdo_free[st], ph ^ 1);
            mbar_expect_tx(&qdo_full[st], 2 * ATT_PSUB_BYTES);
            const int qrow = row0 + it * 64;
            tma_load_2d(&tm_q64, &qdo_full[st], sQ + st * ATT_PSUB_BYTES, qcol, qrow);
            tma_load_2d(&tm_q64, &qdo_full[st], sQ + st * ATT_PSUB_BYTES + 8192, qcol + 64, qrow);
            tma_load_2d(&tm_dy64, &qdo_full[st], sDO + st * ATT_PSUB_BYTES, qcol, qrow);
            tma_load_2d(&tm_dy64, &qdo_full[st], sDO + st * ATT_PSUB_BYTES + 8192, qcol + 64, qrow);
            if (++st == ATT_PQDO_STAGES) { st = 0; ph ^= 1; }
          }
          item = next;
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      const bool leader = elect_one();
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      int g_n = 0;  // relevant sub-tiles issued so far, all items (Q/dO ring position, score buffer, phases)
      for (int k = 0;; ++k) {
        const int item = __shfl_sync(0xffffffffu, wait_item(k), 0);
        if (item < 0) break;
        const int sl = k % ATT_P_SLOTS;
        int n_total = 0;
        for (int it = 0; it < nq; ++it) n_total += (slot_rel[sl * 4 + (it >> 5)] >> (it & 31)) & 1u;
        n_total = __shfl_sync(0xffffffffu, n_total, 0);
        mbar_wait(kv_full, k & 1);
        auto issue_scores = [&](int n) {  // Q/dO stage (g_n + n) % 4, TMEM score buffer (g_n + n) & 1
          const int gn = g_n + n, st = gn % ATT_PQDO_STAGES, tb = gn & 1;
          mbar_wait(&qdo_full[st], (gn / ATT_PQDO_STAGES) & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t d = tmem_base + tb * 128;
            issue_scores_128x64(d, k_addr, 16384, smem_u32(sQ + st * ATT_PSUB_BYTES), 8192);        // S^T  = K Q^T
            issue_scores_128x64(d + 64, v_addr, 16384, smem_u32(sDO + st * ATT_PSUB_BYTES), 8192);  // dP^T = V dO^T
            umma_commit(&sdp_full[tb]);
            if (n == n_total - 1) umma_commit(kv_free);  // no later MMA of this item reads sK / sV
          }
          __syncwarp();
        };
        if (n_total > 0) {
          issue_scores(0);
        } else {
          tc_fence_after();
          if (leader) umma_commit(kv_free);
          __syncwarp();
        }
        for (int n = 0; n < n_total; ++n) {
          if (n + 1 < n_total) issue_scores(n + 1);
          const int gn = g_n + n, st = gn % ATT_PQDO_STAGES, tb = gn & 1;
          mbar_wait(&pds_full[tb], (gn >> 1) & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t buf = tmem_base + tb * 128;
            issue_grad_ts_128x128x64(tmem_base + 256, buf, buf + 32, smem_u32(sDO + st * ATT_PSUB_BYTES), 8192, n > 0);      // dV += P^T dO
            issue_grad_ts_128x128x64(tmem_base + 384, buf + 64, buf + 96, smem_u32(sQ + st * ATT_PSUB_BYTES), 8192, n > 0);  // dK += dS^T Q
            umma_commit(&qdo_free[st]);
          }
          __syncwarp();
        }
        if (leader) umma_commit(grads_done);  // once per item (immediately when the item had no relevant sub-tile)
        __syncwarp();
        g_n += n_total;
      }
    }
  } else {
    // ===================== compute warps: two threads per key row, 32 query columns each =====================
    reg_alloc<224>();
    const int q = warp & 3;
    const int hh = (warp - ATT_BWD_FIRST_COMPUTE_WARP) >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int g_n = 0;
    for (int k = 0;; ++k) {
      const int item = wait_item(k);
      if (item < 0) break;
      int tk, h, b;
      decode(item, tk, h, b);
      const int j0 = tk * ATT_BN;
      const int row0 = b * T;
      const volatile uint32_t* relw = slot_rel + (k % ATT_P_SLOTS) * 4;
      auto relevant = [&](int it) -> bool { return (relw[it >> 5] >> (it & 31)) & 1u; };
    const int r = q * 32 + lane;     // key row within the tile
    const int j = j0 + r;
    const bool key_ok = j < T;
    const int kq0 = j0 + q * 32;     // first key of this warp
    const long long bh = static_cast<long long>(b) * p.H + h;
    const float inv_keep = kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const float keep_frac = kDrop ? 1.0f - p.drop_p : 1.0f;
    const float sc2 = p.scale * LOG2E;
    const uint32_t mybit = 1u << keep_bit_pos(lane);  // this key's bit inside the keep word of its 32-key group
    uint8_t* wpar = sPar + (warp - ATT_BWD_FIRST_COMPUTE_WARP) * 2 * ATT_PWPAR_BYTES;
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
      uint8_t* base = wpar + buf * ATT_PWPAR_BYTES;
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
      const int st = n & 1;            // this warp's parameter buffer (private: local parity)
      const int gn = g_n + n;          // score buffer / barrier phase: counted across items
      const int tb = gn & 1;
      const int i0 = it * 64 + hh * 32;
      const int nx2 = nx < nq ? next_relevant(nx) : nq;
      QParams zn2 = {};
      if (nx2 < nq) zn2 = load_params(nx2);  // in flight during the math of this AND the next sub-tile
      const uint8_t* base = wpar + st * ATT_PWPAR_BYTES;
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
      mbar_wait(&sdp_full[tb], (gn >> 1) & 1);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + tb * 128 + hh * 32, sv);
      tmem_ld_32x32(lane_addr + tb * 128 + 64 + hh * 32, dv);
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
      // P^T over the first 16 of the S^T columns this thread has read, dS^T over the first 16 of its dP^T columns
      __syncwarp();
      tmem_st_32x16(lane_addr + tb * 128 + hh * 32, ptw);
      tmem_st_32x16(lane_addr + tb * 128 + 64 + hh * 32, dsw);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[tb]);
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
    mbar_wait(grads_done, k & 1);  // committed once per item, also for items without a relevant sub-tile
    tc_fence_after();
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
      tc_fence_before();  // accumulator reads ordered before the next item's first dV / dK MMA (pds_full arrival)
      g_n += n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 512);
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&sched[1], 1) == static_cast<int>(gridDim.x) - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

}  // namespace obt

using namespace obt;

int launch_attn_tc_dq_persist(const CUtensorMap& tm_qkv, const void* qkv, long long ld, const void* dy, long long lddy,
                              const AttnTcParams& p, int C, int* sched, bool drop, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_tc_dq_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          AttnPDqSmem::BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(attn_tc_dq_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          AttnPDqSmem::BYTES);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      set_last_error("obt_attn_tc_bwd: cudaFuncSetAttribute(dq persistent): %s",
                     cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      return OBT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int n_items = ((p.T + ATT_BM - 1) / ATT_BM) * p.H * p.B;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  if (drop)
    attn_tc_dq_persist_kernel<true><<<grid, ATT_BWD_THREADS, AttnPDqSmem::BYTES, stream>>>(
        tm_qkv, static_cast<const __nv_bfloat16*>(qkv), ld, static_cast<const __nv_bfloat16*>(dy), lddy, p, C, sched);
  else
    attn_tc_dq_persist_kernel<false><<<grid, ATT_BWD_THREADS, AttnPDqSmem::BYTES, stream>>>(
        tm_qkv, static_cast<const __nv_bfloat16*>(qkv), ld, static_cast<const __nv_bfloat16*>(dy), lddy, p, C, sched);
  return check_launch("attn_tc_dq_persist");
}

int launch_attn_tc_dkv_persist(const CUtensorMap& tm_qkv, const CUtensorMap& tm_q64, const CUtensorMap& tm_dy64,
                               const AttnTcParams& p, int C, int* sched, bool drop, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_tc_dkv_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          AttnPDkvSmem::BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(attn_tc_dkv_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          AttnPDkvSmem::BYTES);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      set_last_error("obt_attn_tc_bwd: cudaFuncSetAttribute(dkv persistent): %s",
                     cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      return OBT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int n_items = ((p.T + ATT_BN - 1) / ATT_BN) * p.H * p.B;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  if (drop)
    attn_tc_dkv_persist_kernel<true><<<grid, ATT_BWD_THREADS, AttnPDkvSmem::BYTES, stream>>>(tm_qkv, tm_q64, tm_dy64, p, C, sched);
  else
    attn_tc_dkv_persist_kernel<false><<<grid, ATT_BWD_THREADS, AttnPDkvSmem::BYTES, stream>>>(tm_qkv, tm_q64, tm_dy64, p, C, sched);
  return check_launch("attn_tc_dkv_persist");
}
