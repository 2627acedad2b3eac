// bf16 GEMM on tcgen05 tensor cores for sm_100a:  D[M,N] (+)= op(A)[M,K] * op(B)[N,K]^T, fp32 accumulation in TMEM.
//
// Replaces every nn.Linear forward / dgrad / wgrad of the reference encoder block
// (training/model.py:102 c_attn, :151 attn c_proj, :163 c_fc, :166 mlp c_proj, :253 lm_head via MuReadout).
//
// Design (one persistent CTA per SM, or one CTA pair per 2 SMs with cta_group::2):
//   warp 0      : TMA producer  - cp.async.bulk.tensor tiles into a 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1      : MMA issuer    - one thread issues tcgen05.mma (128|256 x 256 x 16), accumulators in TMEM,
//                                 2 accumulator stages (2 x 256 columns) so the epilogue overlaps the next tile
//   warps 2..9  : epilogue      - 8 warps, two per TMEM lane quadrant (128 columns each): tcgen05.ld (row per thread)
//                                 -> swizzled smem transpose -> fused epilogue -> coalesced 16-byte global stores
// Operand layouts: either operand may be K-major (K contiguous in memory) or MN-major (M/N contiguous), which covers
// forward (K,K), dgrad (K,MN) and wgrad (MN,MN) without materialising transposes.
#include <math.h>
#include "common.cuh"
#include "ptx.cuh"
#include "gemm_epilogue.cuh"

namespace obt {

struct TileCoord {
  int m_blk, n_blk, split;
};

__device__ __forceinline__ TileCoord tile_coord(int tile, int num_m, int num_n) {
  const int per_split = num_m * num_n;
  TileCoord c;
  c.split = tile / per_split;
  int t = tile - c.split * per_split;
  const int group_sz = GEMM_GROUP_M * num_n;
  int g = t / group_sz;
  int first_m = g * GEMM_GROUP_M;
  int gm = min(GEMM_GROUP_M, num_m - first_m);
  int r = t - g * group_sz;
  c.m_blk = first_m + r % gm;
  c.n_blk = r / gm;
  return c;
}

// kMode 1: one CTA per tile (128 x 256), cta_group::1 MMA.
// kMode 2: CTA pair per 256 x 256 tile, cta_group::2 MMA (each CTA holds half of B, accumulators in both TMEMs).
// kMode 3: cluster of 2 CTAs on vertically adjacent 128 x 256 tiles sharing the B tile: each CTA TMA-loads half of
//          B and MULTICASTS it into both CTAs' smem (L2 -> SM operand traffic 48 KB -> 32 KB per k-block per CTA);
//          the MMAs stay cta_group::1, a smem stage is recycled only when BOTH CTAs' MMAs have consumed it.
template <int kMode>
struct GemmCfg {
  static constexpr int CG = kMode == 2 ? 2 : 1;        // MMA cta_group
  static constexpr int CLUSTER = kMode == 1 ? 1 : 2;   // CTAs per cluster = 128-row blocks per tile
  static constexpr int STAGES = kMode == 2 ? 6 : 4;
  static constexpr int BNL = GEMM_BN / CG;             // rows of B resident per CTA
  static constexpr uint32_t A_BYTES = GEMM_BM_CTA * GEMM_BK * 2;
  static constexpr uint32_t B_BYTES = BNL * GEMM_BK * 2;
  static constexpr uint32_t SMEM_BYTES =
      STAGES * (A_BYTES + B_BYTES) + 256 + GEMM_EPI_WARPS * GEMM_STAGE_BYTES_PER_WARP + 1024;
};

template <int kMode, bool kAMN, bool kBMN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ GemmParams p) {
  // (__grid_constant__: the out-of-line slow-path epilogue takes `p` by reference; without it the whole struct was
  // copied to the stack and EVERY field access in the epilogue loops became a local-memory load)
  using Cfg = GemmCfg<kMode>;
  constexpr int kCG = Cfg::CG;
  constexpr int kCluster = Cfg::CLUSTER;
  constexpr int STAGES = Cfg::STAGES;
  constexpr uint32_t A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES;

  // 1024-byte alignment (128B-swizzle atoms) is requested from the toolchain, so every smem address below is a
  // link-time constant instead of a live register (the run-time round-up cost registers / spill reloads in the loops)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(bars) + 256;  // 4 x 4 KB, 128-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCluster == 2) ? cluster_ctarank() : 0u;
  const int cluster_id = (kCluster == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_clusters = (kCluster == 2) ? (gridDim.x >> 1) : gridDim.x;
  const int total_tiles = p.num_m * p.num_n * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      // mode 2: ONE arrival (the leader's expect_tx covering the bytes of both CTAs); the peer's TMA loads signal
      // the leader's barrier through their transaction bytes alone (an explicit remote arrive per k-block needed a
      // release fence = MEMBAR + ERRBAR in the peer's producer and halved the kernel's throughput)
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kMode == 3 ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], GEMM_EPI_WARPS * kCG);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<kCG>(tmem_slot, 512);
  }
  tc_fence_before();
  if constexpr (kCluster == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  // shuffled from lane 0 so that the compiler knows the value is warp-uniform (tcgen05 operands in uniform registers)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(tmem_slot), 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const TileCoord tc = tile_coord(tile, p.num_m, p.num_n);
        const int m0 = tc.m_blk * (GEMM_BM_CTA * kCluster) + static_cast<int>(cta_rank) * GEMM_BM_CTA;
        const int n0 = tc.n_blk * GEMM_BN + (kMode == 2 ? static_cast<int>(cta_rank) * Cfg::BNL : 0);
        const int kb0 = tc.split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_dst = sA + stage * A_BYTES;
          uint8_t* b_dst = sB + stage * B_BYTES;
          const int k0 = kb * GEMM_BK;
          if constexpr (kMode == 3) {
            mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
            if constexpr (!kAMN) {
              tma_load_2d(&tmA, &full[stage], a_dst, k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < GEMM_BM_CTA / 64; ++j) tma_load_2d(&tmA, &full[stage], a_dst + j * 8192, m0 + j * 64, k0);
            }
            // this CTA's half of the B tile, multicast into both CTAs (same smem offset, each CTA's own barrier)
            if constexpr (!kBMN) {
              tma_load_2d_mc(&tmB, &full[stage], b_dst + cta_rank * 16384, k0, n0 + static_cast<int>(cta_rank) * 128, 0x3);
            } else {
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const int j = static_cast<int>(cta_rank) * 2 + jj;
                tma_load_2d_mc(&tmB, &full[stage], b_dst + j * 8192, n0 + j * 64, k0, 0x3);
              }
            }
          } else if constexpr (kCG == 1) {
            mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
            if constexpr (!kAMN) {
              tma_load_2d(&tmA, &full[stage], a_dst, k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < GEMM_BM_CTA / 64; ++j) tma_load_2d(&tmA, &full[stage], a_dst + j * 8192, m0 + j * 64, k0);
            }
            if constexpr (!kBMN) {
              tma_load_2d(&tmB, &full[stage], b_dst, k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < Cfg::BNL / 64; ++j) tma_load_2d(&tmB, &full[stage], b_dst + j * 8192, n0 + j * 64, k0);
            }
          } else {
            if (cta_rank == 0) mbar_expect_tx(&full[stage], (A_BYTES + B_BYTES) * 2);
            if constexpr (!kAMN) {
              tma_load_2d_2sm(&tmA, &full[stage], a_dst, k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < GEMM_BM_CTA / 64; ++j)
                tma_load_2d_2sm(&tmA, &full[stage], a_dst + j * 8192, m0 + j * 64, k0);
            }
            if constexpr (!kBMN) {
              tma_load_2d_2sm(&tmB, &full[stage], b_dst, k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < Cfg::BNL / 64; ++j)
                tma_load_2d_2sm(&tmB, &full[stage], b_dst + j * 8192, n0 + j * 64, k0);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp runs the warp-uniform control flow and waits on the barriers; one elected lane issues. With
    // `if (lane == 0)` around the loop the compiler wrapped every tcgen05.mma in a per-lane R2UR/ELECT waterfall loop
    // (~13 instructions per MMA); with provably uniform operands the MMAs of a k-block issue back to back.
    if (kMode != 2 || cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM_CTA * kCG, GEMM_BN, kAMN, kBMN);
      constexpr uint32_t A_LBO = kAMN ? GEMM_BK * 128 : 0;
      constexpr uint32_t B_LBO = kBMN ? GEMM_BK * 128 : 0;
      constexpr uint32_t A_KSTEP = kAMN ? 16 * 128 : 32;  // bytes per UMMA_K=16 step
      constexpr uint32_t B_KSTEP = kBMN ? 16 * 128 : 32;
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
        const TileCoord tc = tile_coord(tile, p.num_m, p.num_n);
        const int kb0 = tc.split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int acc_stage = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_stage * GEMM_BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (leader) {
            const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
            const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              const uint64_t a_desc = make_smem_desc_sw128(a_addr + k * A_KSTEP, A_LBO, 1024);
              const uint64_t b_desc = make_smem_desc_sw128(b_addr + k * B_KSTEP, B_LBO, 1024);
              umma_bf16_ss<kCG>(d_tmem, a_desc, b_desc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if constexpr (kMode == 1) umma_commit(&empty[stage]);
            else if constexpr (kMode == 2) umma_commit_2sm(&empty[stage], 0x3);
            else umma_commit_mc(&empty[stage], 0x3);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader) {
          if constexpr (kCG == 1) umma_commit(&tfull[acc_stage]); else umma_commit_2sm(&tfull[acc_stage], 0x3);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // 8 epilogue warps: two per TMEM lane quadrant (a warp may only touch lanes 32*(warp%4)..+31); the pair splits
    // the 256 accumulator columns in halves, which keeps the GELU / dropout epilogues shorter than the main loop.
    const int q = warp & 3;
    const int ew = warp - 2;       // 0..7
    const int chalf = ew >> 2;     // which 128-column half this warp stores
    uint8_t* stage = epi_stage + ew * GEMM_STAGE_BYTES_PER_WARP;
    int it = 0;
    // second-operand slice of a tile -> L2, one tile ahead of its use (see epilogue_prefetch_aux)
    auto prefetch_tile = [&](int t) {
      if (t >= total_tiles) return;
      const TileCoord tp = tile_coord(t, p.num_m, p.num_n);
      const long long rb = static_cast<long long>(tp.m_blk) * (GEMM_BM_CTA * kCluster) + cta_rank * GEMM_BM_CTA + q * 32;
      epilogue_prefetch_aux(p, rb + lane, tp.n_blk * GEMM_BN + chalf * 128);
    };
    prefetch_tile(cluster_id);
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
      prefetch_tile(tile + num_clusters);
      const TileCoord tc = tile_coord(tile, p.num_m, p.num_n);
      const int acc_stage = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long row_base =
          static_cast<long long>(tc.m_blk) * (GEMM_BM_CTA * kCluster) + cta_rank * GEMM_BM_CTA + q * 32;
      mbar_wait(&tfull[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_stage * GEMM_BN;
      epilogue_warp_tile(p, taddr, stage, lane, row_base, tc.n_blk * GEMM_BN, tc.split, chalf * 2, chalf * 2 + 2);
      // release this accumulator stage back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // (TMEM reads are ordered by the tcgen05 fences; the remote arrive itself needs no release fence)
        if constexpr (kCG == 1) mbar_arrive(&tempty[acc_stage]); else mbar_arrive_cluster_relaxed(&tempty[acc_stage], 0);
      }
    }
  }

  tc_fence_before();
  if constexpr (kCluster == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<kCG>(tmem_base, 512);
  }
}

// out[m,n] = rb( (accumulate ? float(out[m,n]) : 0) + float(rb(sum_s partial[s][m][n])) )
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, __nv_bfloat16* __restrict__ out, long long ldd,
                                     int M, int N, int splits, int accumulate) {
  const long long total4 = static_cast<long long>(M) * N / 4;
  const long long stride = static_cast<long long>(M) * N;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(partial)[i];
    for (int s = 1; s < splits; ++s) {
      float4 t = reinterpret_cast<const float4*>(partial + s * stride)[i];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    const long long e = i * 4;
    const long long m = e / N;
    const int n = static_cast<int>(e - m * N);
    __nv_bfloat16* dst = out + m * ldd + n;
    float o[4] = {rb(acc.x), rb(acc.y), rb(acc.z), rb(acc.w)};
    if (accumulate) {
      uint2 old = *reinterpret_cast<const uint2*>(dst);
      o[0] += bf16_lo(old.x); o[1] += bf16_hi(old.x); o[2] += bf16_lo(old.y); o[3] += bf16_hi(old.y);
    }
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
  }
}

template <int kMode, bool kAMN, bool kBMN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<kMode>;
  constexpr int kCG = Cfg::CLUSTER;  // CTAs per cluster
  auto kern = gemm_bf16_kernel<kMode, kAMN, kBMN>;
  static bool attr_set = false;  // benign race: idempotent
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(gemm smem=%u): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return OBT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int total_tiles = p.num_m * p.num_n * p.splits;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // Persistent grid = number of CTAs (clusters) that can be CO-RESIDENT. For CTA pairs this can be fewer than
  // SMs/2 (a GPC with an odd number of enabled SMs strands one): launching more would run a second wave and
  // double the kernel time (first cta_group::2 measurements, profiles/r01_*).
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = sm_count() / kCG;
    if (kCG > 1) {
      cfg.gridDim = dim3(sm_count() / kCG * kCG);
      int q = 0;
      if (cudaOccupancyMaxActiveClusters(&q, kern, &cfg) == cudaSuccess && q > 0 && q < n) n = q;
      (void)cudaGetLastError();
    }
    max_clusters = n;
  }
  const int clusters = total_tiles < max_clusters ? total_tiles : max_clusters;
  cfg.gridDim = dim3(clusters * kCG);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
  if (e != cudaSuccess) {
    set_last_error("gemm launch: %s", cudaGetErrorString(e));
    return OBT_ERR_CUDA;
  }
  return OBT_OK;
}

static int g_force_cta_group = 0;  // 0 = auto, 1 / 2 / 3 = forced kernel mode (tests / ablations)

}  // namespace obt

using namespace obt;

extern "C" void obt_gemm_set_cta_group(int cg) { obt::g_force_cta_group = cg; }

// Number of fp32 elements of workspace obt_gemm_bf16 wants for this problem when split-K is allowed.
extern "C" long long obt_gemm_workspace_elems(long long M, long long N, long long K) {
  (void)K;
  // at most 8 splits
  return 8 * M * N;
}

extern "C" int obt_gemm_bf16(const void* A, const void* B, void* D, long long M, long long N, long long K,
                             long long lda, long long ldb, long long ldd, int a_mn_major, int b_mn_major,
                             int epilogue, const void* aux_in, long long ld_aux_in, void* aux_out,
                             long long ld_aux_out, int gelu_mode, float drop_p, unsigned long long seed,
                             unsigned long long offset, void* workspace, long long workspace_elems,
                             const float* rope_cos, const float* rope_sin, int rope_T, int rope_head_dim,
                             int rope_cols, cudaStream_t stream) {
  OBT_REQUIRE(A && B && D, "obt_gemm_bf16: null operand");
  OBT_REQUIRE(M > 0 && N > 0 && K > 0, "obt_gemm_bf16: empty problem M=%lld N=%lld K=%lld", M, N, K);
  OBT_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "obt_gemm_bf16: dims exceed int32");
  OBT_REQUIRE(((epilogue >= EPI_PLAIN && epilogue <= EPI_RESID_DROPOUT) || (epilogue >= EPI_ROPE && epilogue <= EPI_DELTA)) &&
                  epilogue != EPI_PARTIAL,
              "obt_gemm_bf16: bad epilogue %d", epilogue);
  if (epilogue == EPI_DELTA)
    // workspace = delta fp32 [M / rope_T, N / 128, rope_T]; the vectorised epilogue path is required (aligned, N % 8)
    OBT_REQUIRE(aux_in != nullptr && workspace != nullptr && rope_T > 0 && M % rope_T == 0 && N % 128 == 0 &&
                    workspace_elems >= M * (N / 128) && ld_aux_in % 8 == 0 && ldd % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(aux_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0,
                "obt_gemm_bf16: epilogue 11 needs aux_in (y), workspace = delta [B, N/128, T] and rope_T = T");
  if (epilogue == EPI_ROPE)
    OBT_REQUIRE(rope_cos != nullptr && rope_T > 0 && rope_head_dim > 0 && rope_head_dim % 8 == 0 && rope_cols % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(rope_cos) & 15) == 0 && (reinterpret_cast<uintptr_t>(rope_sin) & 15) == 0,
                "obt_gemm_bf16: rotary epilogue needs 16-byte aligned fp32 tables, head_dim %% 8 == 0 (got %d), "
                "rope_cols %% 8 == 0 (got %d)", rope_head_dim, rope_cols);
  if (epilogue == EPI_RESID || epilogue == EPI_GELU_BWD || epilogue == EPI_RESID_DROPOUT || epilogue == EPI_ROWMASK ||
      epilogue == EPI_MUL)
    OBT_REQUIRE(aux_in != nullptr, "obt_gemm_bf16: epilogue %d needs aux_in", epilogue);
  if (epilogue == EPI_GELU || epilogue == EPI_GELU_DG)
    OBT_REQUIRE(aux_out != nullptr, "obt_gemm_bf16: GELU epilogue needs aux_out");

  int cg = g_force_cta_group;
  // auto: one CTA per SM with 128x256 tiles measured 1.17-1.40 PFLOP/s on the block/head shapes (86-95 % of cuBLAS);
  // the CTA-pair variant stays selectable for ablations (obt_gemm_set_cta_group).
  // mode 3 = cluster of two such CTAs sharing (multicasting) the B tile.
  // auto: CTA pairs (cta_group::2, 256 x 256 tiles, each CTA holding half of B) measured 1.29-1.47 PFLOP/s on every
  // block / head shape (profiles/r01_gemm_modes_v3.txt), level with cuBLAS and 3-15 % above the one-CTA (mode 1) and
  // multicast-cluster (mode 3) kernels, which stay selectable for ablations (obt_gemm_set_cta_group).
  if (cg == 0) cg = 2;
  if (cg == 2 && M <= GEMM_BM_CTA) cg = 1;  // a single 128-row block has no partner
  if (cg == 3 && M <= GEMM_BM_CTA) cg = 1;  // a single row block has no partner to share B with
  const int cluster = (cg == 1) ? 1 : 2;     // 128-row blocks per tile / CTAs per cluster
  const int bm = GEMM_BM_CTA * cluster;

  GemmParams p = {};
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.num_m = static_cast<int>((M + bm - 1) / bm);
  p.num_n = static_cast<int>((N + GEMM_BN - 1) / GEMM_BN);
  p.num_kb = static_cast<int>((K + GEMM_BK - 1) / GEMM_BK);
  p.epi = (epilogue == EPI_GELU && gelu_mode == 1) ? EPI_GELU_EAGER : epilogue;
  p.gelu_mode = gelu_mode;
  p.D = static_cast<__nv_bfloat16*>(D);
  p.ldd = ldd;
  p.aux_in = static_cast<const __nv_bfloat16*>(aux_in);
  p.ld_aux_in = ld_aux_in;
  p.aux_out = static_cast<__nv_bfloat16*>(aux_out);
  p.ld_aux_out = ld_aux_out;
  p.drop_p = drop_p;
  p.drop_thr = static_cast<uint32_t>(ceilf(drop_p * 16777216.0f));
  p.seed = seed;
  p.offset = offset;
  p.rope_cos = rope_cos;
  p.rope_sin = rope_sin;
  p.rope_T = rope_T;
  p.rope_d = rope_head_dim;
  p.rope_cols = rope_cols;
  p.rope_T_mask = (rope_T > 0 && (rope_T & (rope_T - 1)) == 0) ? rope_T - 1 : -1;
  p.rope_d_mask = (rope_head_dim > 0 && (rope_head_dim & (rope_head_dim - 1)) == 0) ? rope_head_dim - 1 : -1;
  p.delta = epilogue == EPI_DELTA ? static_cast<float*>(workspace) : nullptr;
  p.delta_T = rope_T;
  auto aligned16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  p.vec_ok = (N % 8 == 0) && (ldd % 8 == 0) && aligned16(D) &&
             (aux_in == nullptr || epilogue == EPI_ROWMASK || (ld_aux_in % 8 == 0 && aligned16(aux_in))) &&
             (aux_out == nullptr || (ld_aux_out % 8 == 0 && aligned16(aux_out)));

  // split-K when the output tile grid cannot fill the machine and the reduction is long (wgrad shapes).
  p.splits = 1;
  const int tiles = p.num_m * p.num_n;
  const int slots = sm_count() / cluster;
  const bool splittable = (epilogue == EPI_PLAIN || (epilogue == EPI_RESID && aux_in == D && ld_aux_in == ldd)) &&
                          workspace != nullptr && (M * N) % 4 == 0 && (N % 4 == 0) && (ldd % 4 == 0);
  if (splittable && tiles < slots && p.num_kb >= 16) {
    // Pick the split count with the smallest estimated time: tensor time / wave efficiency + the fp32 partials'
    // round trip through HBM. (c_attn's weight gradient, 48 tiles on 74 CTA pairs, ran at 65 % occupancy unsplit:
    // 3 splits = 144 tiles fill 1.95 waves.)
    const double t_math = 2.0 * static_cast<double>(M) * N * K / 1.45e15;
    double best = 1e30;
    int best_s = 1;
    for (int s = 1; s <= 8; ++s) {
      if (s > 1 && ((p.num_kb / s) < 4 || static_cast<long long>(s) * M * N > workspace_elems)) break;
      const int work = tiles * s;
      const int waves = (work + slots - 1) / slots;
      const double eff = static_cast<double>(work) / (static_cast<double>(waves) * slots);
      const double t_io = s > 1 ? (2.0 * s * static_cast<double>(M) * N * 4.0) / 5.0e12 + 4e-6 : 0.0;
      const double t = t_math / eff + t_io;
      if (t < best * 0.98) {  // prefer fewer splits unless the gain is real
        best = t;
        best_s = s;
      }
    }
    p.splits = best_s;
  }
  p.kb_per_split = (p.num_kb + p.splits - 1) / p.splits;
  // all splits must be non-empty
  while (p.splits > 1 && (p.splits - 1) * p.kb_per_split >= p.num_kb) {
    --p.splits;
    p.kb_per_split = (p.num_kb + p.splits - 1) / p.splits;
  }
  int final_epi = epilogue;
  if (p.splits > 1) {
    p.partial = static_cast<float*>(workspace);
    p.epi = EPI_PARTIAL;
  }

  CUtensorMap tmA, tmB;
  int rc;
  const int bnl = GEMM_BN / cluster;  // B rows per TMA box: modes 2 and 3 load half of the tile per CTA
  if (!a_mn_major) {
    OBT_REQUIRE(lda % 8 == 0, "obt_gemm_bf16: lda=%lld must be a multiple of 8 elements", lda);
    rc = get_tensor_map_2d(&tmA, A, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(lda), 64,
                           GEMM_BM_CTA);
  } else {
    OBT_REQUIRE(lda % 8 == 0, "obt_gemm_bf16: lda=%lld must be a multiple of 8 elements", lda);
    rc = get_tensor_map_2d(&tmA, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), 64,
                           64);
  }
  if (rc != OBT_OK) return rc;
  OBT_REQUIRE(ldb % 8 == 0, "obt_gemm_bf16: ldb=%lld must be a multiple of 8 elements", ldb);
  if (!b_mn_major) {
    rc = get_tensor_map_2d(&tmB, B, static_cast<uint64_t>(K), static_cast<uint64_t>(N), static_cast<uint64_t>(ldb), 64,
                           static_cast<uint32_t>(bnl));
  } else {
    rc = get_tensor_map_2d(&tmB, B, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldb), 64,
                           64);
  }
  if (rc != OBT_OK) return rc;

#define OBT_LAUNCH(CG, AM, BM_)                                                  \
  if (cg == CG && (a_mn_major != 0) == AM && (b_mn_major != 0) == BM_) {          \
    rc = launch_gemm<CG, AM, BM_>(tmA, tmB, p, stream);                          \
  } else
  OBT_LAUNCH(1, false, false)
  OBT_LAUNCH(1, false, true)
  OBT_LAUNCH(1, true, false)
  OBT_LAUNCH(1, true, true)
  OBT_LAUNCH(2, false, false)
  OBT_LAUNCH(2, false, true)
  OBT_LAUNCH(2, true, false)
  OBT_LAUNCH(2, true, true)
  OBT_LAUNCH(3, false, false)
  OBT_LAUNCH(3, false, true)
  OBT_LAUNCH(3, true, false)
  OBT_LAUNCH(3, true, true) {
    set_last_error("obt_gemm_bf16: no kernel for cta_group=%d", cg);
    rc = OBT_ERR_UNSUPPORTED;
  }
#undef OBT_LAUNCH
  if (rc != OBT_OK) return rc;

  if (p.splits > 1) {
    const long long total4 = M * N / 4;
    int blocks = static_cast<int>((total4 + 255) / 256);
    const int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p.partial, p.D, ldd, p.M, p.N, p.splits,
                                                    final_epi == EPI_RESID ? 1 : 0);
    return check_launch("splitk_reduce");
  }
  return OBT_OK;
}
