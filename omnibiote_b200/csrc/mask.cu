// Attention-mask producers / compressors for the hot path's input contract.
//
//  * obt_doc_mask_intervals : per-token visible key interval [lo,hi) straight from token ids, reproducing the
//    reference builder `create_attention_mask` (training/train_encoder.py:25-57) including its quirks: for every
//    batch row except the one handled first by the `if` branch (row 0), the first two documents are merged; with
//    padding=False an EOS is virtually appended; with padding=True a row without any EOS attends everywhere and
//    tokens after the last EOS are fully masked (finite -1e9 everywhere => uniform attention).
//  * obt_pad_mask_intervals : `pad_attn` of evals/gue.py:15-21 (rows/cols after first_pad+1 masked).
//  * obt_mask_from_intervals: materialises the dense additive bf16 (B,T,T) tensor {0, -1e9} for drop-in callers.
//  * obt_mask_compress      : the inverse: dense additive mask -> intervals + a flag telling whether the mask is
//    exactly interval-structured with values {0, bf16(-1e9)} (otherwise the dense path must be used).
#include "common.cuh"
#include "ptx.cuh"

namespace obt {

__global__ void doc_mask_intervals_kernel(const long long* __restrict__ ids, int* __restrict__ lo, int* __restrict__ hi,
                                          int T, long long eos, int padding) {
  extern __shared__ int s_eos[];  // positions of EOS tokens in this row (at most T+1)
  __shared__ int s_n;
  const int b = blockIdx.x;
  const long long* row = ids + static_cast<long long>(b) * T;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int t = 0; t < T; ++t)
      if (row[t] == eos) s_eos[n++] = t;
    if (!padding) s_eos[n++] = T;  // virtual EOS appended at position T (train_encoder.py:33-37)
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    int l = 0, h = 0;
    if (n == 0) {
      l = 0; h = T;  // no EOS in the row: everything visible (train_encoder.py:53-55)
    } else {
      // k = index of the first EOS at or after i
      int a = 0, c = n;
      while (a < c) {
        int m = (a + c) >> 1;
        if (s_eos[m] >= i) c = m; else a = m + 1;
      }
      const int k = a;
      if (k < n) {
        l = (k == 0) ? 0 : s_eos[k - 1] + 1;
        h = min(s_eos[k] + 1, T);
        if (b > 0 && k <= 1) {
          // quirk: rows entered through the `else` branch never advance prev_index on their first EOS, so the
          // block of the second EOS starts at 0 again: documents 0 and 1 are merged (train_encoder.py:44-51)
          l = 0;
          h = min(s_eos[n > 1 ? 1 : 0] + 1, T);
        }
      }  // else: after the last EOS -> fully masked row (l == h == 0)
    }
    lo[static_cast<long long>(b) * T + i] = l;
    hi[static_cast<long long>(b) * T + i] = h;
  }
}

__global__ void pad_mask_intervals_kernel(const long long* __restrict__ ids, int* __restrict__ lo, int* __restrict__ hi,
                                          int T, long long pad) {
  __shared__ int s_first;
  const int b = blockIdx.x;
  const long long* row = ids + static_cast<long long>(b) * T;
  if (threadIdx.x == 0) s_first = T;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    if (row[t] == pad) atomicMin(&s_first, t);
  __syncthreads();
  // gue.py:18-19: rows and columns from first_pad + 1 on are masked; the first PAD itself stays attended.
  const int vis = min(T, s_first + 1);
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const bool live = i < vis;
    lo[static_cast<long long>(b) * T + i] = 0;
    hi[static_cast<long long>(b) * T + i] = live ? vis : 0;
  }
}

// mask[b,i,j] = (lo <= j < hi) ? 0 : -1e9 (bf16: 0xCE6E); 8 elements per thread
__global__ void mask_from_intervals_kernel(const int* __restrict__ lo, const int* __restrict__ hi,
                                           __nv_bfloat16* __restrict__ mask, long long rows, int T) {
  const int chunks = T / 8;
  const long long total = rows * chunks;
  const uint32_t NEG = 0xCE6Eu;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = e / chunks;
    const int j0 = static_cast<int>(e - r * chunks) * 8;
    const int l = lo[r], h = hi[r];
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + 2 * k;
      const uint32_t a = (j >= l && j < h) ? 0u : NEG;
      const uint32_t c = (j + 1 >= l && j + 1 < h) ? 0u : NEG;
      w[k] = a | (c << 16);
    }
    reinterpret_cast<uint4*>(mask + r * T)[j0 / 8] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// One warp per (b,i) row of a dense additive bf16 mask: find [lo,hi) of zeros and verify the structure.
__global__ void mask_compress_kernel(const __nv_bfloat16* __restrict__ mask, long long msb, long long msq,
                                     int* __restrict__ lo, int* __restrict__ hi, int* __restrict__ not_interval, int B,
                                     int T) {
  const int warps = blockDim.x >> 5;
  const long long r = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5);
  if (r >= static_cast<long long>(B) * T) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(r / T), i = static_cast<int>(r % T);
  const unsigned short* row = reinterpret_cast<const unsigned short*>(mask + b * msb + i * msq);
  int first = T, last = -1, zeros = 0, bad = 0;
  for (int j = lane; j < T; j += 32) {
    const unsigned short v = row[j];
    if (v == 0 || v == 0x8000u) {
      first = min(first, j);
      last = max(last, j);
      ++zeros;
    } else if (v != 0xCE6Eu) {
      bad = 1;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (lane == 0) {
    if (zeros == 0) {
      lo[r] = 0; hi[r] = 0;
    } else {
      lo[r] = first; hi[r] = last + 1;
      if (zeros != last + 1 - first) bad = 1;
    }
    if (bad) atomicExch(not_interval, 1);
  }
}

}  // namespace obt

using namespace obt;

extern "C" int obt_doc_mask_intervals(const long long* ids, int* lo, int* hi, int B, int T, long long eos_token,
                                      int padding, cudaStream_t stream) {
  OBT_REQUIRE(ids && lo && hi, "obt_doc_mask_intervals: null pointer");
  OBT_REQUIRE(B > 0 && T > 0 && T <= 32768, "obt_doc_mask_intervals: bad shape B=%d T=%d", B, T);
  doc_mask_intervals_kernel<<<B, 256, (T + 1) * sizeof(int), stream>>>(ids, lo, hi, T, eos_token, padding);
  return check_launch("doc_mask_intervals");
}

extern "C" int obt_pad_mask_intervals(const long long* ids, int* lo, int* hi, int B, int T, long long pad_token,
                                      cudaStream_t stream) {
  OBT_REQUIRE(ids && lo && hi, "obt_pad_mask_intervals: null pointer");
  OBT_REQUIRE(B > 0 && T > 0, "obt_pad_mask_intervals: bad shape");
  pad_mask_intervals_kernel<<<B, 256, 0, stream>>>(ids, lo, hi, T, pad_token);
  return check_launch("pad_mask_intervals");
}

extern "C" int obt_mask_from_intervals(const int* lo, const int* hi, void* mask, int B, int T, cudaStream_t stream) {
  OBT_REQUIRE(lo && hi && mask, "obt_mask_from_intervals: null pointer");
  OBT_REQUIRE(T % 8 == 0, "obt_mask_from_intervals: T=%d must be a multiple of 8", T);
  const long long rows = static_cast<long long>(B) * T;
  const long long total = rows * (T / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  mask_from_intervals_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(lo, hi, static_cast<__nv_bfloat16*>(mask),
                                                                              rows, T);
  return check_launch("mask_from_intervals");
}

// mask: additive bf16 (b, i, j) with strides (msb, msq, 1). not_interval (device int, pre-zeroed) is set to 1 when any
// row is not {0,-1e9}-valued with a single contiguous run of zeros.
extern "C" int obt_mask_compress(const void* mask, long long msb, long long msq, int* lo, int* hi, int* not_interval,
                                 int B, int T, cudaStream_t stream) {
  OBT_REQUIRE(mask && lo && hi && not_interval, "obt_mask_compress: null pointer");
  const long long rows = static_cast<long long>(B) * T;
  const int warps = 8;
  mask_compress_kernel<<<static_cast<unsigned>((rows + warps - 1) / warps), warps * 32, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(mask), msb, msq, lo, hi, not_interval, B, T);
  return check_launch("mask_compress");
}

// ---------------------------------------------------------------------------------------------
// MLM input masking on device (train_encoder.py:273-279): mask = Bernoulli(p) & id != PAD & id != EOS;
// masked inputs get MASK_TOKEN (no 80/10/10 split in the reference). Philox instead of numpy's host RNG.
// ---------------------------------------------------------------------------------------------
namespace obt {
__global__ void mlm_mask_kernel(const long long* __restrict__ ids, long long* __restrict__ masked,
                                unsigned char* __restrict__ mask, long long n, float prob, unsigned long long seed,
                                unsigned long long offset, long long pad, long long eos, long long mask_token,
                                float* __restrict__ counters) {
  int n_masked = 0, n_tokens = 0;
  for (long long i4 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i4 * 4 < n;
       i4 += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 r = philox4x32(seed, static_cast<unsigned long long>(i4), offset);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long i = i4 * 4 + e;
      if (i >= n) break;
      const long long id = ids[i];
      const bool m = ((w[e] >> 8) * (1.0f / 16777216.0f) < prob) && id != pad && id != eos;
      mask[i] = m ? 1 : 0;
      masked[i] = m ? mask_token : id;
      n_masked += m ? 1 : 0;
      n_tokens += (id != pad) ? 1 : 0;
    }
  }
  // step bookkeeping of train_encoder.py:350 (`(input_ids != PAD_TOKEN).sum()`) without a host round trip: integer
  // counts accumulated in fp32 (exact below 2^24, so the atomic order does not matter)
  if (counters != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_masked += __shfl_xor_sync(0xffffffffu, n_masked, o);
      n_tokens += __shfl_xor_sync(0xffffffffu, n_tokens, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (n_masked) atomicAdd(counters + 0, static_cast<float>(n_masked));
      if (n_tokens) atomicAdd(counters + 1, static_cast<float>(n_tokens));
    }
  }
}
}  // namespace obt

extern "C" int obt_mlm_mask(const long long* ids, long long* masked_ids, unsigned char* mask, long long n, float prob,
                            unsigned long long seed, unsigned long long offset, long long pad_token, long long eos_token,
                            long long mask_token, float* counters, cudaStream_t stream) {
  OBT_REQUIRE(ids && masked_ids && mask, "obt_mlm_mask: null pointer");
  OBT_REQUIRE(prob >= 0.f && prob <= 1.f, "obt_mlm_mask: prob=%f", prob);
  if (n == 0) return OBT_OK;
  long long blocks = ((n + 3) / 4 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  obt::mlm_mask_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(ids, masked_ids, mask, n, prob, seed, offset,
                                                                        pad_token, eos_token, mask_token, counters);
  return check_launch("mlm_mask");
}

// ---------------------------------------------------------------------------------------------
// Tile metadata of an interval mask, computed ONCE per micro-batch and shared by every layer, head and attention
// kernel (forward, dQ, dK/dV): each of their CTAs used to re-derive it from the per-row intervals in its prologue
// (loads + shared-memory atomics + a block barrier in front of the first TMA / MMA; ~45 % of a backward CTA's life is
// such fixed latency, profiles/r02c_attn_dq_w8.source.txt).
//   qmeta[b][tq] = {min lo, max hi, any fully-masked row, 0} over the 128 query rows of tile tq
//   kmeta[b][tk] = 128 relevance bits: bit it set <=> the 64-query sub-tile `it` has a row that sees a key of the
//                  128-key tile tk, or a fully-masked row (those attend to every key)
// ---------------------------------------------------------------------------------------------
namespace obt {
__global__ void attn_tile_meta_kernel(const int* __restrict__ row_lo, const int* __restrict__ row_hi, int T,
                                      int* __restrict__ qmeta, unsigned int* __restrict__ kmeta) {
  __shared__ int s_q[3];
  __shared__ unsigned int s_rel[4];
  const int tile = blockIdx.x, b = blockIdx.y, nT = gridDim.x;
  if (threadIdx.x == 0) {
    s_q[0] = T; s_q[1] = 0; s_q[2] = 0;
    s_rel[0] = s_rel[1] = s_rel[2] = s_rel[3] = 0u;
  }
  __syncthreads();
  const int* lo = row_lo + static_cast<long long>(b) * T;
  const int* hi = row_hi + static_cast<long long>(b) * T;
  const int i = tile * 128 + threadIdx.x;
  if (threadIdx.x < 128 && i < T) {
    const int l = lo[i], h = hi[i];
    if (l >= h) {
      atomicExch(&s_q[2], 1);
    } else {
      atomicMin(&s_q[0], l);
      atomicMax(&s_q[1], h);
    }
  }
  const int j0 = tile * 128;
  for (int r = threadIdx.x; r < T; r += blockDim.x) {
    const int l = lo[r], h = hi[r];
    if ((l >= h) || (l < j0 + 128 && h > j0)) atomicOr(&s_rel[(r >> 6) >> 5], 1u << ((r >> 6) & 31));
  }
  __syncthreads();
  const long long o = (static_cast<long long>(b) * nT + tile) * 4;
  if (threadIdx.x < 4) {
    qmeta[o + threadIdx.x] = threadIdx.x < 3 ? s_q[threadIdx.x] : 0;
    kmeta[o + threadIdx.x] = s_rel[threadIdx.x];
  }
}
}  // namespace obt

extern "C" int obt_attn_tile_meta(const int* row_lo, const int* row_hi, int B, int T, int* qmeta, unsigned int* kmeta,
                                  cudaStream_t stream) {
  OBT_REQUIRE(row_lo && row_hi && qmeta && kmeta, "obt_attn_tile_meta: null pointer");
  OBT_REQUIRE(B > 0 && T > 0 && T <= 128 * 64, "obt_attn_tile_meta: bad B=%d T=%d", B, T);
  dim3 grid((T + 127) / 128, B);
  obt::attn_tile_meta_kernel<<<grid, 256, 0, stream>>>(row_lo, row_hi, T, qmeta, kmeta);
  return check_launch("attn_tile_meta");
}

// ---------------------------------------------------------------------------------------------
// Row compaction for the masked-rows-only head (d loss / d logits is exactly zero on the ~85 % of
// rows outside the MLM mask, so the head GEMMs, the CE and their backward only need the masked rows).
//   compact_rows : stable list of the rows with mask != 0 -> idx[0..count), padded with -1 up to `cap`; the targets of
//                  those rows and a validity byte per slot; meta = {count, overflow (count > cap)}.
//   gather_rows  : dst[s] = src[idx[s]] (zeros for idx < 0);  scatter_rows : dst = 0, then dst[idx[s]] = src[s].
// ---------------------------------------------------------------------------------------------
namespace obt {

constexpr int COMPACT_THREADS = 1024;

__global__ void __launch_bounds__(COMPACT_THREADS, 1)
compact_rows_kernel(const unsigned char* __restrict__ mask, const long long* __restrict__ targets, long long M, int cap,
                    int* __restrict__ idx, long long* __restrict__ tgt_c, unsigned char* __restrict__ valid_c,
                    int* __restrict__ meta) {
  __shared__ int warp_tot[COMPACT_THREADS / 32];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  // chunks of COMPACT_THREADS rows, processed in order so that the output order is the row order (deterministic sums)
  for (long long c0 = 0; c0 < M; c0 += COMPACT_THREADS) {
    const long long i = c0 + tid;
    const bool m = i < M && mask[i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    const int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < COMPACT_THREADS / 32; ++w) {
      const int t = warp_tot[w];
      if (w < warp) wbase += t;
      total += t;
    }
    const int base = base_s;
    if (m) {
      const int slot = base + wbase + before;
      if (slot < cap) {
        idx[slot] = static_cast<int>(i);
        tgt_c[slot] = targets[i];
        valid_c[slot] = 1;
      }
    }
    __syncthreads();
    if (tid == 0) base_s = base + total;
    __syncthreads();
  }
  const int count = base_s;
  for (int s = count + tid; s < cap; s += COMPACT_THREADS) {
    idx[s] = -1;
    tgt_c[s] = 0;
    valid_c[s] = 0;
  }
  if (tid == 0) {
    meta[0] = count;
    meta[1] = count > cap ? 1 : 0;
  }
}

// one warp per destination row, 16-byte chunks
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, const int* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ dst, long long ldd, int n_slots, int C) {
  const int warps = blockDim.x >> 5;
  const int s = blockIdx.x * warps + (threadIdx.x >> 5);
  if (s >= n_slots) return;
  const int lane = threadIdx.x & 31;
  const int r = idx[s];
  for (int c = lane; c < C / 8; c += 32) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r >= 0) v = reinterpret_cast<const uint4*>(src + static_cast<long long>(r) * lds)[c];
    reinterpret_cast<uint4*>(dst + static_cast<long long>(s) * ldd)[c] = v;
  }
}

__global__ void scatter_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, const int* __restrict__ idx,
                                    __nv_bfloat16* __restrict__ dst, long long ldd, int n_slots, int C) {
  const int warps = blockDim.x >> 5;
  const int s = blockIdx.x * warps + (threadIdx.x >> 5);
  if (s >= n_slots) return;
  const int lane = threadIdx.x & 31;
  const int r = idx[s];
  if (r < 0) return;
  for (int c = lane; c < C / 8; c += 32)
    reinterpret_cast<uint4*>(dst + static_cast<long long>(r) * ldd)[c] =
        reinterpret_cast<const uint4*>(src + static_cast<long long>(s) * lds)[c];
}

}  // namespace obt

extern "C" int obt_compact_rows(const unsigned char* mask, const long long* targets, long long M, int cap, int* idx,
                                long long* targets_c, unsigned char* valid_c, int* meta, cudaStream_t stream) {
  OBT_REQUIRE(mask && targets && idx && targets_c && valid_c && meta, "obt_compact_rows: null pointer");
  OBT_REQUIRE(M > 0 && M < (1ll << 31) && cap > 0, "obt_compact_rows: bad M=%lld cap=%d", M, cap);
  obt::compact_rows_kernel<<<1, obt::COMPACT_THREADS, 0, stream>>>(mask, targets, M, cap, idx, targets_c, valid_c, meta);
  return check_launch("compact_rows");
}

extern "C" int obt_gather_rows(const void* src, long long lds, const int* idx, void* dst, long long ldd, int n_slots,
                               int C, cudaStream_t stream) {
  OBT_REQUIRE(src && idx && dst, "obt_gather_rows: null pointer");
  OBT_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && n_slots > 0, "obt_gather_rows: C=%d must be a multiple of 8", C);
  const int warps = 8;
  obt::gather_rows_kernel<<<(n_slots + warps - 1) / warps, warps * 32, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(src), lds, idx, static_cast<__nv_bfloat16*>(dst), ldd, n_slots, C);
  return check_launch("gather_rows");
}

extern "C" int obt_scatter_rows(const void* src, long long lds, const int* idx, void* dst, long long ldd, long long M,
                                int n_slots, int C, cudaStream_t stream) {
  OBT_REQUIRE(src && idx && dst, "obt_scatter_rows: null pointer");
  OBT_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldd == C && n_slots > 0 && M > 0,
              "obt_scatter_rows: C=%d must be a multiple of 8 and dst contiguous", C);
  cudaError_t e = cudaMemsetAsync(dst, 0, static_cast<size_t>(M) * C * 2, stream);
  if (e != cudaSuccess) {
    set_last_error("obt_scatter_rows: memset: %s", cudaGetErrorString(e));
    return OBT_ERR_CUDA;
  }
  const int warps = 8;
  obt::scatter_rows_kernel<<<(n_slots + warps - 1) / warps, warps * 32, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(src), lds, idx, static_cast<__nv_bfloat16*>(dst), ldd, n_slots, C);
  return check_launch("scatter_rows");
}
