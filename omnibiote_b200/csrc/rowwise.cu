// Memory-bound row-wise kernels of the encoder block: embedding gather / scatter-add, LayerNorm forward / backward,
// rotary ("RoPE" / cosine-scale) application, dropout, encode() pooling.
// All of them are HBM-bound: 16-byte vector accesses, one warp per row, fp32 math, one bf16 rounding per reference
// rounding point.
#include "common.cuh"
#include "ptx.cuh"

namespace obt {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// keep/scale decision for element `e` of a flat tensor; 4 consecutive elements share one Philox call.
// keep <=> (r >> 8) * 2^-24 >= p  <=>  (r >> 8) >= ceil(p * 2^24): both sides of the float comparison are exact (a 24-bit
// integer scaled by a power of two, p scaled by a power of two), so the integer form selects the same elements as the
// GEMM epilogue's float form with one compare instead of convert + multiply + compare per element.
__device__ __forceinline__ uint32_t dropout_threshold(float p) { return static_cast<uint32_t>(ceilf(p * 16777216.0f)); }
__device__ __forceinline__ void dropout4(unsigned long long seed, unsigned long long offset, unsigned long long e4,
                                         float p, bool (&keep)[4]) {
  const uint32_t thr = dropout_threshold(p);  // loop-invariant: hoisted by the compiler
  uint4 r = rand4x32(seed, e4, offset);
  keep[0] = (r.x >> 8) >= thr;
  keep[1] = (r.y >> 8) >= thr;
  keep[2] = (r.z >> 8) >= thr;
  keep[3] = (r.w >> 8) >= thr;
}

// ---------------------------------------------------------------------------------------------
// Embedding gather (+ in-place dropout): model.py:241-242
// ---------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const long long* __restrict__ idx, const __nv_bfloat16* __restrict__ wte,
                                 __nv_bfloat16* __restrict__ out, long long M, int C, int V, float p,
                                 unsigned long long seed, unsigned long long offset, int* __restrict__ err) {
  const int warps_per_block = blockDim.x >> 5;
  const long long row = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  long long id = idx[row];
  if (id < 0 || id >= V) {
    if (lane == 0) atomicExch(err, 1);
    id = 0;
  }
  const uint4* src = reinterpret_cast<const uint4*>(wte + id * C);
  uint4* dst = reinterpret_cast<uint4*>(out + row * C);
  const float scale = 1.0f / (1.0f - p);
  for (int c = lane; c < C / 8; c += 32) {
    uint4 u = src[c];
    if (p > 0.f) {
      float f[8];
      unpack8(u, f);
      bool k0[4], k1[4];
      const unsigned long long e4 = (static_cast<unsigned long long>(row) * C + c * 8) >> 2;
      dropout4(seed, offset, e4, p, k0);
      dropout4(seed, offset, e4 + 1, p, k1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[j] = k0[j] ? f[j] * scale : 0.f;
        f[4 + j] = k1[j] ? f[4 + j] * scale : 0.f;
      }
      u = pack8(f);
    }
    dst[c] = u;
  }
}

// scatter-add of token-row gradients into an fp32 scratch (V,C) + touched flags
__global__ void embed_bwd_scatter_kernel(const long long* __restrict__ idx, const __nv_bfloat16* __restrict__ dout,
                                         float* __restrict__ scratch, int* __restrict__ touched, long long M, int C,
                                         int V, float p, unsigned long long seed, unsigned long long offset) {
  const int warps_per_block = blockDim.x >> 5;
  const long long row = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  long long id = idx[row];
  if (id < 0 || id >= V) return;
  if (lane == 0) touched[id] = 1;
  const uint4* src = reinterpret_cast<const uint4*>(dout + row * C);
  float* dst = scratch + id * C;
  const float scale = 1.0f / (1.0f - p);
  for (int c = lane; c < C / 8; c += 32) {
    float f[8];
    unpack8(src[c], f);
    if (p > 0.f) {
      bool k0[4], k1[4];
      const unsigned long long e4 = (static_cast<unsigned long long>(row) * C + c * 8) >> 2;
      dropout4(seed, offset, e4, p, k0);
      dropout4(seed, offset, e4 + 1, p, k1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[j] = k0[j] ? rb(f[j] * scale) : 0.f;
        f[4 + j] = k1[j] ? rb(f[4 + j] * scale) : 0.f;
      }
    }
    atomicAdd(reinterpret_cast<float4*>(dst + c * 8), make_float4(f[0], f[1], f[2], f[3]));
    atomicAdd(reinterpret_cast<float4*>(dst + c * 8 + 4), make_float4(f[4], f[5], f[6], f[7]));
  }
}

// dwte[v,:] = rb( (accumulate ? dwte[v,:] : 0) + rb(scratch[v,:]) ) for touched rows; re-zeroes scratch and flags.
__global__ void embed_bwd_commit_kernel(float* __restrict__ scratch, int* __restrict__ touched,
                                        __nv_bfloat16* __restrict__ dwte, int C, int V, int accumulate) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int v = blockIdx.x * warps_per_block + (threadIdx.x >> 5); v < V; v += gridDim.x * warps_per_block) {
    const int t = touched[v];
    if (!t) {
      if (!accumulate) {
        for (int c = lane; c < C / 8; c += 32) reinterpret_cast<uint4*>(dwte + static_cast<long long>(v) * C)[c] = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    float* s = scratch + static_cast<long long>(v) * C;
    uint4* g = reinterpret_cast<uint4*>(dwte + static_cast<long long>(v) * C);
    for (int c = lane; c < C / 8; c += 32) {
      float4 a = reinterpret_cast<float4*>(s)[2 * c], b = reinterpret_cast<float4*>(s)[2 * c + 1];
      float f[8] = {rb(a.x), rb(a.y), rb(a.z), rb(a.w), rb(b.x), rb(b.y), rb(b.z), rb(b.w)};
      if (accumulate) {
        float o[8];
        unpack8(g[c], o);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += o[j];
      }
      g[c] = pack8(f);
      reinterpret_cast<float4*>(s)[2 * c] = make_float4(0, 0, 0, 0);
      reinterpret_cast<float4*>(s)[2 * c + 1] = make_float4(0, 0, 0, 0);
    }
    __syncwarp();
    if (lane == 0) touched[v] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward: F.layer_norm(x, (C,), weight, None, 1e-5)  (model.py:63-72)
//   y = rb((x - mean) * rstd * gamma); optional second output z = rb(float(y) / readout_div)  (MuReadout prologue)
// One warp per row, row cached in registers (C <= 8*32*kMaxChunks).
// ---------------------------------------------------------------------------------------------
template <int kMaxChunks>
__global__ void ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                              __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ z, float* __restrict__ mean_out,
                              float* __restrict__ rstd_out, long long M, int C, float eps, float readout_div) {
  const int warps_per_block = blockDim.x >> 5;
  const long long row = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int nchunks = C / 8;
  const uint4* src = reinterpret_cast<const uint4*>(x + row * C);
  float v[kMaxChunks][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      unpack8(src[c], v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = v[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  uint4* dst = reinterpret_cast<uint4*>(y + row * C);
  uint4* zdst = z ? reinterpret_cast<uint4*>(z + row * C) : nullptr;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      float g[8], o[8];
      unpack8(reinterpret_cast<const uint4*>(gamma)[c], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rb((v[i][j] - mean) * rstd * g[j]);
      dst[c] = pack8(o);
      if (zdst) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = o[j] / readout_div;
        zdst[c] = pack8(o);
      }
    }
  }
}

// LayerNorm backward (dy is first divided by dy_div and rounded: the MuReadout 1/width_mult adjoint).
//   dx = rb( (dres ? dres : 0) + rb(rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat))) )
//   dx_drop (optional) = dropout-replay of dx with (drop_p, seed, offset) keyed by the flat element index: the
//     gradient entering the `x + dropout(f(x))` branch that produced this LayerNorm's input (model.py:151,167), i.e.
//     the operand of that branch's dgrad / wgrad GEMMs. Writing it here saves the stand-alone dropout kernel's
//     re-read of dx (profiles/r01_launches_v7.txt: 2 x 24 us per layer).
//   dgamma (+)= rb(sum_rows dy * xhat): per-block fp32 partials go to dgamma_partial[gridDim.x][C]; the first
//     ceil(C/32) blocks then wait until every block has published its partial and reduce 32 columns each, in block
//     order (deterministic). This tail replaces the separate ln_dgamma_reduce launch (10.5 us x 17 per micro-batch).
//     No deadlock: a waiting block only occupies its own slot, every other block runs to completion without it.
template <int kMaxChunks>
__global__ void __launch_bounds__(256, kMaxChunks == 4 ? 2 : 1)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
              const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dres,
              __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dx_drop, float* __restrict__ dgamma_partial,
              unsigned int* __restrict__ sync_counter, __nv_bfloat16* __restrict__ dgamma, int accumulate_dgamma,
              long long M, int C, float dy_div, float drop_p, unsigned long long seed, unsigned long long offset) {
  // dgamma partial sums: 32 fp32 accumulators per lane in registers (fixed lane -> column mapping), written once per
  // block to shared memory [warp][chunk i][j][lane] (lane fastest: conflict-free) for the block reduction. The kernel
  // is instruction-issue bound since the dropout replay moved in (60 % issue slots busy, 3.5 TB/s,
  // profiles/r02c_ln_bwd.details.txt): the shared-memory read-modify-write per element cost 3 instructions, the
  // accumulators fit in the 128 registers two resident blocks per SM allow.
  extern __shared__ float s_dg[];
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nchunks = C / 8;
  const float drop_scale = 1.0f / (1.0f - drop_p);
  float* my_dg = s_dg + static_cast<size_t>(warp) * kMaxChunks * 256;
  float dg_acc[kMaxChunks][8];
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dg_acc[i][j] = 0.f;

  for (long long row = static_cast<long long>(blockIdx.x) * warps_per_block + warp; row < M;
       row += static_cast<long long>(gridDim.x) * warps_per_block) {
    const uint4* dsrc = reinterpret_cast<const uint4*>(dy + row * C);
    const uint4* xsrc = reinterpret_cast<const uint4*>(x + row * C);
    const uint4* rsrc = dres ? reinterpret_cast<const uint4*>(dres + row * C) : nullptr;
    uint4 pd[kMaxChunks], px[kMaxChunks], pr[kMaxChunks];
    // all loads of the row first
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int c = lane + 32 * i;
      pd[i] = px[i] = pr[i] = make_uint4(0, 0, 0, 0);
      if (c < nchunks) {
        pd[i] = dsrc[c];
        px[i] = xsrc[c];
        if (rsrc) pr[i] = rsrc[c];
      }
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float d[8], g[8], xv[8];
        unpack8(pd[i], d);
        unpack8(px[i], xv);
        unpack8(reinterpret_cast<const uint4*>(gamma)[c], g);
        if (dy_div != 1.0f) {
          // MuReadout adjoint: the reference divides by width_mult and rounds to bf16 before ln_f's backward
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = rb(d[j] / dy_div);
          pd[i] = pack8(d);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float a = d[j] * g[j];
          dg_acc[i][j] = fmaf(d[j], xh, dg_acc[i][j]);
          s1 += a;
          s2 += a * xh;
        }
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    uint4* dst = reinterpret_cast<uint4*>(dx + row * C);
    uint4* ddst = dx_drop ? reinterpret_cast<uint4*>(dx_drop + row * C) : nullptr;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float d[8], g[8], xv[8], r[8], o[8];
        unpack8(pd[i], d);
        unpack8(px[i], xv);
        unpack8(pr[i], r);
        unpack8(reinterpret_cast<const uint4*>(gamma)[c], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          o[j] = rb(rstd * (d[j] * g[j] - s1 - xh * s2)) + r[j];
        }
        const uint4 packed = pack8(o);
        dst[c] = packed;
        if (ddst) {
          // same mask as dropout_kernel / the GEMM's EPI_RESID_DROPOUT: one RNG call per 4 consecutive flat elements
          float q[8];
          unpack8(packed, q);  // the bf16-rounded dx, exactly what the stand-alone kernel would read back
          bool k0[4], k1[4];
          const unsigned long long e4 = (static_cast<unsigned long long>(row) * C + c * 8) >> 2;
          dropout4(seed, offset, e4, drop_p, k0);
          dropout4(seed, offset, e4 + 1, drop_p, k1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            q[j] = k0[j] ? q[j] * drop_scale : 0.f;
            q[4 + j] = k1[j] ? q[4 + j] * drop_scale : 0.f;
          }
          ddst[c] = pack8(q);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) my_dg[(i * 8 + j) * 32 + lane] = dg_acc[i][j];
  __syncthreads();
  // block reduction of the per-warp dgamma partials; column of (i, j, lane) = (lane + 32 i) * 8 + j
  for (int t = threadIdx.x; t < kMaxChunks * 256; t += blockDim.x) {
    const int ln = t & 31, ij = t >> 5;
    const int col = (ln + 32 * (ij >> 3)) * 8 + (ij & 7);
    if (col < C) {
      float acc = 0.f;
      for (int w = 0; w < warps_per_block; ++w) acc += s_dg[static_cast<size_t>(w) * kMaxChunks * 256 + t];
      dgamma_partial[static_cast<long long>(blockIdx.x) * C + col] = acc;
    }
  }
  // publish, then the first ceil(C/32) blocks reduce 32 columns each once all partials are visible
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(sync_counter, 1u);
  const int n_reducers = (C + 31) / 32;
  if (static_cast<int>(blockIdx.x) >= n_reducers) return;
  if (threadIdx.x == 0) {
    // bounded: a stale counter (only possible after an aborted launch) must not hang the device
    unsigned int spins = 0;
    while (*reinterpret_cast<volatile unsigned int*>(sync_counter) < gridDim.x) {
      __nanosleep(64);
      if (++spins > (1u << 24)) __trap();
    }
    __threadfence();
  }
  __syncthreads();
  {
    // 32 columns x (blockDim.x / 32) row groups; 128-byte coalesced reads per partial row
    float* red = s_dg;  // reuse: [nry][33]
    const int tx = threadIdx.x & 31, ry = threadIdx.x >> 5, nry = blockDim.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    float t = 0.f;
    if (col < C)
      for (int bk = ry; bk < static_cast<int>(gridDim.x); bk += nry)
        t += __ldcg(dgamma_partial + static_cast<long long>(bk) * C + col);
    red[ry * 33 + tx] = t;
    __syncthreads();
    if (ry == 0 && col < C) {
      float tot = 0.f;
      for (int k = 0; k < nry; ++k) tot += red[k * 33 + tx];
      float o = rb(tot);
      if (accumulate_dgamma) o += __bfloat162float(dgamma[col]);
      dgamma[col] = __float2bfloat16_rn(o);
    }
  }
  // the last reducer to finish re-arms the counter for the next launch on this stream
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(sync_counter + 1, 1u) + 1u;
    if (done == static_cast<unsigned int>(min(n_reducers, static_cast<int>(gridDim.x)))) {
      sync_counter[0] = 0u;
      sync_counter[1] = 0u;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Rotary embedding on the q and k column ranges of the fused qkv buffer (model.py:39-50,108), in place.
//   sin_tab == nullptr : bf16 model, freqs_cis was cast to a REAL bf16 table -> cosine scaling
//   sin_tab != nullptr : complex table -> true interleaved rotation; `inverse` applies the adjoint (backward)
// qkv: [M, ld] with q at column 0 and k at column C; positions t = row % T; pairs are adjacent elements (2i, 2i+1).
// ---------------------------------------------------------------------------------------------
__global__ void rope_kernel(__nv_bfloat16* __restrict__ qkv, const float* __restrict__ cos_tab,
                            const float* __restrict__ sin_tab, long long M, int T, int C, int d, long long ld,
                            int inverse) {
  const int warps_per_block = blockDim.x >> 5;
  const long long row = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int t = static_cast<int>(row % T);
  const int half = d / 2;
  const float* ct = cos_tab + static_cast<long long>(t) * half;
  const float* st = sin_tab ? sin_tab + static_cast<long long>(t) * half : nullptr;
  // 2*C/8 chunks: first C/8 are q, next C/8 are k (k starts at column C)
  for (int c = lane; c < 2 * (C / 8); c += 32) {
    uint4* ptr = reinterpret_cast<uint4*>(qkv + row * ld) + c;
    float f[8];
    unpack8(*ptr, f);
    const int col = (c * 8) % C;       // column within q or k
    const int i0 = (col % d) / 2;      // pair index within the head
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float cs = ct[i0 + j];
      if (st) {
        const float sn = inverse ? -st[i0 + j] : st[i0 + j];
        const float a = f[2 * j], b = f[2 * j + 1];
        f[2 * j] = a * cs - b * sn;
        f[2 * j + 1] = a * sn + b * cs;
      } else {
        f[2 * j] *= cs;
        f[2 * j + 1] *= cs;
      }
    }
    *ptr = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// Element-wise dropout (forward and backward share the mask through (seed, offset, flat index)).
// out = keep ? rb(in * 1/(1-p)) : 0
// ---------------------------------------------------------------------------------------------
__global__ void dropout_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n8,
                               float p, unsigned long long seed, unsigned long long offset) {
  const float scale = 1.0f / (1.0f - p);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(reinterpret_cast<const uint4*>(in)[i], f);
    bool k0[4], k1[4];
    dropout4(seed, offset, static_cast<unsigned long long>(i) * 2, p, k0);
    dropout4(seed, offset, static_cast<unsigned long long>(i) * 2 + 1, p, k1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[j] = k0[j] ? f[j] * scale : 0.f;
      f[4 + j] = k1[j] ? f[4 + j] * scale : 0.f;
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// encode() pooling over the token axis (model.py:269-278): mean (fp32 accumulate, one rounding) and max (exact).
// grid (C/8/128 , B, nsplit): each block reduces rows [z*rows_per, ...) of batch b for 128 column-chunks.
// ---------------------------------------------------------------------------------------------
__global__ void pool_partial_kernel(const __nv_bfloat16* __restrict__ emb, float* __restrict__ partial, int T, int C,
                                    int rows_per, int mode) {
  const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= C / 8) return;
  const int b = blockIdx.y, z = blockIdx.z;
  const int t0 = z * rows_per, t1 = min(T, t0 + rows_per);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = mode == 0 ? 0.f : -INFINITY;
  const __nv_bfloat16* base = emb + (static_cast<long long>(b) * T) * C + chunk * 8;
  for (int t = t0; t < t1; ++t) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(base + static_cast<long long>(t) * C), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = mode == 0 ? acc[j] + f[j] : fmaxf(acc[j], f[j]);
  }
  float* dst = partial + ((static_cast<long long>(b) * gridDim.z + z) * C) + chunk * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[j] = acc[j];
}

__global__ void pool_final_kernel(const float* __restrict__ partial, __nv_bfloat16* __restrict__ out, int B, int T,
                                  int C, int nsplit, int mode) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(B) * C) return;
  const int b = static_cast<int>(i / C), c = static_cast<int>(i % C);
  float acc = mode == 0 ? 0.f : -INFINITY;
  for (int z = 0; z < nsplit; ++z) {
    float v = partial[(static_cast<long long>(b) * nsplit + z) * C + c];
    acc = mode == 0 ? acc + v : fmaxf(acc, v);
  }
  if (mode == 0) acc = acc / T;
  out[i] = __float2bfloat16_rn(acc);
}

// backward of mean pooling: demb[b,t,:] = rb(dout[b,:] / T) ; of max pooling: gradient to the arg-max position
// (first maximal position, as torch.max(dim) returns).
__global__ void pool_bwd_kernel(const __nv_bfloat16* __restrict__ emb, const __nv_bfloat16* __restrict__ pooled,
                                const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ demb, int B, int T,
                                int C, int mode) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(B) * C) return;
  const int b = static_cast<int>(i / C), c = static_cast<int>(i % C);
  const float g = __bfloat162float(dout[i]);
  if (mode == 0) {
    const __nv_bfloat16 v = __float2bfloat16_rn(g / T);
    for (int t = 0; t < T; ++t) demb[(static_cast<long long>(b) * T + t) * C + c] = v;
  } else {
    const __nv_bfloat16 mx = pooled[i];
    bool done = false;
    for (int t = 0; t < T; ++t) {
      const long long e = (static_cast<long long>(b) * T + t) * C + c;
      const bool hit = !done && (emb[e] == mx);
      demb[e] = hit ? dout[i] : __float2bfloat16_rn(0.f);
      done |= hit;
    }
  }
}

// out = rb(float(in) / div): the MuReadout input scaling `output_mult * x / width_mult` and its adjoint.
__global__ void scale_div_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n8,
                                 float div) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(reinterpret_cast<const uint4*>(in)[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] / div;
    reinterpret_cast<uint4*>(out)[i] = pack8(f);
  }
}

}  // namespace obt

using namespace obt;

extern "C" int obt_scale_div(const void* in, void* out, long long n, float div, cudaStream_t stream) {
  OBT_REQUIRE(in && out, "obt_scale_div: null pointer");
  OBT_REQUIRE(n % 8 == 0, "obt_scale_div: n=%lld must be a multiple of 8", n);
  if (n == 0) return OBT_OK;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  scale_div_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), n8, div);
  return check_launch("scale_div");
}

extern "C" int obt_embed_fwd(const long long* idx, const void* wte, void* out, long long M, int C, int V, float drop_p,
                             unsigned long long seed, unsigned long long offset, int* err_flag, cudaStream_t stream) {
  OBT_REQUIRE(idx && wte && out && err_flag, "obt_embed_fwd: null pointer");
  OBT_REQUIRE(C % 8 == 0, "obt_embed_fwd: C=%d must be a multiple of 8", C);
  if (M == 0) return OBT_OK;
  const int wpb = 8;
  embed_fwd_kernel<<<static_cast<unsigned>((M + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      idx, static_cast<const __nv_bfloat16*>(wte), static_cast<__nv_bfloat16*>(out), M, C, V, drop_p, seed, offset,
      err_flag);
  return check_launch("embed_fwd");
}

extern "C" int obt_embed_bwd(const long long* idx, const void* dout, void* dwte, float* scratch, int* touched,
                             long long M, int C, int V, int accumulate, float drop_p, unsigned long long seed,
                             unsigned long long offset, cudaStream_t stream) {
  OBT_REQUIRE(idx && dout && dwte && scratch && touched, "obt_embed_bwd: null pointer");
  OBT_REQUIRE(C % 8 == 0, "obt_embed_bwd: C=%d must be a multiple of 8", C);
  const int wpb = 8;
  if (M > 0) {
    embed_bwd_scatter_kernel<<<static_cast<unsigned>((M + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
        idx, static_cast<const __nv_bfloat16*>(dout), scratch, touched, M, C, V, drop_p, seed, offset);
    int rc = check_launch("embed_bwd_scatter");
    if (rc) return rc;
  }
  int blocks = (V + wpb - 1) / wpb;
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  embed_bwd_commit_kernel<<<blocks, wpb * 32, 0, stream>>>(scratch, touched, static_cast<__nv_bfloat16*>(dwte), C, V,
                                                           accumulate);
  return check_launch("embed_bwd_commit");
}

extern "C" int obt_layernorm_fwd(const void* x, const void* gamma, void* y, void* z, float* mean, float* rstd,
                                 long long M, int C, float eps, float readout_div, cudaStream_t stream) {
  OBT_REQUIRE(x && gamma && y, "obt_layernorm_fwd: null pointer");
  OBT_REQUIRE(C % 8 == 0 && C <= 8 * 32 * 16, "obt_layernorm_fwd: C=%d must be a multiple of 8 and <= 4096", C);
  if (M == 0) return OBT_OK;
  const int wpb = 4;
  const unsigned grid = static_cast<unsigned>((M + wpb - 1) / wpb);
  auto xx = static_cast<const __nv_bfloat16*>(x);
  auto gg = static_cast<const __nv_bfloat16*>(gamma);
  auto yy = static_cast<__nv_bfloat16*>(y);
  auto zz = static_cast<__nv_bfloat16*>(z);
  if (C <= 1024)
    ln_fwd_kernel<4><<<grid, wpb * 32, 0, stream>>>(xx, gg, yy, zz, mean, rstd, M, C, eps, readout_div);
  else if (C <= 2048)
    ln_fwd_kernel<8><<<grid, wpb * 32, 0, stream>>>(xx, gg, yy, zz, mean, rstd, M, C, eps, readout_div);
  else
    ln_fwd_kernel<16><<<grid, wpb * 32, 0, stream>>>(xx, gg, yy, zz, mean, rstd, M, C, eps, readout_div);
  return check_launch("ln_fwd");
}

// workspace: fp32 [32 + obt_layernorm_bwd_workspace_rows() * C]; the first two words are the grid-sync counters of the
// fused dgamma reduction and MUST be zero before the first call (the kernel leaves them zero again); the per-block
// partial sums start at word 32 (128-byte aligned).
extern "C" int obt_layernorm_bwd_workspace_rows(void) { return sm_count() * 4; }

extern "C" int obt_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                                 const void* dres, void* dx, void* dgamma, int accumulate_dgamma, float* workspace,
                                 long long M, int C, float dy_div, void* dx_drop, float drop_p,
                                 unsigned long long seed, unsigned long long offset, cudaStream_t stream) {
  OBT_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && workspace, "obt_layernorm_bwd: null pointer");
  OBT_REQUIRE(C % 8 == 0 && C <= 2048, "obt_layernorm_bwd: C=%d must be a multiple of 8 and <= 2048", C);
  OBT_REQUIRE(M > 0, "obt_layernorm_bwd: empty input");
  OBT_REQUIRE(dx_drop == nullptr || (drop_p > 0.f && drop_p < 1.f), "obt_layernorm_bwd: drop_p=%f with dx_drop", drop_p);
  const int wpb = (C <= 1024) ? 8 : 4;  // 32 KB of dgamma staging per block either way
  const size_t smem_bytes = static_cast<size_t>(wpb) * ((C <= 1024) ? 4 : 8) * 256 * sizeof(float);
  const int max_grid = sm_count() * 4;
  // ONE wave: every block must be resident at once. The reducer blocks of the fused dgamma tail hold their slot while
  // they wait for the last partial; with more blocks than slots (592 on 296) those 32 held slots pushed the remaining
  // blocks into a third wave (59.7 -> 105.9 us, profiles/r02a_launches.txt). The row loop is grid-strided anyway.
  static int occ4 = 0, occ8 = 0;
  int& occ = (C <= 1024) ? occ4 : occ8;
  if (occ == 0) {
    cudaError_t e = (C <= 1024)
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_bwd_kernel<4>, wpb * 32, smem_bytes)
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_bwd_kernel<8>, wpb * 32, smem_bytes);
    if (e != cudaSuccess || occ < 1) {
      occ = 0;
      set_last_error("obt_layernorm_bwd: occupancy query failed: %s", cudaGetErrorString(e));
      return OBT_ERR_CUDA;
    }
  }
  int grid = sm_count() * occ;
  if (grid > max_grid) grid = max_grid;
  if (M < static_cast<long long>(grid) * wpb) grid = static_cast<int>((M + wpb - 1) / wpb);
  if (grid < 1) grid = 1;
  // the reducer blocks (the first ceil(C/32)) must exist: with very few rows the grid is padded (row loop is a no-op)
  const int n_reducers = (C + 31) / 32;
  if (grid < n_reducers) grid = n_reducers;
  OBT_REQUIRE(grid <= max_grid && grid <= sm_count() * occ,
              "obt_layernorm_bwd: C=%d needs more reducer blocks than can be resident", C);
  const size_t smem = smem_bytes;
  auto a = static_cast<const __nv_bfloat16*>(dy);
  auto b = static_cast<const __nv_bfloat16*>(x);
  auto g = static_cast<const __nv_bfloat16*>(gamma);
  auto r = static_cast<const __nv_bfloat16*>(dres);
  auto o = static_cast<__nv_bfloat16*>(dx);
  auto od = static_cast<__nv_bfloat16*>(dx_drop);
  auto dg = static_cast<__nv_bfloat16*>(dgamma);
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
  float* partial = workspace + 32;
  if (C <= 1024)
    ln_bwd_kernel<4><<<grid, wpb * 32, smem, stream>>>(a, b, g, mean, rstd, r, o, od, partial, counter, dg,
                                                       accumulate_dgamma, M, C, dy_div, drop_p, seed, offset);
  else
    ln_bwd_kernel<8><<<grid, wpb * 32, smem, stream>>>(a, b, g, mean, rstd, r, o, od, partial, counter, dg,
                                                       accumulate_dgamma, M, C, dy_div, drop_p, seed, offset);
  return check_launch("ln_bwd");
}

extern "C" int obt_rope(void* qkv, const float* cos_tab, const float* sin_tab, long long M, int T, int C, int head_dim,
                        long long ld, int inverse, cudaStream_t stream) {
  OBT_REQUIRE(qkv && cos_tab, "obt_rope: null pointer");
  OBT_REQUIRE(C % 8 == 0 && head_dim % 8 == 0 && ld % 8 == 0 && C % head_dim == 0,
              "obt_rope: C=%d head_dim=%d ld=%lld must be multiples of 8", C, head_dim, ld);
  if (M == 0) return OBT_OK;
  const int wpb = 8;
  rope_kernel<<<static_cast<unsigned>((M + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      static_cast<__nv_bfloat16*>(qkv), cos_tab, sin_tab, M, T, C, head_dim, ld, inverse);
  return check_launch("rope");
}

extern "C" int obt_dropout(const void* in, void* out, long long n, float p, unsigned long long seed,
                           unsigned long long offset, cudaStream_t stream) {
  OBT_REQUIRE(in && out, "obt_dropout: null pointer");
  OBT_REQUIRE(n % 8 == 0, "obt_dropout: n=%lld must be a multiple of 8", n);
  OBT_REQUIRE(p >= 0.f && p < 1.f, "obt_dropout: p=%f out of range", p);
  if (n == 0) return OBT_OK;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  dropout_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), n8, p, seed, offset);
  return check_launch("dropout");
}

// mode 0 = mean, 1 = max. workspace: fp32 [B * obt_pool_splits(T) * C]
extern "C" int obt_pool_splits(int T) {
  int s = (T + 63) / 64;
  return s < 1 ? 1 : s;
}

extern "C" int obt_pool(const void* emb, void* out, float* workspace, int B, int T, int C, int mode,
                        cudaStream_t stream) {
  OBT_REQUIRE(emb && out && workspace, "obt_pool: null pointer");
  OBT_REQUIRE(C % 8 == 0, "obt_pool: C=%d must be a multiple of 8", C);
  OBT_REQUIRE(mode == 0 || mode == 1, "obt_pool: mode %d", mode);
  OBT_REQUIRE(B > 0 && T > 0, "obt_pool: empty input");
  const int nsplit = obt_pool_splits(T);
  dim3 grid((C / 8 + 127) / 128, B, nsplit);
  pool_partial_kernel<<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(emb), workspace, T, C, 64, mode);
  int rc = check_launch("pool_partial");
  if (rc) return rc;
  const long long n = static_cast<long long>(B) * C;
  pool_final_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      workspace, static_cast<__nv_bfloat16*>(out), B, T, C, nsplit, mode);
  return check_launch("pool_final");
}

extern "C" int obt_pool_bwd(const void* emb, const void* pooled, const void* dout, void* demb, int B, int T, int C,
                            int mode, cudaStream_t stream) {
  OBT_REQUIRE(emb && pooled && dout && demb, "obt_pool_bwd: null pointer");
  const long long n = static_cast<long long>(B) * C;
  pool_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(emb), static_cast<const __nv_bfloat16*>(pooled),
      static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(demb), B, T, C, mode);
  return check_launch("pool_bwd");
}
