// Generic-shape attention (any head_dim multiple of 8, any T) on CUDA cores, fp32 math.
// This is the path for head dims other than 128 and the on-device cross-check for the tcgen05 kernel
// (attn_tc.cu); the BASELINE shapes (d = 128) run on tensor cores.
// Semantics: F.scaled_dot_product_attention(q, k, v, attn_mask, dropout_p, scale=8/n_embd, is_causal=False)
// (training/model.py:111-138): P = softmax(q k^T * scale + mask) row-wise in fp32, dropout on P, y = P v.
// A fully-masked row carries a finite -1e9 bias on every key and therefore attends uniformly; the
// backward here is the mathematically exact adjoint.
#include "common.cuh"
#include "ptx.cuh"
#include "dropmask.cuh"

namespace obt {

struct AttnSimtParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  long long ld;  // row pitch (elements) of q/k/v rows (token-major: row = b*T + t), head h at column h*d
  const __nv_bfloat16* mask;  // additive, nullable; element (b,h,i,j) at mask[b*msb + h*msh + i*msq + j]
  long long msb, msh, msq;
  const int* row_lo;  // optional interval mask: key j visible to query (b,i) iff lo <= j < hi; lo==hi -> uniform row
  const int* row_hi;
  __nv_bfloat16* y;  // [M, ldy], head h at column h*d
  long long ldy;
  float* lse;  // [B,H,T,2] = (row max, log of the exp-sum): kept apart so a -1e9 row max cannot absorb log(sum)
  int B, H, T, d;
  float scale, drop_p;
  const uint32_t* keep;  // dropout keep bits [B,H,T,ceil(T/32)] (dropmask.cuh), required when drop_p > 0
  // backward
  const __nv_bfloat16* dy;
  long long lddy;
  float* delta;  // [B,H,T]
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  long long ldd;
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float dot_row(const float* __restrict__ a, const __nv_bfloat16* __restrict__ b, int d) {
  float s = 0.f;
  for (int c = 0; c < d / 8; ++c) {
    uint4 u = reinterpret_cast<const uint4*>(b)[c];
    s += a[c * 8 + 0] * bf16_lo(u.x) + a[c * 8 + 1] * bf16_hi(u.x) + a[c * 8 + 2] * bf16_lo(u.y) +
         a[c * 8 + 3] * bf16_hi(u.y) + a[c * 8 + 4] * bf16_lo(u.z) + a[c * 8 + 5] * bf16_hi(u.z) +
         a[c * 8 + 6] * bf16_lo(u.w) + a[c * 8 + 7] * bf16_hi(u.w);
  }
  return s;
}

// additive bias for (b,h,i,j)
__device__ __forceinline__ float mask_bias(const AttnSimtParams& p, int b, int h, int i, int j) {
  if (p.mask) return __bfloat162float(p.mask[b * p.msb + h * p.msh + i * p.msq + j]);
  if (p.row_lo) {
    const int lo = p.row_lo[b * p.T + i], hi = p.row_hi[b * p.T + i];
    if (lo >= hi) return -1e9f;  // fully masked row: uniform over all keys
    return (j >= lo && j < hi) ? 0.f : -1e9f;
  }
  return 0.f;
}

__device__ __forceinline__ float keep_scale(const AttnSimtParams& p, int b, int h, int i, int j) {
  if (p.drop_p <= 0.f) return 1.0f;
  const uint32_t w = p.keep[((static_cast<long long>(b) * p.H + h) * p.T + i) * keep_words(p.T) + (j >> 5)];
  return ((w >> keep_bit_pos(j & 31)) & 1u) ? 1.0f / (1.0f - p.drop_p) : 0.f;
}

// grid (ceil(T / warps), H, B); one warp per query row. smem per warp: d floats (q) + T floats (scores).
__global__ void attn_simt_fwd_kernel(const AttnSimtParams p) {
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * warps + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (i >= p.T) return;
  float* sq = sm + static_cast<size_t>(warp) * (p.d + p.T);
  float* sc = sq + p.d;
  const long long row = static_cast<long long>(b) * p.T + i;
  const __nv_bfloat16* qrow = p.q + row * p.ld + h * p.d;
  for (int e = lane; e < p.d; e += 32) sq[e] = __bfloat162float(qrow[e]);
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < p.T; j += 32) {
    const __nv_bfloat16* krow = p.k + (static_cast<long long>(b) * p.T + j) * p.ld + h * p.d;
    const float s = dot_row(sq, krow, p.d) * p.scale + mask_bias(p, b, h, i, j);
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max_f(mx);
  float sum = 0.f;
  for (int j = lane; j < p.T; j += 32) {
    const float e = __expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum_f(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < p.T; j += 32) sc[j] = sc[j] * inv * keep_scale(p, b, h, i, j);
  __syncwarp();
  if (lane == 0 && p.lse) {
    float* l = p.lse + 2 * ((static_cast<long long>(b) * p.H + h) * p.T + i);
    l[0] = mx;
    l[1] = logf(sum);
  }
  __nv_bfloat16* yrow = p.y + row * p.ldy + h * p.d;
  for (int e = lane; e < p.d; e += 32) {
    float o = 0.f;
    const __nv_bfloat16* vcol = p.v + (static_cast<long long>(b) * p.T) * p.ld + h * p.d + e;
    for (int j = 0; j < p.T; ++j) o += sc[j] * __bfloat162float(vcol[static_cast<long long>(j) * p.ld]);
    yrow[e] = __float2bfloat16_rn(o);
  }
}

// dQ pass: one warp per query row; also emits delta_i = dO_i . O_i for the dK/dV pass.
__global__ void attn_simt_bwd_dq_kernel(const AttnSimtParams p) {
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * warps + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (i >= p.T) return;
  float* sq = sm + static_cast<size_t>(warp) * (2 * p.d + p.T);
  float* sdo = sq + p.d;
  float* sds = sdo + p.d;
  const long long row = static_cast<long long>(b) * p.T + i;
  const __nv_bfloat16* qrow = p.q + row * p.ld + h * p.d;
  const __nv_bfloat16* dorow = p.dy + row * p.lddy + h * p.d;
  const __nv_bfloat16* yrow = p.y + row * p.ldy + h * p.d;
  float dl = 0.f;
  for (int e = lane; e < p.d; e += 32) {
    sq[e] = __bfloat162float(qrow[e]);
    const float g = __bfloat162float(dorow[e]);
    sdo[e] = g;
    dl += g * __bfloat162float(yrow[e]);
  }
  dl = warp_sum_f(dl);
  __syncwarp();
  const long long bh = static_cast<long long>(b) * p.H + h;
  const float lmx = p.lse[2 * (bh * p.T + i)], lsum = p.lse[2 * (bh * p.T + i) + 1];
  if (lane == 0) p.delta[bh * p.T + i] = dl;
  for (int j = lane; j < p.T; j += 32) {
    const long long krow_i = (static_cast<long long>(b) * p.T + j) * p.ld + h * p.d;
    const float s = dot_row(sq, p.k + krow_i, p.d) * p.scale + mask_bias(p, b, h, i, j);
    const float pr = __expf((s - lmx) - lsum);
    const float dp = dot_row(sdo, p.v + krow_i, p.d) * keep_scale(p, b, h, i, j);
    sds[j] = pr * (dp - dl);
  }
  __syncwarp();
  __nv_bfloat16* dqrow = p.dq + row * p.ldd + h * p.d;
  for (int e = lane; e < p.d; e += 32) {
    float o = 0.f;
    const __nv_bfloat16* kcol = p.k + (static_cast<long long>(b) * p.T) * p.ld + h * p.d + e;
    for (int j = 0; j < p.T; ++j) o += sds[j] * __bfloat162float(kcol[static_cast<long long>(j) * p.ld]);
    dqrow[e] = __float2bfloat16_rn(o * p.scale);
  }
}

// dK/dV pass: one warp per key row j, loops over all queries i.
__global__ void attn_simt_bwd_dkv_kernel(const AttnSimtParams p) {
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * warps + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (j >= p.T) return;
  float* sk = sm + static_cast<size_t>(warp) * (2 * p.d + 2 * p.T);
  float* sv = sk + p.d;
  float* sds = sv + p.d;
  float* spd = sds + p.T;
  const long long row = static_cast<long long>(b) * p.T + j;
  const __nv_bfloat16* krow = p.k + row * p.ld + h * p.d;
  const __nv_bfloat16* vrow = p.v + row * p.ld + h * p.d;
  for (int e = lane; e < p.d; e += 32) {
    sk[e] = __bfloat162float(krow[e]);
    sv[e] = __bfloat162float(vrow[e]);
  }
  __syncwarp();
  const long long bh = static_cast<long long>(b) * p.H + h;
  for (int i = lane; i < p.T; i += 32) {
    const long long qi = static_cast<long long>(b) * p.T + i;
    const float s = dot_row(sk, p.q + qi * p.ld + h * p.d, p.d) * p.scale + mask_bias(p, b, h, i, j);
    const float pr = __expf((s - p.lse[2 * (bh * p.T + i)]) - p.lse[2 * (bh * p.T + i) + 1]);
    const float ks = keep_scale(p, b, h, i, j);
    const float dp = dot_row(sv, p.dy + qi * p.lddy + h * p.d, p.d) * ks;
    sds[i] = pr * (dp - p.delta[bh * p.T + i]);
    spd[i] = pr * ks;
  }
  __syncwarp();
  __nv_bfloat16* dkrow = p.dk + row * p.ldd + h * p.d;
  __nv_bfloat16* dvrow = p.dv + row * p.ldd + h * p.d;
  for (int e = lane; e < p.d; e += 32) {
    float ok = 0.f, ov = 0.f;
    const __nv_bfloat16* qcol = p.q + (static_cast<long long>(b) * p.T) * p.ld + h * p.d + e;
    const __nv_bfloat16* docol = p.dy + (static_cast<long long>(b) * p.T) * p.lddy + h * p.d + e;
    for (int i = 0; i < p.T; ++i) {
      ok += sds[i] * __bfloat162float(qcol[static_cast<long long>(i) * p.ld]);
      ov += spd[i] * __bfloat162float(docol[static_cast<long long>(i) * p.lddy]);
    }
    dkrow[e] = __float2bfloat16_rn(ok * p.scale);
    dvrow[e] = __float2bfloat16_rn(ov);
  }
}

static int pick_warps(size_t per_warp_bytes) {
  int w = 4;
  while (w > 1 && per_warp_bytes * w > 200 * 1024) w >>= 1;
  return w;
}

template <typename K>
static int launch_simt(K kern, const AttnSimtParams& p, size_t per_warp_floats, const char* what, cudaStream_t stream) {
  const size_t per_warp = per_warp_floats * sizeof(float);
  if (per_warp > 200 * 1024) {
    set_last_error("%s: T=%d too long for the generic attention kernel", what, p.T);
    return OBT_ERR_UNSUPPORTED;
  }
  const int warps = pick_warps(per_warp);
  const size_t smem = per_warp * warps;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) {
    set_last_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return OBT_ERR_CUDA;
  }
  dim3 grid((p.T + warps - 1) / warps, p.H, p.B);
  kern<<<grid, warps * 32, smem, stream>>>(p);
  return check_launch(what);
}

}  // namespace obt

using namespace obt;

static int fill_common(AttnSimtParams& p, const void* q, const void* k, const void* v, long long ld, const void* mask,
                       long long msb, long long msh, long long msq, const int* row_lo, const int* row_hi, int B, int H,
                       int T, int d, float scale, float drop_p, const unsigned int* keep) {
  OBT_REQUIRE(q && k && v, "attention: null q/k/v");
  OBT_REQUIRE(d % 8 == 0 && d > 0, "attention: head_dim=%d must be a positive multiple of 8", d);
  OBT_REQUIRE(ld % 8 == 0, "attention: ld=%lld must be a multiple of 8", ld);
  OBT_REQUIRE(B > 0 && H > 0 && T > 0, "attention: empty problem");
  OBT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attention: dropout p=%f", drop_p);
  OBT_REQUIRE(drop_p == 0.f || keep != nullptr, "attention: dropout needs the keep mask (obt_attn_keep_mask)");
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ld = ld;
  p.mask = static_cast<const __nv_bfloat16*>(mask);
  p.msb = msb; p.msh = msh; p.msq = msq;
  p.row_lo = row_lo; p.row_hi = row_hi;
  p.B = B; p.H = H; p.T = T; p.d = d;
  p.scale = scale; p.drop_p = drop_p; p.keep = keep;
  return OBT_OK;
}

extern "C" int obt_attn_simt_fwd(const void* q, const void* k, const void* v, long long ld, const void* mask,
                                 long long msb, long long msh, long long msq, const int* row_lo, const int* row_hi,
                                 void* y, long long ldy, float* lse, int B, int H, int T, int d, float scale,
                                 float drop_p, const unsigned int* keep, cudaStream_t stream) {
  AttnSimtParams p = {};
  int rc = fill_common(p, q, k, v, ld, mask, msb, msh, msq, row_lo, row_hi, B, H, T, d, scale, drop_p, keep);
  if (rc) return rc;
  OBT_REQUIRE(y && lse, "obt_attn_simt_fwd: null output");
  p.y = static_cast<__nv_bfloat16*>(y);
  p.ldy = ldy;
  p.lse = lse;
  return launch_simt(attn_simt_fwd_kernel, p, static_cast<size_t>(d) + T, "attn_simt_fwd", stream);
}

extern "C" int obt_attn_simt_bwd(const void* q, const void* k, const void* v, long long ld, const void* mask,
                                 long long msb, long long msh, long long msq, const int* row_lo, const int* row_hi,
                                 const void* y, long long ldy, const void* dy, long long lddy, const float* lse,
                                 float* delta, void* dq, void* dk, void* dv, long long ldd, int B, int H, int T, int d,
                                 float scale, float drop_p, const unsigned int* keep, cudaStream_t stream) {
  AttnSimtParams p = {};
  int rc = fill_common(p, q, k, v, ld, mask, msb, msh, msq, row_lo, row_hi, B, H, T, d, scale, drop_p, keep);
  if (rc) return rc;
  OBT_REQUIRE(y && dy && lse && delta && dq && dk && dv, "obt_attn_simt_bwd: null pointer");
  p.y = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(y));
  p.ldy = ldy;
  p.dy = static_cast<const __nv_bfloat16*>(dy);
  p.lddy = lddy;
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.dq = static_cast<__nv_bfloat16*>(dq);
  p.dk = static_cast<__nv_bfloat16*>(dk);
  p.dv = static_cast<__nv_bfloat16*>(dv);
  p.ldd = ldd;
  rc = launch_simt(attn_simt_bwd_dq_kernel, p, 2 * static_cast<size_t>(d) + T, "attn_simt_bwd_dq", stream);
  if (rc) return rc;
  return launch_simt(attn_simt_bwd_dkv_kernel, p, 2 * static_cast<size_t>(d) + 2 * static_cast<size_t>(T),
                     "attn_simt_bwd_dkv", stream);
}
