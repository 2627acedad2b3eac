// MLM cross-entropy (forward + backward over materialised logits), global grad-norm clipping and the fused
// muP AdamW step.  Reference arithmetic: training/train_encoder.py:301-305 (loss), :316 (clip_grad_norm_),
// :195-201,317 (MuAdamW = torch.optim.AdamW over muP-scaled param groups, bf16 state).
#include "common.cuh"
#include "ptx.cuh"

namespace obt {

__device__ __forceinline__ float block_reduce_sum(float v, float* sbuf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sbuf[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? sbuf[threadIdx.x] : 0.f;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sbuf[0] = t;
  }
  __syncthreads();
  t = sbuf[0];
  __syncthreads();
  return t;
}

__device__ __forceinline__ float block_reduce_max(float v, float* sbuf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) sbuf[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? sbuf[threadIdx.x] : -INFINITY;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (lane == 0) sbuf[0] = t;
  }
  __syncthreads();
  t = sbuf[0];
  __syncthreads();
  return t;
}

// One block per token row. Rows whose weight is zero are skipped entirely (their loss term is exactly 0 in the
// reference: `loss *= mask.float()`), which removes ~85 % of the logits reads for 15 % MLM masking.
//   tok_loss[row] = rb(lse - logit[target])   (bf16-rounded like F.cross_entropy(reduction="none") on bf16 logits)
__global__ void ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ld,
                              const long long* __restrict__ targets, const unsigned char* __restrict__ row_mask,
                              float* __restrict__ lse_out, float* __restrict__ tok_loss, int V) {
  __shared__ float sbuf[32];
  const long long row = blockIdx.x;
  if (row_mask && row_mask[row] == 0) {
    if (threadIdx.x == 0) {
      lse_out[row] = 0.f;
      tok_loss[row] = 0.f;
    }
    return;
  }
  const __nv_bfloat16* x = logits + row * ld;
  float m = -INFINITY, s = 0.f;
  for (int c = threadIdx.x; c < V / 8; c += blockDim.x) {
    uint4 u = reinterpret_cast<const uint4*>(x)[c];
    float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                  bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
    float cm = f[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) cm = fmaxf(cm, f[j]);
    if (cm > m) {
      s *= __expf(m - cm);
      m = cm;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf(f[j] - m);
  }
  const float gm = block_reduce_max(m, sbuf);
  s = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
  const float gs = block_reduce_sum(s, sbuf);
  if (threadIdx.x == 0) {
    const float lse = gm + logf(gs);
    const long long y = targets[row];
    const float xy = (y >= 0 && y < V) ? __bfloat162float(x[y]) : 0.f;
    lse_out[row] = lse;
    tok_loss[row] = rb(lse - xy);
  }
}

// scalars[0] = loss = rb( rb(sum_t rb(rb(tok_t / n_acc) * m_t)) / count )   (train_encoder.py:301-305, all bf16)
// scalars[1] = count = sum_t m_t
// scalars[2] = g = rb(rb(1/count) / n_acc)   gradient of the loss w.r.t. every masked token's CE term
// scalars[3] = n_acc (kept for the backward: an upstream factor u turns g into rb(rb(u/count) / n_acc))
__global__ void ce_reduce_kernel(const float* __restrict__ tok_loss, const unsigned char* __restrict__ row_mask,
                                 long long M, float n_acc, float* __restrict__ scalars) {
  __shared__ float sbuf[32];
  float s = 0.f, cnt = 0.f;
  for (long long i = threadIdx.x; i < M; i += blockDim.x) {
    const float m = row_mask ? static_cast<float>(row_mask[i] != 0) : 1.f;
    s += rb(rb(tok_loss[i] / n_acc) * m);
    cnt += m;
  }
  s = block_reduce_sum(s, sbuf);
  cnt = block_reduce_sum(cnt, sbuf);
  if (threadIdx.x == 0) {
    scalars[0] = rb(rb(s) / cnt);
    scalars[1] = cnt;
    scalars[2] = rb(rb(1.0f / cnt) / n_acc);
    scalars[3] = n_acc;
  }
}

// dlogits[row, j] = rb( (exp(logit_j - lse) - [j == target]) * g ) for masked rows, exact zeros otherwise, with
// g = rb(rb(u / count) / n_acc) and u = upstream * (upstream_dev ? float(*upstream_dev) : 1): the autograd chain of
// train_encoder.py:301-305 for an incoming d loss = u (u = 1 reproduces scalars[2]). upstream_dev is a bf16 device
// scalar (the dtype of the loss), so that `(k * loss).backward()` needs no host synchronisation.
// Written in place over the logits buffer.
__global__ void ce_bwd_kernel(__nv_bfloat16* __restrict__ logits, long long ld, const long long* __restrict__ targets,
                              const unsigned char* __restrict__ row_mask, const float* __restrict__ lse_in,
                              const float* __restrict__ scalars, float upstream,
                              const __nv_bfloat16* __restrict__ upstream_dev, int V, int unmasked_rows_zero) {
  const long long row = blockIdx.x;
  __nv_bfloat16* x = logits + row * ld;
  if (row_mask && row_mask[row] == 0) {
    if (unmasked_rows_zero) return;  // the head GEMM's row-mask epilogue already stored the zeros
    for (int c = threadIdx.x; c < V / 8; c += blockDim.x) reinterpret_cast<uint4*>(x)[c] = make_uint4(0, 0, 0, 0);
    return;
  }
  const float up = upstream_dev ? upstream * __bfloat162float(*upstream_dev) : upstream;
  const float g = (up == 1.0f) ? scalars[2] : rb(rb(up / scalars[1]) / scalars[3]);
  const float lse = lse_in[row];
  const int y = static_cast<int>(targets[row]);
  for (int c = threadIdx.x; c < V / 8; c += blockDim.x) {
    uint4 u = reinterpret_cast<const uint4*>(x)[c];
    float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                  bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p = __expf(f[j] - lse);
      if (c * 8 + j == y) p -= 1.0f;
      f[j] = p * g;
    }
    reinterpret_cast<uint4*>(x)[c] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

// ---------------------------------------------------------------------------------------------
// Multi-tensor grad-norm + AdamW
// ---------------------------------------------------------------------------------------------
struct ParamMeta {
  __nv_bfloat16* p;
  __nv_bfloat16* g;
  __nv_bfloat16* m;
  __nv_bfloat16* v;
  long long numel;
  float lr;   // base lr of the param group (already muP-scaled: lr / width_mult for matrix-like params)
  float wd;   // weight decay of the group (wd * width_mult for matrix-like params)
};

constexpr int OPT_CHUNK = 16384;  // elements per block

// partial[b] = sum over the block's chunk of rb(g * gscale)^2
__global__ void gradnorm_partial_kernel(const ParamMeta* __restrict__ metas, const int* __restrict__ blk_tensor,
                                        const long long* __restrict__ blk_off, float gscale,
                                        float* __restrict__ partial) {
  __shared__ float sbuf[32];
  const ParamMeta pm = metas[blk_tensor[blockIdx.x]];
  const long long off = blk_off[blockIdx.x];
  const long long n = min(static_cast<long long>(OPT_CHUNK), pm.numel - off);
  const __nv_bfloat16* g = pm.g + off;
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const long long n8 = n / 8;
    for (long long i = threadIdx.x; i < n8; i += blockDim.x) {
      uint4 u = reinterpret_cast<const uint4*>(g)[i];
      float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                    bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = gscale == 1.0f ? f[j] : rb(f[j] * gscale);
        s += t * t;
      }
    }
    for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
      float t = __bfloat162float(g[i]);
      t = gscale == 1.0f ? t : rb(t * gscale);
      s += t * t;
    }
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      float t = __bfloat162float(g[i]);
      t = gscale == 1.0f ? t : rb(t * gscale);
      s += t * t;
    }
  }
  s = block_reduce_sum(s, sbuf);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[0] = total L2 norm, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6))  (torch clip_grad_norm_)
__global__ void gradnorm_final_kernel(const float* __restrict__ partial, int n, float max_norm, float* __restrict__ out) {
  __shared__ float sbuf[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_reduce_sum(s, sbuf);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(s);
    out[0] = norm;
    // torch.clamp(coef, max=1.0) propagates a NaN norm (fminf would turn it into 1): a non-finite gradient then
    // poisons the step visibly, as it does in the reference, instead of being applied unclipped
    const float coef = max_norm / (norm + 1e-6f);
    out[1] = (max_norm > 0.f) ? ((coef != coef) ? coef : fminf(coef, 1.0f)) : 1.0f;
  }
}

// torch.optim.AdamW single-step arithmetic with bf16 state, rounding after every foreach primitive like the
// reference:
//   g  = rb(rb(g*gscale) * clip)
//   p  = rb(p * (1 - lr*wd))
//   m  = rb(m + (1-b1)*(g - m))
//   v  = rb(rb(v*b2) + (1-b2)*g*g)
//   den= rb(rb(rb(sqrt(v)) / sqrt(1-b2^t)) + eps)     (the division as a multiplication by the host-side reciprocal)
//   p  = rb(p - (lr/(1-b1^t)) * (m / den))
__global__ void __launch_bounds__(256)
adamw_kernel(const ParamMeta* __restrict__ metas, const int* __restrict__ blk_tensor,
             const long long* __restrict__ blk_off, const float* __restrict__ clip_scalars,
             const int* __restrict__ skip_flag, float gscale, float lr_mult, float one_minus_b1, float beta2,
             float one_minus_b2, float eps, float bc1, float inv_bc2_sqrt, int zero_grad) {
  const ParamMeta pm = metas[blk_tensor[blockIdx.x]];
  const long long off = blk_off[blockIdx.x];
  const long long n = min(static_cast<long long>(OPT_CHUNK), pm.numel - off);
  const float clip = clip_scalars ? clip_scalars[1] : 1.0f;
  // skip_flag != 0: the gradients of this step are known to be incomplete (masked-rows head overflow); parameters and
  // moments stay untouched, the gradient buffer is still cleared
  const bool skip = skip_flag != nullptr && *skip_flag != 0;
  const float lr = pm.lr * lr_mult;
  const float decay = 1.0f - lr * pm.wd;
  const float step = lr / bc1;
  __nv_bfloat16* P = pm.p + off;
  __nv_bfloat16* G = pm.g + off;
  __nv_bfloat16* Mm = pm.m + off;
  __nv_bfloat16* Vv = pm.v + off;
  // The kernel was instruction-issue bound at 51 % of the HBM peak (IEEE sqrt and two IEEE divisions per element):
  // sqrt.approx / rcp-based division are exact to ~2^-22, i.e. invisible after the bf16 rounding that follows each
  // of them except on rounding ties (<= 1 bf16 ulp, the tolerance of the reference-golden test).
  auto upd = [&](float& p, float g, float& m, float& v) {
    g = gscale == 1.0f ? g : rb(g * gscale);
    g = rb(g * clip);
    p = rb(p * decay);
    m = rb(m + one_minus_b1 * (g - m));
    v = rb(rb(v * beta2) + one_minus_b2 * g * g);
    float sq;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v));
    const float den = rb(rb(rb(sq) * inv_bc2_sqrt) + eps);
    p = rb(p - step * __fdividef(m, den));
  };
  auto upd8 = [&](uint4& up, const uint4& ug, uint4& um, uint4& uv) {
    uint32_t* wp = &up.x; const uint32_t* wg = &ug.x; uint32_t* wm = &um.x; uint32_t* wv = &uv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float p0 = bf16_lo(wp[j]), p1 = bf16_hi(wp[j]);
      float m0 = bf16_lo(wm[j]), m1 = bf16_hi(wm[j]);
      float v0 = bf16_lo(wv[j]), v1 = bf16_hi(wv[j]);
      upd(p0, bf16_lo(wg[j]), m0, v0);
      upd(p1, bf16_hi(wg[j]), m1, v1);
      wp[j] = pack_bf16x2(p0, p1);
      wm[j] = pack_bf16x2(m0, m1);
      wv[j] = pack_bf16x2(v0, v1);
    }
  };
  const bool aligned = ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(G) |
                         reinterpret_cast<uintptr_t>(Mm) | reinterpret_cast<uintptr_t>(Vv)) & 15) == 0;
  const long long n8 = aligned ? n / 8 : 0;
  uint4* P4 = reinterpret_cast<uint4*>(P);
  uint4* G4 = reinterpret_cast<uint4*>(G);
  uint4* M4 = reinterpret_cast<uint4*>(Mm);
  uint4* V4 = reinterpret_cast<uint4*>(Vv);
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  if (skip) {
    if (zero_grad) {
      for (long long i = threadIdx.x; i < n8; i += blockDim.x) G4[i] = zero4;
      for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) G[i] = __float2bfloat16_rn(0.f);
    }
    return;
  }
  // two 16-byte vectors of each of the four streams in flight per thread (8 independent loads before any math)
  long long i = threadIdx.x;
  for (; i + blockDim.x < n8; i += 2 * blockDim.x) {
    const long long k = i + blockDim.x;
    uint4 up0 = P4[i], ug0 = G4[i], um0 = M4[i], uv0 = V4[i];
    uint4 up1 = P4[k], ug1 = G4[k], um1 = M4[k], uv1 = V4[k];
    upd8(up0, ug0, um0, uv0);
    P4[i] = up0; M4[i] = um0; V4[i] = uv0;
    if (zero_grad) G4[i] = zero4;
    upd8(up1, ug1, um1, uv1);
    P4[k] = up1; M4[k] = um1; V4[k] = uv1;
    if (zero_grad) G4[k] = zero4;
  }
  for (; i < n8; i += blockDim.x) {
    uint4 up = P4[i], ug = G4[i], um = M4[i], uv = V4[i];
    upd8(up, ug, um, uv);
    P4[i] = up; M4[i] = um; V4[i] = uv;
    if (zero_grad) G4[i] = zero4;
  }
  for (long long j = n8 * 8 + threadIdx.x; j < n; j += blockDim.x) {
    float p = __bfloat162float(P[j]), m = __bfloat162float(Mm[j]), v = __bfloat162float(Vv[j]);
    upd(p, __bfloat162float(G[j]), m, v);
    P[j] = __float2bfloat16_rn(p);
    Mm[j] = __float2bfloat16_rn(m);
    Vv[j] = __float2bfloat16_rn(v);
    if (zero_grad) G[j] = __float2bfloat16_rn(0.f);
  }
}

}  // namespace obt

using namespace obt;

extern "C" int obt_ce_fwd(const void* logits, long long ld, const long long* targets, const unsigned char* row_mask,
                          float* lse, float* tok_loss, float* scalars, long long M, int V, float n_acc,
                          cudaStream_t stream) {
  OBT_REQUIRE(logits && targets && lse && tok_loss && scalars, "obt_ce_fwd: null pointer");
  OBT_REQUIRE(V % 8 == 0 && ld % 8 == 0, "obt_ce_fwd: V=%d ld=%lld must be multiples of 8", V, ld);
  OBT_REQUIRE(M > 0 && M < (1ll << 31), "obt_ce_fwd: bad M=%lld", M);
  ce_fwd_kernel<<<static_cast<unsigned>(M), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), ld, targets,
                                                             row_mask, lse, tok_loss, V);
  int rc = check_launch("ce_fwd");
  if (rc) return rc;
  ce_reduce_kernel<<<1, 1024, 0, stream>>>(tok_loss, row_mask, M, n_acc, scalars);
  return check_launch("ce_reduce");
}

extern "C" int obt_ce_bwd(void* logits, long long ld, const long long* targets, const unsigned char* row_mask,
                          const float* lse, const float* scalars, float upstream, const void* upstream_dev,
                          long long M, int V, int unmasked_rows_zero, cudaStream_t stream) {
  OBT_REQUIRE(logits && targets && lse && scalars, "obt_ce_bwd: null pointer");
  OBT_REQUIRE(V % 8 == 0 && ld % 8 == 0, "obt_ce_bwd: V=%d ld=%lld must be multiples of 8", V, ld);
  OBT_REQUIRE(M > 0 && M < (1ll << 31), "obt_ce_bwd: bad M=%lld", M);
  ce_bwd_kernel<<<static_cast<unsigned>(M), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(logits), ld, targets,
                                                             row_mask, lse, scalars, upstream,
                                                             static_cast<const __nv_bfloat16*>(upstream_dev), V,
                                                             unmasked_rows_zero);
  return check_launch("ce_bwd");
}

extern "C" int obt_opt_chunk_elems(void) { return OPT_CHUNK; }
extern "C" int obt_opt_meta_bytes(void) { return static_cast<int>(sizeof(ParamMeta)); }

// norm_out: device float[2] = {total norm, clip coefficient}; partial: device float[n_blocks]
extern "C" int obt_grad_norm(const void* metas, const int* blk_tensor, const long long* blk_off, int n_blocks,
                             float gscale, float max_norm, float* partial, float* norm_out, cudaStream_t stream) {
  OBT_REQUIRE(metas && blk_tensor && blk_off && partial && norm_out, "obt_grad_norm: null pointer");
  OBT_REQUIRE(n_blocks > 0, "obt_grad_norm: no blocks");
  gradnorm_partial_kernel<<<n_blocks, 256, 0, stream>>>(static_cast<const ParamMeta*>(metas), blk_tensor, blk_off,
                                                        gscale, partial);
  int rc = check_launch("gradnorm_partial");
  if (rc) return rc;
  gradnorm_final_kernel<<<1, 1024, 0, stream>>>(partial, n_blocks, max_norm, norm_out);
  return check_launch("gradnorm_final");
}

extern "C" int obt_adamw_step(const void* metas, const int* blk_tensor, const long long* blk_off, int n_blocks,
                              const float* clip_scalars, const int* skip_flag, float gscale, float lr_mult,
                              double beta1, double beta2, double eps, int step, int zero_grad, cudaStream_t stream) {
  OBT_REQUIRE(metas && blk_tensor && blk_off, "obt_adamw_step: null pointer");
  OBT_REQUIRE(n_blocks > 0 && step >= 1, "obt_adamw_step: bad n_blocks=%d step=%d", n_blocks, step);
  const float bc1 = static_cast<float>(1.0 - pow(beta1, static_cast<double>(step)));
  const float inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(1.0 - pow(beta2, static_cast<double>(step))));
  // (1 - beta) is formed in double like the reference's Python floats, then narrowed once
  adamw_kernel<<<n_blocks, 256, 0, stream>>>(static_cast<const ParamMeta*>(metas), blk_tensor, blk_off, clip_scalars,
                                             skip_flag, gscale, lr_mult, static_cast<float>(1.0 - beta1),
                                             static_cast<float>(beta2), static_cast<float>(1.0 - beta2),
                                             static_cast<float>(eps), bc1, inv_bc2_sqrt, zero_grad);
  return check_launch("adamw");
}
