/*
 * omnibiote_b200 — C ABI of the B200 (sm_100a) kernels behind the OmniBioTA encoder hot path.
 *
 * The reference (nyuolab/OmniBioTE) has no native/FFI layer: its boundary is the Python class API of
 * training/model.py, and every device op is a PyTorch library call. Each entry point below therefore cites the
 * reference Python line(s) whose computation it replaces. The Python module omnibiote_b200/model.py mirrors the
 * reference classes and calls these functions through ctypes with raw device pointers and the current CUDA stream.
 *
 * Conventions
 *  - plain pointers and sizes only; all tensors are bf16 (uint16 storage) unless stated; row-major.
 *  - the caller owns every buffer (including workspaces); the library never allocates or frees device memory and
 *    keeps no reference to caller buffers (the only global state is a mutex-guarded TMA-descriptor cache).
 *  - every function enqueues on `stream` and returns immediately: 0 on success, negative on error
 *    (-1 invalid argument, -2 CUDA error, -3 unsupported); obt_last_error() returns a thread-local message.
 *  - no CPU fallback: pointers must be device pointers on the current device.
 *  - rb(x) below = round-to-nearest-even to bf16.
 */
#ifndef OMNIBIOTE_B200_H_
#define OMNIBIOTE_B200_H_

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* obt_last_error(void);
int obt_version(void);
void obt_clear_descriptor_cache(void);

/* ---- GEMM (tcgen05 / TMEM / TMA) ------------------------------------------------------------------------------
 * D[M,N] = epilogue( op(A)[M,K] * op(B)[N,K]^T ), fp32 accumulation.
 *   a_mn_major = 0: A stored [M][K] (pitch lda)   = 1: A stored [K][M] (pitch lda)
 *   b_mn_major = 0: B stored [N][K] (pitch ldb)   = 1: B stored [K][N] (pitch ldb)
 * Replaces nn.Linear forward (model.py:102,151,163,166; MuReadout model.py:208,253) = (0,0);
 * its input gradient (autograd, train_encoder.py:308) = (0,1); its weight gradient = (1,1).
 * epilogue: 0 plain rb(acc)
 *           1 D = rb(aux_in + rb(acc))                 residual add (model.py:179-180) / .grad accumulation
 *           2 aux_out = U = rb(acc); D = rb(gelu(U))   fused_gelu (model.py:23-25)
 *           3 D = rb(rb(acc) * gelu'(aux_in))          backward of fused_gelu, aux_in = U
 *           5 D = rb(aux_in + dropout(rb(acc)))        resid_dropout + residual (model.py:151,167,179-180)
 *           9 aux_out = rb(gelu'(U)), D = rb(gelu(U)) with U = rb(acc): forward of fused_gelu that saves the DERIVATIVE
 *          10 D = rb(rb(acc) * aux_in): its backward (aux_in = the saved derivative); 9 + 10 replace 2 + 3 in the block
 *          11 D = rb(acc) and workspace[b, h, t] (fp32 [M / rope_T, N / 128, rope_T]) = sum over the 128 columns of
 *             head h of D * aux_in: the attention backward's delta = rowsum(dO * O) emitted by the GEMM that produces
 *             dO = d_a Wo (aux_in = y of model.py:148; rope_T = sequence length; N %% 128 == 0)
 *           8 D = aux_in[row] ? rb(acc) : 0 with aux_in a uint8 row mask [M]: MLM head whose unmasked rows are never
 *             read (zero loss weight and gradient, train_encoder.py:301-305); pairs with obt_ce_bwd(unmasked_rows_zero)
 *           7 D = rb(rotary(rb(acc))) on columns < rope_cols: apply_rotary_emb (model.py:39-50,108) fused into
 *             c_attn; rope_cos / rope_sin are fp32 [rope_T, rope_head_dim/2] tables (position = row % rope_T),
 *             rope_sin = NULL for the real bf16 freqs_cis buffer a bf16 model carries (cosine scaling)
 * gelu_mode: 0 = single rounding (TorchScript-fused execution), 1 = one bf16 rounding per primitive (eager CPU).
 * workspace (fp32, optional): enables split-K for small-output / long-reduction shapes (weight gradients).
 */
void obt_gemm_set_cta_group(int cta_group); /* 0 auto, 1 or 2 forced (tests, ablations) */
long long obt_gemm_workspace_elems(long long M, long long N, long long K);
int obt_gemm_bf16(const void* A, const void* B, void* D, long long M, long long N, long long K, long long lda,
                  long long ldb, long long ldd, int a_mn_major, int b_mn_major, int epilogue, const void* aux_in,
                  long long ld_aux_in, void* aux_out, long long ld_aux_out, int gelu_mode, float drop_p,
                  unsigned long long seed, unsigned long long offset, void* workspace, long long workspace_elems,
                  const float* rope_cos, const float* rope_sin, int rope_T, int rope_head_dim, int rope_cols,
                  cudaStream_t stream);

/* ---- embedding: nn.Embedding + nn.Dropout(inplace) (model.py:241-242) ----------------------------------------- */
int obt_embed_fwd(const long long* idx, const void* wte, void* out, long long M, int C, int V, float drop_p,
                  unsigned long long seed, unsigned long long offset, int* err_flag, cudaStream_t stream);
/* dense (V,C) gradient like embedding_dense_backward; scratch fp32 [V*C] and touched int32 [V] must be zero on
 * entry and are zero again on exit. */
int obt_embed_bwd(const long long* idx, const void* dout, void* dwte, float* scratch, int* touched, long long M, int C,
                  int V, int accumulate, float drop_p, unsigned long long seed, unsigned long long offset,
                  cudaStream_t stream);

/* ---- LayerNorm: F.layer_norm(x, (C,), weight, None, 1e-5) (model.py:63-72) -------------------------------------
 * z (optional) = rb(y / readout_div): the MuReadout input scaling x / width_mult (mup MuReadout.forward). */
int obt_layernorm_fwd(const void* x, const void* gamma, void* y, void* z, float* mean, float* rstd, long long M, int C,
                      float eps, float readout_div, cudaStream_t stream);
int obt_layernorm_bwd_workspace_rows(void);
/* dx = rb(dres + rb(ln_bwd(rb(dy / dy_div)))); dgamma (+)= rb(sum_rows dy * xhat), reduced inside the same launch.
 * workspace fp32 [32 + rows*C], words 0..1 zero before the first call (the kernel re-zeroes them).
 * dx_drop (optional, NULL to skip) = dropout replay of dx with (drop_p, seed, offset), the mask of obt_dropout /
 * GEMM epilogue 5 on the same flat indices: the upstream gradient of the residual branch `x + dropout(f(x))`
 * (model.py:151,167) whose output this LayerNorm normalised. */
int obt_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                      const void* dres, void* dx, void* dgamma, int accumulate_dgamma, float* workspace, long long M,
                      int C, float dy_div, void* dx_drop, float drop_p, unsigned long long seed,
                      unsigned long long offset, cudaStream_t stream);

/* ---- rotary embedding, in place on the q and k column ranges of the fused qkv buffer (model.py:39-50,108) ------
 * sin_tab == NULL: bf16 model whose complex freqs_cis buffer was cast to a real bf16 table (cosine scaling).
 * tables are fp32 [>=T][head_dim/2]. inverse != 0 applies the adjoint (backward). */
int obt_rope(void* qkv, const float* cos_tab, const float* sin_tab, long long M, int T, int C, int head_dim,
             long long ld, int inverse, cudaStream_t stream);

/* ---- dropout (nn.Dropout): out = keep ? rb(in/(1-p)) : 0, mask defined by (seed, offset, flat index) ---------- */
int obt_dropout(const void* in, void* out, long long n, float p, unsigned long long seed, unsigned long long offset,
                cudaStream_t stream);

/* ---- MuReadout input scaling: out = rb(in / div), div = width_mult / output_mult (mup MuReadout.forward) ------- */
int obt_scale_div(const void* in, void* out, long long n, float div, cudaStream_t stream);

/* ---- encode() pooling (model.py:269-278): mode 0 = mean over T, 1 = max over T -------------------------------- */
int obt_pool_splits(int T);
int obt_pool(const void* emb, void* out, float* workspace, int B, int T, int C, int mode, cudaStream_t stream);
int obt_pool_bwd(const void* emb, const void* pooled, const void* dout, void* demb, int B, int T, int C, int mode,
                 cudaStream_t stream);

/* ---- attention: F.scaled_dot_product_attention(q,k,v,attn_mask,dropout_p,scale=8/n_embd) (model.py:111-138) ---
 * q,k,v are column ranges of the fused qkv buffer (row pitch ld); y is written head-major into [M, ldy] so the
 * transpose+contiguous of model.py:148 disappears. mask: additive bf16, element (b,h,i,j) at
 * mask[b*msb + h*msh + i*msq + j] (msh = 0 for the head-expanded view of train_encoder.py:292), or NULL.
 * row_lo/row_hi (optional, instead of mask): per (b,i) visible key interval; lo >= hi marks a fully-masked row
 * (uniform attention). lse: fp32 [B,H,T,2] = (row max, log exp-sum).
 * Attention dropout (dropout_p of model.py:118,134) is a precomputed bit matrix keep[B,H,T,ceil(T/32)] drawn ONCE
 * per layer and micro-batch by obt_attn_keep_mask from (seed, offset) and only read by the forward and backward
 * kernels (key j of query (b,h,i): word j/32, bit 8*(j&3) + 7 - ((j&31)>>2)); keep may be NULL when drop_p == 0. */
/* row_lo / row_hi (optional, int32 [B,T]): the interval mask the kernels will use; words of a row entirely outside its
 * visible interval are then stored as all-ones instead of being drawn (they only ever multiply zero probabilities). */
int obt_attn_keep_mask(unsigned int* keep, int B, int H, int T, float drop_p, unsigned long long seed,
                       unsigned long long offset, const int* row_lo, const int* row_hi, cudaStream_t stream);
int obt_attn_simt_fwd(const void* q, const void* k, const void* v, long long ld, const void* mask, long long msb,
                      long long msh, long long msq, const int* row_lo, const int* row_hi, void* y, long long ldy,
                      float* lse, int B, int H, int T, int d, float scale, float drop_p, const unsigned int* keep,
                      cudaStream_t stream);
int obt_attn_simt_bwd(const void* q, const void* k, const void* v, long long ld, const void* mask, long long msb,
                      long long msh, long long msq, const int* row_lo, const int* row_hi, const void* y, long long ldy,
                      const void* dy, long long lddy, const float* lse, float* delta, void* dq, void* dk, void* dv,
                      long long ldd, int B, int H, int T, int d, float scale, float drop_p, const unsigned int* keep,
                      cudaStream_t stream);

/* tensor-core (tcgen05/TMEM/TMA) forward for head_dim == 128; qkv is the fused [M,3C] buffer (q | k | v). */
int obt_attn_tc_fwd(const void* qkv, long long ld, const void* mask, long long msb, long long msh, long long msq,
                    const int* row_lo, const int* row_hi, void* y, long long ldy, float* lse, int B, int H, int T, int d,
                    float scale, float drop_p, const unsigned int* keep, const int* qmeta, cudaStream_t stream);

/* Tile metadata of an interval mask (row_lo / row_hi, int32 [B,T]), computed once per micro-batch and shared by every
 * layer, head and tensor-core attention kernel instead of a scan of the intervals in each of their CTAs:
 *   qmeta int32 [B, ceil(T/128), 4] = {min lo, max hi, any fully-masked row, 0} per 128-query tile
 *   kmeta uint32 [B, ceil(T/128), 4] = relevance bits of the 64-query sub-tiles per 128-key tile
 * Both are optional (NULL) inputs of obt_attn_tc_fwd / obt_attn_tc_bwd. */
int obt_attn_tile_meta(const int* row_lo, const int* row_hi, int B, int T, int* qmeta, unsigned int* kmeta,
                       cudaStream_t stream);

/* tensor-core backward (head_dim == 128): delta pre-pass + dQ kernel + dK/dV kernel. dqkv is the fused [M,3C]
 * gradient buffer (dq | dk | dv, pitch ldd); delta is fp32 [B,H,T]: scratch written by the pre-pass, or, with
 * delta_ready != 0, an INPUT already holding rowsum(dy * y) per head (obt_gemm_bf16 epilogue 11 emits it while
 * producing dy, and the pre-pass is skipped); autograd adjoint of model.py:111-148.
 * qmeta / kmeta: optional tile metadata (obt_attn_tile_meta).
 * ds_scratch: optional bf16 [B*H * T * roundup(T,64)] scratch (128-byte aligned, contents irrelevant): when given, the
 * dK/dV kernel also stores its dS^T tiles there and dQ = scale/(1-p) * dS K is accumulated from them by a score-free
 * kernel (no second evaluation of Q K^T, dO V^T and the exponentials); NULL = the stand-alone dQ kernel.
 * rope_cos / rope_sin (optional, fp32 [>=T, 64]): when given, the adjoint of apply_rotary_emb is applied to dq and dk
 * in the kernels' epilogues (rope_sin = NULL: cosine scaling), so dqkv is the gradient of the PRE-rotary c_attn output. */
int obt_attn_tc_bwd(const void* qkv, long long ld, const void* mask, long long msb, long long msh, long long msq,
                    const int* row_lo, const int* row_hi, const void* y, long long ldy, const void* dy, long long lddy,
                    const float* lse, float* delta, int delta_ready, void* dqkv, long long ldd, int B, int H, int T,
                    int d, float scale, float drop_p, const unsigned int* keep, const float* rope_cos,
                    const float* rope_sin, const int* qmeta, const unsigned int* kmeta, void* ds_scratch,
                    cudaStream_t stream);

/* ---- attention-mask producers / compressors (input contract of the hot path) ------------------------------------
 * obt_doc_mask_intervals : per (b,i) visible key interval [lo,hi) from token ids = create_attention_mask
 *                          (train_encoder.py:25-57) incl. its quirks; lo >= hi marks a fully-masked row.
 * obt_pad_mask_intervals : pad_attn (evals/gue.py:15-21).
 * obt_mask_from_intervals: dense additive bf16 (B,T,T) {0,-1e9} tensor (what train_encoder.py:290-291 builds).
 * obt_mask_compress      : dense additive mask (strides msb, msq, 1) -> intervals; *not_interval (pre-zeroed device
 *                          int) is set when the mask is not exactly interval-structured. */
int obt_doc_mask_intervals(const long long* ids, int* lo, int* hi, int B, int T, long long eos_token, int padding,
                           cudaStream_t stream);
int obt_pad_mask_intervals(const long long* ids, int* lo, int* hi, int B, int T, long long pad_token,
                           cudaStream_t stream);
int obt_mask_from_intervals(const int* lo, const int* hi, void* mask, int B, int T, cudaStream_t stream);
int obt_mask_compress(const void* mask, long long msb, long long msq, int* lo, int* hi, int* not_interval, int B, int T,
                      cudaStream_t stream);

/* ---- MLM input masking (train_encoder.py:273-279): mask = Bernoulli(prob) & id != PAD & id != EOS;
 * masked_ids = mask ? MASK : id. Device Philox stream instead of the reference's host numpy RNG.
 * counters (optional fp32[2]): += {number of masked positions, number of non-PAD tokens} (train_encoder.py:350). */
int obt_mlm_mask(const long long* ids, long long* masked_ids, unsigned char* mask, long long n, float prob,
                 unsigned long long seed, unsigned long long offset, long long pad_token, long long eos_token,
                 long long mask_token, float* counters, cudaStream_t stream);

/* ---- masked-rows-only head (optional path of the training step): d loss / d logits of train_encoder.py:301-305 is
 * exactly zero outside the MLM mask, so ln_f's output rows inside the mask are compacted, the head GEMM, the CE and
 * their backward run on those rows only, and the input gradient is scattered back (zeros elsewhere). Results are
 * identical to the dense path; the dense path stays the default (it is what the reference executes).
 * obt_compact_rows: idx[cap] = rows with mask != 0 in row order, -1 padded; targets_c / valid_c per slot;
 *                   meta = {count, count > cap}. */
int obt_compact_rows(const unsigned char* mask, const long long* targets, long long M, int cap, int* idx,
                     long long* targets_c, unsigned char* valid_c, int* meta, cudaStream_t stream);
int obt_gather_rows(const void* src, long long lds, const int* idx, void* dst, long long ldd, int n_slots, int C,
                    cudaStream_t stream);
/* dst [M, C] contiguous: zero-filled, then dst[idx[s]] = src[s] for idx[s] >= 0 */
int obt_scatter_rows(const void* src, long long lds, const int* idx, void* dst, long long ldd, long long M, int n_slots,
                     int C, cudaStream_t stream);

/* ---- MLM loss (train_encoder.py:301-305) over materialised logits ----------------------------------------------
 * scalars (device fp32[4]): [0] loss, [1] number of masked tokens, [2] d loss / d CE_t, [3] n_acc.
 * row_mask: uint8 [M] or NULL. */
int obt_ce_fwd(const void* logits, long long ld, const long long* targets, const unsigned char* row_mask, float* lse,
               float* tok_loss, float* scalars, long long M, int V, float n_acc, cudaStream_t stream);
/* overwrites logits with d loss / d logits (exact zeros on unmasked rows) for an incoming d loss =
 * upstream * (upstream_dev ? *upstream_dev : 1); upstream_dev is a bf16 device scalar (autograd's grad of the bf16
 * loss) or NULL. unmasked_rows_zero != 0: the caller
 * guarantees those rows already hold zeros (head GEMM with epilogue 8), so they are neither read nor written. */
int obt_ce_bwd(void* logits, long long ld, const long long* targets, const unsigned char* row_mask, const float* lse,
               const float* scalars, float upstream, const void* upstream_dev, long long M, int V,
               int unmasked_rows_zero, cudaStream_t stream);

/* ---- clip_grad_norm_ (train_encoder.py:316) + MuAdamW step (train_encoder.py:199,317) --------------------------
 * metas: device array of {void* p, g, m, v; int64 numel; float lr, wd} (obt_opt_meta_bytes() each);
 * blk_tensor/blk_off: one entry per block of obt_opt_chunk_elems() elements. */
int obt_opt_chunk_elems(void);
int obt_opt_meta_bytes(void);
int obt_grad_norm(const void* metas, const int* blk_tensor, const long long* blk_off, int n_blocks, float gscale,
                  float max_norm, float* partial, float* norm_out, cudaStream_t stream);
int obt_adamw_step(const void* metas, const int* blk_tensor, const long long* blk_off, int n_blocks,
                   const float* clip_scalars, const int* skip_flag, float gscale, float lr_mult, double beta1,
                   double beta2, double eps, int step, int zero_grad, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OMNIBIOTE_B200_H_ */
