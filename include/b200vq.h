/* b200vq.h -- C ABI of libb200vq.so: the B200 (sm_100a) kernels behind the ViT-VQGAN hot path of
 * thuanz123/enhancing-transformers.
 *
 * Boundary.  The reference has no FFI of its own on this path: its
 * ViTEncoder / ViTDecoder / VectorQuantizer are nn.Modules whose arithmetic lives in ATen.  The
 * drop-in boundary is therefore the nn.Module surface (enhancing_transformers_b200/layers.py,
 * quantizers.py) and *this* header is what those modules bind with ctypes -- each entry point
 * names the reference line(s) whose ATen calls it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless typed otherwise; nothing is allocated or
 *     freed inside the library, workspaces are sized by the *_workspace_bytes queries and owned
 *     by the caller (PyTorch's caching allocator on the Python side);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronises the device, and returns 0 on success or a negative code; the message is
 *     available from b200vq_last_error() (thread local);
 *   - matrices are row-major with an explicit leading dimension in elements;
 *   - `round_out != 0` stores values rounded to tf32 (round-to-nearest, fp32 container) because
 *     the tensor is only ever consumed as a tcgen05 kind::tf32 operand, which would otherwise
 *     truncate the low 13 mantissa bits;
 *   - fp16 buffers (void*) are plain IEEE binary16 arrays; they only ever hold GEMM operands.
 */
#ifndef B200VQ_H_
#define B200VQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VQ_VERSION 202 /* 0.2.2 */

int b200vq_version(void);
const char* b200vq_last_error(void);
/* compiled-for architecture string, e.g. "sm_100a" */
const char* b200vq_arch(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
long long b200vq_launch_count(void);
/* cap the grid of the persistent (one CTA per SM) GEMM / attention kernels at n CTAs, 0 = no cap.  Leaves SMs to a
 * concurrently running NCCL all-reduce (reference main.py:54-57 strategy="ddp").  Also read once from the environment
 * variable B200VQ_SM_LIMIT. */
int b200vq_set_sm_limit(int n);

/* ---- GEMM: C[M,N] = epi( alpha * A . B^T ) on tcgen05 tensor cores, fp32 accumulate ------------------
 * Replaces nn.Linear / Conv2d(k=s) / ConvTranspose2d(k=s) forward, dgrad and wgrad:
 *   layers.py:99-101 (FeedForward), :118,:120 (to_qkv, to_out), :169 (patch embed), :204 (to_pixel),
 *   vitvqgan.py:38-39 (pre_quant / post_quant).
 * a_major/b_major: 0 = operand stored [rows, K] (K contiguous); 1 = stored [K_total, rows]
 * (rows contiguous).  splits > 1 contracts K_total = K*splits in `splits` slices and writes
 * slice z to C + z*c_split_stride (reduce with b200vq_splitk_reduce).
 * Epilogue, in this order: * alpha; + bias[N]; act (0 none, 1 tanh -- layers.py:100);
 * * (1 - aux^2) (tanh backward); + res[row % res_row_mod or row] (residual add layers.py:147-148
 * or positional table :179,:210); tf32 rounding / fp16 conversion.
 * colsum_part (nullable, [ceil(M/32)][N], splits must be 1): receives the column sums of every 32-row group of
 * the stored C -- summed over the groups (b200vq_colsum) they are the bias gradient of the Linear whose
 * pre-activation gradient this GEMM produced (net.0, layers.py:99), without another pass over C.
 * cta_group: 1 = one CTA per 128 x bn tile, 2 = CTA pair per 256 x bn tile; bn in {0=auto,64,128,192,256}.
 *
 * Three operand flavours:
 *   gemm_tf32   A, B fp32, kind::tf32 (10-bit mantissa; the hardware truncates, so producers round to nearest).
 *   gemm_3xtf32 A, B fp32 plus their truncation residues A_lo, B_lo (b200vq_split_tf32_lo): the
 *               error-compensated product A_lo.B + A.B_lo + A.B -- fp32-grade, used by precision="parity".
 *   gemm_f16    A, B fp16 (kind::f16: the same 11-bit significand as tf32 at twice the tensor rate and half the
 *               operand bytes); C fp32, or fp16 when out_half (then aux is fp16 too; no residual / split-K).
 *               alpha (device scalar, nullable) undoes the power-of-two scale fp16 gradient operands carry. */
int b200vq_gemm_tf32(const float* A, long long lda, int a_major, const float* B, long long ldb, int b_major,
                     float* C, long long ldc, int M, int N, int K, int splits, long long c_split_stride,
                     const float* bias, const float* res, long long ldres, int res_row_mod,
                     const float* aux, long long ldaux, float* colsum_part, int act, int round_out, int cta_group,
                     int bn, void* stream);
int b200vq_gemm_3xtf32(const float* A, const float* A_lo, long long lda, int a_major, const float* B, const float* B_lo,
                       long long ldb, int b_major, float* C, long long ldc, int M, int N, int K, int splits,
                       long long c_split_stride, const float* bias, const float* res, long long ldres, int res_row_mod,
                       const float* aux, long long ldaux, float* colsum_part, int act, int cta_group, int bn, void* stream);
int b200vq_gemm_f16(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major, void* C,
                    long long ldc, int out_half, int M, int N, int K, int splits, long long c_split_stride,
                    const float* bias, const float* res, long long ldres, int res_row_mod, const void* aux,
                    long long ldaux, float* colsum_part, int act, int round_out, const float* alpha, int cta_group,
                    int bn, void* stream);
/* out = alpha * sum_z part[z] (alpha: device scalar or NULL) */
int b200vq_splitk_reduce(const float* part, int splits, long long n, long long split_stride, const float* alpha,
                         float* out, void* stream);

/* ---- LayerNorm (nn.LayerNorm(dim), eps 1e-5: layers.py:88,143) ---------------------------------
 * y (fp32, optionally tf32-rounded) and/or y16 (fp16) receive the normalised rows; either may be NULL. */
int b200vq_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, void* y16, float* mean,
                         float* rstd, int M, int D, int round_out, void* stream);
size_t b200vq_layernorm_bwd_workspace_bytes(int D);
/* dx = LN'(dy) (+ dres, the skip-connection gradient of layers.py:147-148).  dxsum (nullable, [D]) receives the
 * column sums of dx: dx is also the gradient at the bias of the Linear that wrote this residual stream
 * (to_out layers.py:118, net.2 layers.py:101), so that bias gradient costs no extra pass over dx.
 * dx16 (nullable): fp16 copy of dx multiplied by *dx16_scale, the operand of the next block's fp16 GEMMs.
 * dy: fp32, or (dy_half = 1) fp16 carrying a gradient scale -- the fp16 output of a dgrad GEMM -- multiplied by
 * *dy_scale (device scalar 1/S, or NULL) as it is read. */
int b200vq_layernorm_bwd(const void* dy, int dy_half, const float* dy_scale, const float* x, const float* mean,
                         const float* rstd, const float* gamma, const float* dres, float* dx, void* dx16,
                         const float* dx16_scale, float* dgamma, float* dbeta, float* dxsum, int M, int D, int round_out,
                         void* workspace, size_t ws_bytes, void* stream);

/* ---- attention core (layers.py:124-130): softmax(q k^T * scale) v, no mask -----------------------
 * qkv is the [B*N, 3*heads*dh] output of to_qkv (q | k | v thirds, head h at columns h*dh);
 * out is [B*N, heads*dh] (fp32, or fp16 when out_half); lse [B*heads*N] holds log-sum-exp of the scaled scores
 * (saved for backward instead of the [B,h,N,N] probabilities the reference keeps).  dh must be 64 or 32. */
int b200vq_attention_fwd(const float* qkv, void* out, int out_half, float* lse, int B, int N, int heads, int dh,
                         float scale, int round_out, void* stream);
/* dqkv [B*N, 3*heads*dh] from dout (fp32); `out` as the forward stored it; delta [B*heads*N] is scratch
 * (rowsum(dout*out)).  dqkv_half: store fp16, multiplied by *dqkv_scale (nullable). */
int b200vq_attention_bwd(const float* qkv, const void* out, int out_half, const float* lse, const float* dout,
                         void* dqkv, int dqkv_half, const float* dqkv_scale, float* delta, int B, int N, int heads,
                         int dh, float scale, int round_out, void* stream);
/* fp16-operand core (dh == 64): qkv16 / out16 / dout16 / dqkv16 are fp16 matrices of the shapes above; tensor-core
 * operands fp16 (kind::f16), scores / softmax / accumulation fp32.  dout16 carries the power-of-two gradient scale of
 * the backward segment and dqkv16 comes out with the same scale (everything is linear in dout). */
int b200vq_attention_f16_fwd(const void* qkv16, void* out16, float* lse, int B, int N, int heads, int dh, float scale,
                             void* stream);
int b200vq_attention_f16_bwd(const void* qkv16, const void* out16, const float* lse, const void* dout16, void* dqkv16,
                             float* delta, int B, int N, int heads, int dh, float scale, void* stream);
/* the same contract with every product in error-compensated 3xTF32 (fp32-grade; precision="parity") */
int b200vq_attention_exact_fwd(const float* qkv, float* out, float* lse, int B, int N, int heads, int dh, float scale,
                               void* stream);
int b200vq_attention_exact_bwd(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv,
                               float* delta, int B, int N, int heads, int dh, float scale, void* stream);

/* ---- vector quantiser (quantizers.py:38-92) ------------------------------------------------------
 * z [M,D], codebook E [K,D] (un-normalised nn.Embedding weight); out [M,D] = straight-through
 * value z + (z_q - z); idx int64 [M, depth]; loss: device scalar.  depth = 1 for the plain
 * quantiser, num_quantizers for use_residual=True.  use_norm: quantizers.py:24 (1 = l2-normalise). */
size_t b200vq_vq_workspace_bytes(int M, int K, int depth);
int b200vq_vq_fwd(const float* z, const float* E, float* out, long long* idx, float* loss, int M, int K, int D,
                  int depth, float beta, int use_norm, void* workspace, size_t ws_bytes, void* stream);
/* g_out [M,D] or NULL, g_loss device scalar or NULL; gz [M,D] written, gE [K,D] zero-filled then
 * accumulated (nn.Embedding dense backward). */
int b200vq_vq_bwd(const float* z, const float* E, const long long* idx, const float* g_out, const float* g_loss,
                  float* gz, float* gE, int M, int K, int D, int depth, int residual, float beta, int use_norm,
                  void* stream);
/* decode_codes (vitvqgan.py:81-86): out[m] = sum_t norm(E[codes[m,t]]) */
int b200vq_vq_embed(const float* E, const long long* codes, float* out, int M, int K, int D, int depth, int use_norm,
                    void* stream);

/* ---- re-layout / reductions ------------------------------------------------------------------------*/
/* [B,C,H,W] -> [B*(H/ph)*(W/pw), C*ph*pw], patch vector ordered (c,row,col) (layers.py:157-171); pw % 4 == 0 */
int b200vq_patchify(const float* img, float* patches, int B, int C, int H, int W, int ph, int pw, int round_out,
                    void* stream);
/* inverse, adding bias[c] if non-NULL (ConvTranspose2d bias, layers.py:204) */
int b200vq_unpatchify(const float* tokens, const float* bias, float* img, int B, int C, int H, int W, int ph, int pw,
                      void* stream);
size_t b200vq_colsum_workspace_bytes(int N);
/* out[n] = sum_m X[m,n] (bias gradients) */
int b200vq_colsum(const float* X, long long ld, int M, int N, float* out, void* workspace, size_t ws_bytes, void* stream);
/* out = tf32-rounded copy of in (weights shadow for the tensor-core path) */
int b200vq_round_tf32(const float* in, float* out, long long n, void* stream);

/* out[m,:] = x[m,:] + table[m % R,:]  (token + de_pos_embedding, layers.py:210) */
int b200vq_add_rows_mod(const float* x, const float* table, float* out, long long M, int D, int R, void* stream);


/* ---- operand preparation for the 3xTF32 and fp16 data paths --------------------------------------*/
/* lo = in - trunc_tf32(in): the bits kind::tf32 drops (exact in fp32) */
int b200vq_split_tf32_lo(const float* in, float* lo, long long n, void* stream);
/* out16 = fp16(in * *scale), saturating at +-65504 (scale: device scalar or NULL) */
int b200vq_to_half(const float* in, void* out16, long long n, const float* scale, void* stream);
/* scale2 = {S, 1/S}, S = 2^(target_log2 - ceil(log2(max|g|))): the power-of-two gradient scale of one backward
 * segment, chosen on the device from the incoming gradient (no host synchronisation) */
size_t b200vq_grad_scale_workspace_bytes(void);
int b200vq_grad_scale(const float* g, long long n, int target_log2, float* scale2, void* workspace, size_t ws_bytes,
                      void* stream);

/* ---- the loss stage's two native ops ------------------------------------------
 * bias_act: out = act'(x + bias[(i / step_b) % size_b], ref) * scale -- losses/op/fused_bias_act_kernel.cu:18-65.
 *   act 1 linear / 3 leaky ReLU (slope alpha); grad 0 value, 1 first derivative applied to x (ref = forward output),
 *   2 second derivative (zero).  bias / ref nullable.
 * upfirdn2d: per plane [in_h, in_w]: zero-insert upsample (up), pad / crop, FIR with `kernel` [kh, kw] (true
 *   convolution), decimate (down) -- losses/op/upfirdn2d_kernel.cu:107-207, upfirdn2d.py:168-206.
 *   out_h = (in_h*up_y + pad_y0 + pad_y1 - kh + down_y) / down_y, likewise out_w. */
int b200vq_bias_act(const float* x, const float* bias, const float* ref, float* out, long long n, int step_b, int size_b,
                    int act, int grad, float alpha, float scale, void* stream);
int b200vq_upfirdn2d(const float* in, const float* kernel, float* out, long long planes, int in_h, int in_w, int kh, int kw,
                     int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                     void* stream);

/* ---- stage-2 transformer (reference enhancing/modules/stage2/layers.py) --------------------
 * attention_causal: the stage-1 attention contract (same packed qkv / out / lse / delta layouts, fp32) with the stage-2
 *   mask of layers.py:43-48,83-85: query q attends to key k iff k <= max(q, cond_len - 1) -- causal, the cond_len
 *   condition tokens fully visible to each other.  exact != 0: error-compensated 3xTF32 (precision="parity");
 *   otherwise the tcgen05 kind::tf32 kernels (round_out as in b200vq_attention_fwd).  dh (the head size
 *   embed_dim / n_heads) must be 32 or 64.
 * time_mix: y = x * w + shift(x) * (1 - w), shift = one step along T with a zero first row (layers.py:50-58), bit-identical
 *   to the reference's four fp32 tensor ops.  x, y [M = B*T, C]; w [C].
 *   bwd: gx likewise; gw_part [ceil(M / 64), C] holds per-64-row partial sums of g * (x - shift(x)) -- b200vq_colsum over
 *   it is the gradient of w.
 * sqrelu: grad 0: y = relu(x)^2 (layers.py:108); grad 1: y = g * 2 relu(x), x the pre-activation.
 * token_embed: x[b] = cat(Wc[conds[b]] + pos_c, Wi[codes[b]] + pos_i) (layers.py:199-206); conds int64 [B, Tc], codes
 *   int64 [B, Ti], x [B, Tc + Ti, C].  bwd: gWc / gWi are zero-filled then accumulated, gpos_* = sum over the batch.
 * copy_rows: dst[b, t] = src[b, t - off_dst + off_src] for off_dst <= t < off_dst + n, else 0 (the window
 *   x[:, cond-1:-1] of layers.py:210 and its zero-padded gradient).
 * decode_attention: one sampling step of layers.py:66-81 (use_cache with layer_past): appends the k / v thirds of
 *   qkv [B, 3*C] to cache_k / cache_v [B, Tmax, C] at row `pos` and writes softmax(q K^T * scale) V over rows 0..pos
 *   to out [B, C]. */
int b200vq_attention_causal_fwd(const float* qkv, float* out, float* lse, int B, int N, int heads, int dh, float scale,
                                int cond_len, int exact, int round_out, void* stream);
int b200vq_attention_causal_bwd(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv,
                                float* delta, int B, int N, int heads, int dh, float scale, int cond_len, int exact,
                                int round_out, void* stream);
int b200vq_time_mix_fwd(const float* x, const float* w, float* y, long long M, int T, int C, int round_out, void* stream);
size_t b200vq_time_mix_bwd_workspace_bytes(long long M, int C);
int b200vq_time_mix_bwd(const float* g, const float* x, const float* w, float* gx, float* gw_part, long long M, int T, int C,
                        void* stream);
int b200vq_sqrelu(const float* x, const float* g, float* y, long long n, int grad, int round_out, void* stream);
int b200vq_token_embed_fwd(const long long* conds, const long long* codes, const float* Wc, const float* pos_c, const float* Wi,
                           const float* pos_i, float* x, int B, int Tc, int Ti, int C, int Vc, int Vi, void* stream);
int b200vq_token_embed_bwd(const long long* conds, const long long* codes, const float* g, float* gWc, float* gpos_c, float* gWi,
                           float* gpos_i, int B, int Tc, int Ti, int C, int Vc, int Vi, void* stream);
int b200vq_copy_rows(const float* src, float* dst, int B, int T_src, int T_dst, int off_src, int off_dst, int n, int C,
                     void* stream);
int b200vq_decode_attention(const float* qkv, float* cache_k, float* cache_v, float* out, int B, int heads, int hs, int Tmax,
                            int pos, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VQ_H_ */
