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
 *     truncate the low 13 mantissa bits.
 */
#ifndef B200VQ_H_
#define B200VQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VQ_VERSION 100 /* 0.1.0 */

int b200vq_version(void);
const char* b200vq_last_error(void);
/* compiled-for architecture string, e.g. "sm_100a" */
const char* b200vq_arch(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
long long b200vq_launch_count(void);

/* ---- GEMM: C[M,N] = epi( A . B^T ) on tcgen05 tensor cores, tf32 in / fp32 accumulate ---------
 * Replaces nn.Linear / Conv2d(k=s) / ConvTranspose2d(k=s) forward, dgrad and wgrad:
 *   layers.py:99-101 (FeedForward), :118,:120 (to_qkv, to_out), :169 (patch embed), :204 (to_pixel).
 * a_major/b_major: 0 = operand stored [rows, K] (K contiguous); 1 = stored [K_total, rows]
 * (rows contiguous).  splits > 1 contracts K_total = K*splits in `splits` slices and writes
 * slice z to C + z*c_split_stride (reduce with b200vq_splitk_reduce).
 * Epilogue, in this order: + bias[N]; act (0 none, 1 tanh -- layers.py:100);
 * * (1 - aux^2) (tanh backward); + res[row % res_row_mod or row] (residual add layers.py:147-148
 * or positional table :179,:210); tf32 rounding.
 * colsum_part (nullable, [ceil(M/32)][N], splits must be 1): receives the column sums of every 32-row group of
 * the stored C -- summed over the groups (b200vq_colsum) they are the bias gradient of the Linear whose
 * pre-activation gradient this GEMM produced (net.0, layers.py:99), without another pass over C.
 * cta_group: 1 = one CTA per 128 x bn tile, 2 = CTA pair per 256 x bn tile; bn in {0=auto,64,128,192,256}. */
int b200vq_gemm_tf32(const float* A, long long lda, int a_major, const float* B, long long ldb, int b_major,
                     float* C, long long ldc, int M, int N, int K, int splits, long long c_split_stride,
                     const float* bias, const float* res, long long ldres, int res_row_mod,
                     const float* aux, long long ldaux, float* colsum_part, int act, int round_out, int cta_group,
                     int bn, void* stream);
int b200vq_splitk_reduce(const float* part, int splits, long long n, long long split_stride, float* out, void* stream);

/* ---- LayerNorm (nn.LayerNorm(dim), eps 1e-5: layers.py:88,143) ---------------------------------*/
int b200vq_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                         int M, int D, int round_out, void* stream);
size_t b200vq_layernorm_bwd_workspace_bytes(int D);
/* dx = LN'(dy) (+ dres, the skip-connection gradient of layers.py:147-148).  dxsum (nullable, [D]) receives the
 * column sums of dx: dx is also the gradient at the bias of the Linear that wrote this residual stream
 * (to_out layers.py:118, net.2 layers.py:101), so that bias gradient costs no extra pass over dx. */
int b200vq_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                         const float* dres, float* dx, float* dgamma, float* dbeta, float* dxsum, int M, int D,
                         int round_out, void* workspace, size_t ws_bytes, void* stream);

/* ---- attention core (layers.py:124-130): softmax(q k^T * scale) v, no mask -----------------------
 * qkv is the [B*N, 3*heads*dh] output of to_qkv (q | k | v thirds, head h at columns h*dh);
 * out is [B*N, heads*dh]; lse [B*heads*N] holds log-sum-exp of the scaled scores (saved for
 * backward instead of the [B,h,N,N] probabilities the reference keeps).  dh must be 64 or 32. */
int b200vq_attention_fwd(const float* qkv, float* out, float* lse, int B, int N, int heads, int dh, float scale,
                         int round_out, void* stream);
/* dqkv [B*N, 3*heads*dh] from dout; delta [B*heads*N] is scratch (rowsum(dout*out)) */
int b200vq_attention_bwd(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv,
                         float* delta, int B, int N, int heads, int dh, float scale, int round_out, void* stream);

/* ---- vector quantiser (quantizers.py:38-92) ------------------------------------------------------
 * z [M,D], codebook E [K,D] (un-normalised nn.Embedding weight); out [M,D] = straight-through
 * value z + (z_q - z); idx int64 [M, depth]; loss: device scalar.  depth = 1 for the plain
 * quantiser, num_quantizers for use_residual=True. */
size_t b200vq_vq_workspace_bytes(int M, int K, int depth);
int b200vq_vq_fwd(const float* z, const float* E, float* out, long long* idx, float* loss, int M, int K, int D,
                  int depth, float beta, void* workspace, size_t ws_bytes, void* stream);
/* g_out [M,D] or NULL, g_loss device scalar or NULL; gz [M,D] written, gE [K,D] zero-filled then
 * accumulated (nn.Embedding dense backward). */
int b200vq_vq_bwd(const float* z, const float* E, const long long* idx, const float* g_out, const float* g_loss,
                  float* gz, float* gE, int M, int K, int D, int depth, int residual, float beta, void* stream);
/* decode_codes (vitvqgan.py:81-86): out[m] = sum_t normalize(E[codes[m,t]]) */
int b200vq_vq_embed(const float* E, const long long* codes, float* out, int M, int K, int D, int depth, void* stream);

/* ---- re-layout / reductions ------------------------------------------------------------------------*/
/* [B,C,H,W] -> [B*(H/p)*(W/p), C*p*p], patch vector ordered (c,ph,pw) (layers.py:168-171) */
int b200vq_patchify(const float* img, float* patches, int B, int C, int H, int W, int p, int round_out, void* stream);
/* inverse, adding bias[c] if non-NULL (ConvTranspose2d bias, layers.py:204) */
int b200vq_unpatchify(const float* tokens, const float* bias, float* img, int B, int C, int H, int W, int p, void* stream);
size_t b200vq_colsum_workspace_bytes(int N);
/* out[n] = sum_m X[m,n] (bias gradients) */
int b200vq_colsum(const float* X, long long ld, int M, int N, float* out, void* workspace, size_t ws_bytes, void* stream);
/* out = tf32-rounded copy of in (weights shadow for the tensor-core path) */
int b200vq_round_tf32(const float* in, float* out, long long n, void* stream);

/* out[m,:] = x[m,:] + table[m % R,:]  (token + de_pos_embedding, layers.py:210) */
int b200vq_add_rows_mod(const float* x, const float* table, float* out, long long M, int D, int R, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VQ_H_ */
