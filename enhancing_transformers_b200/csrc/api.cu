// extern "C" surface of libb200vq.so (see include/b200vq.h).  Thin: argument forwarding,
// error text, launch accounting.  No allocation, no synchronisation.
#include "../../include/b200vq.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// implemented in the other translation units
int gemm_tf32(const float*, long long, int, const float*, long long, int, float*, long long, int, int, int, int, long long,
              const float*, const float*, long long, int, const float*, long long, float*, int, int, int, int, cudaStream_t);
int splitk_reduce(const float*, int, long long, long long, float*, cudaStream_t);
int layernorm_forward(const float*, const float*, const float*, float*, float*, float*, int, int, int, cudaStream_t);
size_t layernorm_bwd_workspace_bytes(int);
int layernorm_backward(const float*, const float*, const float*, const float*, const float*, const float*, float*, float*,
                       float*, float*, int, int, int, void*, size_t, cudaStream_t);
int attention_forward(const float*, float*, float*, int, int, int, int, float, int, cudaStream_t);
int attention_backward(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, float, int,
                       cudaStream_t);
size_t vq_workspace_bytes(int, int, int);
int vq_forward(const float*, const float*, float*, long long*, float*, int, int, int, int, float, void*, size_t, cudaStream_t);
int vq_backward(const float*, const float*, const long long*, const float*, const float*, float*, float*, int, int, int, int,
                int, float, cudaStream_t);
int vq_embed(const float*, const long long*, float*, int, int, int, int, cudaStream_t);
int patchify(const float*, float*, int, int, int, int, int, int, cudaStream_t);
int unpatchify(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
size_t colsum_workspace_bytes(int);
int colsum(const float*, long long, int, int, float*, void*, size_t, cudaStream_t);
int round_tf32_copy(const float*, float*, long long, cudaStream_t);
int add_rows_mod(const float*, const float*, float*, long long, int, int, cudaStream_t);

}  // namespace b200

using namespace b200;
#define S(stream) static_cast<cudaStream_t>(stream)

extern "C" {

int b200vq_version(void) { return B200VQ_VERSION; }
const char* b200vq_last_error(void) { return g_err; }
const char* b200vq_arch(void) { return "sm_100a"; }
long long b200vq_launch_count(void) { return g_launches.load(); }

int b200vq_gemm_tf32(const float* A, long long lda, int a_major, const float* B, long long ldb, int b_major, float* C,
                     long long ldc, int M, int N, int K, int splits, long long c_split_stride, const float* bias,
                     const float* res, long long ldres, int res_row_mod, const float* aux, long long ldaux,
                     float* colsum_part, int act, int round_out, int cta_group, int bn, void* stream) {
  return gemm_tf32(A, lda, a_major, B, ldb, b_major, C, ldc, M, N, K, splits, c_split_stride, bias, res, ldres, res_row_mod,
                   aux, ldaux, colsum_part, act, round_out, cta_group, bn, S(stream));
}
int b200vq_splitk_reduce(const float* part, int splits, long long n, long long split_stride, float* out, void* stream) {
  return splitk_reduce(part, splits, n, split_stride, out, S(stream));
}
int b200vq_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int M,
                         int D, int round_out, void* stream) {
  return layernorm_forward(x, gamma, beta, y, mean, rstd, M, D, round_out, S(stream));
}
size_t b200vq_layernorm_bwd_workspace_bytes(int D) { return layernorm_bwd_workspace_bytes(D); }
int b200vq_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                         const float* dres, float* dx, float* dgamma, float* dbeta, float* dxsum, int M, int D,
                         int round_out, void* workspace, size_t ws_bytes, void* stream) {
  return layernorm_backward(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, dxsum, M, D, round_out, workspace, ws_bytes,
                            S(stream));
}
int b200vq_attention_fwd(const float* qkv, float* out, float* lse, int B, int N, int heads, int dh, float scale,
                         int round_out, void* stream) {
  return attention_forward(qkv, out, lse, B, N, heads, dh, scale, round_out, S(stream));
}
int b200vq_attention_bwd(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv, float* delta,
                         int B, int N, int heads, int dh, float scale, int round_out, void* stream) {
  return attention_backward(qkv, out, lse, dout, dqkv, delta, B, N, heads, dh, scale, round_out, S(stream));
}
size_t b200vq_vq_workspace_bytes(int M, int K, int depth) { return vq_workspace_bytes(M, K, depth); }
int b200vq_vq_fwd(const float* z, const float* E, float* out, long long* idx, float* loss, int M, int K, int D, int depth,
                  float beta, void* workspace, size_t ws_bytes, void* stream) {
  return vq_forward(z, E, out, idx, loss, M, K, D, depth, beta, workspace, ws_bytes, S(stream));
}
int b200vq_vq_bwd(const float* z, const float* E, const long long* idx, const float* g_out, const float* g_loss, float* gz,
                  float* gE, int M, int K, int D, int depth, int residual, float beta, void* stream) {
  return vq_backward(z, E, idx, g_out, g_loss, gz, gE, M, K, D, depth, residual, beta, S(stream));
}
int b200vq_vq_embed(const float* E, const long long* codes, float* out, int M, int K, int D, int depth, void* stream) {
  return vq_embed(E, codes, out, M, K, D, depth, S(stream));
}
int b200vq_patchify(const float* img, float* patches, int B, int C, int H, int W, int p, int round_out, void* stream) {
  return patchify(img, patches, B, C, H, W, p, round_out, S(stream));
}
int b200vq_unpatchify(const float* tokens, const float* bias, float* img, int B, int C, int H, int W, int p, void* stream) {
  return unpatchify(tokens, bias, img, B, C, H, W, p, S(stream));
}
size_t b200vq_colsum_workspace_bytes(int N) { return colsum_workspace_bytes(N); }
int b200vq_colsum(const float* X, long long ld, int M, int N, float* out, void* workspace, size_t ws_bytes, void* stream) {
  return colsum(X, ld, M, N, out, workspace, ws_bytes, S(stream));
}
int b200vq_round_tf32(const float* in, float* out, long long n, void* stream) { return round_tf32_copy(in, out, n, S(stream)); }

int b200vq_add_rows_mod(const float* x, const float* table, float* out, long long M, int D, int R, void* stream) {
  return add_rows_mod(x, table, out, M, D, R, S(stream));
}

}  // extern "C"
