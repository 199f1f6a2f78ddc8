// Internal interface of the row-wise kernels (see rowops.cu).
#pragma once
#include "common.cuh"

namespace madtp {

struct LayerNormArgs {
  const float* x;   // [rows, ldx] fp32 input
  long long ldx;
  int rows;
  int d;            // row width (multiple of 128, <= 1024)
  const float* gamma;  // nullptr => no normalisation, only the x_hi/x_lo split is produced
  const float* beta;
  float eps;
  float* y_f32;     // optional outputs, each [rows, d] contiguous
  __half* y_hi;     // fp16 hi/lo split of y (operands of an F16x3 GEMM): hi = fp16(y), lo = fp16(y - hi)
  __half* y_lo;
  __half* y_f16;    // fp16 copy of y (operand of an fp16 GEMM)
  __half* x_hi;     // fp16 hi/lo split of the raw input row
  __half* x_lo;
  const int* n_dev; // dynamic row count: rows = min(rows, *n_dev * n_mult) read on the device (nullptr: rows is exact)
  int n_mult;
};

// LayerNorm of the packed token stream x [B * N, d] written as the fp16 key/value operand of the cross-attention
// kernels: sequence b lands at y16 + (b / per_group) * group_stride + ((b % per_group) * P + t) * d with P = N rounded
// up to 8 rows (TMA box origins must be 16-byte aligned); rows N..P-1 are written as zeros. *p_out receives P.
struct LayerNormPackArgs {
  const float* x;
  int B, N, d;               // N: tokens per sequence (capacity when n_dev is given)
  const int* n_dev;
  const float* gamma; const float* beta; float eps;
  float* y_f32;              // optional [B * N, d] packed fp32 output (the module's return value)
  __half* y16;
  int per_group; long long group_stride;
  int* p_out;                // optional device scalar
};
int launch_layernorm_pack(const LayerNormPackArgs& a, cudaStream_t stream);
// out[b, :] = x[(b * N + token) * d ...]: one token of every sequence of a packed stream (N from *n_dev when given).
int launch_take_token(const float* x, int B, int N, const int* n_dev, int token, int d, float* out, cudaStream_t stream);

int launch_layernorm(const LayerNormArgs& a, cudaStream_t stream);
int launch_split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t stream);
int launch_split_f16(const float* x, void* hi, void* lo, long long n, float scale, cudaStream_t stream);
int launch_cast_f16(const float* x, void* y, long long n, cudaStream_t stream);
int launch_patchify(const float* img, void* hi, void* lo, int B, int C, int H, int W, int P, cudaStream_t stream);
int launch_assemble_tokens(const float* patches, const float* cls, const float* pos, float* x, int B, int n, int d,
                           cudaStream_t stream);
int launch_lm_nll(const float* logits, long long ld, int R, int V, const long long* labels, float eps, float* loss,
                  float* lse, cudaStream_t stream);
int launch_bert_embed(const long long* ids, const float* word, const float* posemb, float* out, int B, int L, int d,
                      int vocab, int n_pos, cudaStream_t stream);

}  // namespace madtp
