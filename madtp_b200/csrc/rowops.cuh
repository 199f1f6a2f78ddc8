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
};

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
                      int vocab, cudaStream_t stream);

}  // namespace madtp
