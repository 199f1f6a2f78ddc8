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
  float* y_hi;      // tf32 hi/lo split of y (operands of a TF32x3 GEMM)
  float* y_lo;
  __half* y_f16;    // fp16 copy of y (operand of an fp16 GEMM)
  float* x_hi;      // tf32 hi/lo split of the raw input row
  float* x_lo;
};

int launch_layernorm(const LayerNormArgs& a, cudaStream_t stream);
int launch_split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t stream);
int launch_cast_f16(const float* x, void* y, long long n, cudaStream_t stream);
int launch_patchify(const float* img, float* hi, float* lo, int B, int C, int H, int W, int P, cudaStream_t stream);
int launch_assemble_tokens(const float* patches, const float* cls, const float* pos, float* x, int B, int n, int d,
                           cudaStream_t stream);
int launch_bert_embed(const long long* ids, const float* word, const float* posemb, float* out, int B, int L, int d,
                      int vocab, cudaStream_t stream);

}  // namespace madtp
