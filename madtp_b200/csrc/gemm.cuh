// Internal interface of the GEMM kernels (not part of the C ABI; see include/madtp_b200.h).
#pragma once
#include "common.cuh"

namespace madtp {

// Epilogue applied to every accumulator element:  out = act(alpha * acc + bias[col]) + residual[row, col]
struct GemmEpilogue {
  void* c;                // output [M, ldc] (fp32 or fp16)
  long long ldc;          // elements
  int c_f16;              // 0: fp32 output, 1: fp16 output
  const float* bias;      // [N] or nullptr
  const float* residual;  // fp32 [M, ldr] or nullptr
  long long ldr;
  int act;                // 0 none, 1 GELU(erf), 2 ReLU, 3 QuickGELU (x * sigmoid(1.702 x))
  float alpha;
  // ---- mode 1 (split-operand kernel only): fused q|k|v projection feeding the tensor-core attention kernels ----
  // columns [0, qk_cols) are written as fp16 hi/lo planes of kQkPlaneScale * value into c / c_lo ([M, ldc] halves
  // each); columns [qk_cols, N) (the value projection, head h = (col - qk_cols) / 64) are written TRANSPOSED per
  // sequence as fp16 hi/lo planes of kVPlaneScale * value, vt[((b * heads + h) * 64 + d) * ld_vt + i] for token i of
  // sequence b = row / n_tok, so that keys are contiguous.
  int chunk_kb;           // split-operand kernels: k-blocks accumulated in TMEM per drained chunk (0 = default)
  int mode;
  __half* c_lo;
  __half* vt_hi;
  __half* vt_lo;
  long long ld_vt;
  int n_tok;
  int heads;
  int qk_cols;
  // ---- dynamic extents (device-resident token counts; nullptr = the host value is exact) ----
  // M = *m_dev * m_mult rows / N = *n_dev * n_mult columns are read at kernel start; the host M / N are then only
  // CAPACITIES (grid size, tensor-map extents). Operand rows beyond the dynamic extent are read but feed only output
  // rows / columns that are never stored. mode 1 additionally takes n_tok from *m_dev.
  const int* m_dev;
  int m_mult;
  const int* n_dev;
  int n_mult;
};

// Effective extents of a launch (device code).
__device__ __forceinline__ int gemm_dyn_m(const GemmEpilogue& ep, int M) {
  return ep.m_dev ? min(M, load_len(ep.m_dev) * ep.m_mult) : M;
}
__device__ __forceinline__ int gemm_dyn_n(const GemmEpilogue& ep, int N) {
  return ep.n_dev ? min(N, load_len(ep.n_dev) * ep.n_mult) : N;
}

// Power-of-two scales of the q/k and v operand planes (keep the lo plane of typical activations a normal fp16 number;
// the attention kernels fold them back into the softmax scale and the output normalisation).
// k-blocks (64 k-elements each) the fp16-plane kernel lets the tensor core accumulate before a chunk is drained into
// fp32 registers. The accumulator truncates on every add, so the error grows with the chunk; the TMEM read-back of a
// 128 x 256 chunk (128 KB at ~64 B/clk) costs more than its MMAs, so the speed grows with it too (measured: DESIGN.md).
constexpr int kF16ChunkKb = 1;
constexpr float kQkPlaneScale = 8.0f;
constexpr float kVPlaneScale = 16.0f;

enum GemmPrecision : int {
  kGemmF16 = 0,      // fp16 operands, fp32 accumulate (kind::f16), one MMA per k-step
  kGemmTF32x3 = 1,   // fp32 operands pre-split into tf32 hi/lo; hi*hi + hi*lo + lo*hi (kind::tf32)
  kGemmSimtF32 = 2,  // CUDA-core fp32 FFMA (device-side checker and tiny shapes)
  kGemmF16x3 = 3,    // fp32 operands pre-split into fp16 hi/lo planes (madtp_split_f16); same three products, kind::f16
};

// GELU(x) = x * Phi(x) with Phi from the complementary error function in the Abramowitz-Stegun 7.1.26 form
// erfc(z) = P(t) exp(-z^2), t = 1 / (1 + p z): |error| <= 1.5e-7 on erf, and because the tail is formed as a product
// (never as 1 - erf) it keeps its relative accuracy for negative x. 16 instructions and two SFU operations per element
// instead of erff's ~28: the fc1 epilogue (one GELU per accumulator element) was issue-bound on erff, not on the MMAs.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  const float q = 0.5f * p * ex2_approx(-1.4426950408889634f * z * z);   // 0.5 erfc(|x| / sqrt 2)
  return x * (x < 0.f ? q : 1.0f - q);
}

// erf GELU through ONE SFU operation: Phi(x) = 0.5 (1 + erf(x / sqrt 2)) ~ 0.5 (1 + tanh(x (a + b x^2 + c x^4))) with a
// minimax fit over |x| <= 8 (beyond that x^2 is clamped: tanh is saturated). |GELU error| <= 2.6e-5 from the fit plus
// tanh.approx's 2^-11 relative error -- the size of the fp16 rounding the value lane applies to this output anyway -- for
// 8 instructions per element instead of gelu_erf's 16: the fc1 epilogue (one GELU per accumulator element) bounded that
// GEMM at 44 % tensor-pipe activity against 65 % for fc2 with the same FLOPs (profiles/r2_kernels.md).
__device__ __forceinline__ float gelu_tanh5(float x) {
  const float x2 = fminf(x * x, 64.0f);
  const float p = fmaf(x2, fmaf(x2, -3.5151678902194784e-4f, 3.700564602133838e-2f), 0.7975078842880567f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case 1:
      return gelu_erf(x);
    case 4:
      return gelu_tanh5(x);
    case 2:
      return fmaxf(x, 0.0f);
    case 3:
      return x / (1.0f + __expf(-1.702f * x));
    default:
      return x;
  }
}

// C[M,N] = epilogue(A[M,K] * B[N,K]^T). A and B are row-major with the reduction dimension contiguous
// (activations [tokens, features] and nn.Linear weights [out, in]).
// a_lo/b_lo are only read for kGemmTF32x3. Returns a Status.
// 2-D TMA descriptor of a row-major [rows, cols] matrix (leading dimension ld elements, fp32 or fp16) with a
// box of box_rows x 128 bytes and the 128-byte swizzle; out-of-bounds elements read as zero.
int make_tmap(CUtensorMap* map, const void* ptr, bool f32, long long rows, long long cols, long long ld, int box_rows);
int make_tmap_colgroups(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                        int box_groups);

// Fused q|k|v projection (F16x3) with the split / transposed epilogue described at GemmEpilogue::mode.
int launch_gemm_qkv(const void* a_hi, const void* a_lo, long long lda, const void* w_hi, const void* w_lo,
                    long long ldb, const float* bias, float alpha, int M, int K, int n_tok, int heads, void* qk_hi,
                    void* qk_lo, long long ld_qk, void* vt_hi, void* vt_lo, long long ld_vt, const int* n_dev,
                    cudaStream_t stream);

int launch_gemm(int precision, const void* a, const void* a_lo, long long lda, const void* b, const void* b_lo,
                long long ldb, const GemmEpilogue& ep, int M, int N, int K, cudaStream_t stream);

}  // namespace madtp
