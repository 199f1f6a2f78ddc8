// Internal interface of the attention kernels (see attention.cu).
#pragma once
#include "common.cuh"

namespace madtp {

// q, k, v are fp32 with head h at column offset h*64 of each token row. Row strides (ld*) and batch strides (bs*)
// are in elements, so fused QKV buffers ([B, N, 3*H*64]) and separate projections are both expressible.
struct AttnArgs {
  const float* q; long long ldq, bsq;
  const float* k; long long ldk, bsk;
  const float* v; long long ldv, bsv;
  int B, H, Nq, Nk;
  float scale;              // logits = q.k * scale + key_mask
  const float* key_mask;    // additive, [B, Nk] or nullptr (BERT padding mask: 0 / -10000)
  int causal;               // 1: key j is visible to query i only if j <= i (CLIP text transformer, clip/model.py:452-457)
  // ---- outputs of the forward pass ----
  __half* out_f16; long long ldo, bso;   // context, heads merged: out[b, i, h*64 + :]  (ldo/bso in elements)
  float* row_max;           // [B, H, Nq]  max_j logits           (nullptr when statistics are not needed)
  float* row_sum;           // [B, H, Nq]  sum_j exp(logit - max)
  float* out_norm;          // [B, H, Nq]  || context[b, h, i, :] ||_2   ("head importance" before normalisation)
  // ---- outputs of the statistics pass (self-attention only, Nq == Nk) ----
  float* col_part;          // [B, ceil(Nq/64), Nk]  partial column sums over query tiles of max_h P[b,h,i,j], i,j >= 1
  float* cls_attn;          // [B, Nk]  sum_h P[b,h,0,j] * norm[b,h,j] / (sum_h' norm[b,h',j] + 1e-8); entry 0 unused
};

// Tensor-core path (attn_tc.cu). q and k are fp16 hi/lo planes of kQkPlaneScale * value, [B*N, ld_qk] (q of head h at
// column h*64, k at H*64 + h*64); v is stored transposed per (sequence, head) as fp16 hi/lo planes of kVPlaneScale *
// value: vt[((b*H + h)*64 + d) * ld_vt + token]. Both are written by launch_gemm_qkv (gemm.cuh).
struct AttnTcArgs {
  const __half* qk_hi; const __half* qk_lo; long long ld_qk;
  const __half* vt_hi; const __half* vt_lo; long long ld_vt;
  int B, H, N;
  float scale;
  const float* key_mask;                 // additive [B, N] or nullptr
  __half* out_f16; long long ldo, bso;   // context, heads merged
  float* out_f32;                        // optional fp32 copy of the context (same ldo / bso, in elements)
  float* row_lse;                        // [B, H, N]  log sum_j exp(logit_j)   (max + log of the row sum)
  float* out_norm;                       // [B, H, N]  || context[b, h, i, :] ||_2
  float* col_part; int n_parts;          // [B, ceil(N/128), N]
  float* cls_attn;                       // [B, N]
  float* cls_p;                          // [B, H, N] CLS query row: 256 * exp(logit - max of its 64-key tile)
  float* cls_tile_max;                   // [B, H, ceil(N/64)] those maxima (running row maximum at each key tile)
  int causal;                            // 1: key j visible to query i only if j <= i (CLIP text tower, clip/model.py:452-457)
  const int* n_dev;                      // dynamic N read on the device (N above is then the capacity): the sequences are
                                         // packed with the dynamic N in every buffer, V^T keeps its pitch ld_vt
};
int launch_attn_fwd_tc(const AttnTcArgs& a, cudaStream_t stream);
int launch_attn_stats_tc(const AttnTcArgs& a, cudaStream_t stream);

// Tensor-core cross-attention (cross_attn_tc.cu), value lane: q [B*Lq, ldq] and k [B*Nk or Nk, ldk] fp16 row-major
// (head h at column h*64), V^T [H*64, ld_vt] fp16 with the keys of sequence b at columns b*vt_cols_per_batch + j.
// k_rows_per_batch / vt_cols_per_batch (>= Nk, the latter a multiple of 8: TMA box origins are 16-byte aligned) are
// the per-sequence pitches, or 0 when every sequence attends to the same keys (broadcast).
struct CrossTcArgs {
  const __half* q; long long ldq;
  const __half* k; long long ldk; int k_rows_per_batch;
  const __half* vt; long long ld_vt; int vt_cols_per_batch;
  const float* v_bias;                   // [H*64] added to the normalised output, or nullptr
  int B, H, Lq, Nk;
  float scale;
  const float* key_mask;                 // additive [B, Nk] or nullptr
  __half* out_f16; long long ldo, bso;
  const int* lq_dev;                     // dynamic Lq (queries and output packed with it)
  const int* nk_dev;                     // dynamic Nk; per-sequence pitches become Nk rounded up to 8 (when not 0)
  // ragged keys (all three or none): sequence b attends to k_len[b] keys starting at row / V^T column k_start[b] (a
  // multiple of 8) of the packed K / V^T, and key 0 carries the extra logit key0_bias[b] -- ln(multiplicity) of the
  // first key, which is how "pad the shorter image with copies of its CLS token" (compress_retrieval_dtp.py:142-154)
  // is evaluated without materialising the copies. Nk is then the capacity (max_b k_len[b]).
  const int* k_start; const int* k_len; const float* key0_bias;
};
int launch_cross_attn_tc(const CrossTcArgs& a, cudaStream_t stream);

int launch_attn_fwd(const AttnArgs& a, cudaStream_t stream);
// Short-sequence self-attention (small_attn.cu): Nq == Nk <= 64, context + (optionally) col_sum[B,L] =
// sum_{i>=1} max_h P and cls_attn[B,L]; scratch holds B*H*L*(L+1) floats when statistics are requested.
int launch_small_self_attn(const AttnArgs& a, float* col_sum, float* cls_attn, float* scratch, const int* n_dev,
                           cudaStream_t stream);
int launch_attn_stats(const AttnArgs& a, cudaStream_t stream);

}  // namespace madtp
