// Internal interface of the Dynamic Token Pruning kernels (see dtp.cu).
#pragma once
#include "common.cuh"

namespace madtp {

constexpr int kDtpMaxTokens = 1024;  // n = N-1 prunable tokens per sequence handled by one CTA

// Column (over-token) softmax statistics of token_att / divisor, per codebook entry t:
//   col_max[b,t] = max_j x[b,j,t],  col_sum[b,t] = sum_j exp(x[b,j,t] - col_max[b,t]),  x = token_att / divisor
// Dynamic lengths (every launcher below): `n_dev` (may be nullptr) points to the device-resident token count N of the
// packed sequences INCLUDING the n_sub leading tokens that are not prunable (CLS / [ENC]); the kernel then uses
// n = N - n_sub, recomputes every per-sequence stride as N * (row pitch), and the host-side n / strides are only
// capacities.
int launch_token_colstats(const float* token_att, long long ld_ta, long long bs_ta, int B, int n, int T, float divisor,
                          float* col_max, float* col_sum, const int* n_dev, int n_sub, cudaStream_t stream);

// Query_model's aggregated feature (reference models/utils.py:174-178):
//   sd_ft[b,t,:] (+)= sum_j softmax_j(token_att[b,j,t] / divisor) * x[b,j,:]
int launch_query_sdft(const float* token_att, long long ld_ta, long long bs_ta, const float* col_max,
                      const float* col_sum, const float* x, long long ldx, long long bsx, int B, int n, int T, int d,
                      float divisor, float* sd_ft, int accumulate, const int* n_dev, int n_sub, cudaStream_t stream);

// Tensor-core variant (sdft_tc.cu): x is the dense fp32 matrix [x_rows, d]; token j of batch b sits at row
// b*row_stride + first_row + j.
int launch_query_sdft_tc(const float* token_att, long long ld_ta, long long bs_ta, const float* col_max,
                         const float* col_sum, const float* x, long long x_rows, int row_stride, int first_row, int B,
                         int n, int T, int d, float divisor, float* sd_ft, int accumulate, const int* n_dev,
                         cudaStream_t stream);

// MN-major tensor-core variant (sdft_tc.cu): x as fp16 hi/lo planes [x_rows, d] of x / x_unscale, no transposition.
int launch_query_sdft_planes(const float* token_att, long long ld_ta, long long bs_ta, const float* col_max,
                             const float* col_sum, const __half* x_hi, const __half* x_lo, float x_unscale,
                             long long x_rows, int row_stride, int first_row, int B, int n, int T, int d, float divisor,
                             float* sd_ft, int accumulate, const int* n_dev, cudaStream_t stream);

struct DtpScoreArgs {
  int B, n, T;                // n prunable tokens (sequence position 1..n), T codebook entries
  const float* col_part;      // [B, n_parts, n+1] partial column sums from attn_stats (index 0 = CLS, unused)
  int n_parts;
  const float* cls_attn;      // [B, n+1] (index 0 unused)
  const float* token_att;     // [B, *, ld_ta]; row j (0-based over prunable tokens) at token_att + b*bs_ta + j*ld_ta
  long long ld_ta, bs_ta;
  float temperature;
  float* score;               // [B, n]  Importance_score
  float* threshold;           // [B]
  int* count;                 // [B]     #(score > threshold)
  int* topk;                  // [1]     max_b count, must be zeroed before the launch (atomicMax)
  const int* n_dev;           // dynamic N = n + 1 (see above); sequences are then packed: strides N * ld_ta, N
  int parts_tile;             // with n_dev: n_parts = ceil(N / parts_tile) (0: n_parts as given)
};
int launch_dtp_score(const DtpScoreArgs& a, cudaStream_t stream);

struct DtpSelectArgs {
  int B, n;
  const float* score;         // [B, n]
  const int* topk;            // device scalar k (batch max count)
  unsigned char* keep;        // [B, n]  1 = survivor
  int* dst;                   // [B, n]  slot of survivor j among the survivors (ascending token index), -1 if pruned
  float* tail_w;              // [B, n]  merge weight of pruned token j (0 for survivors)
  int* tail_idx;              // [B, n]  pruned token indices, ascending; the first n-k entries are valid
  // optional additive key-mask bookkeeping for text (mask_mode 0: none)
  int mask_mode;              // 1: nlvr_encoder semantics (mask of the r-th ranked token at slot r, r <= k)
                              // 2: med.py semantics (mask travels with its token; merged slot = mask of rank k)
  const float* mask_in;       // [B, n+1] additive mask incl. position 0
  float* mask_out;            // [B, n+1] worst case; entries [0, k+2) are written
  int max_keep;               // nothing is pruned when k <= max_keep (0 everywhere but CLIP, clip/model.py:220)
  const int* n_dev;           // dynamic N = n + 1; mask_in is packed [B, N] and mask_out is written packed [B, N_out]
  int* n_out;                 // with n_dev: receives N_out = k + 2 (or N when nothing is pruned) -- the next layer's N
  int* k_out;                 // optional trajectory record: k, or -1 when nothing is pruned
};
int launch_dtp_select(const DtpSelectArgs& a, cudaStream_t stream);

struct DtpGatherArgs {
  int B, n, d;
  const float* x;             // [B, n+1, d] tokens incl. position 0 (CLS / [ENC]) which always survives
  long long bsx;              // batch stride of x (elements)
  const int* topk;
  const int* dst;             // from dtp_select
  const float* tail_w;
  const int* tail_idx;
  float* out;                 // [B, k+2, d] packed survivors: [cls, survivors ascending, merged]
  long long bso;              // batch stride of out (elements) -- the caller sizes it with the k it read back
  __half* out_f16;            // optional fp16 copy of out with the same batch stride (operand of the next GEMM)
  int max_keep;
  const int* n_dev;           // dynamic N = n + 1: x is packed [B, N, d], out is written packed [B, N_out, d]
};
int launch_dtp_gather(const DtpGatherArgs& a, cudaStream_t stream);

// Fused selection + compaction + merged token + LayerNorm (dtp_apply.cu): what dtp_select + dtp_gather + the following
// layernorm launch do, in one kernel with a radix select.
struct DtpApplyArgs {
  int B, n, d;
  const float* score;         // [B, n]
  const int* topk;            // device scalar k (batch max count)
  const float* x;             // [B, n+1, d] tokens incl. position 0 (always survives)
  long long bsx;              // batch stride of x (elements, a multiple of d)
  float* out;                 // [B, k+2, d]: [position 0, survivors ascending, merged]
  long long bso;
  __half* out_f16;            // optional fp16 copy of out
  const float* ln_gamma;      // optional LayerNorm applied to every output row, written as fp16 to ln_out
  const float* ln_beta;
  float ln_eps;
  __half* ln_out;             // [B, k+2, d] fp16, same batch stride as out
  unsigned char* keep;        // optional [B, n]: 1 = survivor
  int mask_mode;              // as DtpSelectArgs
  const float* mask_in;
  float* mask_out;
  int max_keep;
  const int* n_dev;           // device-resident N = n + 1 (packed input / output), as DtpSelectArgs / DtpGatherArgs
  int* n_out;
  int* k_out;
};
int launch_dtp_apply(const DtpApplyArgs& a, cudaStream_t stream);

// vector_gather (reference models/utils.py:13-33): out[b,i,:] = x[b, idx[b,i], :]; one warp per output row.
int launch_gather_rows(const float* x, long long bsx, const int* idx, float* out, int B, int L, int K, int d,
                       cudaStream_t stream);

}  // namespace madtp
