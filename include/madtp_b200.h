/*
 * madtp_b200 -- C ABI of the B200 (sm_100a) kernels behind MADTP's pruned vision-language forward path.
 *
 * The reference (double125/MADTP) is pure PyTorch and has no native interface; every entry point below replaces a
 * group of library calls on its hot path. The citation on each function is the reference code it stands in for
 * (paths under the reference tree). Python binds these with ctypes (madtp_b200/_lib.py); see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error (1 invalid argument, 2 CUDA error, 3 workspace,
 *     4 unsupported); madtp_last_error_string() describes the last failure on the calling thread.
 *   - all pointers are DEVICE pointers unless stated otherwise; the library never allocates, frees or retains them.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls only enqueue work.
 *   - leading dimensions (ld*) and batch strides (bs*) are in ELEMENTS.
 *   - fp16 buffers are IEEE binary16 (`__half`), passed as void*.
 */
#ifndef MADTP_B200_H_
#define MADTP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MADTP_B200_ABI_VERSION 4   /* 3: device-resident token counts (`*_dev` arguments), cross-attention over any Nk; 4: madtp_query_sdft_planes, out_f32 of madtp_attn_tc_fwd */

/*
 * Device-resident lengths (ABI 3). The number of tokens a layer keeps (topk_num, vit.py:145) decides every later
 * shape; the reference reads it back with `.item()` once per pruned layer. Every entry point whose work depends on a
 * per-sequence token count therefore also accepts that count from DEVICE memory: an `int32_t*` argument named `*_dev`
 * (NULL = the host value is exact, the ABI-2 behaviour). With a non-NULL `*_dev`
 *   - the kernel reads the count when it starts; the host-side count is only a CAPACITY (>= the real one) used for
 *     the grid size, tensor-map extents and argument checks;
 *   - sequences are PACKED with the dynamic count: sequence b of a [B, N, ...] buffer starts at b * N_dynamic rows, and
 *     every batch stride the signature carries is recomputed as N_dynamic * (row pitch) on the device;
 *   - madtp_dtp_select writes the NEXT layer's count (and the per-layer trajectory) to device memory,
 * so a whole pruned forward can be enqueued (or captured in a CUDA graph) without a single host read-back.
 */

/* GEMM operand precision */
#define MADTP_GEMM_F16 0     /* fp16 operands, fp32 accumulate, tcgen05 kind::f16 */
#define MADTP_GEMM_TF32X3 1  /* fp32 operands pre-split into tf32 hi/lo, 3 tcgen05 kind::tf32 MMAs per k-step */
#define MADTP_GEMM_SIMT 2    /* fp32 FFMA on CUDA cores (device-side checker, tiny shapes) */
#define MADTP_GEMM_F16X3 3   /* fp32 operands pre-split into fp16 hi/lo planes, 3 tcgen05 kind::f16 MMAs per k-step */

/* GEMM epilogue activation */
#define MADTP_ACT_NONE 0
#define MADTP_ACT_GELU 1       /* erf GELU: nn.GELU (vit.py:18) / ACT2FN["gelu"] (med.py:308) */
#define MADTP_ACT_RELU 2       /* cls_head ReLU (blip_nlvr.py:58) */
#define MADTP_ACT_QUICKGELU 3  /* clip/model.py QuickGELU */
#define MADTP_ACT_GELU_FAST 4  /* the same erf GELU through one tanh.approx: |error| <= 2.6e-5 + 2^-11 relative (the fp16
                                  rounding of the value lane); half the epilogue instructions of MADTP_ACT_GELU */

int madtp_abi_version(void);
const char* madtp_last_error_string(void);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
long long madtp_launch_count(void);

/*
 * C[M,N] = act(alpha * A[M,K] . B[N,K]^T + bias[N]) + residual[M,N]
 * Replaces every nn.Linear on the path: vit.py:48-49,77,92 (qkv, proj), vit.py:24-26 (fc1, fc2),
 * nlvr_encoder.py:103-109 (query/key/value), :247-263 (dense0/dense1/merge_layer), :371,385 (intermediate, output),
 * models/utils.py:170 (token . codebook^T), blip_nlvr.py:56-60 (cls_head), timm PatchEmbed conv (vit.py:241).
 * a_lo / b_lo: residual planes, only for MADTP_GEMM_TF32X3 (fp32 planes, madtp_split_tf32) and MADTP_GEMM_F16X3 (fp16
 * planes, madtp_split_f16 / madtp_layernorm; fold the weight's power-of-two scale into alpha).
 * c_f16 != 0 writes fp16 output. bias, residual may be NULL. MADTP_GEMM_F16 / _F16X3 read fp16 a, b; the others fp32.
 */
int madtp_gemm(int precision, const void* a, const void* a_lo, int64_t lda, const void* b, const void* b_lo,
               int64_t ldb, void* c, int64_t ldc, int c_f16, const float* bias, const float* residual, int64_t ldr,
               int act, float alpha, int M, int N, int K, const int32_t* m_dev, int m_mult, const int32_t* n_dev,
               int n_mult, void* stream);
/* m_dev / n_dev (each may be NULL): M = *m_dev * m_mult rows, N = *n_dev * n_mult columns (sequences x tokens). */

/*
 * Row LayerNorm with fused GEMM-operand preparation. Replaces nn.LayerNorm at vit.py:111,115,239 (eps 1e-6) and
 * nlvr_encoder.py:54,243,381 / med.py:54,242,324 (eps 1e-12).
 * Outputs are optional (NULL to skip), each [rows, d] contiguous: y_f32; y_hi/y_lo (fp16 hi/lo split of y: hi = fp16(y),
 * lo = fp16(y - hi), the operand planes of MADTP_GEMM_F16X3); y_f16; x_hi/x_lo (the same split of the un-normalised
 * input row, operand of the Query_model product).
 * gamma == beta == NULL: only x_hi/x_lo are produced. d must be a multiple of 128, <= 1024.
 */
int madtp_layernorm(const float* x, int64_t ldx, int rows, int d, const float* gamma, const float* beta, float eps,
                    float* y_f32, void* y_hi, void* y_lo, void* y_f16, void* x_hi, void* x_lo, const int32_t* n_dev,
                    int n_mult, void* stream);
/* n_dev: rows = *n_dev * n_mult. */

/*
 * Final LayerNorm of the image encoder written straight into the operand layout of the cross-attention K / V^T
 * projections (vit.py:309 followed by nlvr_encoder.py:103-109 / med.py:103-109 with is_cross_attention): x is the packed
 * stream [B * N, d]; sequence b is written as fp16 at y16 + (b / per_group) * group_stride + ((b % per_group) * P + t) * d
 * with P = N rounded up to 8 rows (TMA box origins are 16-byte aligned), rows N..P-1 zero. per_group / group_stride
 * (elements) split the batch into equal groups with their own base (BLIP-NLVR: image0 / image1 halves, blip_nlvr.py:67).
 * y_f32 (optional): the packed fp32 result [B * N, d]. p_out_dev (optional) receives P. n_dev: dynamic N.
 */
int madtp_layernorm_pack(const float* x, int B, int N, int d, const float* gamma, const float* beta, float eps,
                         float* y_f32, void* y16, int per_group, int64_t group_stride, int32_t* p_out_dev,
                         const int32_t* n_dev, void* stream);
/* out[b, :] = x[(b * N + token) * d ...]: one token of every sequence of a packed stream (`last_hidden_state[:, 0, :]`,
 * blip_nlvr.py:80). n_dev: dynamic N. */
int madtp_take_token(const float* x, int B, int N, int token, int d, float* out, const int32_t* n_dev, void* stream);

/* hi = round_to_tf32(x), lo = x - hi (exact). Used once per weight at load time. */
int madtp_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
/* fp16 split for MADTP_GEMM_F16X3: hi = fp16(scale*x), lo = fp16(scale*x - hi); scale is a power of two (weights are
 * scaled up at load time so that lo stays a normal fp16 number; the GEMM's alpha takes the scale out again). */
int madtp_split_f16(const float* x, void* hi_f16, void* lo_f16, int64_t n, float scale, void* stream);
/* y = (fp16) x */
int madtp_cast_f16(const float* x, void* y_f16, int64_t n, void* stream);

/*
 * ViT stem (timm PatchEmbed = Conv2d(C, D, P, stride P), vit.py:241-242,284): non-overlapping patches of
 * img[B,C,H,W] as GEMM rows [B*(H/P)*(W/P), C*P*P] in Conv2d weight order, written as fp16 hi/lo planes.
 */
int madtp_patchify(const float* img, void* rows_hi, void* rows_lo, int B, int C, int H, int W, int P, void* stream);
/* x[b,0,:] = cls + pos[0]; x[b,1+p,:] = patches[b,p,:] + pos[1+p]   (vit.py:286-289) */
int madtp_assemble_tokens(const float* patches, const float* cls, const float* pos, float* x, int B, int n, int d,
                          void* stream);
/* out[b,l,:] = word[ids[b,l]] + position[l]   (nlvr_encoder.py:61-85; the LayerNorm is a madtp_layernorm call) */
/* n_pos: rows of the position table; L > n_pos is an argument error (ids outside [0, vocab) are clamped). */
int madtp_bert_embed(const int64_t* ids, const float* word, const float* position, float* out, int B, int L, int d,
                     int vocab, int n_pos, void* stream);

/*
 * Language-model head statistics per logits row (VQA answer ranking: models/med.py:1040-1047 CrossEntropyLoss with
 * label_smoothing 0.1 and reduction 'none'; models/blip_vqa.py:168-171 softmax of the first-token logits):
 *   lse[r] = log sum_v exp(logits[r, v]);   loss[r] = (1 - eps)(lse - logits[r, label]) + eps (lse - mean_v logits[r, v]),
 * 0 where label < 0 (ignore_index). labels/loss come in pairs and may be NULL; lse may be NULL.
 */
int madtp_lm_nll(const float* logits, int64_t ld, int R, int V, const int64_t* labels, float label_smoothing,
                 float* loss, float* lse, void* stream);

/*
 * Multi-head attention, head dim 64: context = softmax(q.k^T * scale + key_mask) v, heads merged into
 * out_f16[b, i, h*64 + :]. Replaces vit.py:79-91, nlvr_encoder.py:174-219 / med.py:175-217 (self and cross
 * attention) without materialising the [B,H,N,N] probabilities.
 * key_mask: additive [B, Nk] (0 / -10000, nlvr_encoder.py:870-871) or NULL. causal != 0 additionally hides keys
 * j > i from query i (the CLIP text transformer's mask, clip/model.py:452-457 cropped by clip/mock.py:309-310).
 * row_max/row_sum/out_norm ([B,H,Nq] each, all three or none): softmax row statistics and ||context[b,h,i,:]||_2,
 * the un-normalised "head importance" of vit.py:97 / nlvr_encoder.py:231.
 */
int madtp_attn_fwd(const float* q, int64_t ldq, int64_t bsq, const float* k, int64_t ldk, int64_t bsk, const float* v,
                   int64_t ldv, int64_t bsv, int B, int H, int Nq, int Nk, float scale, const float* key_mask,
                   void* out_f16, int64_t ldo, int64_t bso, float* row_max, float* row_sum, float* out_norm,
                   int causal, void* stream);

/*
 * Self-attention of a SHORT sequence (L <= 64, the text encoders) with every pruning statistic in one launch:
 * context as madtp_attn_fwd, col_sum[b, j] = sum_{i >= 1} max_h P[b,h,i,j] and cls_attn[b, j] as madtp_attn_stats
 * (col_sum / cls_attn both NULL: context only; otherwise `scratch` provides B*H*L*(L+1) floats).
 * nlvr_encoder.py:174-235,404-406 / med.py:175-235,348-350.
 */
int madtp_attn_small_self(const float* q, int64_t ldq, int64_t bsq, const float* k, int64_t ldk, int64_t bsk,
                          const float* v, int64_t ldv, int64_t bsv, int B, int H, int L, float scale,
                          const float* key_mask, void* out_f16, int64_t ldo, int64_t bso, float* col_sum,
                          float* cls_attn, float* scratch, int causal, const int32_t* l_dev, void* stream);
/* l_dev: dynamic L (packed q / k / v / out / key_mask / col_sum / cls_attn). */

/*
 * Pruning statistics of a self-attention (Nq == Nk == N), from the row statistics of madtp_attn_fwd:
 *   col_part[b, it, j] = sum_{i in query tile it, i >= 1} max_h P[b,h,i,j]      (vit.py:126-127; sum `it` in order)
 *   cls_attn[b, j]     = sum_h P[b,h,0,j] * norm[b,h,j] / (sum_h' norm[b,h',j] + 1e-8)   (vit.py:96-100)
 * col_part is [B, ceil(N/64), N], cls_attn is [B, N]; index j = 0 (CLS) is not meaningful.
 */
int madtp_attn_stats(const float* q, int64_t ldq, int64_t bsq, const float* k, int64_t ldk, int64_t bsk, int B, int H,
                     int N, float scale, const float* key_mask, const float* row_max, const float* row_sum,
                     const float* out_norm, float* col_part, float* cls_attn, int causal, void* stream);

/*
 * Query_model (models/utils.py:147-183), given token_att = ft . sd^T from madtp_gemm:
 *   colstats: col_max[b,t], col_sum[b,t] of softmax over tokens of token_att / divisor   (divisor = sqrt(sd_dim))
 *   sdft:     sd_ft[b,t,:] (+)= sum_j softmax_j(token_att[b,j,t] / divisor) * ft[b,j,:]
 * token_att row j of batch b is at token_att + b*bs_ta + j*ld_ta; ft row j at ft + b*bs_ft + j*ld_ft.
 * n_dev / n_sub: *n_dev is the packed token count N INCLUDING the n_sub leading non-prunable tokens (n = N - n_sub).
 */
int madtp_token_colstats(const float* token_att, int64_t ld_ta, int64_t bs_ta, int B, int n, int T, float divisor,
                         float* col_max, float* col_sum, const int32_t* n_dev, int n_sub, void* stream);
int madtp_query_sdft(const float* token_att, int64_t ld_ta, int64_t bs_ta, const float* col_max, const float* col_sum,
                     const float* ft, int64_t ld_ft, int64_t bs_ft, int B, int n, int T, int d, float divisor,
                     float* sd_ft, int accumulate, const int32_t* n_dev, int n_sub, void* stream);

/* madtp_query_sdft on the tensor cores: ft is the dense fp32 matrix x [x_rows, d] (token j of batch b at row
 * b*row_stride + first_row + j); both operands are re-laid out K-major in shared memory, nothing transposed touches HBM. */
int madtp_query_sdft_tc(const float* token_att, int64_t ld_ta, int64_t bs_ta, const float* col_max, const float* col_sum,
                        const float* x, int64_t x_rows, int row_stride, int first_row, int B, int n, int T, int d,
                        float divisor, float* sd_ft, int accumulate, const int32_t* n_dev, void* stream);
/* n_dev: *n_dev = tokens per sequence (row_stride = *n_dev, n = *n_dev - first_row). */

/* The same aggregation from the fp16 hi/lo planes of ft (x = x_unscale * (x_hi + x_lo), dense [x_rows, d]; the planes
 * the LayerNorm entry point writes for the codebook product): the tensor core reads both operands MN-major, i.e. with
 * the token index as the row index they have in HBM, so nothing is transposed anywhere (models/utils.py:174-178). */
int madtp_query_sdft_planes(const float* token_att, int64_t ld_ta, int64_t bs_ta, const float* col_max,
                            const float* col_sum, const void* x_hi, const void* x_lo, float x_unscale, int64_t x_rows,
                            int row_stride, int first_row, int B, int n, int T, int d, float divisor, float* sd_ft,
                            int accumulate, const int32_t* n_dev, void* stream);

/*
 * DTP scoring (vit.py:123-145 / nlvr_encoder.py:400-432 / med.py:345-369): Importance_score, threshold,
 * per-row count and topk = max_b count (atomicMax into *topk, which the caller zeroes first).
 * n = number of prunable tokens (positions 1..n); col_part/cls_attn as produced by madtp_attn_stats with N = n+1.
 */
int madtp_dtp_score(int B, int n, int T, const float* col_part, int n_parts, const float* cls_attn,
                    const float* token_att, int64_t ld_ta, int64_t bs_ta, float temperature, float* score,
                    float* threshold, int32_t* count, int32_t* topk, const int32_t* n_dev, int parts_tile, void* stream);
/* n_dev: *n_dev = n + 1 (packed col_part / cls_attn / token_att / score); parts_tile > 0: n_parts = ceil(N / parts_tile)
 * (128 for madtp_attn_tc_stats, 64 for madtp_attn_stats), 0: n_parts as given. */

/*
 * DTP selection (vit.py:153-158): exact top-k of score per row (k read from *topk on the device), survivors keep
 * ascending token order. keep[B,n] (1 = survivor), dst[B,n] (slot among survivors or -1), tail_w[B,n] (merge weight
 * score_j / (sum_tail + 1e-8) of pruned tokens), tail_idx[B,n] (pruned token indices, first n-k valid).
 * If k <= max_keep or n - k <= 1 nothing is pruned and dst is the identity: max_keep = 0 is vit.py:148-149 /
 * nlvr_encoder.py:434 / med.py:371; CLIP passes its EOT guard (clip/model.py:220,492).
 * mask_mode 1 (nlvr_encoder.py:451-452) / 2 (med.py:377-390) additionally gathers the additive key mask
 * mask_in[B,n+1] -> mask_out[B, 0..k+1]; mask_mode 0 ignores both.
 */
int madtp_dtp_select(int B, int n, const float* score, const int32_t* topk, uint8_t* keep, int32_t* dst, float* tail_w,
                     int32_t* tail_idx, int mask_mode, const float* mask_in, float* mask_out, int max_keep,
                     const int32_t* n_dev, int32_t* n_out_dev, int32_t* k_out_dev, void* stream);
/* n_dev: *n_dev = n + 1; mask_in is then packed [B, N] and mask_out is written packed [B, N_out]. n_out_dev receives
 * N_out = k + 2 (N when nothing is pruned) -- the next layer's `*_dev`; k_out_dev (optional) receives k or -1. */

/*
 * DTP gather + merge (vit.py:154-161,202; models/utils.py:13-33 vector_gather): out[b] = [x[b,0], survivors in
 * ascending token order, sum_j tail_w[j] x[b,1+j]] -- shape [B, k+2, d] with batch stride bso. out_f16 (may be NULL):
 * an fp16 copy with the same layout, the operand of the GEMM that follows in the text encoders.
 */
int madtp_dtp_gather(int B, int n, int d, const float* x, int64_t bsx, const int32_t* topk, const int32_t* dst,
                     const float* tail_w, const int32_t* tail_idx, float* out, int64_t bso, void* out_f16, int max_keep,
                     const int32_t* n_dev, void* stream);
/* n_dev: *n_dev = n + 1; x is packed [B, N, d] and out is written packed [B, N_out, d] (a full copy when nothing is
 * pruned, so that the next layer always reads `out`). */

/*
 * madtp_dtp_select + madtp_dtp_gather + the LayerNorm that follows the pruning (vit.py:205 norm2) as ONE kernel:
 * radix select of the k-th largest score (ties -> lower token index), stream compaction in ascending token order, merged
 * token, and -- when ln_gamma / ln_beta / ln_out_f16 are given -- LayerNorm(out) written as fp16 (the operand of the
 * GEMM that follows), computed while the row is in registers. keep (optional) [B, n]. out_f16 (optional): fp16 copy of
 * out. Everything else as madtp_dtp_select / madtp_dtp_gather; results are bit-identical to that pair + madtp_layernorm.
 * d must be a multiple of 128, <= 1024; n <= 1024 (mask modes: n <= 256).
 */
int madtp_dtp_apply(int B, int n, int d, const float* score, const int32_t* topk, const float* x, int64_t bsx, float* out,
                    int64_t bso, void* out_f16, const float* ln_gamma, const float* ln_beta, float ln_eps,
                    void* ln_out_f16, uint8_t* keep, int mask_mode, const float* mask_in, float* mask_out, int max_keep,
                    const int32_t* n_dev, int32_t* n_out_dev, int32_t* k_out_dev, void* stream);

/*
 * Tensor-core self-attention for the scoring lane (vit.py:75-103 without materialising P). madtp_gemm_qkv is the
 * fused q|k|v projection (vit.py:77; a_hi/a_lo and w_hi/w_lo are fp16 hi/lo planes as for MADTP_GEMM_F16X3, alpha
 * removes the weight's scale) with an epilogue that writes q and k as fp16 hi/lo planes of MADTP_QK_PLANE_SCALE * value,
 * [M, ld_qk] (q of head h at column h*64, k at heads*64 + h*64), and v transposed per (sequence, head) as fp16 hi/lo
 * planes of MADTP_V_PLANE_SCALE * value, vt[((b*heads + h)*64 + d) * ld_vt + token] (keys contiguous; ld_vt a multiple
 * of 8). M = B * n_tok rows, K = model width. hi = fp16(s*x), lo = fp16(s*x - hi): hi + lo carries 22 mantissa bits.
 * madtp_attn_tc_fwd: context (fp16, heads merged), row_lse[b,h,i] = log sum_j exp(logit) and out_norm[b,h,i];
 * optionally (both or neither) the CLS query row of every head as cls_p[b,h,j] = 256 exp(logit_j - cls_tile_max[b,h,
 * j/64]) with cls_tile_max [B,H,ceil(N/64)] the running row maximum at each 64-key tile.
 * madtp_attn_tc_stats: col_part[b, it, j] = sum_{i in 128-query tile it, i >= 1} max_h P[b,h,i,j]  (n_parts =
 * ceil(N/128)) and cls_attn[b, j] as in madtp_attn_stats, from cls_p / cls_tile_max of the forward pass. Head dim 64.
 */
#define MADTP_QK_PLANE_SCALE 8.0f
#define MADTP_V_PLANE_SCALE 16.0f
int madtp_gemm_qkv(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo, int64_t ldb,
                   const float* bias, float alpha, int M, int K, int n_tok, int heads, void* qk_hi, void* qk_lo,
                   int64_t ld_qk, void* vt_hi, void* vt_lo, int64_t ld_vt, const int32_t* n_dev, void* stream);
int madtp_attn_tc_fwd(const void* qk_hi, const void* qk_lo, int64_t ld_qk, const void* vt_hi, const void* vt_lo,
                      int64_t ld_vt, int B, int H, int N, float scale, const float* key_mask, void* out_f16,
                      int64_t ldo, int64_t bso, float* row_lse, float* out_norm, float* cls_p, float* cls_tile_max,
                      int causal, const int32_t* n_dev, float* out_f32, void* stream);
/* out_f32 (may be NULL): fp32 copy of the context with the same ldo / bso (in elements) -- the operand of the opt-in
 * value lane at scoring precision (madtp_b200.functional.value_lane_split). */
int madtp_attn_tc_stats(const void* qk_hi, const void* qk_lo, int64_t ld_qk, int B, int H, int N, float scale,
                        const float* key_mask, const float* row_lse, const float* out_norm, float* col_part,
                        int n_parts, float* cls_attn, const float* cls_p, const float* cls_tile_max, int causal,
                        const int32_t* n_dev, void* stream);
/* causal != 0 (both passes): key j is visible to query i only if j <= i -- the CLIP text tower's mask (clip/model.py:
 * 452-457, cropped to the current length by clip/mock.py:309-310). */
/* n_dev (all three): dynamic n_tok / N; every [.., N] statistics buffer and the q/k planes are packed with it, V^T keeps
 * its pitch ld_vt (>= the capacity). */

/*
 * Asynchronous read-back of a few bytes (the per-layer topk_num, vit.py:145 `.item()`): begin records an event on
 * `stream` (after the score kernel), a library-owned side stream waits for it and copies src_dev -> dst_pinned (pinned
 * host memory); wait blocks the host until that copy is done. Work queued on `stream` after begin (the attention output
 * projection, independent of the pruning decision) runs while the host waits. slot in [0, 8): one outstanding read-back
 * per slot. The side stream and its events are the only resources the library keeps (one set per device).
 */
int madtp_readback_begin(const void* src_dev, void* dst_pinned, int64_t bytes, int slot, void* stream);
int madtp_readback_wait(int slot);

/*
 * Tensor-core cross-attention of text queries over image tokens (nlvr_encoder.py:174-219, med.py:175-217 with
 * is_cross_attention; value lane, fp16 operands). Lq <= 128, any Nk (walked in blocks of 128 keys with an online
 * softmax), head dim 64.
 * q [B*Lq, ldq] and k [B*Nk, ldk] fp16 row-major (head h at column h*64); V^T [H*64, ld_vt] fp16 with the keys of
 * sequence b at columns b*vt_cols_per_batch + j (the value projection run as W_v . X^T); v_bias [H*64] is added to the
 * normalised output (rows of P sum to one). k_rows_per_batch / vt_cols_per_batch: per-sequence pitch (>= Nk; for V^T
 * a multiple of 8 so that every TMA box origin is 16-byte aligned), or 0 when every sequence attends to the same keys
 * (ITM rerank of one image against many captions). out[b, i, h*64 + d] fp16.
 */
int madtp_attn_cross_tc(const void* q_f16, int64_t ldq, const void* k_f16, int64_t ldk, int k_rows_per_batch,
                        const void* vt_f16, int64_t ld_vt, int vt_cols_per_batch, const float* v_bias, int B, int H,
                        int Lq, int Nk, float scale, const float* key_mask, void* out_f16, int64_t ldo, int64_t bso,
                        const int32_t* lq_dev, const int32_t* nk_dev, const int32_t* k_start_dev,
                        const int32_t* k_len_dev, const float* key0_bias_dev, void* stream);
/* Ragged keys (k_start_dev / k_len_dev [B], both or neither; key0_bias_dev [B] optional): sequence b attends to
 * k_len[b] keys that start at row / V^T column k_start[b] (a multiple of 8) of a PACKED K / V^T -- the t2i ITM rerank of
 * one caption against k_test images whose pruned lengths differ (compress_retrieval_dtp.py:186-200). k_rows_per_batch /
 * vt_cols_per_batch then carry the packed totals and Nk = max_b k_len[b]. key0_bias[b] is added to the logit of key 0:
 * ln(1 + pad_b) evaluates "pad the shorter image with pad_b copies of its CLS token" (:142-154) exactly, without the
 * copies. */
/* lq_dev / nk_dev: dynamic Lq (packed q / out) and Nk (packed key_mask; non-zero per-sequence pitches become Nk rounded
 * up to 8, the layout madtp_layernorm_pack + the K / V^T projections produce). */

/* vector_gather (models/utils.py:13-33): out[b,i,:] = x[b, idx[b,i], :], x [B,L,d] with batch stride bsx, idx [B,K]
 * (indices are clamped to [0, L)), out [B,K,d] contiguous. */
int madtp_gather_rows(const float* x, int64_t bsx, const int32_t* idx, float* out, int B, int L, int K, int d,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MADTP_B200_H_ */
