// extern "C" boundary of libmadtp_b200.so -- see include/madtp_b200.h for the contract of every entry point.
#include <stdlib.h>
#include "../../include/madtp_b200.h"

#include <stdarg.h>

#include <atomic>

#include "attention.cuh"
#include "common.cuh"
#include "dtp.cuh"
#include "gemm.cuh"
#include "rowops.cuh"

namespace madtp {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
    // development switch: size the persistent kernels' grids for fewer SMs than the device has, leaving the rest to
    // kernels of another stream (bench.py --streams 2: the latency-bound text encoder of the other batch)
    const char* lim = getenv("MADTP_SM_LIMIT");
    if (lim != nullptr && atoi(lim) > 0 && atoi(lim) < sms) sms = atoi(lim);
  }
  return sms;
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int counted(int status, int launches = 1) {
  if (status == kOk) g_launches.fetch_add(launches, std::memory_order_relaxed);
  return status;
}

}  // namespace madtp

using namespace madtp;

extern "C" {

int madtp_abi_version(void) { return MADTP_B200_ABI_VERSION; }
const char* madtp_last_error_string(void) { return last_error(); }
long long madtp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int madtp_gemm(int precision, const void* a, const void* a_lo, int64_t lda, const void* b, const void* b_lo,
               int64_t ldb, void* c, int64_t ldc, int c_f16, const float* bias, const float* residual, int64_t ldr,
               int act, float alpha, int M, int N, int K, const int32_t* m_dev, int m_mult, const int32_t* n_dev,
               int n_mult, void* stream) {
  GemmEpilogue ep = {};
  ep.m_dev = m_dev; ep.m_mult = m_mult;
  ep.n_dev = n_dev; ep.n_mult = n_mult;
  MADTP_CHECK_ARG((m_dev == nullptr || m_mult > 0) && (n_dev == nullptr || n_mult > 0), "gemm: m_mult / n_mult must be > 0");
  ep.c = c;
  ep.ldc = ldc;
  ep.c_f16 = c_f16;
  ep.bias = bias;
  ep.residual = residual;
  ep.ldr = ldr;
  ep.act = act;
  ep.alpha = alpha;
  MADTP_CHECK_ARG(act >= 0 && act <= 4, "gemm: unknown activation %d", act);
  return counted(launch_gemm(precision, a, a_lo, lda, b, b_lo, ldb, ep, M, N, K, as_stream(stream)), M > 0 ? 1 : 0);
}

int madtp_layernorm(const float* x, int64_t ldx, int rows, int d, const float* gamma, const float* beta, float eps,
                    float* y_f32, void* y_hi, void* y_lo, void* y_f16, void* x_hi, void* x_lo, const int32_t* n_dev,
                    int n_mult, void* stream) {
  LayerNormArgs a;
  a.n_dev = n_dev;
  a.n_mult = n_mult;
  a.x = x;
  a.ldx = ldx;
  a.rows = rows;
  a.d = d;
  a.gamma = gamma;
  a.beta = beta;
  a.eps = eps;
  a.y_f32 = y_f32;
  a.y_hi = static_cast<__half*>(y_hi);
  a.y_lo = static_cast<__half*>(y_lo);
  a.y_f16 = static_cast<__half*>(y_f16);
  a.x_hi = static_cast<__half*>(x_hi);
  a.x_lo = static_cast<__half*>(x_lo);
  return counted(launch_layernorm(a, as_stream(stream)), rows > 0 ? 1 : 0);
}

int madtp_layernorm_pack(const float* x, int B, int N, int d, const float* gamma, const float* beta, float eps,
                         float* y_f32, void* y16, int per_group, int64_t group_stride, int32_t* p_out_dev,
                         const int32_t* n_dev, void* stream) {
  LayerNormPackArgs a;
  a.x = x; a.B = B; a.N = N; a.d = d; a.n_dev = n_dev;
  a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.y_f32 = y_f32; a.y16 = static_cast<__half*>(y16);
  a.per_group = per_group; a.group_stride = group_stride; a.p_out = p_out_dev;
  return counted(launch_layernorm_pack(a, as_stream(stream)), B > 0 ? 1 : 0);
}
int madtp_take_token(const float* x, int B, int N, int token, int d, float* out, const int32_t* n_dev, void* stream) {
  return counted(launch_take_token(x, B, N, n_dev, token, d, out, as_stream(stream)), B > 0 ? 1 : 0);
}

int madtp_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
  return counted(launch_split_tf32(x, hi, lo, n, as_stream(stream)), n > 0 ? 1 : 0);
}
int madtp_split_f16(const float* x, void* hi_f16, void* lo_f16, int64_t n, float scale, void* stream) {
  return counted(launch_split_f16(x, hi_f16, lo_f16, n, scale, as_stream(stream)), n > 0 ? 1 : 0);
}
int madtp_cast_f16(const float* x, void* y_f16, int64_t n, void* stream) {
  return counted(launch_cast_f16(x, y_f16, n, as_stream(stream)), n > 0 ? 1 : 0);
}

int madtp_patchify(const float* img, void* rows_hi, void* rows_lo, int B, int C, int H, int W, int P, void* stream) {
  return counted(launch_patchify(img, rows_hi, rows_lo, B, C, H, W, P, as_stream(stream)), B > 0 ? 1 : 0);
}
int madtp_assemble_tokens(const float* patches, const float* cls, const float* pos, float* x, int B, int n, int d,
                          void* stream) {
  return counted(launch_assemble_tokens(patches, cls, pos, x, B, n, d, as_stream(stream)), B > 0 ? 1 : 0);
}
int madtp_bert_embed(const int64_t* ids, const float* word, const float* position, float* out, int B, int L, int d,
                     int vocab, int n_pos, void* stream) {
  return counted(launch_bert_embed(reinterpret_cast<const long long*>(ids), word, position, out, B, L, d, vocab, n_pos,
                                   as_stream(stream)),
                 B > 0 ? 1 : 0);
}

int madtp_lm_nll(const float* logits, int64_t ld, int R, int V, const int64_t* labels, float label_smoothing,
                 float* loss, float* lse, void* stream) {
  return counted(launch_lm_nll(logits, ld, R, V, reinterpret_cast<const long long*>(labels), label_smoothing, loss, lse,
                               as_stream(stream)),
                 R > 0 ? 1 : 0);
}

int madtp_attn_fwd(const float* q, int64_t ldq, int64_t bsq, const float* k, int64_t ldk, int64_t bsk, const float* v,
                   int64_t ldv, int64_t bsv, int B, int H, int Nq, int Nk, float scale, const float* key_mask,
                   void* out_f16, int64_t ldo, int64_t bso, float* row_max, float* row_sum, float* out_norm,
                   int causal, void* stream) {
  AttnArgs a = {};
  a.q = q; a.ldq = ldq; a.bsq = bsq;
  a.k = k; a.ldk = ldk; a.bsk = bsk;
  a.v = v; a.ldv = ldv; a.bsv = bsv;
  a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk;
  a.scale = scale;
  a.key_mask = key_mask;
  a.out_f16 = static_cast<__half*>(out_f16); a.ldo = ldo; a.bso = bso;
  a.row_max = row_max; a.row_sum = row_sum; a.out_norm = out_norm;
  a.causal = causal;
  return counted(launch_attn_fwd(a, as_stream(stream)), B > 0 ? 1 : 0);
}

int madtp_attn_stats(const float* q, int64_t ldq, int64_t bsq, const float* k, int64_t ldk, int64_t bsk, int B, int H,
                     int N, float scale, const float* key_mask, const float* row_max, const float* row_sum,
                     const float* out_norm, float* col_part, float* cls_attn, int causal, void* stream) {
  AttnArgs a = {};
  a.q = q; a.ldq = ldq; a.bsq = bsq;
  a.k = k; a.ldk = ldk; a.bsk = bsk;
  a.v = k; a.ldv = ldk; a.bsv = bsk;  // unused by the statistics pass
  a.B = B; a.H = H; a.Nq = N; a.Nk = N;
  a.scale = scale;
  a.key_mask = key_mask;
  a.row_max = const_cast<float*>(row_max);
  a.row_sum = const_cast<float*>(row_sum);
  a.out_norm = const_cast<float*>(out_norm);
  a.col_part = col_part;
  a.cls_attn = cls_attn;
  a.causal = causal;
  return counted(launch_attn_stats(a, as_stream(stream)), B > 0 ? 1 : 0);
}

int madtp_attn_small_self(const float* q, int64_t ldq, int64_t bsq, const float* k, int64_t ldk, int64_t bsk,
                          const float* v, int64_t ldv, int64_t bsv, int B, int H, int L, float scale,
                          const float* key_mask, void* out_f16, int64_t ldo, int64_t bso, float* col_sum,
                          float* cls_attn, float* scratch, int causal, const int32_t* l_dev, void* stream) {
  AttnArgs a = {};
  a.q = q; a.ldq = ldq; a.bsq = bsq;
  a.k = k; a.ldk = ldk; a.bsk = bsk;
  a.v = v; a.ldv = ldv; a.bsv = bsv;
  a.B = B; a.H = H; a.Nq = L; a.Nk = L;
  a.scale = scale;
  a.key_mask = key_mask;
  a.out_f16 = static_cast<__half*>(out_f16); a.ldo = ldo; a.bso = bso;
  a.causal = causal;
  return counted(launch_small_self_attn(a, col_sum, cls_attn, scratch, l_dev, as_stream(stream)),
                 B > 0 ? (col_sum ? 2 : 1) : 0);
}

int madtp_token_colstats(const float* token_att, int64_t ld_ta, int64_t bs_ta, int B, int n, int T, float divisor,
                         float* col_max, float* col_sum, const int32_t* n_dev, int n_sub, void* stream) {
  return counted(launch_token_colstats(token_att, ld_ta, bs_ta, B, n, T, divisor, col_max, col_sum, n_dev, n_sub,
                                       as_stream(stream)),
                 B > 0 ? 1 : 0);
}
int madtp_query_sdft(const float* token_att, int64_t ld_ta, int64_t bs_ta, const float* col_max, const float* col_sum,
                     const float* ft, int64_t ld_ft, int64_t bs_ft, int B, int n, int T, int d, float divisor,
                     float* sd_ft, int accumulate, const int32_t* n_dev, int n_sub, void* stream) {
  return counted(launch_query_sdft(token_att, ld_ta, bs_ta, col_max, col_sum, ft, ld_ft, bs_ft, B, n, T, d, divisor,
                                   sd_ft, accumulate, n_dev, n_sub, as_stream(stream)),
                 B > 0 ? 1 : 0);
}

int madtp_query_sdft_tc(const float* token_att, int64_t ld_ta, int64_t bs_ta, const float* col_max, const float* col_sum,
                        const float* x, int64_t x_rows, int row_stride, int first_row, int B, int n, int T, int d,
                        float divisor, float* sd_ft, int accumulate, const int32_t* n_dev, void* stream) {
  return counted(launch_query_sdft_tc(token_att, ld_ta, bs_ta, col_max, col_sum, x, x_rows, row_stride,
                                      first_row, B, n, T, d, divisor, sd_ft, accumulate, n_dev, as_stream(stream)),
                 B > 0 ? 1 : 0);
}

int madtp_query_sdft_planes(const float* token_att, int64_t ld_ta, int64_t bs_ta, const float* col_max,
                            const float* col_sum, const void* x_hi, const void* x_lo, float x_unscale, int64_t x_rows,
                            int row_stride, int first_row, int B, int n, int T, int d, float divisor, float* sd_ft,
                            int accumulate, const int32_t* n_dev, void* stream) {
  return counted(launch_query_sdft_planes(token_att, ld_ta, bs_ta, col_max, col_sum, static_cast<const __half*>(x_hi),
                                          static_cast<const __half*>(x_lo), x_unscale, x_rows, row_stride, first_row, B,
                                          n, T, d, divisor, sd_ft, accumulate, n_dev, as_stream(stream)),
                 B > 0 ? 1 : 0);
}

int madtp_dtp_score(int B, int n, int T, const float* col_part, int n_parts, const float* cls_attn,
                    const float* token_att, int64_t ld_ta, int64_t bs_ta, float temperature, float* score,
                    float* threshold, int32_t* count, int32_t* topk, const int32_t* n_dev, int parts_tile,
                    void* stream) {
  DtpScoreArgs a;
  a.n_dev = n_dev; a.parts_tile = parts_tile;
  a.B = B; a.n = n; a.T = T;
  a.col_part = col_part; a.n_parts = n_parts;
  a.cls_attn = cls_attn;
  a.token_att = token_att; a.ld_ta = ld_ta; a.bs_ta = bs_ta;
  a.temperature = temperature;
  a.score = score; a.threshold = threshold; a.count = count; a.topk = topk;
  return counted(launch_dtp_score(a, as_stream(stream)), B > 0 ? 1 : 0);
}

int madtp_dtp_select(int B, int n, const float* score, const int32_t* topk, uint8_t* keep, int32_t* dst, float* tail_w,
                     int32_t* tail_idx, int mask_mode, const float* mask_in, float* mask_out, int max_keep,
                     const int32_t* n_dev, int32_t* n_out_dev, int32_t* k_out_dev, void* stream) {
  DtpSelectArgs a;
  a.n_dev = n_dev; a.n_out = n_out_dev; a.k_out = k_out_dev;
  MADTP_CHECK_ARG(n_dev == nullptr || n_out_dev != nullptr, "dtp_select: n_dev needs n_out_dev");
  a.B = B; a.n = n;
  a.score = score; a.topk = topk;
  a.keep = keep; a.dst = dst; a.tail_w = tail_w; a.tail_idx = tail_idx;
  a.mask_mode = mask_mode; a.mask_in = mask_in; a.mask_out = mask_out;
  a.max_keep = max_keep;
  return counted(launch_dtp_select(a, as_stream(stream)), B > 0 ? 1 : 0);
}

int madtp_dtp_gather(int B, int n, int d, const float* x, int64_t bsx, const int32_t* topk, const int32_t* dst,
                     const float* tail_w, const int32_t* tail_idx, float* out, int64_t bso, void* out_f16, int max_keep,
                     const int32_t* n_dev, void* stream) {
  DtpGatherArgs a;
  a.n_dev = n_dev;
  a.B = B; a.n = n; a.d = d;
  a.x = x; a.bsx = bsx;
  a.topk = topk; a.dst = dst; a.tail_w = tail_w; a.tail_idx = tail_idx;
  a.out = out; a.bso = bso;
  a.out_f16 = static_cast<__half*>(out_f16);
  a.max_keep = max_keep;
  return counted(launch_dtp_gather(a, as_stream(stream)), B > 0 ? 1 : 0);
}

int madtp_dtp_apply(int B, int n, int d, const float* score, const int32_t* topk, const float* x, int64_t bsx, float* out,
                    int64_t bso, void* out_f16, const float* ln_gamma, const float* ln_beta, float ln_eps,
                    void* ln_out_f16, uint8_t* keep, int mask_mode, const float* mask_in, float* mask_out, int max_keep,
                    const int32_t* n_dev, int32_t* n_out_dev, int32_t* k_out_dev, void* stream) {
  DtpApplyArgs a;
  a.B = B; a.n = n; a.d = d;
  a.score = score; a.topk = topk;
  a.x = x; a.bsx = bsx; a.out = out; a.bso = bso;
  a.out_f16 = static_cast<__half*>(out_f16);
  a.ln_gamma = ln_gamma; a.ln_beta = ln_beta; a.ln_eps = ln_eps; a.ln_out = static_cast<__half*>(ln_out_f16);
  a.keep = keep;
  a.mask_mode = mask_mode; a.mask_in = mask_in; a.mask_out = mask_out; a.max_keep = max_keep;
  a.n_dev = n_dev; a.n_out = n_out_dev; a.k_out = k_out_dev;
  return counted(launch_dtp_apply(a, as_stream(stream)), B > 0 ? 1 : 0);
}

int madtp_gemm_qkv(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo, int64_t ldb,
                   const float* bias, float alpha, int M, int K, int n_tok, int heads, void* qk_hi, void* qk_lo,
                   int64_t ld_qk, void* vt_hi, void* vt_lo, int64_t ld_vt, const int32_t* n_dev, void* stream) {
  return counted(launch_gemm_qkv(a_hi, a_lo, lda, w_hi, w_lo, ldb, bias, alpha, M, K, n_tok, heads, qk_hi, qk_lo, ld_qk, vt_hi,
                                 vt_lo, ld_vt, n_dev, as_stream(stream)),
                 M > 0 ? 1 : 0);
}

int madtp_attn_tc_fwd(const void* qk_hi, const void* qk_lo, int64_t ld_qk, const void* vt_hi, const void* vt_lo,
                      int64_t ld_vt, int B, int H, int N, float scale, const float* key_mask, void* out_f16,
                      int64_t ldo, int64_t bso, float* row_lse, float* out_norm, float* cls_p, float* cls_tile_max,
                      int causal, const int32_t* n_dev, float* out_f32, void* stream) {
  AttnTcArgs a = {};
  a.n_dev = n_dev;
  a.causal = causal;
  a.qk_hi = static_cast<const __half*>(qk_hi); a.qk_lo = static_cast<const __half*>(qk_lo); a.ld_qk = ld_qk;
  a.vt_hi = static_cast<const __half*>(vt_hi); a.vt_lo = static_cast<const __half*>(vt_lo); a.ld_vt = ld_vt;
  a.B = B; a.H = H; a.N = N; a.scale = scale; a.key_mask = key_mask;
  a.out_f16 = static_cast<__half*>(out_f16); a.ldo = ldo; a.bso = bso; a.out_f32 = out_f32;
  a.row_lse = row_lse; a.out_norm = out_norm;
  a.cls_p = cls_p; a.cls_tile_max = cls_tile_max;
  return counted(launch_attn_fwd_tc(a, as_stream(stream)), B > 0 ? 1 : 0);
}

int madtp_attn_tc_stats(const void* qk_hi, const void* qk_lo, int64_t ld_qk, int B, int H, int N, float scale,
                        const float* key_mask, const float* row_lse, const float* out_norm, float* col_part,
                        int n_parts, float* cls_attn, const float* cls_p, const float* cls_tile_max, int causal,
                        const int32_t* n_dev, void* stream) {
  AttnTcArgs a = {};
  a.n_dev = n_dev;
  a.causal = causal;
  a.qk_hi = static_cast<const __half*>(qk_hi); a.qk_lo = static_cast<const __half*>(qk_lo); a.ld_qk = ld_qk;
  a.B = B; a.H = H; a.N = N; a.scale = scale; a.key_mask = key_mask;
  a.row_lse = const_cast<float*>(row_lse); a.out_norm = const_cast<float*>(out_norm);
  a.col_part = col_part; a.n_parts = n_parts; a.cls_attn = cls_attn;
  a.cls_p = const_cast<float*>(cls_p); a.cls_tile_max = const_cast<float*>(cls_tile_max);
  return counted(launch_attn_stats_tc(a, as_stream(stream)), B > 0 ? 2 : 0);
}

int madtp_attn_cross_tc(const void* q_f16, int64_t ldq, const void* k_f16, int64_t ldk, int k_rows_per_batch,
                        const void* vt_f16, int64_t ld_vt, int vt_cols_per_batch, const float* v_bias, int B, int H,
                        int Lq, int Nk, float scale, const float* key_mask, void* out_f16, int64_t ldo, int64_t bso,
                        const int32_t* lq_dev, const int32_t* nk_dev, const int32_t* k_start_dev,
                        const int32_t* k_len_dev, const float* key0_bias_dev, void* stream) {
  CrossTcArgs a = {};
  a.lq_dev = lq_dev; a.nk_dev = nk_dev;
  a.k_start = k_start_dev; a.k_len = k_len_dev; a.key0_bias = key0_bias_dev;
  a.q = static_cast<const __half*>(q_f16); a.ldq = ldq;
  a.k = static_cast<const __half*>(k_f16); a.ldk = ldk; a.k_rows_per_batch = k_rows_per_batch;
  a.vt = static_cast<const __half*>(vt_f16); a.ld_vt = ld_vt; a.vt_cols_per_batch = vt_cols_per_batch;
  a.v_bias = v_bias;
  a.B = B; a.H = H; a.Lq = Lq; a.Nk = Nk; a.scale = scale; a.key_mask = key_mask;
  a.out_f16 = static_cast<__half*>(out_f16); a.ldo = ldo; a.bso = bso;
  return counted(launch_cross_attn_tc(a, as_stream(stream)), B > 0 ? 1 : 0);
}

// ---- asynchronous scalar read-back (the per-layer `topk_num`) --------------------------------------------------
// The reference reads topk_num with .item() (models/vit.py:145), a device->host copy ordered behind everything already
// queued on the compute stream. Here the copy runs on a library-owned side stream behind an event recorded right after
// the score kernel, so kernels queued later on the compute stream (the attention output projection, which does not
// depend on the pruning decision) execute while the host waits for the four bytes.
namespace {
constexpr int kRbSlots = 8, kRbDevices = 16;
struct ReadbackState {
  cudaStream_t side = nullptr;
  cudaEvent_t ready[kRbSlots] = {}, done[kRbSlots] = {};
};
ReadbackState g_rb[kRbDevices];
int readback_state(ReadbackState** out) {
  int dev = 0;
  MADTP_CUDA(cudaGetDevice(&dev));
  MADTP_CHECK_ARG(dev >= 0 && dev < kRbDevices, "readback: device index %d out of range", dev);
  ReadbackState& st = g_rb[dev];
  if (st.side == nullptr) {
    MADTP_CUDA(cudaStreamCreateWithFlags(&st.side, cudaStreamNonBlocking));
    for (int i = 0; i < kRbSlots; ++i) {
      MADTP_CUDA(cudaEventCreateWithFlags(&st.ready[i], cudaEventDisableTiming));
      MADTP_CUDA(cudaEventCreateWithFlags(&st.done[i], cudaEventDisableTiming));
    }
  }
  *out = &st;
  return kOk;
}
}  // namespace

int madtp_readback_begin(const void* src_dev, void* dst_pinned, int64_t bytes, int slot, void* stream) {
  MADTP_CHECK_ARG(src_dev && dst_pinned && bytes > 0 && slot >= 0 && slot < kRbSlots, "readback_begin: bad arguments");
  ReadbackState* st = nullptr;
  int rc = readback_state(&st);
  if (rc != kOk) return rc;
  MADTP_CUDA(cudaEventRecord(st->ready[slot], as_stream(stream)));
  MADTP_CUDA(cudaStreamWaitEvent(st->side, st->ready[slot], 0));
  MADTP_CUDA(cudaMemcpyAsync(dst_pinned, src_dev, static_cast<size_t>(bytes), cudaMemcpyDeviceToHost, st->side));
  MADTP_CUDA(cudaEventRecord(st->done[slot], st->side));
  return kOk;
}

int madtp_readback_wait(int slot) {
  MADTP_CHECK_ARG(slot >= 0 && slot < kRbSlots, "readback_wait: bad slot");
  ReadbackState* st = nullptr;
  int rc = readback_state(&st);
  if (rc != kOk) return rc;
  MADTP_CUDA(cudaEventSynchronize(st->done[slot]));
  return kOk;
}

int madtp_gather_rows(const float* x, int64_t bsx, const int32_t* idx, float* out, int B, int L, int K, int d,
                      void* stream) {
  return counted(launch_gather_rows(x, bsx, idx, out, B, L, K, d, as_stream(stream)), (B > 0 && K > 0) ? 1 : 0);
}

}  // extern "C"
