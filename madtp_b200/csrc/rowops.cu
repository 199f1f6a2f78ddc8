// Row-wise HBM-bound kernels around the GEMMs: LayerNorm with fused operand preparation (fp16 hi/lo split, fp16
// copy), the splits themselves, patch extraction for the ViT stem, token assembly and the BERT embedding lookup.
// One warp owns one row; rows are d <= 1024 floats held in registers, accessed with 128-bit loads/stores.
#include "rowops.cuh"

namespace madtp {

// ------------------------------------------------------------------------------------------------
// LayerNorm (two-pass in registers: mean, then centred variance) + optional operand preparation.
// ------------------------------------------------------------------------------------------------
// fp16 hi/lo split of four consecutive values: hi = fp16(v), lo = fp16(v - hi); 8-byte stores.
__device__ __forceinline__ void store_split4(__half* hi, __half* lo, long long idx4, const float4& v) {
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
  uint2 ph, pl;
  ph.x = *reinterpret_cast<const uint32_t*>(&h0);
  ph.y = *reinterpret_cast<const uint32_t*>(&h1);
  pl.x = *reinterpret_cast<const uint32_t*>(&l0);
  pl.y = *reinterpret_cast<const uint32_t*>(&l1);
  reinterpret_cast<uint2*>(hi)[idx4] = ph;
  reinterpret_cast<uint2*>(lo)[idx4] = pl;
}

template <int V>  // V float4 per lane: d == 128 * V
__global__ void __launch_bounds__(256)
layernorm_kernel(LayerNormArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int rows = a.n_dev ? min(a.rows, load_len(a.n_dev) * a.n_mult) : a.rows;
  if (warp >= rows) return;
  const long long row = warp;
  const float4* xin = reinterpret_cast<const float4*>(a.x + row * a.ldx);
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = xin[lane + 32 * i];

  if (a.x_hi) {  // hi/lo split of the *input* row (operand of the token/codebook product)
#pragma unroll
    for (int i = 0; i < V; ++i) store_split4(a.x_hi + row * a.d, a.x_lo + row * a.d, lane + 32 * i, v[i]);
  }
  if (a.gamma == nullptr) return;  // split-only call

  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / static_cast<float>(a.d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float var = warp_sum(q) / static_cast<float>(a.d);
  const float rstd = 1.0f / sqrtf(var + a.eps);

  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(a.beta);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    const int c = lane + 32 * i;
    if (a.y_f32) reinterpret_cast<float4*>(a.y_f32 + row * a.d)[c] = y;
    if (a.y_hi) store_split4(a.y_hi + row * a.d, a.y_lo + row * a.d, c, y);
    if (a.y_f16) {
      __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      reinterpret_cast<uint2*>(a.y_f16 + row * a.d)[c] = pk;
    }
  }
}

int launch_layernorm(const LayerNormArgs& a, cudaStream_t stream) {
  MADTP_CHECK_ARG(a.rows >= 0 && a.d > 0 && a.d % 128 == 0 && a.d <= 1024,
                  "layernorm: d must be a multiple of 128 and <= 1024 (d=%d)", a.d);
  MADTP_CHECK_ARG(a.x != nullptr && a.ldx % 4 == 0, "layernorm: bad input");
  MADTP_CHECK_ARG((a.x_hi == nullptr) == (a.x_lo == nullptr) && (a.y_hi == nullptr) == (a.y_lo == nullptr),
                  "layernorm: hi/lo outputs come in pairs");
  MADTP_CHECK_ARG((a.gamma == nullptr) == (a.beta == nullptr), "layernorm: gamma/beta come in pairs");
  if (a.rows == 0) return kOk;
  const int blocks = (a.rows + 7) / 8;
  switch (a.d / 128) {
#define MADTP_LN_CASE(V)                                  \
  case V:                                                 \
    layernorm_kernel<V><<<blocks, 256, 0, stream>>>(a);   \
    break;
    MADTP_LN_CASE(1) MADTP_LN_CASE(2) MADTP_LN_CASE(3) MADTP_LN_CASE(4) MADTP_LN_CASE(5) MADTP_LN_CASE(6)
    MADTP_LN_CASE(7) MADTP_LN_CASE(8)
#undef MADTP_LN_CASE
  }
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// LayerNorm -> fp16, repacked per sequence with the row padding the cross-attention operands need (see rowops.cuh).
template <int V>
__global__ void __launch_bounds__(256)
layernorm_pack_kernel(LayerNormPackArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int N = a.n_dev ? min(a.N, load_len(a.n_dev)) : a.N;
  const int P = (N + 7) & ~7;
  if (warp == 0 && lane == 0 && a.p_out) *a.p_out = P;
  const int Pcap = (a.N + 7) & ~7;
  const int b = warp / Pcap, t = warp - b * Pcap;
  if (b >= a.B || t >= P) return;
  __half* dst = a.y16 + (b / a.per_group) * a.group_stride + (static_cast<long long>(b % a.per_group) * P + t) * a.d;
  if (t >= N) {
#pragma unroll
    for (int i = 0; i < V; ++i) reinterpret_cast<uint2*>(dst)[lane + 32 * i] = make_uint2(0u, 0u);
    return;
  }
  const long long row = static_cast<long long>(b) * N + t;
  const float4* xin = reinterpret_cast<const float4*>(a.x + row * a.d);
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = xin[lane + 32 * i];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / static_cast<float>(a.d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(a.d) + a.eps);
  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(a.beta);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = lane + 32 * i;
    const float4 g = __ldg(g4 + c), be = __ldg(b4 + c);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + be.x;
    y.y = (v[i].y - mean) * rstd * g.y + be.y;
    y.z = (v[i].z - mean) * rstd * g.z + be.z;
    y.w = (v[i].w - mean) * rstd * g.w + be.w;
    if (a.y_f32) reinterpret_cast<float4*>(a.y_f32 + row * a.d)[c] = y;
    __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0);
    pk.y = *reinterpret_cast<uint32_t*>(&h1);
    reinterpret_cast<uint2*>(dst)[c] = pk;
  }
}

int launch_layernorm_pack(const LayerNormPackArgs& a, cudaStream_t stream) {
  MADTP_CHECK_ARG(a.x && a.gamma && a.beta && a.y16, "layernorm_pack: null pointer");
  MADTP_CHECK_ARG(a.B >= 0 && a.N > 0 && a.d > 0 && a.d % 128 == 0 && a.d <= 1024 && a.per_group > 0 &&
                      a.group_stride % 8 == 0,
                  "layernorm_pack: bad shape (d=%d must be a multiple of 128, <= 1024)", a.d);
  if (a.B == 0) return kOk;
  const long long warps = static_cast<long long>(a.B) * ((a.N + 7) & ~7);
  const int blocks = static_cast<int>((warps + 7) / 8);
  switch (a.d / 128) {
#define MADTP_LNP_CASE(V)                                     \
  case V:                                                     \
    layernorm_pack_kernel<V><<<blocks, 256, 0, stream>>>(a);  \
    break;
    MADTP_LNP_CASE(1) MADTP_LNP_CASE(2) MADTP_LNP_CASE(3) MADTP_LNP_CASE(4) MADTP_LNP_CASE(5) MADTP_LNP_CASE(6)
    MADTP_LNP_CASE(7) MADTP_LNP_CASE(8)
#undef MADTP_LNP_CASE
  }
  MADTP_LAUNCH_CHECK();
  return kOk;
}

__global__ void __launch_bounds__(256)
take_token_kernel(const float* __restrict__ x, int B, int N_cap, const int* __restrict__ n_dev, int token, int d4,
                  float* __restrict__ out) {
  const int N = n_dev ? min(N_cap, load_len(n_dev)) : N_cap;
  const long long total = static_cast<long long>(B) * d4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / d4;
    const int c = static_cast<int>(i - b * d4);
    reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(x)[(b * N + token) * d4 + c];
  }
}

int launch_take_token(const float* x, int B, int N, const int* n_dev, int token, int d, float* out, cudaStream_t stream) {
  MADTP_CHECK_ARG(x && out && B >= 0 && N > 0 && token >= 0 && token < N && d > 0 && d % 4 == 0, "take_token: bad arguments");
  if (B == 0) return kOk;
  const long long total = static_cast<long long>(B) * (d / 4);
  take_token_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(x, B, N, n_dev, token, d / 4, out);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Elementwise tf32 split / fp16 cast (weights at load time, small activations)
// ------------------------------------------------------------------------------------------------
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                  long long n) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float v = x[i];
    const float h = tf32_hi(v);
    hi[i] = h;
    lo[i] = v - h;
  }
}
// fp16 hi/lo split of scale * x: hi = fp16(scale*x), lo = fp16(scale*x - hi). scale is a power of two chosen by the
// caller so that hi stays finite and lo stays a normal fp16 number for all but negligible elements.
__global__ void split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                                 long long n, float scale) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float v = x[i] * scale;
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}
__global__ void cast_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, long long n) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) y[i] = __float2half_rn(x[i]);
}

int launch_split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t stream) {
  MADTP_CHECK_ARG(x && hi && lo && n >= 0, "split_tf32: bad arguments");
  if (n == 0) return kOk;
  long long blocks = (n + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  split_tf32_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, hi, lo, n);
  MADTP_LAUNCH_CHECK();
  return kOk;
}
int launch_split_f16(const float* x, void* hi, void* lo, long long n, float scale, cudaStream_t stream) {
  MADTP_CHECK_ARG(x && hi && lo && n >= 0 && scale > 0.f, "split_f16: bad arguments");
  if (n == 0) return kOk;
  long long blocks = (n + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  split_f16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, static_cast<__half*>(hi), static_cast<__half*>(lo),
                                                                 n, scale);
  MADTP_LAUNCH_CHECK();
  return kOk;
}
int launch_cast_f16(const float* x, void* y, long long n, cudaStream_t stream) {
  MADTP_CHECK_ARG(x && y && n >= 0, "cast_f16: bad arguments");
  if (n == 0) return kOk;
  long long blocks = (n + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  cast_f16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, static_cast<__half*>(y), n);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// ViT stem: non-overlapping PxP patches of [B,C,H,W] -> rows [B*gh*gw, C*P*P] (Conv2d weight order c,py,px),
// written as fp16 hi/lo planes so the projection runs on the error-compensated tensor-core path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, __half* __restrict__ hi, __half* __restrict__ lo, int B, int C, int H,
                int W, int P) {
  const int gh = H / P, gw = W / P;
  const int kdim = C * P * P;
  const long long total4 = static_cast<long long>(B) * gh * gw * kdim / 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const long long e = i * 4;
    const int kk = static_cast<int>(e % kdim);
    const long long prow = e / kdim;
    const int px = kk % P, py = (kk / P) % P, c = kk / (P * P);
    const int gx = static_cast<int>(prow % gw), gy = static_cast<int>((prow / gw) % gh);
    const int b = static_cast<int>(prow / (static_cast<long long>(gw) * gh));
    const float4 v = *reinterpret_cast<const float4*>(
        img + ((static_cast<long long>(b) * C + c) * H + (gy * P + py)) * W + gx * P + px);
    store_split4(hi, lo, i, v);
  }
}

int launch_patchify(const float* img, void* hi, void* lo, int B, int C, int H, int W, int P, cudaStream_t stream) {
  MADTP_CHECK_ARG(img && hi && lo, "patchify: null pointer");
  MADTP_CHECK_ARG(P % 4 == 0 && H % P == 0 && W % P == 0 && W % 4 == 0, "patchify: H,W must be multiples of P, P of 4");
  const long long total4 = static_cast<long long>(B) * (H / P) * (W / P) * C * P * P / 4;
  if (total4 == 0) return kOk;
  long long blocks = (total4 + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  patchify_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(img, static_cast<__half*>(hi), static_cast<__half*>(lo), B, C,
                                                                H, W, P);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// x[b,0,:] = cls + pos[0];  x[b,1+p,:] = patches[b,p,:] + pos[1+p]
__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const float* __restrict__ patches, const float* __restrict__ cls,
                       const float* __restrict__ pos, float* __restrict__ x, int B, int n, int d) {
  const int d4 = d / 4;
  const long long total = static_cast<long long>(B) * (n + 1) * d4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % d4);
    const long long r = i / d4;
    const int t = static_cast<int>(r % (n + 1));
    const long long b = r / (n + 1);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + static_cast<long long>(t) * d4 + c);
    float4 v;
    if (t == 0)
      v = __ldg(reinterpret_cast<const float4*>(cls) + c);
    else
      v = reinterpret_cast<const float4*>(patches)[(b * n + (t - 1)) * d4 + c];
    reinterpret_cast<float4*>(x)[i] = make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
  }
}

int launch_assemble_tokens(const float* patches, const float* cls, const float* pos, float* x, int B, int n, int d,
                           cudaStream_t stream) {
  MADTP_CHECK_ARG(patches && cls && pos && x && d % 4 == 0, "assemble_tokens: bad arguments");
  const long long total = static_cast<long long>(B) * (n + 1) * (d / 4);
  if (total == 0) return kOk;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  assemble_tokens_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(patches, cls, pos, x, B, n, d);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// BERT embeddings: out[b,l,:] = word[ids[b,l]] + position[l]  (LayerNorm follows as a separate launch)
__global__ void __launch_bounds__(256)
bert_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ posemb,
                  float* __restrict__ out, int B, int L, int d, int vocab) {
  const int d4 = d / 4;
  const long long total = static_cast<long long>(B) * L * d4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % d4);
    const long long r = i / d4;
    const int l = static_cast<int>(r % L);
    long long id = ids[r];
    if (id < 0) id = 0;
    if (id >= vocab) id = vocab - 1;
    const float4 w = __ldg(reinterpret_cast<const float4*>(word) + id * d4 + c);
    const float4 p = __ldg(reinterpret_cast<const float4*>(posemb) + static_cast<long long>(l) * d4 + c);
    reinterpret_cast<float4*>(out)[i] = make_float4(w.x + p.x, w.y + p.y, w.z + p.z, w.w + p.w);
  }
}

int launch_bert_embed(const long long* ids, const float* word, const float* posemb, float* out, int B, int L, int d,
                      int vocab, int n_pos, cudaStream_t stream) {
  MADTP_CHECK_ARG(ids && word && posemb && out && d % 4 == 0 && vocab > 0, "bert_embed: bad arguments");
  MADTP_CHECK_ARG(L <= n_pos, "bert_embed: sequence length %d exceeds the position table (%d rows)", L, n_pos);
  const long long total = static_cast<long long>(B) * L * (d / 4);
  if (total == 0) return kOk;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  bert_embed_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(ids, word, posemb, out, B, L, d, vocab);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Language-model head statistics of one logits row (VQA answer ranking, models/med.py:1040-1047 and
// models/blip_vqa.py:168-171): lse = log sum_v exp(z_v) and, when labels are given, the label-smoothed cross entropy
//   loss = (1 - eps) (lse - z_label) + eps (lse - mean_v z_v),   0 for ignored positions (label < 0).
// One 256-thread CTA per row, two passes over the (L2-resident) row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lm_nll_kernel(const float* __restrict__ logits, long long ld, int V, const long long* __restrict__ labels, float eps,
              float* __restrict__ loss, float* __restrict__ lse_out) {
  __shared__ float red[8];
  __shared__ float red2[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* z = logits + static_cast<long long>(blockIdx.x) * ld;
  float mx = -INFINITY;
  for (int v = tid; v < V; v += 256) mx = fmaxf(mx, z[v]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float se = 0.f, sz = 0.f;
  for (int v = tid; v < V; v += 256) {
    const float x = z[v];
    se += expf(x - mx);
    sz += x;
  }
  se = warp_sum(se);
  sz = warp_sum(sz);
  if (lane == 0) {
    red[warp] = se;
    red2[warp] = sz;
  }
  __syncthreads();
  if (tid == 0) {
    float tse = 0.f, tsz = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      tse += red[w];
      tsz += red2[w];
    }
    const float lse = mx + logf(tse);
    if (lse_out) lse_out[blockIdx.x] = lse;
    if (loss) {
      const long long lab = labels[blockIdx.x];
      float l = 0.f;
      if (lab >= 0 && lab < V) l = (1.0f - eps) * (lse - z[lab]) + eps * (lse - tsz / static_cast<float>(V));
      loss[blockIdx.x] = l;
    }
  }
}

int launch_lm_nll(const float* logits, long long ld, int R, int V, const long long* labels, float eps, float* loss,
                  float* lse, cudaStream_t stream) {
  MADTP_CHECK_ARG(logits && R >= 0 && V > 0 && ld >= V, "lm_nll: bad arguments");
  MADTP_CHECK_ARG((labels == nullptr) == (loss == nullptr), "lm_nll: labels and loss come in pairs");
  MADTP_CHECK_ARG(loss || lse, "lm_nll: nothing to compute");
  MADTP_CHECK_ARG(eps >= 0.f && eps < 1.f, "lm_nll: label smoothing must be in [0, 1)");
  if (R == 0) return kOk;
  lm_nll_kernel<<<R, 256, 0, stream>>>(logits, ld, V, labels, eps, loss, lse);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

}  // namespace madtp
