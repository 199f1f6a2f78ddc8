// Attention kernels for the SHORT sequences of the text encoders (fp32 CUDA cores; everything fits in shared memory):
//
//   small_self_attn_kernel   one CTA per (head, sequence), L <= 64 tokens: logits, softmax, context, ||context||; the
//                            probabilities of every head go to a small scratch buffer
//   small_self_stats_kernel  one CTA per sequence: max over heads, column sums over the non-CLS queries and the
//                            head-importance-weighted CLS row (models/nlvr_encoder.py:213-235,404-406)
// replacing an attn_fwd + attn_stats pair whose 64x64 tiles are mostly padding at L ~ 20-35.
#include "attention.cuh"

namespace madtp {

namespace {
constexpr int SL = 64;        // max self-attention length
constexpr int SP = 65;        // shared-memory pitch (floats)
}  // namespace

struct SmallSelfArgs {
  const float* q; long long ldq, bsq;
  const float* k; long long ldk, bsk;
  const float* v; long long ldv, bsv;
  int B, H, L;
  float scale;
  const float* key_mask;     // additive [B, L] or nullptr
  int causal;
  __half* out_f16; long long ldo, bso;
  float* col_sum;            // [B, L]  sum_{i>=1} max_h P[b,h,i,j]   (one "part": n_parts = 1); nullptr = no statistics
  float* cls_attn;           // [B, L]
  float* p_scratch;          // [B, H, L, L] probabilities of every head (statistics only)
  float* n_scratch;          // [B, H, L]    ||context[b,h,i]||
  const int* n_dev;          // dynamic L (packed sequences: every batch stride becomes L * row pitch)
};

__device__ __forceinline__ void small_self_dyn(SmallSelfArgs& a) {
  if (a.n_dev == nullptr) return;
  const int L = min(a.L, load_len(a.n_dev));
  a.L = L;
  a.bsq = L * a.ldq;
  a.bsk = L * a.ldk;
  a.bsv = L * a.ldv;
  a.bso = L * a.ldo;
}

// grid (H, B), block 128: one head of one sequence
__global__ void __launch_bounds__(128)
small_self_attn_kernel(SmallSelfArgs a) {
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + SL * SP;
  float* Vs = Ks + SL * SP;
  float* Ps = Vs + SL * SP;
  float* Msk = Ps + SL * SP;
  float* Nsq = Msk + SL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  small_self_dyn(a);
  const int h = blockIdx.x, b = blockIdx.y, L = a.L, H = a.H;
  const bool stats = a.col_sum != nullptr;
  if (tid < SL) Msk[tid] = (a.key_mask != nullptr && tid < L) ? a.key_mask[static_cast<long long>(b) * L + tid] : 0.f;
  for (int x = tid; x < L * 16; x += 128) {
    const int r = x >> 4, c4 = x & 15;
    const float4 qv = *reinterpret_cast<const float4*>(a.q + b * a.bsq + r * a.ldq + h * 64 + c4 * 4);
    const float4 kv = *reinterpret_cast<const float4*>(a.k + b * a.bsk + r * a.ldk + h * 64 + c4 * 4);
    const float4 vv = *reinterpret_cast<const float4*>(a.v + b * a.bsv + r * a.ldv + h * 64 + c4 * 4);
    float* qd = Qs + r * SP + c4 * 4;
    float* kd = Ks + r * SP + c4 * 4;
    float* vd = Vs + r * SP + c4 * 4;
    qd[0] = qv.x; qd[1] = qv.y; qd[2] = qv.z; qd[3] = qv.w;
    kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
    vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
  }
  __syncthreads();
  for (int x = tid; x < L * L; x += 128) {
    const int i = x / L, j = x - i * L;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; d += 2) {
      s0 = fmaf(Qs[i * SP + d], Ks[j * SP + d], s0);
      s1 = fmaf(Qs[i * SP + d + 1], Ks[j * SP + d + 1], s1);
    }
    const bool vis = !a.causal || j <= i;
    Ps[i * SP + j] = vis ? fmaf(s0 + s1, a.scale, Msk[j]) : -INFINITY;
  }
  __syncthreads();
  for (int i = warp; i < L; i += 4) {   // softmax: one warp per row
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) mx = fmaxf(mx, Ps[i * SP + j]);
    mx = warp_max(mx);
    float e0 = 0.f, e1 = 0.f, sum = 0.f;
    if (lane < L) { e0 = expf(Ps[i * SP + lane] - mx); sum += e0; }
    if (lane + 32 < L) { e1 = expf(Ps[i * SP + lane + 32] - mx); sum += e1; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    if (lane < L) Ps[i * SP + lane] = e0 * inv;
    if (lane + 32 < L) Ps[i * SP + lane + 32] = e1 * inv;
  }
  __syncthreads();
  // context: thread = (row, dim); a row is covered by two consecutive warps
  const int rows_pad = (L + 1) & ~1;
  for (int x = tid; x < rows_pad * 64; x += 128) {
    const int i = x >> 6, d = x & 63;
    float o = 0.f;
    if (i < L) {
      for (int j = 0; j < L; ++j) o = fmaf(Ps[i * SP + j], Vs[j * SP + d], o);
      a.out_f16[b * a.bso + static_cast<long long>(i) * a.ldo + h * 64 + d] = __float2half_rn(o);
    }
    if (stats) {
      const float q2 = warp_sum(o * o);
      if (lane == 0 && i < L) Nsq[i * 2 + (d >> 5)] = q2;
    }
  }
  if (!stats) return;
  __syncthreads();
  const long long bh = static_cast<long long>(b) * H + h;
  if (tid < L) a.n_scratch[bh * L + tid] = sqrtf(Nsq[tid * 2] + Nsq[tid * 2 + 1]);
  for (int x = tid; x < L * L; x += 128) {
    const int i = x / L, j = x - i * L;
    a.p_scratch[(bh * L + i) * L + j] = Ps[i * SP + j];
  }
}

// grid B, block 256: max over heads, column sums over the non-CLS queries, head-importance-weighted CLS row.
// Thread (i-slice s = tid / 64, column j = tid % 64): partial column sums over rows i = 1 + s, 5 + s, ... combined in
// slice order (fixed order, deterministic).
__global__ void __launch_bounds__(256)
small_self_stats_kernel(SmallSelfArgs a) {
  __shared__ float part[4][SL];
  small_self_dyn(a);
  const int j = threadIdx.x & 63, sl = threadIdx.x >> 6, b = blockIdx.x, L = a.L, H = a.H;
  const float* P = a.p_scratch + static_cast<long long>(b) * H * L * L;
  const float* Nr = a.n_scratch + static_cast<long long>(b) * H * L;
  float col = 0.f;
  if (j < L) {
    for (int i = 1 + sl; i < L; i += 4) {
      float mx = 0.f;
      for (int h = 0; h < H; ++h) mx = fmaxf(mx, __ldg(P + (h * L + i) * L + j));
      col += mx;
    }
  }
  part[sl][j] = col;
  __syncthreads();
  if (sl != 0 || j >= L) return;
  a.col_sum[static_cast<long long>(b) * L + j] = (part[0][j] + part[1][j]) + (part[2][j] + part[3][j]);
  float hs = 0.f;
  for (int h = 0; h < H; ++h) hs += Nr[h * L + j];
  hs += 1e-8f;
  float acc = 0.f;
  for (int h = 0; h < H; ++h) acc += P[(h * L) * L + j] * (Nr[h * L + j] / hs);
  a.cls_attn[static_cast<long long>(b) * L + j] = acc;
}

int launch_small_self_attn(const AttnArgs& g, float* col_sum, float* cls_attn, float* scratch, const int* n_dev,
                           cudaStream_t stream) {
  MADTP_CHECK_ARG(g.q && g.k && g.v && g.out_f16, "small_self_attn: null pointer");
  MADTP_CHECK_ARG(g.Nq == g.Nk && g.Nq >= 1 && g.Nq <= SL && g.H >= 1 && g.H <= 65535 && g.B <= 65535,
                  "small_self_attn: needs Nq == Nk <= %d", SL);
  MADTP_CHECK_ARG(g.ldq % 4 == 0 && g.ldk % 4 == 0 && g.ldv % 4 == 0 && g.bsq % 4 == 0 && g.bsk % 4 == 0 &&
                      g.bsv % 4 == 0,
                  "small_self_attn: strides must be multiples of 4 elements");
  MADTP_CHECK_ARG((col_sum == nullptr) == (cls_attn == nullptr), "small_self_attn: statistics come together");
  MADTP_CHECK_ARG(col_sum == nullptr || scratch != nullptr, "small_self_attn: statistics need the scratch buffer");
  if (g.B == 0) return kOk;
  SmallSelfArgs a;
  a.q = g.q; a.ldq = g.ldq; a.bsq = g.bsq;
  a.k = g.k; a.ldk = g.ldk; a.bsk = g.bsk;
  a.v = g.v; a.ldv = g.ldv; a.bsv = g.bsv;
  a.B = g.B; a.H = g.H; a.L = g.Nq; a.scale = g.scale; a.key_mask = g.key_mask; a.causal = g.causal;
  a.out_f16 = g.out_f16; a.ldo = g.ldo; a.bso = g.bso;
  a.col_sum = col_sum; a.cls_attn = cls_attn;
  a.p_scratch = scratch;
  a.n_scratch = scratch ? scratch + static_cast<long long>(g.B) * g.H * g.Nq * g.Nq : nullptr;
  a.n_dev = n_dev;
  const int smem = (4 * SL * SP + 3 * SL) * sizeof(float);
  MADTP_SMEM_ATTR_ONCE(smem, small_self_attn_kernel);
  small_self_attn_kernel<<<dim3(g.H, g.B), 128, smem, stream>>>(a);
  MADTP_LAUNCH_CHECK();
  if (col_sum != nullptr) {
    small_self_stats_kernel<<<g.B, 256, 0, stream>>>(a);
    MADTP_LAUNCH_CHECK();
  }
  return kOk;
}

}  // namespace madtp
