// Attention with fused token-pruning statistics (head dim 64), fp32 on the CUDA cores.
//
// The DTP score needs three quantities that the reference reads off a materialised P = softmax(QK^T) of shape
// [B,H,N,N] (reference models/vit.py:81-83,96-100,126-128):
//   (1) || (P V)[b,h,j,:] ||          per head and token             -> attn_fwd   (out_norm)
//   (2) P[b,h,0,j]                    the CLS query row of each head -> attn_stats (cls_attn, weighted by (1))
//   (3) sum_i max_h P[b,h,i,j]        over non-CLS queries i         -> attn_stats (col_part)
// Nothing here materialises P. attn_fwd is a flash-style pass (online softmax) that also emits the row maxima and
// row sums; attn_stats recomputes the logits tile by tile for all heads of one (query tile, key tile), normalises
// them with the saved row statistics, takes the max over heads in registers and column-sums in a fixed order
// (no floating-point atomics, so the statistics are deterministic).
//
// Both kernels compute the logits with the same FMA order, so P in the second pass is exactly consistent with the
// row statistics of the first. Everything is fp32: keep-mask decisions hinge on score gaps of ~1e-7 relative.
//
// Thread layout (128 threads, one 64x64 tile): ty = tid/8 owns rows {ty, ty+16, ty+32, ty+48}, tx = tid%8 owns
// columns {tx, tx+8, ..., tx+56}. Shared-memory tiles are row-major with a 68-float pitch, which makes every
// 128-bit operand read of a warp conflict-free without transposing anything.
#include "attention.cuh"

namespace madtp {

namespace {

constexpr int T = 64;     // tile edge (queries, keys) and head dim
constexpr int LDS = 68;   // shared-memory row pitch in floats
constexpr int NTHREADS = 128;

// Copy a [64 x 64] fp32 tile (rows row0.. of a [n_rows, ld] matrix) into shared memory, zero-filling past n_rows.
template <int ROWS = T>
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, long long ld, int row0,
                                          int n_rows, int tid) {
#pragma unroll
  for (int it = 0; it < (ROWS * T / 4) / NTHREADS; ++it) {
    const int idx = tid + it * NTHREADS;
    const int r = idx >> 4, c4 = idx & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * ld) + c4);
    *reinterpret_cast<float4*>(dst + r * LDS + c4 * 4) = v;
  }
}

// s[r][c] = sum_k Qs[ty+16r][k] * Ks[tx+8c][k], k ascending (the order is part of the contract between passes).
template <int RT>
__device__ __forceinline__ void qk_tile(const float* Qs, const float* Ks, int ty, int tx, float (&s)[RT][8]) {
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) s[r][c] = 0.f;
#pragma unroll 2
  for (int k4 = 0; k4 < T / 4; ++k4) {
    float4 qa[RT], kb[8];
#pragma unroll
    for (int r = 0; r < RT; ++r) qa[r] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * r) * LDS + k4 * 4);
#pragma unroll
    for (int c = 0; c < 8; ++c) kb[c] = *reinterpret_cast<const float4*>(Ks + (tx + 8 * c) * LDS + k4 * 4);
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        s[r][c] = fmaf(qa[r].x, kb[c].x, s[r][c]);
        s[r][c] = fmaf(qa[r].y, kb[c].y, s[r][c]);
        s[r][c] = fmaf(qa[r].z, kb[c].z, s[r][c]);
        s[r][c] = fmaf(qa[r].w, kb[c].w, s[r][c]);
      }
  }
}

// logits = s * scale + mask[j];  keys past Nk get -inf
template <int RT>
__device__ __forceinline__ void finish_logits(float (&s)[RT][8], float scale, const float* Ms, int j0, int Nk, int tx,
                                              int causal, int i_first) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int jl = tx + 8 * c;
    const bool valid = (j0 + jl) < Nk;
    const float mk = Ms[jl];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const bool vis = valid && (!causal || (j0 + jl) <= (i_first + 16 * r));
      s[r][c] = vis ? fmaf(s[r][c], scale, mk) : -INFINITY;
    }
  }
}

__device__ __forceinline__ float group8_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Pass 1: context = softmax(QK^T * scale + mask) V, row statistics, context row norms.
// grid = (ceil(Nq/64), H, B)
// ------------------------------------------------------------------------------------------------
// RT = query-row groups per thread: the query tile has 16*RT rows (RT = 4: 64 rows; RT = 2: 32 rows for the
// cross-attention of a few text queries over many image keys, where a 64-row tile would be mostly padding).
template <int RT>
__global__ void __launch_bounds__(NTHREADS)
attn_fwd_kernel(AttnArgs a) {
  constexpr int BQ = 16 * RT;
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + BQ * LDS;
  float* Vs = Ks + T * LDS;
  float* Ps = Vs + T * LDS;
  float* Ms = Ps + BQ * LDS;  // [64] additive key mask of the current key tile

  const int tid = threadIdx.x, ty = tid >> 3, tx = tid & 7;
  const int i0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const float* q = a.q + b * a.bsq + h * T;
  const float* k = a.k + b * a.bsk + h * T;
  const float* v = a.v + b * a.bsv + h * T;

  load_tile<BQ>(Qs, q, a.ldq, i0, a.Nq, tid);

  float m[RT], l[RT], o[RT][8];
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[r][c] = 0.f;
  }

  for (int j0 = 0; j0 < a.Nk; j0 += T) {
    __syncthreads();  // previous tile fully consumed (also orders the Q load on the first iteration)
    load_tile(Ks, k, a.ldk, j0, a.Nk, tid);
    load_tile(Vs, v, a.ldv, j0, a.Nk, tid);
    if (tid < T) Ms[tid] = (a.key_mask != nullptr && j0 + tid < a.Nk) ? a.key_mask[b * a.Nk + j0 + tid] : 0.f;
    __syncthreads();

    float s[RT][8];
    qk_tile<RT>(Qs, Ks, ty, tx, s);
    finish_logits<RT>(s, a.scale, Ms, j0, a.Nk, tx, a.causal, i0 + ty);

#pragma unroll
    for (int r = 0; r < RT; ++r) {
      float mx = s[r][0];
#pragma unroll
      for (int c = 1; c < 8; ++c) mx = fmaxf(mx, s[r][c]);
      mx = group8_max(mx);
      const float m_new = fmaxf(m[r], mx);
      const float corr = (m[r] == -INFINITY) ? 0.f : expf(m[r] - m_new);
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float p = expf(s[r][c] - m_new);  // -inf logits give exactly 0
        ps += p;
        Ps[(ty + 16 * r) * LDS + tx + 8 * c] = p;
      }
      ps = group8_sum(ps);
      l[r] = l[r] * corr + ps;
      m[r] = m_new;
#pragma unroll
      for (int c = 0; c < 8; ++c) o[r][c] *= corr;
    }
    __syncthreads();

    // o[r][0..3] -> d = tx*4 + {0..3};  o[r][4..7] -> d = 32 + tx*4 + {0..3}
#pragma unroll 2
    for (int j4 = 0; j4 < T / 4; ++j4) {
      float4 pa[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) pa[r] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * r) * LDS + j4 * 4);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 v0 = *reinterpret_cast<const float4*>(Vs + (j4 * 4 + jj) * LDS + tx * 4);
        const float4 v1 = *reinterpret_cast<const float4*>(Vs + (j4 * 4 + jj) * LDS + 32 + tx * 4);
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const float p = jj == 0 ? pa[r].x : jj == 1 ? pa[r].y : jj == 2 ? pa[r].z : pa[r].w;
          o[r][0] = fmaf(p, v0.x, o[r][0]);
          o[r][1] = fmaf(p, v0.y, o[r][1]);
          o[r][2] = fmaf(p, v0.z, o[r][2]);
          o[r][3] = fmaf(p, v0.w, o[r][3]);
          o[r][4] = fmaf(p, v1.x, o[r][4]);
          o[r][5] = fmaf(p, v1.y, o[r][5]);
          o[r][6] = fmaf(p, v1.z, o[r][6]);
          o[r][7] = fmaf(p, v1.w, o[r][7]);
        }
      }
    }
  }

#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int i = i0 + ty + 16 * r;
    const float inv = 1.0f / l[r];
    float nsq = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      o[r][c] *= inv;
      nsq = fmaf(o[r][c], o[r][c], nsq);
    }
    nsq = group8_sum(nsq);
    if (i < a.Nq) {
      __half* dst = a.out_f16 + b * a.bso + i * a.ldo + h * T;
      __half2 h0 = __floats2half2_rn(o[r][0], o[r][1]), h1 = __floats2half2_rn(o[r][2], o[r][3]);
      __half2 h2 = __floats2half2_rn(o[r][4], o[r][5]), h3 = __floats2half2_rn(o[r][6], o[r][7]);
      uint2 p0, p1;
      p0.x = *reinterpret_cast<uint32_t*>(&h0);
      p0.y = *reinterpret_cast<uint32_t*>(&h1);
      p1.x = *reinterpret_cast<uint32_t*>(&h2);
      p1.y = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint2*>(dst + tx * 4) = p0;
      *reinterpret_cast<uint2*>(dst + 32 + tx * 4) = p1;
      if (tx == 0 && a.row_max != nullptr) {
        const long long sidx = (static_cast<long long>(b) * a.H + h) * a.Nq + i;
        a.row_max[sidx] = m[r];
        a.row_sum[sidx] = l[r];
        a.out_norm[sidx] = sqrtf(nsq);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pass 2: max over heads of the normalised probabilities, column-summed over non-CLS queries, plus the
// head-importance-weighted CLS row. grid = (ceil(N/64) key tiles, ceil(N/64) query tiles, B)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS)
attn_stats_kernel(AttnArgs a) {
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + T * LDS;
  float* Ms = Ks + T * LDS;    // [64]
  float* Hs = Ms + T;          // [64] 1 / (sum_h norm[b,h,j] + 1e-8)   (query tile 0 only)
  float* Red = Hs + T;         // [4][64] cross-warp column-sum staging

  const int tid = threadIdx.x, ty = tid >> 3, tx = tid & 7;
  const int j0 = blockIdx.x * T, it = blockIdx.y, i0 = it * T, b = blockIdx.z;
  const int N = a.Nq;
  const bool cls_tile = (it == 0);

  if (tid < T) {
    const int j = j0 + tid;
    Ms[tid] = (a.key_mask != nullptr && j < N) ? a.key_mask[b * N + j] : 0.f;
    if (cls_tile) {
      float hs = 0.f;
      if (j < N)
        for (int h = 0; h < a.H; ++h) hs += a.out_norm[(static_cast<long long>(b) * a.H + h) * N + j];
      Hs[tid] = hs + 1e-8f;
    }
  }

  float mx[4][8];
  float cacc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    cacc[c] = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) mx[r][c] = 0.f;  // probabilities are >= 0
  }

  for (int h = 0; h < a.H; ++h) {
    __syncthreads();
    load_tile(Qs, a.q + b * a.bsq + h * T, a.ldq, i0, N, tid);
    load_tile(Ks, a.k + b * a.bsk + h * T, a.ldk, j0, N, tid);
    __syncthreads();
    float s[4][8];
    qk_tile<4>(Qs, Ks, ty, tx, s);
    finish_logits<4>(s, a.scale, Ms, j0, N, tx, a.causal, i0 + ty);
    const long long sbase = (static_cast<long long>(b) * a.H + h) * N;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 16 * r;
      float rm = 0.f, rl = 1.f;
      if (i < N) {
        rm = a.row_max[sbase + i];
        rl = a.row_sum[sbase + i];
      }
      const float inv = 1.0f / rl;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float p = expf(s[r][c] - rm) * inv;
        mx[r][c] = fmaxf(mx[r][c], p);
        if (cls_tile && r == 0 && ty == 0) {  // query row 0 of this batch element
          const int jl = tx + 8 * c;
          const int j = j0 + jl;
          if (j < N) cacc[c] += p * (a.out_norm[sbase + j] / Hs[jl]);
        }
      }
    }
  }

  // column sums over the query rows of this tile, excluding the CLS query (i == 0) and rows past N
  float cs[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 16 * r;
      if (i >= 1 && i < N) t += mx[r][c];
    }
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    cs[c] = t;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  if (lane < 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) Red[warp * T + lane + 8 * c] = cs[c];
  }
  __syncthreads();
  if (tid < T) {
    const int j = j0 + tid;
    if (j < N) {
      const float t = (Red[tid] + Red[T + tid]) + (Red[2 * T + tid] + Red[3 * T + tid]);
      a.col_part[(static_cast<long long>(b) * gridDim.y + it) * N + j] = t;
    }
  }
  if (cls_tile && ty == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = j0 + tx + 8 * c;
      if (j < N) a.cls_attn[static_cast<long long>(b) * N + j] = cacc[c];
    }
  }
}

static int check_common(const AttnArgs& a) {
  MADTP_CHECK_ARG(a.q && a.k && a.v, "attention: null q/k/v");
  MADTP_CHECK_ARG(a.B >= 0 && a.H > 0 && a.Nq > 0 && a.Nk > 0, "attention: bad shape B=%d H=%d Nq=%d Nk=%d", a.B, a.H,
                  a.Nq, a.Nk);
  MADTP_CHECK_ARG(a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 && a.bsq % 4 == 0 && a.bsk % 4 == 0 &&
                      a.bsv % 4 == 0,
                  "attention: q/k/v strides must be multiples of 4 elements");
  MADTP_CHECK_ARG(a.B <= 65535 && a.H <= 65535, "attention: B and H must fit the grid y/z limits");
  return kOk;
}

int launch_attn_fwd(const AttnArgs& a, cudaStream_t stream) {
  int st = check_common(a);
  if (st != kOk) return st;
  MADTP_CHECK_ARG(a.out_f16 != nullptr && a.ldo % 4 == 0 && a.bso % 4 == 0, "attn_fwd: bad output");
  MADTP_CHECK_ARG((a.row_max == nullptr) == (a.row_sum == nullptr) && (a.row_max == nullptr) == (a.out_norm == nullptr),
                  "attn_fwd: row_max/row_sum/out_norm come together");
  if (a.B == 0) return kOk;
  const bool small_q = a.Nq <= 32 && a.row_max == nullptr;   // few queries: 32-row tiles waste far less
  MADTP_SMEM_ATTR_ONCE((4 * T * LDS + T) * (int)sizeof(float), attn_fwd_kernel<4>);
  MADTP_SMEM_ATTR_ONCE(((2 * 32 + 2 * T) * LDS + T) * (int)sizeof(float), attn_fwd_kernel<2>);
  if (small_q) {
    const int smem = ((2 * 32 + 2 * T) * LDS + T) * sizeof(float);
    dim3 grid((a.Nq + 31) / 32, a.H, a.B);
    attn_fwd_kernel<2><<<grid, NTHREADS, smem, stream>>>(a);
  } else {
    const int smem = (4 * T * LDS + T) * sizeof(float);
    dim3 grid((a.Nq + T - 1) / T, a.H, a.B);
    attn_fwd_kernel<4><<<grid, NTHREADS, smem, stream>>>(a);
  }
  MADTP_LAUNCH_CHECK();
  return kOk;
}

int launch_attn_stats(const AttnArgs& a, cudaStream_t stream) {
  int st = check_common(a);
  if (st != kOk) return st;
  MADTP_CHECK_ARG(a.Nq == a.Nk, "attn_stats: self-attention only (Nq == Nk)");
  MADTP_CHECK_ARG(a.row_max && a.row_sum && a.out_norm && a.col_part && a.cls_attn, "attn_stats: null statistics buffer");
  MADTP_CHECK_ARG((a.Nq + T - 1) / T <= 65535, "attn_stats: sequence too long");
  if (a.B == 0) return kOk;
  const int smem = (2 * T * LDS + 2 * T + 4 * T) * sizeof(float);
  dim3 grid((a.Nk + T - 1) / T, (a.Nq + T - 1) / T, a.B);
  attn_stats_kernel<<<grid, NTHREADS, smem, stream>>>(a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

}  // namespace madtp
