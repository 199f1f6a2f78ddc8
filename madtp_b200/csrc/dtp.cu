// Dynamic Token Pruning kernels: alignment statistics over the codebook, importance score + threshold + count,
// top-k selection, and the gather that compacts [B, N, d] -> [B, k+2, d] (survivors in ascending token order plus
// the merged token). Restates reference models/vit.py:123-163, models/nlvr_encoder.py:400-454,
// models/med.py:345-391 and models/utils.py:147-183 -- see oracle/dtp_oracle.py for the CPU statement of the same.
//
// All reductions run in a fixed order (no floating-point atomics). Sums that feed a comparison (normalisers, the
// softmax-weighted threshold, the tail weight sum) are accumulated in fp64 and rounded once, so the only
// discrepancy against the fp32 reference is the reference's own rounding.
#include <cooperative_groups.h>

#include "dtp.cuh"

namespace cg = cooperative_groups;

namespace madtp {

namespace {

template <typename Tv>
__device__ __forceinline__ Tv warp_sum_t(Tv v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum in fp64; every thread receives the result. `red` holds >= 32 doubles.
__device__ __forceinline__ double block_sum_d(double v, double* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  v = warp_sum_t(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}
__device__ __forceinline__ int block_sum_i(int v, int* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  v = warp_sum_i(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  int t = 0;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Column softmax statistics over tokens. grid = (ceil(T/32), B), block = 512: one CTA owns 32 codebook columns of one
// sequence. Its [n x 32] slab of token_att is read from global memory ONCE (lane = column, warp = row split, four
// independent row loads in flight per warp) into shared memory; the maximum and the exponential sum run from there.
// max_j (x_j / divisor) = (max_j x_j) / divisor because IEEE division by a positive number is monotone, so the
// maximum is taken on the raw values and divided once.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kStatThreads = 512, kStatWarps = kStatThreads / 32, kStatInflight = 4;
constexpr int kSlabMaxRows = 1024;   // 128 KB of dynamic shared memory at most

// Loads this CTA's 32 columns (t = 32 * group + lane, valid when `tok`) of rows 0..n-1 into slab[j * 32 + lane]
// (-inf where t >= T); returns the running column maximum over the rows this warp owns (j = warp mod 16).
// row_fn(j, v) sees every row (warp-uniform j, v = this lane's column).
//   vec: rows are 16-byte aligned and T is a multiple of 4 -> the whole slab is requested with 16-byte cp.async
//        copies before anything waits (one memory round trip); otherwise four scalar row loads in flight per warp.
template <typename RowFn>
__device__ __forceinline__ float load_slab(const float* __restrict__ ta, long long ld, int n, int T, int group,
                                           bool vec, float* __restrict__ slab, RowFn row_fn) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = group * 32 + lane;
  const bool tok = t < T;
  float cmx = -INFINITY;
  if (vec) {
    const int g4 = (threadIdx.x & 7) * 4;                 // eight threads per row, 64 rows per sweep
    const int tt = group * 32 + g4;
    for (int j = threadIdx.x >> 3; j < n; j += kStatThreads / 8) {
      float* dst = slab + j * 32 + g4;
      if (tt < T) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(ta + j * ld + tt) : "memory");
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int j = warp; j < n; j += kStatWarps) {
      const float v = slab[j * 32 + lane];
      cmx = fmaxf(cmx, v);
      row_fn(j, v);
    }
    return cmx;
  }
  for (int j0 = warp; j0 < n; j0 += kStatWarps * kStatInflight) {
    float v[kStatInflight];
#pragma unroll
    for (int u = 0; u < kStatInflight; ++u) {
      const int j = j0 + u * kStatWarps;
      v[u] = (tok && j < n) ? ta[j * ld + t] : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < kStatInflight; ++u) {
      const int j = j0 + u * kStatWarps;
      if (j < n) {
        slab[j * 32 + lane] = v[u];
        cmx = fmaxf(cmx, v[u]);
        row_fn(j, v[u]);
      }
    }
  }
  return cmx;
}

__host__ inline bool slab_vec_ok(const float* ta, long long ld, long long bs, int T) {
  return (reinterpret_cast<uintptr_t>(ta) % 16 == 0) && ld % 4 == 0 && bs % 4 == 0 && T % 4 == 0;
}
}  // namespace

__global__ void __launch_bounds__(kStatThreads)
token_colstats_kernel(const float* __restrict__ ta, long long ld, long long bs, int n, int T, float divisor,
                      float* __restrict__ col_max, float* __restrict__ col_sum, const int* __restrict__ n_dev,
                      int n_sub, int vec) {
  extern __shared__ __align__(16) float slab[];
  __shared__ float pmax[kStatWarps][32];
  __shared__ float psum[kStatWarps][32];
  if (n_dev != nullptr) {
    const int N = load_len(n_dev);
    n = min(n, N - n_sub);
    bs = N * ld;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 32 + lane, b = blockIdx.y;
  const bool ok = t < T;
  pmax[warp][lane] = load_slab(ta + b * bs, ld, n, T, blockIdx.x, vec != 0, slab, [](int, float) {});
  __syncthreads();
  float raw = pmax[0][lane];
#pragma unroll
  for (int w = 1; w < kStatWarps; ++w) raw = fmaxf(raw, pmax[w][lane]);
  const float gmx = __fdiv_rn(raw, divisor);
  float s = 0.f;
  if (ok)
    for (int j = warp; j < n; j += kStatWarps) s += expf(__fdiv_rn(slab[j * 32 + lane], divisor) - gmx);
  psum[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && ok) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kStatWarps; ++w) tot += psum[w][lane];   // fixed order
    col_max[b * T + t] = gmx;
    col_sum[b * T + t] = tot;
  }
}

// Sequences longer than the shared-memory slab: the same statistics streamed twice from global memory.
__global__ void __launch_bounds__(256)
token_colstats_stream_kernel(const float* __restrict__ ta, long long ld, long long bs, int n, int T, float divisor,
                             float* __restrict__ col_max, float* __restrict__ col_sum,
                             const int* __restrict__ n_dev, int n_sub) {
  __shared__ float smax[8][32];
  __shared__ float ssum[8][32];
  if (n_dev != nullptr) {
    const int N = load_len(n_dev);
    n = min(n, N - n_sub);
    bs = N * ld;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 32 + lane, b = blockIdx.y;
  const float* base = ta + b * bs;
  const bool ok = t < T;
  float mx = -INFINITY;
  for (int j = warp; j < n; j += 8) mx = fmaxf(mx, ok ? base[j * ld + t] : 0.f);
  smax[warp][lane] = mx;
  __syncthreads();
  float raw = smax[0][lane];
#pragma unroll
  for (int w = 1; w < 8; ++w) raw = fmaxf(raw, smax[w][lane]);
  const float gmx = __fdiv_rn(raw, divisor);
  float s = 0.f;
  for (int j = warp; j < n; j += 8) {
    const float x = ok ? __fdiv_rn(base[j * ld + t], divisor) : 0.f;
    s += expf(x - gmx);
  }
  ssum[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && ok) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += ssum[w][lane];
    col_max[b * T + t] = gmx;
    col_sum[b * T + t] = tot;
  }
}

int launch_token_colstats(const float* token_att, long long ld_ta, long long bs_ta, int B, int n, int T, float divisor,
                          float* col_max, float* col_sum, const int* n_dev, int n_sub, cudaStream_t stream) {
  MADTP_CHECK_ARG(token_att && col_max && col_sum, "token_colstats: null pointer");
  MADTP_CHECK_ARG(B >= 0 && n > 0 && T > 0 && divisor > 0.f && B <= 65535, "token_colstats: bad shape");
  if (B == 0) return kOk;
  dim3 grid((T + 31) / 32, B);
  if (n <= kSlabMaxRows) {
    MADTP_SMEM_ATTR_ONCE(kSlabMaxRows * 32 * 4, token_colstats_kernel);
    token_colstats_kernel<<<grid, kStatThreads, static_cast<size_t>(n) * 32 * 4, stream>>>(
        token_att, ld_ta, bs_ta, n, T, divisor, col_max, col_sum, n_dev, n_sub,
        slab_vec_ok(token_att, ld_ta, bs_ta, T) ? 1 : 0);
  } else {
    token_colstats_stream_kernel<<<grid, 256, 0, stream>>>(token_att, ld_ta, bs_ta, n, T, divisor, col_max, col_sum,
                                                           n_dev, n_sub);
  }
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// sd_ft[b,t,:] (+)= sum_j w[j,t] x[b,j,:],  w = softmax over tokens.  grid = (d/64, B), block = 256.
// Thread (ty = tid/16, tx = tid%16) owns t in [8*ty, 8*ty+8) and d in [4*tx, 4*tx+4) of a 128 x 64 output tile.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
query_sdft_kernel(const float* __restrict__ ta, long long ld_ta, long long bs_ta, const float* __restrict__ col_max,
                  const float* __restrict__ col_sum, const float* __restrict__ x, long long ldx, long long bsx, int n,
                  int T, int d, float divisor, float* __restrict__ out, int accumulate,
                  const int* __restrict__ n_dev, int n_sub) {
  if (n_dev != nullptr) {
    const int N = load_len(n_dev);
    n = min(n, N - n_sub);
    bs_ta = N * ld_ta;
    bsx = N * ldx;
  }
  __shared__ float Ws[32][128 + 4];
  __shared__ float Xs[32][64];
  __shared__ float cmx[128], cinv[128];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int d0 = blockIdx.x * 64, b = blockIdx.y;
  if (tid < 128) {
    const bool ok = tid < T;
    cmx[tid] = ok ? col_max[b * T + tid] : 0.f;
    cinv[tid] = ok ? 1.0f / col_sum[b * T + tid] : 0.f;
  }
  float acc[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  const float* tab = ta + b * bs_ta;
  const float* xb = x + b * bsx;
  for (int j0 = 0; j0 < n; j0 += 32) {
    __syncthreads();
    for (int i = tid; i < 32 * 128; i += 256) {
      const int jj = i >> 7, t = i & 127;
      float w = 0.f;
      if (j0 + jj < n && t < T) w = expf(__fdiv_rn(tab[(j0 + jj) * ld_ta + t], divisor) - cmx[t]) * cinv[t];
      Ws[jj][t] = w;
    }
    for (int i = tid; i < 32 * 16; i += 256) {
      const int jj = i >> 4, c4 = i & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + jj < n && d0 + c4 * 4 < d) v = *reinterpret_cast<const float4*>(xb + (j0 + jj) * ldx + d0 + c4 * 4);
      *reinterpret_cast<float4*>(&Xs[jj][c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < 32; ++jj) {
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[jj][ty * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[jj][ty * 8 + 4]);
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[jj][tx * 4]);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        acc[r][0] = fmaf(w[r], xv.x, acc[r][0]);
        acc[r][1] = fmaf(w[r], xv.y, acc[r][1]);
        acc[r][2] = fmaf(w[r], xv.z, acc[r][2]);
        acc[r][3] = fmaf(w[r], xv.w, acc[r][3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = ty * 8 + r;
    if (t < T && d0 + tx * 4 < d) {
      float4* o = reinterpret_cast<float4*>(out + (static_cast<long long>(b) * T + t) * d + d0 + tx * 4);
      float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      if (accumulate) {
        const float4 p = *o;
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
      *o = v;
    }
  }
}

int launch_query_sdft(const float* token_att, long long ld_ta, long long bs_ta, const float* col_max,
                      const float* col_sum, const float* x, long long ldx, long long bsx, int B, int n, int T, int d,
                      float divisor, float* sd_ft, int accumulate, const int* n_dev, int n_sub, cudaStream_t stream) {
  MADTP_CHECK_ARG(token_att && col_max && col_sum && x && sd_ft, "query_sdft: null pointer");
  MADTP_CHECK_ARG(B >= 0 && n > 0 && T > 0 && T <= 128 && d % 4 == 0 && ldx % 4 == 0 && bsx % 4 == 0 && B <= 65535,
                  "query_sdft: unsupported shape (T=%d must be <= 128, d=%d multiple of 4)", T, d);
  if (B == 0) return kOk;
  dim3 grid((d + 63) / 64, B);
  query_sdft_kernel<<<grid, 256, 0, stream>>>(token_att, ld_ta, bs_ta, col_max, col_sum, x, ldx, bsx, n, T, d, divisor,
                                              sd_ft, accumulate, n_dev, n_sub);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Importance score, threshold and survivor count (reference models/vit.py:126-145).
// grid = (4, B) in clusters of four CTAs, block = 512: the four CTAs of a sequence own 32 codebook columns each (cluster
// rank = column group). Every CTA stages its [n x 32] slab of token_att in shared memory with ONE pass over global
// memory; the row maxima (over all T columns) are exchanged through distributed shared memory, the cheap per-token
// work (a, b, score) is computed redundantly by the four CTAs, the threshold sum -- the only O(n T) exponential pass
// -- is split by columns and its per-CTA minima are exchanged the same way. Rank 0 writes the results.
// ------------------------------------------------------------------------------------------------
constexpr int kScoreCluster = 4;

__global__ void __cluster_dims__(kScoreCluster, 1, 1) __launch_bounds__(kStatThreads, 1)
dtp_score_kernel(DtpScoreArgs a, int vec) {
  extern __shared__ __align__(16) float slab[];            // [n][32]: columns 32*rank .. 32*rank+31 of every token
  __shared__ float S[kDtpMaxTokens];                       // a_j, then the score
  __shared__ float Rpart[kDtpMaxTokens];                   // max over THIS CTA's columns of token j (peers read it)
  __shared__ float Bm[kDtpMaxTokens];                      // b_j = max over all columns
  __shared__ double red[32];
  __shared__ int redi[32];
  __shared__ float pmax[kStatWarps][32];
  __shared__ double pnum[kStatWarps][32], pden[kStatWarps][32];
  __shared__ float thr_part[kScoreCluster];

  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = blockIdx.x;                             // == cluster.block_rank(): gridDim.x is the cluster width
  if (a.n_dev != nullptr) {          // device-resident token count: packed sequences
    const int Nd = min(a.n + 1, load_len(a.n_dev));
    a.n = Nd - 1;
    a.bs_ta = Nd * a.ld_ta;
    if (a.parts_tile > 0) a.n_parts = (Nd + a.parts_tile - 1) / a.parts_tile;
  }
  const int b = blockIdx.y, n = a.n, N = n + 1;
  const int t = rank * 32 + lane;
  const bool tok = t < a.T;

  // (B, first half) slab + this column group's share of b_j = max_t token_att[j,t]
  pmax[warp][lane] = load_slab(a.token_att + b * a.bs_ta, a.ld_ta, n, a.T, rank, vec != 0, slab, [&](int j, float v) {
    const float m = warp_max(v);
    if (lane == 0) Rpart[j] = m;
  });

  // (A) self-attention statistic: a_j = sum over query tiles (fixed order)
  double part = 0.0;
  for (int j = tid; j < n; j += kStatThreads) {
    float s = 0.f;
    for (int p = 0; p < a.n_parts; ++p) s += a.col_part[(static_cast<long long>(b) * a.n_parts + p) * N + 1 + j];
    S[j] = s;
    part += static_cast<double>(s);
  }
  const float a_den = static_cast<float>(block_sum_d(part, red)) + 1e-8f;

  cluster.sync();                    // all four CTAs run and their partial row maxima are complete

  // (B, second half): combine the four column groups through distributed shared memory
  const float* r0 = cluster.map_shared_rank(Rpart, 0);
  const float* r1 = cluster.map_shared_rank(Rpart, 1);
  const float* r2 = cluster.map_shared_rank(Rpart, 2);
  const float* r3 = cluster.map_shared_rank(Rpart, 3);
  part = 0.0;
  for (int j = tid; j < n; j += kStatThreads) {
    const float m = fmaxf(fmaxf(r0[j], r1[j]), fmaxf(r2[j], r3[j]));
    Bm[j] = m;                       // the same thread reads it back in (C)
    part += static_cast<double>(m);
  }
  const float b_den = static_cast<float>(block_sum_d(part, red)) + 1e-8f;

  // (C) Importance_score = (a' + b' + cls_attn) / 3
  for (int j = tid; j < n; j += kStatThreads) {
    const float av = __fdiv_rn(S[j], a_den);
    const float bv = __fdiv_rn(Bm[j], b_den);
    const float cv = a.cls_attn[static_cast<long long>(b) * N + 1 + j];
    const float sc = __fdiv_rn((av + bv) + cv, 3.0f);
    S[j] = sc;
    if (rank == 0) a.score[static_cast<long long>(b) * n + j] = sc;
  }
  __syncthreads();

  // (D) threshold = min_t  sum_j softmax_j(token_att[j,t] / temperature) * score_j   over this CTA's 32 columns
  float raw = pmax[0][lane];
#pragma unroll
  for (int w = 1; w < kStatWarps; ++w) raw = fmaxf(raw, pmax[w][lane]);
  const float gmx = __fdiv_rn(raw, a.temperature);
  double num = 0.0, den = 0.0;
  if (tok)
    for (int j = warp; j < n; j += kStatWarps) {
      const float e = expf(__fdiv_rn(slab[j * 32 + lane], a.temperature) - gmx);
      den += static_cast<double>(e);
      num += static_cast<double>(e) * static_cast<double>(S[j]);
    }
  pnum[warp][lane] = num;
  pden[warp][lane] = den;
  __syncthreads();
  if (warp == 0) {
    float v = INFINITY;
    if (tok) {
      double nn = 0.0, dd = 0.0;
#pragma unroll
      for (int q = 0; q < kStatWarps; ++q) {   // fixed order
        nn += pnum[q][lane];
        dd += pden[q][lane];
      }
      v = static_cast<float>(nn / dd);
    }
    v = warp_min(v);
    if (lane < kScoreCluster) *cluster.map_shared_rank(&thr_part[rank], lane) = v;
  }
  cluster.sync();                    // the last remote access precedes this barrier: CTAs may exit independently after it
  if (rank != 0) return;
  const float thr = fminf(fminf(thr_part[0], thr_part[1]), fminf(thr_part[2], thr_part[3]));

  // (E) count
  int c = 0;
  for (int j = tid; j < n; j += kStatThreads) c += (S[j] > thr) ? 1 : 0;
  c = block_sum_i(c, redi);
  if (tid == 0) {
    a.threshold[b] = thr;
    a.count[b] = c;
    atomicMax(a.topk, c);
  }
}

int launch_dtp_score(const DtpScoreArgs& a, cudaStream_t stream) {
  if (a.B == 0) return kOk;   // empty batch: nothing to do (and torch hands out null pointers for empty tensors)
  MADTP_CHECK_ARG(a.col_part && a.cls_attn && a.token_att && a.score && a.threshold && a.count && a.topk,
                  "dtp_score: null pointer");
  MADTP_CHECK_ARG(a.B >= 0 && a.B <= 65535 && a.n > 0 && a.n <= kDtpMaxTokens,
                  "dtp_score: n=%d out of range (1..%d)", a.n, kDtpMaxTokens);
  MADTP_CHECK_ARG(a.T > 0 && a.T <= 128, "dtp_score: codebook size T=%d must be in 1..128", a.T);
  MADTP_CHECK_ARG(a.temperature > 0.f, "dtp_score: temperature must be > 0");
  MADTP_CHECK_ARG(a.n_parts > 0, "dtp_score: n_parts must be > 0");
  static_assert(kDtpMaxTokens <= kSlabMaxRows, "the score kernel stages every token of the sequence");
  MADTP_SMEM_ATTR_ONCE(kSlabMaxRows * 32 * 4, dtp_score_kernel);
  dtp_score_kernel<<<dim3(kScoreCluster, a.B), kStatThreads, static_cast<size_t>(a.n) * 32 * 4, stream>>>(
      a, slab_vec_ok(a.token_att, a.ld_ta, a.bs_ta, a.T) ? 1 : 0);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Selection: exact top-k by rank (ties -> lower token index first), survivor slots in ascending token order,
// merge weights of the pruned tail, optional mask bookkeeping. grid = B, block = 256.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dtp_select_kernel(DtpSelectArgs a) {
  __shared__ float S[kDtpMaxTokens];
  __shared__ int R[kDtpMaxTokens];         // rank in descending score order
  __shared__ int wsum[8];
  __shared__ double red[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool dyn = a.n_dev != nullptr;
  const int b = blockIdx.x, n = dyn ? min(a.n, load_len(a.n_dev) - 1) : a.n;
  const int k_in = *a.topk;
  const bool identity = (k_in <= a.max_keep) || (n - k_in <= 1);   // reference early-out: nothing is pruned
  const int k = identity ? n : k_in;
  // packed mask rows: the input holds n + 1 entries per sequence; the output k + 2 in dynamic mode (the next layer's
  // packed length), n + 1 (worst case, narrowed by the host) otherwise
  const int mo_pitch = dyn ? (identity ? n + 1 : k + 2) : a.n + 1;
  if (dyn && b == 0 && tid == 0) {
    if (a.n_out) *a.n_out = identity ? n + 1 : k + 2;
    if (a.k_out) *a.k_out = identity ? -1 : k;
  }

  for (int j = tid; j < n; j += 256) S[j] = a.score[static_cast<long long>(b) * n + j];
  __syncthreads();
  for (int j = tid; j < n; j += 256) {
    const float sj = S[j];
    int r = 0;
    for (int i = 0; i < n; ++i) {
      const float si = S[i];
      r += (si > sj || (si == sj && i < j)) ? 1 : 0;
    }
    R[j] = r;
  }
  __syncthreads();

  // exclusive scan of keep flags in token order: thread t owns tokens [4t, 4t+4)
  int flags[4], local = 0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = tid * 4 + u;
    flags[u] = (j < n && R[j] < k) ? 1 : 0;
    local += flags[u];
  }
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int base = incl - local;
  for (int w = 0; w < warp; ++w) base += wsum[w];

  double tail = 0.0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = tid * 4 + u;
    if (j < n && !flags[u]) tail += static_cast<double>(S[j]);
  }
  // block sum of the tail (fp64), every thread gets it
  double tsum;
  {
    double v = warp_sum_t(tail);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    tsum = 0.0;
    for (int w = 0; w < 8; ++w) tsum += red[w];
  }
  const float den = static_cast<float>(tsum) + 1e-8f;

  int run = base;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = tid * 4 + u;
    if (j < n) {
      const long long o = static_cast<long long>(b) * n + j;
      a.keep[o] = static_cast<unsigned char>(flags[u]);
      a.dst[o] = flags[u] ? run : -1;
      a.tail_w[o] = flags[u] ? 0.f : __fdiv_rn(S[j], den);
      if (!flags[u]) a.tail_idx[static_cast<long long>(b) * n + (j - run)] = j;  // pruned tokens, ascending
      if (a.mask_mode != 0 && !identity) {
        const float mj = a.mask_in[static_cast<long long>(b) * (n + 1) + 1 + j];
        float* mo = a.mask_out + static_cast<long long>(b) * mo_pitch;
        if (a.mask_mode == 1) {
          if (R[j] <= k) mo[1 + R[j]] = mj;           // slot r <- mask of the r-th ranked token, r = 0..k
        } else {
          if (flags[u]) mo[1 + run] = mj;             // mask travels with its token
          if (R[j] == k) mo[1 + k] = mj;              // merged slot <- mask of the (k+1)-th ranked token
        }
      }
      run += flags[u];
    }
  }
  if (a.mask_mode != 0) {
    float* mo = a.mask_out + static_cast<long long>(b) * mo_pitch;
    const float* mi = a.mask_in + static_cast<long long>(b) * (n + 1);
    if (identity) {
      for (int j = tid; j < n + 1; j += 256) mo[j] = mi[j];
    } else if (tid == 0) {
      mo[0] = mi[0];
    }
  }
}

int launch_dtp_select(const DtpSelectArgs& a, cudaStream_t stream) {
  if (a.B == 0) return kOk;   // empty batch: nothing to do (and torch hands out null pointers for empty tensors)
  MADTP_CHECK_ARG(a.score && a.topk && a.keep && a.dst && a.tail_w && a.tail_idx, "dtp_select: null pointer");
  MADTP_CHECK_ARG(a.B >= 0 && a.n > 0 && a.n <= kDtpMaxTokens, "dtp_select: n=%d out of range (1..%d)", a.n,
                  kDtpMaxTokens);
  MADTP_CHECK_ARG(a.mask_mode >= 0 && a.mask_mode <= 2, "dtp_select: mask_mode must be 0, 1 or 2");
  MADTP_CHECK_ARG(a.mask_mode == 0 || (a.mask_in && a.mask_out), "dtp_select: mask buffers missing");
  if (a.B == 0) return kOk;
  dtp_select_kernel<<<a.B, 256, 0, stream>>>(a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Gather + merge. grid = (1 + slabs, B), block = 256.
//   blockIdx.x == 0 : merged token  out[b, 1+k, :] = sum_{j pruned} tail_w[j] * x[b, 1+j, :]   (ascending j)
//   blockIdx.x >= 1 : one warp per source row, 128-bit copies of survivors into their slots
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dtp_gather_kernel(DtpGatherArgs a) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool dyn = a.n_dev != nullptr;
  const int b = blockIdx.y, n = dyn ? min(a.n, load_len(a.n_dev) - 1) : a.n, d4 = a.d >> 2;
  const int k_in = *a.topk;
  const bool identity = (k_in <= a.max_keep) || (n - k_in <= 1);
  const int k = identity ? n : k_in;
  if (dyn) {                       // packed input [B, n + 1, d] and packed output [B, N_out, d]
    a.bsx = static_cast<long long>(n + 1) * a.d;
    a.bso = static_cast<long long>(identity ? n + 1 : k + 2) * a.d;
  }
  const float4* xb = reinterpret_cast<const float4*>(a.x + b * a.bsx);
  float4* ob = reinterpret_cast<float4*>(a.out + b * a.bso);
  uint2* ob16 = a.out_f16 ? reinterpret_cast<uint2*>(a.out_f16 + b * a.bso) : nullptr;
  auto pack4 = [](const float4& v) {
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    uint2 p;
    p.x = *reinterpret_cast<const uint32_t*>(&h0);
    p.y = *reinterpret_cast<const uint32_t*>(&h1);
    return p;
  };
  const int* dst = a.dst + static_cast<long long>(b) * n;
  const float* tw = a.tail_w + static_cast<long long>(b) * n;

  if (blockIdx.x == 0) {
    if (identity) return;
    // Each warp reduces a strided subset of the pruned rows over the whole feature dimension (lane owns float4
    // columns lane, lane+32, ...), then the eight partials are combined in warp order.
    extern __shared__ float4 part[];  // [8][d4]
    const int ntail = n - k;
    const int* tidx = a.tail_idx + static_cast<long long>(b) * n;
    constexpr int MAXC = 8;           // d <= 1024
    float4 acc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = warp; i < ntail; i += 8) {
      const int j = tidx[i];
      const float w = tw[j];
      const float4* src = xb + static_cast<long long>(1 + j) * d4;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int cc = lane + 32 * c;
        if (cc < d4) {
          const float4 v = src[cc];
          acc[c].x = fmaf(w, v.x, acc[c].x);
          acc[c].y = fmaf(w, v.y, acc[c].y);
          acc[c].z = fmaf(w, v.z, acc[c].z);
          acc[c].w = fmaf(w, v.w, acc[c].w);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int cc = lane + 32 * c;
      if (cc < d4) part[warp * d4 + cc] = acc[c];
    }
    __syncthreads();
    for (int c = tid; c < d4; c += 256) {
      float4 t = part[c];
      for (int w = 1; w < 8; ++w) {
        const float4 v = part[w * d4 + c];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      ob[static_cast<long long>(1 + k) * d4 + c] = t;
      if (ob16) ob16[static_cast<long long>(1 + k) * d4 + c] = pack4(t);
    }
    return;
  }
  const int slabs = gridDim.x - 1;
  const int rows = n + 1;
  const int per = (rows + slabs - 1) / slabs;
  const int r0 = (blockIdx.x - 1) * per;
  const int r1 = min(rows, r0 + per);
  for (int r = r0 + warp; r < r1; r += 8) {
    int slot;
    if (r == 0) slot = 0;
    else {
      const int s = dst[r - 1];
      if (s < 0) continue;
      slot = 1 + s;
    }
    const float4* src = xb + static_cast<long long>(r) * d4;
    float4* o = ob + static_cast<long long>(slot) * d4;
    // all loads of the row in flight before the first store (d <= 1024: at most eight float4 per lane)
    float4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (lane + 32 * c < d4) v[c] = src[lane + 32 * c];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (lane + 32 * c < d4) {
        o[lane + 32 * c] = v[c];
        if (ob16) ob16[static_cast<long long>(slot) * d4 + lane + 32 * c] = pack4(v[c]);
      }
  }
}

int launch_dtp_gather(const DtpGatherArgs& a, cudaStream_t stream) {
  if (a.B == 0) return kOk;   // empty batch: nothing to do (and torch hands out null pointers for empty tensors)
  MADTP_CHECK_ARG(a.x && a.topk && a.dst && a.tail_w && a.tail_idx && a.out, "dtp_gather: null pointer");
  MADTP_CHECK_ARG(a.B >= 0 && a.n > 0 && a.d > 0 && a.d % 4 == 0 && a.d <= 1024 && a.bsx % 4 == 0 && a.bso % 4 == 0 && a.B <= 65535,
                  "dtp_gather: bad shape");
  if (a.B == 0) return kOk;
  int slabs = (4 * num_sms() + a.B - 1) / a.B;   // ~4 CTAs per SM in total
  if (slabs < 1) slabs = 1;
  if (slabs > (a.n + 8) / 8) slabs = (a.n + 8) / 8;
  dim3 grid(1 + slabs, a.B);
  dtp_gather_kernel<<<grid, 256, 8 * a.d * sizeof(float), stream>>>(a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// vector_gather: one warp per output row, 128-bit copies. grid = ceil(B*K/8), block = 256.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ x, long long bsx, const int* __restrict__ idx, float* __restrict__ out,
                   int B, int L, int K, int d4) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<long long>(B) * K) return;
  const int b = static_cast<int>(row / K);
  int j = idx[row];
  j = j < 0 ? 0 : (j >= L ? L - 1 : j);
  const float4* src = reinterpret_cast<const float4*>(x + b * bsx) + static_cast<long long>(j) * d4;
  float4* dst = reinterpret_cast<float4*>(out) + row * d4;
  for (int c = lane; c < d4; c += 32) dst[c] = src[c];
}

int launch_gather_rows(const float* x, long long bsx, const int* idx, float* out, int B, int L, int K, int d,
                       cudaStream_t stream) {
  MADTP_CHECK_ARG(x && idx && out, "gather_rows: null pointer");
  MADTP_CHECK_ARG(B >= 0 && L > 0 && K >= 0 && d > 0 && d % 4 == 0 && bsx % 4 == 0, "gather_rows: bad shape");
  const long long rows = static_cast<long long>(B) * K;
  if (rows == 0) return kOk;
  gather_rows_kernel<<<static_cast<int>((rows + 7) / 8), 256, 0, stream>>>(x, bsx, idx, out, B, L, K, d / 4);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

}  // namespace madtp
