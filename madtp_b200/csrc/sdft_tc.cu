// Query_model's aggregated feature on the tensor cores (reference models/utils.py:174-178):
//   sd_ft[b, t, :] (+)= sum_j softmax_j(token_att[b, j, t] / divisor) * x[b, j, :]
// i.e. out[T, d] = W^T . X with the contraction over TOKENS, the slow index of both operands (token_att is
// [tokens, T], x is [tokens, d], both row-major). Both operands are re-laid out K-major in shared memory -- nothing
// transposed ever goes through HBM:
//   * X chunks [64 tokens x 128 dims] arrive by TMA (fp32, 128-byte swizzle); eight warps read them column-wise
//     (conflict-free under the swizzle), split into fp16 hi/lo and write [128 dims x 64 tokens] K-major tiles;
//   * W chunks are computed on the fly by the same warps (2^(y - max) / sum in the log2 domain, scaled by 2^10 so
//     that the lo plane of a typical weight ~1/n stays a normal fp16 number) and written as [128 entries x 64 tokens]
//     K-major hi/lo tiles.
// Error-compensated fp16 (3 MMAs per product), accumulated in TMEM over the whole token range -- sd_ft is a model
// output, not part of the scoring lane, so the ~1e-6 accumulation / ex2.approx error is irrelevant.
//
//   grid = (d / 128, B); warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = operand builders, 2..5 epilogue.
#include "dtp.cuh"
#include "gemm.cuh"

namespace madtp {

namespace {
constexpr int TCH = 64;                     // tokens per pipeline stage (one 128-byte K row of halves, 4 k-steps of 16)
constexpr int TM = 128;                     // codebook entries per CTA (T <= 128, zero padded)
constexpr int DN = 128;                     // feature columns per CTA
constexpr int RAW_BYTES = (DN / 32) * TCH * 128;   // X as loaded: 4 boxes of [64 tokens x 32 floats]
constexpr int OP_BYTES = 128 * 128;                // one K-major operand plane: [128 rows x 64 tokens] halves
constexpr int STAGE = RAW_BYTES + 4 * OP_BYTES;    // raw X | W hi | W lo | X^T hi | X^T lo  = 96 KB
constexpr float kWScale = 1024.0f;
constexpr int STAGES = 2;
constexpr int SMEM_TOTAL = STAGES * STAGE + 256;

// byte offset of 16-byte granule k4 (0..7: eight tokens) of a row in a K-major [rows x 128 bytes] tile with the
// 128-byte swizzle
__device__ __forceinline__ int kmajor_off(int row, int k4) {
  return (row >> 3) * 1024 + (row & 7) * 128 + ((k4 ^ (row & 7)) << 4);
}
}  // namespace

struct SdftArgs {
  const float* ta; long long ld_ta, bs_ta;   // token_att row j of batch b at ta + b*bs_ta + j*ld_ta
  const float* col_max; const float* col_sum;
  int n, T, d, row_stride, first_row;        // x row of token j of batch b = b*row_stride + first_row + j
  float divisor;
  float* out; int accumulate;
  const int* n_dev;                          // dynamic tokens per sequence N: row_stride = N, n = N - first_row
};

__global__ void __launch_bounds__(320, 1)
query_sdft_tc_kernel(const __grid_constant__ CUtensorMap tm_x, SdftArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* raw_full = bars;       // [2] TMA landed
  uint64_t* op_full = bars + 2;    // [2] operand tiles built (4 warps)
  uint64_t* empty = bars + 4;      // [2] MMAs of the stage retired
  uint64_t* acc_full = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = blockIdx.x * DN, b = blockIdx.y;
  if (a.n_dev != nullptr) {                  // device-resident token count: packed sequences
    const int Nd = load_len(a.n_dev);
    a.n = min(a.n, Nd - a.first_row);
    a.row_stride = Nd;
    a.bs_ta = Nd * a.ld_ta;
  }
  const int chunks = (a.n + TCH - 1) / TCH;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_x);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&op_full[s], 8);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<DN>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {   // warp-uniform control flow, one elected lane issues
    for (int c = 0; c < chunks; ++c) {
      const int st = c & 1;
      mbar_wait(&empty[st], ((c >> 1) & 1) ^ 1);   // chunk c-2 fully consumed (its builders finished long before)
      uint8_t* xs = smem + st * STAGE;
      const int row = b * a.row_stride + a.first_row + c * TCH;
      if (elect_one()) {
        mbar_arrive_expect_tx(&raw_full[st], RAW_BYTES);
        for (int g = 0; g < DN / 32; ++g) tma_load_2d(&tm_x, &raw_full[st], xs + g * TCH * 128, d0 + g * 32, row);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(0u, TM, DN);
    for (int c = 0; c < chunks; ++c) {
      const int st = c & 1;
      mbar_wait(&op_full[st], (c >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t base = smem_u32(smem + st * STAGE + RAW_BYTES);
      const uint64_t w_hi = make_sw128_kmajor_desc(base), w_lo = make_sw128_kmajor_desc(base + OP_BYTES);
      const uint64_t x_hi = make_sw128_kmajor_desc(base + 2 * OP_BYTES);
      const uint64_t x_lo = make_sw128_kmajor_desc(base + 3 * OP_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, w_lo + 2 * k, x_hi + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, w_hi + 2 * k, x_lo + 2 * k, idesc, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, w_hi + 2 * k, x_hi + 2 * k, idesc, 1u);
        umma_commit(&empty[st]);
        if (c == chunks - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // eight builder warps: thread (t, half) owns row t of both K-major operand tiles and 32 of the chunk's 64 tokens
    const int t = (threadIdx.x - 64) & 127;    // codebook entry (row of W^T) and feature column (row of X^T)
    const int half = (threadIdx.x - 64) >> 7;
    const bool t_ok = t < a.T;
    constexpr float kLog2e = 1.4426950408889634f;
    const float c1 = kLog2e / a.divisor;
    const float off = t_ok ? -a.col_max[b * a.T + t] * kLog2e : 0.f;
    const float cinv = t_ok ? kWScale / a.col_sum[b * a.T + t] : 0.f;
    const float* tab = a.ta + b * a.bs_ta + t;
    const int grp = t >> 5, tl = t & 31;
    auto store_split8 = [](uint8_t* hi, uint8_t* lo, int offb, const float (&v)[8]) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn(v[2 * e] - f.x, v[2 * e + 1] - f.y);
        ph[e] = *reinterpret_cast<const uint32_t*>(&h);
        pl[e] = *reinterpret_cast<const uint32_t*>(&l);
      }
      *reinterpret_cast<uint4*>(hi + offb) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      *reinterpret_cast<uint4*>(lo + offb) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    };
    for (int c = 0; c < chunks; ++c) {
      const int st = c & 1;
      uint8_t* stage = smem + st * STAGE;
      uint8_t* w_hi = stage + RAW_BYTES;
      uint8_t* w_lo = w_hi + OP_BYTES;
      uint8_t* x_hi = w_lo + OP_BYTES;
      uint8_t* x_lo = x_hi + OP_BYTES;
      // token_att values of this thread's 32 tokens: all loads in flight before anything depends on them
      float tv[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const int j = c * TCH + half * 32 + u;
        tv[u] = (t_ok && j < a.n) ? __ldg(tab + j * a.ld_ta) : -INFINITY;
      }
      mbar_wait(&empty[st], ((c >> 1) & 1) ^ 1);  // operand tiles of chunk c-2 no longer read by the tensor core
      // ---- W^T row t: 1024 * softmax weight of each token (2^-inf = 0 past the sequence end / codebook size) ----
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = ex2_approx(fmaf(tv[q * 8 + u], c1, off)) * cinv;
        store_split8(w_hi, w_lo, kmajor_off(t, half * 4 + q), w);
      }
      // ---- X^T row (feature d0 + t): gather the column of the raw [64 tokens x 128 dims] chunk ----
      mbar_wait(&raw_full[st], (c >> 1) & 1);
      const uint8_t* raw = stage + grp * TCH * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = half * 32 + q * 8 + u;   // token row within the chunk
          x[u] = *reinterpret_cast<const float*>(raw + r * 128 + (((tl >> 2) ^ (r & 7)) << 4) + ((tl & 3) << 2));
        }
        store_split8(x_hi, x_lo, kmajor_off(t, half * 4 + q), x);
      }
      fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&op_full[st]);
    }
    if (half == 0) {
    // epilogue: thread = output row t (TMEM lane), 128 columns
    mbar_wait(acc_full, 0);
    tcgen05_fence_after();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;          // TMEM lane = accumulator row = codebook entry
    float* orow = a.out + (static_cast<long long>(b) * a.T + row) * a.d + d0;
#pragma unroll 1
    for (int cc = 0; cc < DN / 32; ++cc) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + cc * 32, v);
      tmem_ld_wait();
      if (row < a.T && d0 + cc * 32 < a.d) {
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          constexpr float kUn = 1.0f / kWScale;
          float4 o = make_float4(__uint_as_float(v[k]) * kUn, __uint_as_float(v[k + 1]) * kUn,
                                 __uint_as_float(v[k + 2]) * kUn, __uint_as_float(v[k + 3]) * kUn);
          float4* p = reinterpret_cast<float4*>(orow + cc * 32 + k);
          if (a.accumulate) {
            const float4 q = *p;
            o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
          }
          *p = o;
        }
      }
    }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<DN>(tmem_base);
  }
}

// x: fp32 rows [x_rows, d] (dense) holding token j of batch b at row b*row_stride + first_row + j.
int launch_query_sdft_tc(const float* token_att, long long ld_ta, long long bs_ta, const float* col_max,
                         const float* col_sum, const float* x, long long x_rows, int row_stride, int first_row, int B,
                         int n, int T, int d, float divisor, float* sd_ft, int accumulate, const int* n_dev,
                         cudaStream_t stream) {
  MADTP_CHECK_ARG(token_att && col_max && col_sum && x && sd_ft, "query_sdft_tc: null pointer");
  MADTP_CHECK_ARG(B >= 0 && n > 0 && T > 0 && T <= TM && d > 0 && d % 32 == 0 && B <= 65535,
                  "query_sdft_tc: unsupported shape (T=%d must be <= 128, d=%d a multiple of 32)", T, d);
  if (B == 0) return kOk;
  CUtensorMap tx;
  int st;
  if ((st = make_tmap(&tx, x, true, x_rows, d, d, TCH)) != kOk) return st;
  MADTP_SMEM_ATTR_ONCE(SMEM_TOTAL, query_sdft_tc_kernel);
  SdftArgs a;
  a.ta = token_att; a.ld_ta = ld_ta; a.bs_ta = bs_ta;
  a.col_max = col_max; a.col_sum = col_sum;
  a.n = n; a.T = T; a.d = d; a.row_stride = row_stride; a.first_row = first_row;
  a.divisor = divisor; a.out = sd_ft; a.accumulate = accumulate;
  a.n_dev = n_dev;
  dim3 grid((d + DN - 1) / DN, B);
  query_sdft_tc_kernel<<<grid, 320, SMEM_TOTAL, stream>>>(tx, a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// The same contraction from the fp16 hi/lo planes of x that the LayerNorm kernel already wrote (ViT / CLIP blocks):
// with MN-MAJOR shared-memory descriptors NOTHING is transposed at all -- the contraction index (tokens) is the row
// index of both operands as they sit in HBM, and the tensor core reads such tiles directly:
//   * X tiles [32 tokens x 64 dims] of both planes arrive by TMA (one 3-D copy per plane and stage: six column
//     groups, each a [32 x 128 bytes] tile with the 128-byte swizzle) and ARE the B operand
//     (MN-major: 128-byte rows along N = dims, 8-token swizzle atoms 1024 bytes apart, 64-dim groups one box apart);
//   * W tiles [32 tokens x 128 entries] are computed by eight builder warps straight in that layout (A operand,
//     MN-major along M = codebook entries): thread = (token, 16 entries), row-contiguous token_att reads.
// One CTA owns 384 feature columns of one sequence (384 TMEM columns; 2 CTAs per sequence at d = 768, one wave), so
// the softmax weights are rebuilt twice per sequence instead of six times and no builder touches X.
//   grid = (ceil(d / 384), B); warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = W builders + epilogue.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int MN_TK = 32;                         // tokens per pipeline stage (two k-steps of 16)
constexpr int MN_DN = 384;                        // feature columns per CTA
constexpr int MN_BOX = MN_TK * 128;               // one [32 tokens x 64 halves] box = four 8-token swizzle atoms
constexpr int MN_X_BYTES = (MN_DN / 64) * MN_BOX; // one plane of X per stage (24 KB)
constexpr int MN_W_BYTES = 2 * MN_BOX;            // one plane of W per stage: entries 0..63 | 64..127 (8 KB)
constexpr int MN_STAGE = 2 * MN_X_BYTES + 2 * MN_W_BYTES;   // X hi | X lo | W hi | W lo = 64 KB
constexpr int MN_STAGES = 3;
constexpr int MN_SMEM = MN_STAGES * MN_STAGE + 1024 + 256;  // operand stages | off[128], cinv[128] | barriers

// MN-major operand tile with the 128-byte swizzle: rows (the contraction index) of 128 bytes = 64 elements along
// M / N, 8-row atoms `sbo` bytes apart, 64-element groups along M / N `lbo` bytes apart.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
constexpr uint32_t kMajorMN = (1u << 15) | (1u << 16);   // instruction descriptor: A and B are MN-major
}  // namespace

struct SdftPlanesArgs {
  const float* ta; long long ld_ta, bs_ta;
  const float* col_max; const float* col_sum;
  int n, T, d, row_stride, first_row;
  float divisor, x_unscale;                  // x planes hold x / x_unscale
  float* out; int accumulate;
  const int* n_dev;
  int ta_vec;                                // token_att rows are 16-byte aligned and T % 4 == 0
};

__global__ void __launch_bounds__(320, 1)
query_sdft_mn_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                     SdftPlanesArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* off_s = reinterpret_cast<float*>(smem + MN_STAGES * MN_STAGE);
  float* cinv_s = off_s + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MN_STAGES * MN_STAGE + 1024);
  uint64_t* x_full = bars;                   // [3] TMA landed
  uint64_t* w_full = bars + 3;               // [3] W tiles built (4 warps)
  uint64_t* empty = bars + 6;                // [3] MMAs of the stage retired
  uint64_t* acc_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = blockIdx.x * MN_DN, b = blockIdx.y;
  if (a.n_dev != nullptr) {                  // device-resident token count: packed sequences
    const int Nd = load_len(a.n_dev);
    a.n = min(a.n, Nd - a.first_row);
    a.row_stride = Nd;
    a.bs_ta = Nd * a.ld_ta;
  }
  const int chunks = (a.n + MN_TK - 1) / MN_TK;
  constexpr float kLog2e = 1.4426950408889634f;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MN_STAGES; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&w_full[s], 8);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 192) {   // per-entry softmax constants in the log2 domain
    const int t = threadIdx.x - 64;
    const bool ok = t < a.T;
    off_s[t] = ok ? -a.col_max[b * a.T + t] * kLog2e : 0.f;
    cinv_s[t] = ok ? kWScale / a.col_sum[b * a.T + t] : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    for (int c = 0; c < chunks; ++c) {
      const int st = c % MN_STAGES;
      mbar_wait(&empty[st], ((c / MN_STAGES) & 1) ^ 1);
      uint8_t* xs = smem + st * MN_STAGE;
      const int row = b * a.row_stride + a.first_row + c * MN_TK;
      if (elect_one()) {
        mbar_arrive_expect_tx(&x_full[st], 2 * MN_X_BYTES);
        tma_load_3d(&tm_hi, &x_full[st], xs, 0, row, d0 / 64);               // six 64-column groups in one copy
        tma_load_3d(&tm_lo, &x_full[st], xs + MN_X_BYTES, 0, row, d0 / 64);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc256 = make_idesc(0u, 128, 256) | kMajorMN;
    constexpr uint32_t idesc128 = make_idesc(0u, 128, 128) | kMajorMN;
    for (int c = 0; c < chunks; ++c) {
      const int st = c % MN_STAGES;
      mbar_wait(&x_full[st], (c / MN_STAGES) & 1);
      mbar_wait(&w_full[st], (c / MN_STAGES) & 1);
      tcgen05_fence_after();
      const uint32_t base = smem_u32(smem + st * MN_STAGE);
      const uint32_t x_hi = base, x_lo = base + MN_X_BYTES;
      const uint32_t w_hi = base + 2 * MN_X_BYTES, w_lo = w_hi + MN_W_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {        // small terms first: lo*hi, hi*lo, hi*hi
          const uint32_t wa = pass == 0 ? w_lo : w_hi;
          const uint32_t xb = pass == 1 ? x_lo : x_hi;
#pragma unroll
          for (int ks = 0; ks < MN_TK / 16; ++ks) {   // 16 tokens = two 8-token atoms
            const uint64_t da = make_sw128_mnmajor_desc(wa + ks * 2048, MN_BOX, 1024);
            const uint64_t db0 = make_sw128_mnmajor_desc(xb + ks * 2048, MN_BOX, 1024);
            const uint64_t db1 = make_sw128_mnmajor_desc(xb + 4 * MN_BOX + ks * 2048, MN_BOX, 1024);
            const uint32_t acc = (c | pass | ks) != 0 ? 1u : 0u;
            umma_f16(tmem_base, da, db0, idesc256, acc);          // feature columns 0..255
            umma_f16(tmem_base + 256, da, db1, idesc128, acc);    // feature columns 256..383
          }
        }
        umma_commit(&empty[st]);
        if (c == chunks - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // W builders (eight warps): thread (jj, tq) owns token jj of the stage and codebook entries [16 tq, 16 tq + 16);
    // the token_att values of the NEXT stage are requested before this stage's exponentials run.
    const int bt = threadIdx.x - 64;
    const int jj = bt >> 3, tq = bt & 7;
    const float c1 = kLog2e / a.divisor;
    const float* tab = a.ta + b * a.bs_ta + tq * 16;
    const int row_off = (tq >> 2) * MN_BOX + (jj >> 3) * 1024 + (jj & 7) * 128;
    float off_r[16], cinv_r[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      off_r[u] = off_s[tq * 16 + u];
      cinv_r[u] = cinv_s[tq * 16 + u];
    }
    auto load_row = [&](int c, float (&tv)[16]) {
      const int j = c * MN_TK + jj;
      if (j < a.n) {
        const float* rowp = tab + j * a.ld_ta;
        if (a.ta_vec) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (tq * 16 + u * 4 < a.T) v = *reinterpret_cast<const float4*>(rowp + u * 4);
            tv[4 * u] = v.x; tv[4 * u + 1] = v.y; tv[4 * u + 2] = v.z; tv[4 * u + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) tv[u] = (tq * 16 + u < a.T) ? rowp[u] : -INFINITY;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) tv[u] = -INFINITY;
      }
    };
    float tv[16], tn[16];
    load_row(0, tv);
    for (int c = 0; c < chunks; ++c) {
      const int st = c % MN_STAGES;
      if (c + 1 < chunks) load_row(c + 1, tn);
      mbar_wait(&empty[st], ((c / MN_STAGES) & 1) ^ 1);   // the stage's previous tiles are no longer read
      uint8_t* w_hi = smem + st * MN_STAGE + 2 * MN_X_BYTES;
      uint8_t* w_lo = w_hi + MN_W_BYTES;
#pragma unroll
      for (int q = 0; q < 2; ++q) {           // eight entries = one 16-byte granule of the token's row
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int u = q * 8 + 2 * e;
          const float w0 = ex2_approx(fmaf(tv[u], c1, off_r[u])) * cinv_r[u];
          const float w1 = ex2_approx(fmaf(tv[u + 1], c1, off_r[u + 1])) * cinv_r[u + 1];
          const __half2 h = __floats2half2_rn(w0, w1);
          const float2 f = __half22float2(h);
          const __half2 l = __floats2half2_rn(w0 - f.x, w1 - f.y);
          ph[e] = *reinterpret_cast<const uint32_t*>(&h);
          pl[e] = *reinterpret_cast<const uint32_t*>(&l);
        }
        const int granule = (tq & 3) * 2 + q;
        const int offb = row_off + ((granule ^ (jj & 7)) << 4);
        *reinterpret_cast<uint4*>(w_hi + offb) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(w_lo + offb) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
      fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&w_full[st]);
#pragma unroll
      for (int u = 0; u < 16; ++u) tv[u] = tn[u];
    }
    // epilogue: thread = codebook entry (TMEM lane), 192 of the 384 feature columns
    mbar_wait(acc_full, 0);
    tcgen05_fence_after();
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const float un = a.x_unscale / kWScale;
    float* orow = a.out + (static_cast<long long>(b) * a.T + row) * a.d + d0;
#pragma unroll 1
    for (int cc = half * (MN_DN / 64); cc < (half + 1) * (MN_DN / 64); ++cc) {
      if (d0 + cc * 32 >= a.d) break;         // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + cc * 32, v);
      tmem_ld_wait();
      if (row < a.T) {
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          float4 o = make_float4(__uint_as_float(v[k]) * un, __uint_as_float(v[k + 1]) * un,
                                 __uint_as_float(v[k + 2]) * un, __uint_as_float(v[k + 3]) * un);
          float4* p = reinterpret_cast<float4*>(orow + cc * 32 + k);
          if (a.accumulate) {
            const float4 q = *p;
            o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
          }
          *p = o;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// x_hi / x_lo: fp16 planes [x_rows, d] (dense) of x / x_unscale; token j of batch b at row b*row_stride + first_row + j.
int launch_query_sdft_planes(const float* token_att, long long ld_ta, long long bs_ta, const float* col_max,
                             const float* col_sum, const __half* x_hi, const __half* x_lo, float x_unscale,
                             long long x_rows, int row_stride, int first_row, int B, int n, int T, int d, float divisor,
                             float* sd_ft, int accumulate, const int* n_dev, cudaStream_t stream) {
  MADTP_CHECK_ARG(token_att && col_max && col_sum && x_hi && x_lo && sd_ft, "query_sdft_planes: null pointer");
  MADTP_CHECK_ARG(B >= 0 && n > 0 && T > 0 && T <= 128 && d > 0 && d % 64 == 0 && B <= 65535,
                  "query_sdft_planes: unsupported shape (T=%d must be <= 128, d=%d a multiple of 64)", T, d);
  if (B == 0) return kOk;
  CUtensorMap th, tl;
  int st;
  if ((st = make_tmap_colgroups(&th, x_hi, x_rows, d, d, MN_TK, MN_DN / 64)) != kOk) return st;
  if ((st = make_tmap_colgroups(&tl, x_lo, x_rows, d, d, MN_TK, MN_DN / 64)) != kOk) return st;
  MADTP_SMEM_ATTR_ONCE(MN_SMEM, query_sdft_mn_kernel);
  SdftPlanesArgs a;
  a.ta = token_att; a.ld_ta = ld_ta; a.bs_ta = bs_ta;
  a.col_max = col_max; a.col_sum = col_sum;
  a.n = n; a.T = T; a.d = d; a.row_stride = row_stride; a.first_row = first_row;
  a.divisor = divisor; a.x_unscale = x_unscale; a.out = sd_ft; a.accumulate = accumulate;
  a.n_dev = n_dev;
  a.ta_vec = (reinterpret_cast<uintptr_t>(token_att) % 16 == 0 && ld_ta % 4 == 0 && bs_ta % 4 == 0 && T % 4 == 0) ? 1 : 0;
  dim3 grid((d + MN_DN - 1) / MN_DN, B);
  query_sdft_mn_kernel<<<grid, 320, MN_SMEM, stream>>>(th, tl, a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

}  // namespace madtp
