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

}  // namespace madtp
