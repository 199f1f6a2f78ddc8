// Tensor-core self-attention with fused token-pruning statistics (head dim 64) for the scoring lane.
//
// Same three outputs as attention.cu (context, ||context|| per head and token, column sums of max_h P, the
// head-importance-weighted CLS row) but the two big contractions run on tcgen05 with error-compensated tf32
// operands (hi/lo planes written by the fused q|k|v projection, gemm.cu launch_gemm_qkv):
//
//   attn_fwd_tc_kernel    one CTA per (128-query tile, head, sequence); 64-key tiles.
//       S   = Q K^T           2 x 12 MMAs per key tile: each 32-wide slice of the head dim accumulates in its own
//                             TMEM buffer and the two partials are added in fp32 registers (the tensor core's
//                             accumulator truncates, so long in-TMEM accumulations lose accuracy -- DESIGN.md section 3)
//       P   = online softmax   4 warps, one query row per thread, fp32, expf
//       O_t = P V              P goes back to TMEM as tf32 hi/lo (tcgen05.st) and is the A operand of 24 MMAs against
//                             V^T tiles (keys contiguous, written transposed by the projection epilogue); every key
//                             tile's partial product is drained and accumulated in fp32 registers with the usual
//                             running-max correction.
//   attn_stats_tc_kernel  one CTA per (128-query tile, 128-key tile, sequence), looping over the heads.
//       log P_h(i,j) = S_h(i,j) * scale + mask_j - lse_h(i) is formed for every head, the running max over heads is
//       kept in registers (exp is monotone, so ONE expf per (i,j) after the head loop replaces one per head), then
//       the tile is column-summed over its query rows in a fixed order.
//   attn_cls_kernel       the CLS query row of every head (tiny, fp32 FFMA) weighted by the per-head context norms.
//
// Warp roles in both tensor-core kernels: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warps 2..5 = consumers (TMEM lane quadrant = warp % 4, thread = query row).
#include "attention.cuh"
#include "gemm.cuh"

namespace madtp {

namespace {

constexpr int BM = 128;        // query rows per CTA
constexpr int BOX = 32;        // floats per 128-byte swizzled row
constexpr int TC_THREADS = 192;

__device__ __forceinline__ void ld2_add(uint32_t ta, uint32_t tb, float (&out)[32]) {
  uint32_t a[32], b[32];
  tmem_ld_32x32b_x32(ta, a);
  tmem_ld_32x32b_x32(tb, b);
  tmem_ld_wait();
#pragma unroll
  for (int k = 0; k < 32; ++k) out[k] = __uint_as_float(a[k]) + __uint_as_float(b[k]);
}

// 12 tf32 MMAs = one 32-wide K slice of an error-compensated product (lo*hi, hi*lo, hi*hi per 8-element k-step)
__device__ __forceinline__ void issue_slice_ss(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                               uint32_t b_lo, uint32_t idesc) {
  const uint64_t dah = make_sw128_kmajor_desc(a_hi), dal = make_sw128_kmajor_desc(a_lo);
  const uint64_t dbh = make_sw128_kmajor_desc(b_hi), dbl = make_sw128_kmajor_desc(b_lo);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, dal + 2 * k, dbh + 2 * k, idesc, k != 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, dah + 2 * k, dbl + 2 * k, idesc, 1u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, 1u);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Pass 1
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 softmax: TMEM lane quadrant = warp % 4 (query rows), column
//   half = (warp - 2) / 4 (32 of the tile's 64 keys, and 32 of the 64 output dims), so every query row is shared by
//   two threads that exchange their tile maxima through shared memory.
//   K tiles ride a 3-deep ring and V^T tiles a 2-deep ring: K(t+1) is requested two tiles ahead of its QK^T, V(t) is
//   only needed after softmax(t), so neither TMA latency sits on the critical path.
// ------------------------------------------------------------------------------------------------
struct FwdSmem {
  static constexpr int Q_BYTES = 4 * BM * 128;              // (slice 0/1) x (hi/lo) boxes of [128 x 32]
  static constexpr int KBOX = 64 * 128;                     // [64 keys x 32] or [64 dims x 32 keys]
  static constexpr int TILE_BYTES = 4 * KBOX;               // one K tile or one V^T tile: (slice 0/1) x (hi/lo)
  static constexpr int K_STAGES = 3, V_STAGES = 2;
  static constexpr int K_OFF = Q_BYTES;
  static constexpr int V_OFF = K_OFF + K_STAGES * TILE_BYTES;
  static constexpr int BAR_OFF = V_OFF + V_STAGES * TILE_BYTES;
  static constexpr int XCH_OFF = BAR_OFF + 256;             // [2][2][128] floats: per-row exchange between column halves
  static constexpr int TOTAL = XCH_OFF + 2 * 2 * BM * 4;
};
constexpr int FWD_THREADS = 320;

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                   const __grid_constant__ CUtensorMap tm_k_hi, const __grid_constant__ CUtensorMap tm_k_lo,
                   const __grid_constant__ CUtensorMap tm_v_hi, const __grid_constant__ CUtensorMap tm_v_lo,
                   AttnTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* q_s = smem;
  uint8_t* k_s = smem + FwdSmem::K_OFF;
  uint8_t* v_s = smem + FwdSmem::V_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR_OFF);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // [3]
  uint64_t* k_empty = bars + 4;    // [3]
  uint64_t* v_full = bars + 7;     // [2]
  uint64_t* v_empty = bars + 9;    // [2]
  uint64_t* s_full = bars + 11;    // S(t) complete in TMEM
  uint64_t* s_empty = bars + 12;   // S(t) copied to registers by all eight softmax warps
  uint64_t* p_full = bars + 13;    // [2] P(t) stored (parity t & 1)
  uint64_t* o_full = bars + 15;    // [2] P(t) V(t) complete (parity t & 1): O partial ready, P buffer free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  float* xch = reinterpret_cast<float*>(smem + FwdSmem::XCH_OFF);   // [parity][half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int N = a.N, HD = a.H * 64;
  const int T = (N + 63) / 64;   // key tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q_hi);
    tma_prefetch_desc(&tm_q_lo);
    tma_prefetch_desc(&tm_k_hi);
    tma_prefetch_desc(&tm_k_lo);
    tma_prefetch_desc(&tm_v_hi);
    tma_prefetch_desc(&tm_v_lo);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < FwdSmem::K_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&p_full[s], 8);
      mbar_init(&o_full[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S (slice A | slice B) at 0; P hi, P lo and the O partial are double-buffered by tile parity so that
  // softmax(t+1) never waits for P(t) V(t): P hi at 128 / 192, P lo at 256 / 320, O partial at 384 / 448
  constexpr uint32_t kP_HI = 128, kP_LO = 256, kO = 384;

  if (warp == 0) {   // warp-uniform control flow, one elected lane issues the TMA
    {
      const int qrow = b * N + i0;
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, FwdSmem::Q_BYTES);
        for (int c = 0; c < 2; ++c) {
          tma_load_2d(&tm_q_hi, q_full, q_s + (c * 2 + 0) * BM * 128, h * 64 + c * BOX, qrow);
          tma_load_2d(&tm_q_lo, q_full, q_s + (c * 2 + 1) * BM * 128, h * 64 + c * BOX, qrow);
        }
      }
      __syncwarp();
      auto load_k = [&](int t) {
        const int st = t % FwdSmem::K_STAGES;
        mbar_wait(&k_empty[st], ((t / FwdSmem::K_STAGES) & 1) ^ 1);
        uint8_t* s = k_s + st * FwdSmem::TILE_BYTES;
        const int krow = b * N + t * 64;
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], FwdSmem::TILE_BYTES);
          for (int c = 0; c < 2; ++c) {
            tma_load_2d(&tm_k_hi, &k_full[st], s + (c * 2 + 0) * FwdSmem::KBOX, HD + h * 64 + c * BOX, krow);
            tma_load_2d(&tm_k_lo, &k_full[st], s + (c * 2 + 1) * FwdSmem::KBOX, HD + h * 64 + c * BOX, krow);
          }
        }
        __syncwarp();
      };
      auto load_v = [&](int t) {
        const int st = t & 1;
        mbar_wait(&v_empty[st], ((t >> 1) & 1) ^ 1);
        uint8_t* s = v_s + st * FwdSmem::TILE_BYTES;
        const int vrow = (b * a.H + h) * 64;
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[st], FwdSmem::TILE_BYTES);
          for (int c = 0; c < 2; ++c) {
            tma_load_2d(&tm_v_hi, &v_full[st], s + (c * 2 + 0) * FwdSmem::KBOX, t * 64 + c * BOX, vrow);
            tma_load_2d(&tm_v_lo, &v_full[st], s + (c * 2 + 1) * FwdSmem::KBOX, t * 64 + c * BOX, vrow);
          }
        }
        __syncwarp();
      };
      load_k(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) load_k(t + 1);
        load_v(t);
      }
    }
  } else if (warp == 1) {
    // The whole warp runs this loop and ONE elected lane issues: with warp-uniform control flow the descriptors stay
    // in uniform registers. (Guarding the loop with `lane == 0` instead makes every tcgen05.mma a ~100-cycle
    // R2UR + waterfall sequence -- measured: that, not the tensor pipe or the softmax, bounded this kernel.)
    constexpr uint32_t idesc = make_idesc(2u, BM, 64);
    const uint32_t q_u = smem_u32(q_s);
    auto issue_qk = [&](int t) {
      const int st = t % FwdSmem::K_STAGES;
      mbar_wait(&k_full[st], (t / FwdSmem::K_STAGES) & 1);
      mbar_wait(s_empty, (t & 1) ^ 1);       // softmax(t-1) has S(t-1) in registers
      tcgen05_fence_after();
      const uint32_t k_u = smem_u32(k_s + st * FwdSmem::TILE_BYTES);
      if (elect_one()) {
        for (int c = 0; c < 2; ++c)
          issue_slice_ss(tmem_base + c * 64, q_u + (c * 2 + 0) * BM * 128, q_u + (c * 2 + 1) * BM * 128,
                         k_u + (c * 2 + 0) * FwdSmem::KBOX, k_u + (c * 2 + 1) * FwdSmem::KBOX, idesc);
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int t = 0; t < T; ++t) {
      if (t + 1 < T) issue_qk(t + 1);
      const int pb = t & 1;
      mbar_wait(&v_full[pb], (t >> 1) & 1);
      mbar_wait(&p_full[pb], (t >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t v_u = smem_u32(v_s + pb * FwdSmem::TILE_BYTES);
      if (elect_one()) {
        // O_t = P_lo V_hi + P_hi V_lo + P_hi V_hi over 8 k-steps of 8 keys
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t pa = tmem_base + (pass == 0 ? kP_LO : kP_HI) + pb * 64;
          const int vplane = (pass == 1) ? 1 : 0;  // pass 1 multiplies by V_lo
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t bd =
                make_sw128_kmajor_desc(v_u + ((ks >> 2) * 2 + vplane) * FwdSmem::KBOX) + 2 * (ks & 3);
            umma_tf32_ts(tmem_base + kO + pb * 64, pa + ks * 8, bd, idesc, (pass | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&o_full[pb]);
        umma_commit(&v_empty[pb]);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int hf = (warp - 2) >> 2;                 // column half
    const int r = quad * 32 + lane;                 // row within the tile
    const int i = i0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float* mask = a.key_mask ? a.key_mask + static_cast<long long>(b) * N : nullptr;
    float m = -INFINITY, l = 0.f;
    float m_hist[2] = {-INFINITY, -INFINITY};   // running max at tiles t-2 / t-1 (indexed by tile parity)
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;

    // drain the partial product of tile u (buffer u & 1, relative to the running max m_u) into o (relative to m_now)
    auto drain = [&](int u, float m_u, float m_now) {
      mbar_wait_spin(&o_full[u & 1], (u >> 1) & 1);
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + lane_off + kO + (u & 1) * 64 + hf * 32, v);
      tmem_ld_wait();
      const float f = expf(m_u - m_now);
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] = fmaf(__uint_as_float(v[k]), f, o[k]);
    };

    for (int t = 0; t < T; ++t) {
      mbar_wait_spin(s_full, t & 1);
      tcgen05_fence_after();
      float s[32];
      ld2_add(tmem_base + lane_off + hf * 32, tmem_base + lane_off + 64 + hf * 32, s);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);

      const int j0 = t * 64 + hf * 32;
      float mx = -INFINITY;
      if (mask == nullptr && j0 + 32 <= N) {        // interior tile: no bounds or mask handling
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          s[k] *= a.scale;
          mx = fmaxf(mx, s[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int j = j0 + k;
          const float mk = (mask != nullptr && j < N) ? __ldg(mask + j) : 0.f;
          s[k] = (j < N) ? fmaf(s[k], a.scale, mk) : -INFINITY;
          mx = fmaxf(mx, s[k]);
        }
      }
      // joint maximum of the row over both column halves
      float* x = xch + (t & 1) * 2 * BM;
      x[hf * BM + r] = mx;
      named_bar_sync(1 + quad, 64);
      mx = fmaxf(mx, x[(hf ^ 1) * BM + r]);

      const float m_new = fmaxf(m, mx);
      const float corr = (m == -INFINITY) ? 0.f : expf(m - m_new);
      float ps = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        s[k] = expf(s[k] - m_new);
        ps += s[k];
      }
      l = l * corr + ps;
      m = m_new;
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] *= corr;     // o is now relative to m_new
      // the P buffer of this parity was last read by P(t-2) V(t-2): drain that partial, which also frees the buffer
      if (t >= 2) drain(t - 2, m_hist[t & 1], m_new);
      m_hist[t & 1] = m_new;

      // P -> TMEM as tf32 hi / lo (A operand of the P V MMAs)
      {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float p = s[k];
          const float ph = tf32_hi(p);
          hi[k] = __float_as_uint(ph);
          lo[k] = __float_as_uint(p - ph);
        }
        tmem_st_32x32b_x32(tmem_base + lane_off + kP_HI + (t & 1) * 64 + hf * 32, hi);
        tmem_st_32x32b_x32(tmem_base + lane_off + kP_LO + (t & 1) * 64 + hf * 32, lo);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t & 1]);
    }
    if (T >= 2) drain(T - 2, m_hist[(T - 2) & 1], m);
    drain(T - 1, m_hist[(T - 1) & 1], m);
    // row sum and squared norm: combine the two column halves (buffers of parity T&1 are free: their last readers
    // passed the barrier of tile T-2 ... T-1 uses the other parity)
    float* x = xch + (T & 1) * 2 * BM;
    x[hf * BM + r] = l;
    named_bar_sync(1 + quad, 64);
    const float l_tot = l + x[(hf ^ 1) * BM + r];
    const float inv = 1.0f / l_tot;
    float nsq = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      o[d] *= inv;
      nsq = fmaf(o[d], o[d], nsq);
    }
    named_bar_sync(1 + quad, 64);                    // both halves have read l before the buffer is reused
    x[hf * BM + r] = nsq;
    named_bar_sync(1 + quad, 64);
    const float nsq_tot = (hf == 0) ? (nsq + x[BM + r]) : (x[r] + nsq);   // dims 0..31 first, then 32..63
    if (i < N) {
      __half* dst = a.out_f16 + b * a.bso + static_cast<long long>(i) * a.ldo + h * 64 + hf * 32;
#pragma unroll
      for (int d = 0; d < 32; d += 8) {
        __half2 h0 = __floats2half2_rn(o[d], o[d + 1]), h1 = __floats2half2_rn(o[d + 2], o[d + 3]);
        __half2 h2 = __floats2half2_rn(o[d + 4], o[d + 5]), h3 = __floats2half2_rn(o[d + 6], o[d + 7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + d) = pk;
      }
      if (hf == 0) {
        const long long sidx = (static_cast<long long>(b) * a.H + h) * N + i;
        a.row_lse[sidx] = m + logf(l_tot);
        a.out_norm[sidx] = sqrtf(nsq_tot);
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

// ------------------------------------------------------------------------------------------------
// Pass 2
// ------------------------------------------------------------------------------------------------
struct StatsSmem {
  static constexpr int BOX_BYTES = BM * 128;                // [128 rows x 32 floats]
  static constexpr int STAGE_BYTES = 4 * BOX_BYTES;         // Q hi, Q lo, K hi, K lo of one (head, 32-wide slice)
  static constexpr int STAGES = 3;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 1024 + 1024;
  static constexpr int RED_LD = 129;                        // column-sum staging pitch (floats), reuses the stages
};

__global__ void __launch_bounds__(TC_THREADS, 1)
attn_stats_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                     AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + StatsSmem::BAR_OFF);
  uint64_t* full = bars;           // [3]
  uint64_t* empty = bars + 3;      // [3]
  uint64_t* s_full = bars + 6;     // [2]
  uint64_t* s_empty = bars + 8;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* colmask = reinterpret_cast<float*>(bars + 12);  // [128] additive mask of this key tile (-inf past N)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * BM, it = blockIdx.y, i0 = it * BM, b = blockIdx.z;
  const int N = a.N, H = a.H, HD = a.H * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < StatsSmem::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  if (warp >= 2) {
    const int c = threadIdx.x - 64;
    const int j = j0 + c;
    colmask[c] = (j < N) ? (a.key_mask ? a.key_mask[static_cast<long long>(b) * N + j] : 0.f) : -INFINITY;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {   // warp-uniform control flow, one elected lane issues
    for (int u = 0; u < 2 * H; ++u) {
      const int st = u % StatsSmem::STAGES, hh = u >> 1, c = u & 1;
      mbar_wait(&empty[st], ((u / StatsSmem::STAGES) & 1) ^ 1);
      uint8_t* s = smem + st * StatsSmem::STAGE_BYTES;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[st], StatsSmem::STAGE_BYTES);
        tma_load_2d(&tm_hi, &full[st], s, hh * 64 + c * BOX, b * N + i0);
        tma_load_2d(&tm_lo, &full[st], s + StatsSmem::BOX_BYTES, hh * 64 + c * BOX, b * N + i0);
        tma_load_2d(&tm_hi, &full[st], s + 2 * StatsSmem::BOX_BYTES, HD + hh * 64 + c * BOX, b * N + j0);
        tma_load_2d(&tm_lo, &full[st], s + 3 * StatsSmem::BOX_BYTES, HD + hh * 64 + c * BOX, b * N + j0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(2u, BM, BM);
    for (int hh = 0; hh < H; ++hh) {
      const int hb = hh & 1;
      mbar_wait(&s_empty[hb], ((hh >> 1) & 1) ^ 1);
      for (int c = 0; c < 2; ++c) {
        const int u = hh * 2 + c, st = u % StatsSmem::STAGES;
        mbar_wait(&full[st], (u / StatsSmem::STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t s = smem_u32(smem + st * StatsSmem::STAGE_BYTES);
        if (elect_one()) {
          issue_slice_ss(tmem_base + hb * 256 + c * 128, s, s + StatsSmem::BOX_BYTES, s + 2 * StatsSmem::BOX_BYTES,
                         s + 3 * StatsSmem::BOX_BYTES, idesc);
          umma_commit(&empty[st]);
          if (c == 1) umma_commit(&s_full[hb]);
        }
        __syncwarp();
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;          // row within the tile
    const int i = i0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const bool row_ok = (i >= 1) && (i < N);  // the CLS query row is excluded (reference vit.py:126)
    float mx[BM];
#pragma unroll
    for (int c = 0; c < BM; ++c) mx[c] = -INFINITY;
    for (int hh = 0; hh < H; ++hh) {
      const int hb = hh & 1;
      const float lse = (i < N) ? a.row_lse[(static_cast<long long>(b) * H + hh) * N + i] : 0.f;
      mbar_wait(&s_full[hb], (hh >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float s[32];
        ld2_add(tmem_base + lane_off + hb * 256 + cc * 32, tmem_base + lane_off + hb * 256 + 128 + cc * 32, s);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float t = fmaf(s[k], a.scale, colmask[cc * 32 + k]) - lse;
          mx[cc * 32 + k] = fmaxf(mx[cc * 32 + k], t);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[hb]);
    }
    // every MMA has retired (the last s_full fired), so the operand stages can be reused as the reduction tile
    float* red = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int c = 0; c < BM; ++c) red[r * StatsSmem::RED_LD + c] = row_ok ? expf(mx[c]) : 0.f;
    named_bar_sync(1, 128);
    const int c = threadIdx.x - 64;
    const int j = j0 + c;
    float sum = 0.f;
    for (int rr = 0; rr < BM; ++rr) sum += red[rr * StatsSmem::RED_LD + c];   // fixed order: deterministic
    if (j < N) a.col_part[(static_cast<long long>(b) * a.n_parts + it) * N + j] = sum;
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// CLS row: cls_attn[b,j] = sum_h softmax_j(q_0 . k_j)_h * norm[b,h,j] / (sum_h' norm[b,h',j] + 1e-8)
//   attn_cls_head_kernel     grid (H, B), block 256: the CLS query row of one head (fp32 FFMA, own softmax) -> scratch
//   attn_cls_combine_kernel  grid (ceil(N/256), B): head-importance weighting and the sum over heads (h ascending)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_cls_head_kernel(AttnTcArgs a) {
  extern __shared__ float sm[];
  const int N = a.N, H = a.H, HD = H * 64;
  float* q0 = sm;        // [64]
  float* red = sm + 64;  // [16]
  float* P = sm + 80;    // [N]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hh = blockIdx.x, b = blockIdx.y;
  const float* row0_hi = a.qk_hi + static_cast<long long>(b) * N * a.ld_qk;
  const float* row0_lo = a.qk_lo + static_cast<long long>(b) * N * a.ld_qk;
  if (tid < 64) q0[tid] = row0_hi[hh * 64 + tid] + row0_lo[hh * 64 + tid];
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < N; j += 256) {
    const float4* kh = reinterpret_cast<const float4*>(row0_hi + static_cast<long long>(j) * a.ld_qk + HD + hh * 64);
    const float4* kl = reinterpret_cast<const float4*>(row0_lo + static_cast<long long>(j) * a.ld_qk + HD + hh * 64);
    float s = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < 16; ++d4) {
      const float4 x = kh[d4], y = kl[d4];
      const float4 q = *reinterpret_cast<const float4*>(q0 + d4 * 4);
      s = fmaf(q.x, x.x + y.x, s);
      s = fmaf(q.y, x.y + y.y, s);
      s = fmaf(q.z, x.z + y.z, s);
      s = fmaf(q.w, x.w + y.w, s);
    }
    const float mk = a.key_mask ? a.key_mask[static_cast<long long>(b) * N + j] : 0.f;
    const float lg = fmaf(s, a.scale, mk);
    P[j] = lg;
    mx = fmaxf(mx, lg);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  float sum = 0.f;
  for (int j = tid; j < N; j += 256) {
    const float e = expf(P[j] - mx);
    P[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[8 + warp] = sum;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += red[8 + w];
  const float inv = 1.0f / tot;
  float* out = a.cls_scratch + (static_cast<long long>(b) * H + hh) * N;
  for (int j = tid; j < N; j += 256) out[j] = P[j] * inv;
}

__global__ void __launch_bounds__(256)
attn_cls_combine_kernel(AttnTcArgs a) {
  const int N = a.N, H = a.H;
  const int j = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (j >= N) return;
  const long long base = static_cast<long long>(b) * H * N + j;
  float hs = 0.f;
  for (int hh = 0; hh < H; ++hh) hs += a.out_norm[base + static_cast<long long>(hh) * N];
  hs += 1e-8f;
  float acc = 0.f;
  for (int hh = 0; hh < H; ++hh)
    acc += a.cls_scratch[base + static_cast<long long>(hh) * N] * (a.out_norm[base + static_cast<long long>(hh) * N] / hs);
  a.cls_attn[static_cast<long long>(b) * N + j] = acc;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int check_tc(const AttnTcArgs& a) {
  MADTP_CHECK_ARG(a.qk_hi && a.qk_lo, "attn_tc: null q/k planes");
  MADTP_CHECK_ARG(a.B >= 0 && a.H > 0 && a.N > 0, "attn_tc: bad shape B=%d H=%d N=%d", a.B, a.H, a.N);
  MADTP_CHECK_ARG(a.ld_qk >= 2LL * a.H * 64 && a.ld_qk % 4 == 0, "attn_tc: bad q/k leading dimension");
  MADTP_CHECK_ARG(a.B <= 65535 && a.H <= 65535, "attn_tc: B and H must fit the grid limits");
  return kOk;
}

int launch_attn_fwd_tc(const AttnTcArgs& a, cudaStream_t stream) {
  int st = check_tc(a);
  if (st != kOk) return st;
  MADTP_CHECK_ARG(a.vt_hi && a.vt_lo && a.ld_vt >= a.N && a.ld_vt % 4 == 0, "attn_fwd_tc: bad V^T planes");
  MADTP_CHECK_ARG(a.out_f16 && a.ldo % 8 == 0 && a.bso % 8 == 0 && a.row_lse && a.out_norm, "attn_fwd_tc: bad outputs");
  if (a.B == 0) return kOk;
  const long long rows = static_cast<long long>(a.B) * a.N;
  CUtensorMap tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo;
  if ((st = make_tmap(&tq_hi, a.qk_hi, true, rows, 2LL * a.H * 64, a.ld_qk, BM)) != kOk) return st;
  if ((st = make_tmap(&tq_lo, a.qk_lo, true, rows, 2LL * a.H * 64, a.ld_qk, BM)) != kOk) return st;
  if ((st = make_tmap(&tk_hi, a.qk_hi, true, rows, 2LL * a.H * 64, a.ld_qk, 64)) != kOk) return st;
  if ((st = make_tmap(&tk_lo, a.qk_lo, true, rows, 2LL * a.H * 64, a.ld_qk, 64)) != kOk) return st;
  const long long vrows = static_cast<long long>(a.B) * a.H * 64;
  if ((st = make_tmap(&tv_hi, a.vt_hi, true, vrows, a.N, a.ld_vt, 64)) != kOk) return st;
  if ((st = make_tmap(&tv_lo, a.vt_lo, true, vrows, a.N, a.ld_vt, 64)) != kOk) return st;
  static bool attr_done = false;
  if (!attr_done) {
    MADTP_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    attr_done = true;
  }
  dim3 grid((a.N + BM - 1) / BM, a.H, a.B);
  attn_fwd_tc_kernel<<<grid, FWD_THREADS, FwdSmem::TOTAL, stream>>>(tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

int launch_attn_stats_tc(const AttnTcArgs& a, cudaStream_t stream) {
  int st = check_tc(a);
  if (st != kOk) return st;
  MADTP_CHECK_ARG(a.row_lse && a.out_norm && a.col_part && a.cls_attn && a.cls_scratch,
                  "attn_stats_tc: null statistics buffer");
  MADTP_CHECK_ARG(a.n_parts == (a.N + BM - 1) / BM, "attn_stats_tc: n_parts must be ceil(N/128)");
  MADTP_CHECK_ARG(a.N <= 8192, "attn_stats_tc: sequence too long for the CLS-row kernel");
  if (a.B == 0) return kOk;
  const long long rows = static_cast<long long>(a.B) * a.N;
  CUtensorMap t_hi, t_lo;
  if ((st = make_tmap(&t_hi, a.qk_hi, true, rows, 2LL * a.H * 64, a.ld_qk, BM)) != kOk) return st;
  if ((st = make_tmap(&t_lo, a.qk_lo, true, rows, 2LL * a.H * 64, a.ld_qk, BM)) != kOk) return st;
  static bool attr_done = false;
  if (!attr_done) {
    MADTP_CUDA(cudaFuncSetAttribute(attn_stats_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    StatsSmem::TOTAL));
    attr_done = true;
  }
  dim3 grid(a.n_parts, a.n_parts, a.B);
  attn_stats_tc_kernel<<<grid, TC_THREADS, StatsSmem::TOTAL, stream>>>(t_hi, t_lo, a);
  MADTP_LAUNCH_CHECK();
  const size_t cls_smem = (80 + static_cast<size_t>(a.N)) * sizeof(float);
  attn_cls_head_kernel<<<dim3(a.H, a.B), 256, cls_smem, stream>>>(a);
  MADTP_LAUNCH_CHECK();
  attn_cls_combine_kernel<<<dim3((a.N + 255) / 256, a.B), 256, 0, stream>>>(a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

}  // namespace madtp
