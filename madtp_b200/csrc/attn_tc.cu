// Tensor-core self-attention with fused token-pruning statistics (head dim 64) for the scoring lane.
//
// Same three outputs as attention.cu (context, ||context|| per head and token, column sums of max_h P, the
// head-importance-weighted CLS row) but the two big contractions run on tcgen05 with error-compensated fp16
// operands: q, k and v arrive as fp16 hi/lo planes (hi = fp16(s x), lo = fp16(s x - hi), 22 mantissa bits together)
// written by the fused q|k|v projection (gemm.cu launch_gemm_qkv) and every product is lo*hi + hi*lo + hi*hi.
//
//   attn_fwd_tc_kernel    one CTA per (128-query tile, head, sequence); 64-key tiles.
//       S   = Q K^T           12 MMAs per key tile (the head dim is one 128-byte operand row) into one of two S buffers
//       P   = online softmax   8 or 16 warps, thread = (query row, 32 or 16 of the tile's keys), fp32, expf
//       O_t = P V              256 P goes back to TMEM as packed fp16 hi/lo (tcgen05.st) and is the A operand of 12
//                             MMAs against V^T tiles (keys contiguous, written transposed by the projection epilogue);
//                             every key tile's partial product is drained and accumulated in fp32 registers with the
//                             usual running-max correction (the tensor core's accumulator truncates, so long in-TMEM
//                             accumulations lose accuracy -- DESIGN.md section 3).
//   attn_stats_tc_kernel  one CTA per (128-query tile, 128-key tile, sequence), looping over the heads.
//       log P_h(i,j) = S_h(i,j) * scale + mask_j - lse_h(i) is formed for every head, the running max over heads is
//       kept in registers (exp is monotone, so ONE expf per (i,j) after the head loop replaces one per head), then
//       the tile is column-summed over its query rows in a fixed order.
//       The same kernel first combines the CLS query row of every head (kept by the forward pass) with the context
//       norms into cls_attn (its consumer warps idle while the operand pipeline fills).
//
// Warp roles in both tensor-core kernels: warp 0 = TMA producer, warp 1 = MMA issuer (warp-uniform loops, one elected
// lane issues), warps 2.. = consumers (TMEM lane quadrant = warp % 4, thread = query row).
#include "attention.cuh"
#include "gemm.cuh"

namespace madtp {

#ifdef MADTP_ATTN_TRACE
// Development aid: clock64() timeline of one CTA of attn_fwd_tc_kernel (roles x tiles x slots), read back through
// madtp_debug_read_attn_trace. Never compiled into the shipped library.
__device__ unsigned long long g_attn_trace[4 * 64 * 8];
#define ATT_TRACE(role, tile, slot)                                                                      \
  do {                                                                                                   \
    if (trace_cta && (tile) < 64) g_attn_trace[((role) * 64 + (tile)) * 8 + (slot)] = clock64();         \
  } while (0)
extern "C" int madtp_debug_read_attn_trace(void* dst) {
  return static_cast<int>(cudaMemcpyFromSymbol(dst, g_attn_trace, sizeof(g_attn_trace)));
}
#else
#define ATT_TRACE(role, tile, slot) do { } while (0)
#endif

namespace {

constexpr int BM = 128;             // query rows per CTA
constexpr float kPScale = 256.0f;   // P is stored as fp16 planes of 256 * exp(s - m): lo stays normal down to P ~ 5e-4

// 12 fp16 MMAs = one 64-wide K slice of an error-compensated product (lo*hi, hi*lo, hi*hi per 16-element k-step)
__device__ __forceinline__ void issue_slice_ss(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                               uint32_t b_lo, uint32_t idesc) {
  const uint64_t dah = make_sw128_kmajor_desc(a_hi), dal = make_sw128_kmajor_desc(a_lo);
  const uint64_t dbh = make_sw128_kmajor_desc(b_hi), dbl = make_sw128_kmajor_desc(b_lo);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dal + 2 * k, dbh + 2 * k, idesc, k != 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dah + 2 * k, dbl + 2 * k, idesc, 1u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, 1u);
}

template <int W>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[W]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_32x32b_x32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld_32x32b_x16(taddr, v); }
template <int W>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&v)[W]);
template <>
__device__ __forceinline__ void tmem_st_cols<16>(uint32_t taddr, const uint32_t (&v)[16]) { tmem_st_32x32b_x16(taddr, v); }
template <>
__device__ __forceinline__ void tmem_st_cols<8>(uint32_t taddr, const uint32_t (&v)[8]) { tmem_st_32x32b_x8(taddr, v); }

}  // namespace

// ------------------------------------------------------------------------------------------------
// Pass 1 (persistent: every CTA walks a strided list of (sequence, head, 128-query tile) work items)
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2.. softmax: TMEM lane quadrant = warp % 4 (query rows), column
//   group = (warp - 2) / 4 (KW of the tile's 64 keys, and KW of the 64 output dims), so every query row is shared by
//   64 / KW threads that exchange their tile maxima through shared memory.
//   All rings and TMEM buffers are indexed by a tile counter g that runs ACROSS work items, so the producer and the
//   MMA warp run ahead into the next item (its Q, K(0), K(1) are in shared memory and S(0) is in TMEM) while the
//   softmax warps finish the current one: the TMA / MMA start-up latency of an item (~3.5k cycles, measured, against
//   ~1.5k per key tile) is paid once per CTA instead of once per item.
//   K tiles ride a 3-deep ring and V^T tiles a 2-deep ring; Q, S, P and the O partial are double-buffered.
// ------------------------------------------------------------------------------------------------
template <int KW, int NB>   // NB = P / O-partial buffers in TMEM: 2 (one CTA per SM) or 1 (two co-resident CTAs per SM)
struct FwdCfg {
  static constexpr int NG = 64 / KW;                       // column groups per tile
  static constexpr int THREADS = 64 + 128 * NG;            // 320 (KW = 32) or 576 (KW = 16)
  static constexpr int Q_BYTES = 2 * BM * 128;             // hi, lo boxes of [128 rows x 64 halves]
  static constexpr int Q_STAGES = NB;                      // 2 query buffers when the CTA owns the SM
  static constexpr int KBOX = 64 * 128;                    // [64 keys x 64 dims] or [64 dims x 64 keys] halves
  static constexpr int TILE_BYTES = 2 * KBOX;              // hi, lo
  static constexpr int K_STAGES = NB == 2 ? 3 : 2, V_STAGES = 2;
  static constexpr int TMEM_COLS = NB == 2 ? 512 : 256;
  static constexpr int CTAS_PER_SM = NB == 2 ? 1 : 2;
  static constexpr int K_OFF = Q_STAGES * Q_BYTES;
  static constexpr int V_OFF = K_OFF + K_STAGES * TILE_BYTES;
  static constexpr int BAR_OFF = V_OFF + V_STAGES * TILE_BYTES;
  static constexpr int XCH_OFF = BAR_OFF + 256;            // [3][NG][128] floats: per-row exchange between groups
  static constexpr int CLS_OFF = XCH_OFF + 3 * NG * BM * 4;   // [64] floats: staging of the CLS query row
  static constexpr int TOTAL = CLS_OFF + 64 * 4;
};

template <int KW, int NB>
__global__ void __launch_bounds__(FwdCfg<KW, NB>::THREADS, FwdCfg<KW, NB>::CTAS_PER_SM)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                   const __grid_constant__ CUtensorMap tm_k_hi, const __grid_constant__ CUtensorMap tm_k_lo,
                   const __grid_constant__ CUtensorMap tm_v_hi, const __grid_constant__ CUtensorMap tm_v_lo,
                   AttnTcArgs a) {
  using Cfg = FwdCfg<KW, NB>;
  constexpr int NG = Cfg::NG;
  constexpr int QS = Cfg::Q_STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* q_s = smem;
  uint8_t* k_s = smem + Cfg::K_OFF;
  uint8_t* v_s = smem + Cfg::V_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2] every QK^T of the item that used this query buffer has retired
  uint64_t* k_full = bars + 4;     // [3]
  uint64_t* k_empty = bars + 7;    // [3]
  uint64_t* v_full = bars + 10;    // [2]
  uint64_t* v_empty = bars + 12;   // [2]
  uint64_t* s_full = bars + 14;    // [2] S(g) complete in TMEM (parity g & 1)
  uint64_t* s_empty = bars + 16;   // [2] S(g) copied to registers by all softmax warps
  uint64_t* p_full = bars + 18;    // [2] P(g) stored
  uint64_t* o_full = bars + 20;    // [2] P(g) V(g) complete: O partial ready, P buffer free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  float* xch = reinterpret_cast<float*>(smem + Cfg::XCH_OFF);   // [parity | item end][group][row]
  float* cls_stage = reinterpret_cast<float*>(smem + Cfg::CLS_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.n_dev ? min(a.N, load_len(a.n_dev)) : a.N, HD = a.H * 64;   // device-resident token count (packed)
  if (a.n_dev) a.bso = static_cast<long long>(N) * a.ldo;
  const int T = (N + 63) / 64;           // key tiles per item
  const int QT = (N + BM - 1) / BM;      // query tiles per (sequence, head)
  const int items = QT * a.H * a.B;
#ifdef MADTP_ATTN_TRACE
  const bool trace_cta = blockIdx.x == gridDim.x / 2 && lane == 0;
  if (warp == 0) ATT_TRACE(0, 63, 0);
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q_hi);
    tma_prefetch_desc(&tm_q_lo);
    tma_prefetch_desc(&tm_k_hi);
    tma_prefetch_desc(&tm_k_lo);
    tma_prefetch_desc(&tm_v_hi);
    tma_prefetch_desc(&tm_v_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 3; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 4 * NG);
      mbar_init(&p_full[s], 4 * NG);
      mbar_init(&o_full[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S double-buffered by tile parity at (g&1)*64; NB buffers (pb = g % NB) of P hi (packed fp16 pairs,
  // 32 columns each) at 128, of P lo behind them and of the O partial (64 columns each) behind those
  constexpr uint32_t kP_HI = 128, kP_LO = 128 + NB * 32, kO = 128 + NB * 64;

  if (warp == 0) {   // warp-uniform control flow, one elected lane issues the TMA
    int g = 0, it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int qt = item % QT, bh = item / QT, h = bh % a.H, b = bh / a.H;
      const int qb = it % QS;
      mbar_wait(&q_empty[qb], ((it / QS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[qb], Cfg::Q_BYTES);
        tma_load_2d(&tm_q_hi, &q_full[qb], q_s + qb * Cfg::Q_BYTES, h * 64, b * N + qt * BM);
        tma_load_2d(&tm_q_lo, &q_full[qb], q_s + qb * Cfg::Q_BYTES + BM * 128, h * 64, b * N + qt * BM);
      }
      __syncwarp();
      auto load_k = [&](int t) {
        const int gk = g + t, st = gk % Cfg::K_STAGES;
        mbar_wait(&k_empty[st], ((gk / Cfg::K_STAGES) & 1) ^ 1);
        ATT_TRACE(0, gk, 0);
        uint8_t* s = k_s + st * Cfg::TILE_BYTES;
        const int krow = b * N + t * 64;
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
          tma_load_2d(&tm_k_hi, &k_full[st], s, HD + h * 64, krow);
          tma_load_2d(&tm_k_lo, &k_full[st], s + Cfg::KBOX, HD + h * 64, krow);
        }
        __syncwarp();
      };
      auto load_v = [&](int t) {
        const int gv = g + t, st = gv & 1;
        mbar_wait(&v_empty[st], ((gv >> 1) & 1) ^ 1);
        ATT_TRACE(0, gv, 1);
        uint8_t* s = v_s + st * Cfg::TILE_BYTES;
        const int vrow = (b * a.H + h) * 64;
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
          tma_load_2d(&tm_v_hi, &v_full[st], s, t * 64, vrow);
          tma_load_2d(&tm_v_lo, &v_full[st], s + Cfg::KBOX, t * 64, vrow);
        }
        __syncwarp();
      };
      load_k(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) load_k(t + 1);
        load_v(t);
      }
      g += T;
    }
  } else if (warp == 1) {
    // The whole warp runs this loop and ONE elected lane issues: with warp-uniform control flow the descriptors stay
    // in uniform registers. (Guarding the loop with `lane == 0` instead makes every tcgen05.mma a ~100-cycle
    // R2UR + waterfall sequence -- measured: that, not the tensor pipe or the softmax, bounded this kernel.)
    constexpr uint32_t idesc = make_idesc(0u, BM, 64);
    // S(g) = Q K(g)^T for tile g of the item whose query sits in buffer qb; `last` releases that buffer
    auto issue_qk = [&](int g, int qb, bool last) {
      const int st = g % Cfg::K_STAGES;
      ATT_TRACE(1, g, 0);
      mbar_wait(&k_full[st], (g / Cfg::K_STAGES) & 1);
      mbar_wait(&s_empty[g & 1], ((g >> 1) & 1) ^ 1);   // softmax(g-2) has S(g-2) in registers
      tcgen05_fence_after();
      ATT_TRACE(1, g, 1);
      const uint32_t q_u = smem_u32(q_s + qb * Cfg::Q_BYTES);
      const uint32_t k_u = smem_u32(k_s + st * Cfg::TILE_BYTES);
      if (elect_one()) {
        issue_slice_ss(tmem_base + (g & 1) * 64, q_u, q_u + BM * 128, k_u, k_u + Cfg::KBOX, idesc);
        umma_commit(&s_full[g & 1]);
        umma_commit(&k_empty[st]);
        if (last) umma_commit(&q_empty[qb]);
      }
      __syncwarp();
      ATT_TRACE(1, g, 2);
    };
    int g = 0, it = 0;
    if (blockIdx.x < items) {
      mbar_wait(&q_full[0], 0);
      issue_qk(0, 0, T == 1);
    }
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int qb = it % QS;
      const bool has_next = item + static_cast<int>(gridDim.x) < items;
      for (int t = 0; t < T; ++t, ++g) {
        // look one tile ahead, across the item boundary
        if (t + 1 < T) {
          issue_qk(g + 1, qb, t + 2 == T);
        } else if (has_next) {
          const int nqb = (it + 1) % QS;
          mbar_wait(&q_full[nqb], ((it + 1) / QS) & 1);
          issue_qk(g + 1, nqb, T == 1);
        }
        const int pb = g % NB, vb = g & 1;
        mbar_wait(&v_full[vb], (g >> 1) & 1);
        ATT_TRACE(1, g, 3);
        mbar_wait(&p_full[pb], (g / NB) & 1);
        tcgen05_fence_after();
        ATT_TRACE(1, g, 4);
        const uint32_t v_u = smem_u32(v_s + vb * Cfg::TILE_BYTES);
        if (elect_one()) {
          // O_g = P_lo V_hi + P_hi V_lo + P_hi V_hi over 4 k-steps of 16 keys (8 packed TMEM columns each)
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t pa = tmem_base + (pass == 0 ? kP_LO : kP_HI) + pb * 32;
            const uint64_t bd = make_sw128_kmajor_desc(v_u + (pass == 1 ? Cfg::KBOX : 0));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_ts(tmem_base + kO + pb * 64, pa + ks * 8, bd + 2 * ks, idesc, (pass | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&o_full[pb]);
          umma_commit(&v_empty[vb]);
        }
        __syncwarp();
        ATT_TRACE(1, g, 5);
      }
    }
  } else {
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;                // column group
    const int r = quad * 32 + lane;                 // row within the tile
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    // p = 256 exp(S sc - m) with the running maximum m of the logits, as d = fma(S, sc, -m) (ONE rounding of a quantity
    // that is small wherever p matters; sc = scale / 64 is a power of two for the usual head dim) followed by
    // 2^(d log2e + 8) on the SFU: two FFMAs and one MUFU.EX2 per element, relative error ~2^-22 -- the accuracy of
    // expf without its range reduction.
    constexpr float kLog2e = 1.4426950408889634f;
    const float sc = a.scale * (1.0f / (kQkPlaneScale * kQkPlaneScale));
    const float mask_to_raw = 1.0f / sc;        // additive key mask expressed in raw-accumulator units
    constexpr float kLogP = 8.0f;               // log2(kPScale): p is produced as 256 * exp(.) directly
    int g = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int qt = item % QT, bh = item / QT, h = bh % a.H, b = bh / a.H;
      const int i = qt * BM + r;
      const float* mask = a.key_mask ? a.key_mask + static_cast<long long>(b) * N : nullptr;
      const bool cls_warp = a.cls_p != nullptr && qt == 0 && quad == 0;
      float m = -INFINITY, l = 0.f;               // m: running maximum of the logits; l = 256 * sum of p
      float m_hist[2] = {-INFINITY, -INFINITY};   // running max at the NB previous tiles (indexed by g % NB)
      float o[KW];
#pragma unroll
      for (int d = 0; d < KW; ++d) o[d] = 0.f;

      // drain the partial product of tile u (buffer u % NB, relative to the running max m_u) into o (relative to m_now)
      auto drain = [&](int u, float m_u, float m_now) {
        mbar_wait_spin(&o_full[u % NB], (u / NB) & 1);
        tcgen05_fence_after();
        uint32_t v[KW];
        tmem_ld_cols<KW>(tmem_base + lane_off + kO + (u % NB) * 64 + grp * KW, v);
        tmem_ld_wait();
        const float f = ex2_approx((m_u - m_now) * kLog2e);
#pragma unroll
        for (int k = 0; k < KW; ++k) o[k] = fmaf(__uint_as_float(v[k]), f, o[k]);
      };

      for (int t = 0; t < T; ++t, ++g) {
        if (warp == 2) ATT_TRACE(2, g, 0);
        mbar_wait_spin(&s_full[g & 1], (g >> 1) & 1);
        tcgen05_fence_after();
        if (warp == 2) ATT_TRACE(2, g, 1);
        float s[KW];
        {
          uint32_t v[KW];
          tmem_ld_cols<KW>(tmem_base + lane_off + (g & 1) * 64 + grp * KW, v);
          tmem_ld_wait();
          ATT_TRACE(3, g, (warp - 2) & 7);
#pragma unroll
          for (int k = 0; k < KW; ++k) s[k] = __uint_as_float(v[k]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[g & 1]);
        if (warp == 2) ATT_TRACE(2, g, 2);

        const int j0 = t * 64 + grp * KW;
        if (mask != nullptr) {                       // additive key mask, in raw units (warp-uniform branch)
#pragma unroll
          for (int k = 0; k < KW; ++k)
            if (j0 + k < N) s[k] = fmaf(__ldg(mask + j0 + k), mask_to_raw, s[k]);
        }
        if (j0 + KW > N) {                           // last tile: keys past the sequence end
#pragma unroll
          for (int k = 0; k < KW; ++k) s[k] = (j0 + k < N) ? s[k] : -INFINITY;
        }
        if (a.causal && j0 + KW - 1 > i) {           // causal: keys after the query (key 0 is always visible)
#pragma unroll
          for (int k = 0; k < KW; ++k) s[k] = (j0 + k <= i) ? s[k] : -INFINITY;
        }
        float mx = s[0];
#pragma unroll
        for (int k = 1; k < KW; ++k) mx = fmaxf(mx, s[k]);
        mx *= sc;                                    // sc > 0: the maximum commutes with the scaling
        // joint maximum of the row over all column groups
        float* x = xch + (g & 1) * NG * BM;
        x[grp * BM + r] = mx;
        named_bar_sync(1 + quad, 32 * NG);
#pragma unroll
        for (int gg = 0; gg < NG; ++gg) mx = fmaxf(mx, x[gg * BM + r]);
        if (warp == 2) ATT_TRACE(2, g, 3);

        const float m_new = fmaxf(m, mx);
        const float corr = (m == -INFINITY) ? 0.f : ex2_approx((m - m_new) * kLog2e);
        const float neg_m = -m_new;
        // The two halves of the work -- the MUFU-bound exponentials and the FMA / TMEM-bound rescale-and-drain of the
        // older partial product -- are done in opposite order by even and odd column groups: the warps of one
        // scheduler run in lockstep behind the barrier above, and this keeps them off the same pipe.
        auto do_exp = [&]() {
          float ps = 0.f;
#pragma unroll
          for (int k = 0; k < KW; ++k) {
            s[k] = ex2_approx(fmaf(fmaf(s[k], sc, neg_m), kLog2e, kLogP));
            ps += s[k];
          }
          l = fmaf(l, corr, ps);
          // the CLS query row (row 0 of query tile 0) feeds cls_attn: keep its 256 p and the maximum they are relative
          // to; the statistics pass turns them into probabilities with the final log-sum-exp (vit.py:96-100)
          if (cls_warp) {               // warp-uniform: only the warps that own row 0 of query tile 0 get here
            // lane 0 holds the row: stage it in shared memory so that the warp writes it with one coalesced store
            float* st = cls_stage + grp * KW;
            if (lane == 0) {
#pragma unroll
              for (int k = 0; k < KW; k += 4) *reinterpret_cast<float4*>(st + k) = make_float4(s[k], s[k + 1], s[k + 2], s[k + 3]);
              if (grp == 0) a.cls_tile_max[static_cast<long long>(bh) * T + t] = m_new;
            }
            __syncwarp();
            if (lane < KW && j0 + lane < N) a.cls_p[static_cast<long long>(bh) * N + j0 + lane] = st[lane];
            __syncwarp();
          }
        };
        auto do_drain = [&]() {
          if (__any_sync(0xffffffffu, corr != 1.0f)) {   // the running maximum rarely moves after the first tiles
#pragma unroll
            for (int k = 0; k < KW; ++k) o[k] *= corr;   // o is now relative to m_new
          }
          // this tile's P buffer was last read by P(g-NB) V(g-NB): drain that partial, which also frees the buffer
          if (t >= NB) drain(g - NB, m_hist[g % NB], m_new);
        };
        if (grp & 1) {
          do_drain();
          do_exp();
        } else {
          do_exp();
          do_drain();
        }
        m = m_new;
        m_hist[g % NB] = m_new;
        if (warp == 2) ATT_TRACE(2, g, 5);

        // 256 p -> TMEM as packed fp16 hi / lo (A operand of the P V MMAs; low half = even key)
        {
          uint32_t hi[KW / 2], lo[KW / 2];
#pragma unroll
          for (int k = 0; k < KW / 2; ++k) {
            const float p0 = s[2 * k], p1 = s[2 * k + 1];
            const __half2 ph = __floats2half2_rn(p0, p1);
            const float2 pf = __half22float2(ph);
            const __half2 pl = __floats2half2_rn(p0 - pf.x, p1 - pf.y);
            hi[k] = *reinterpret_cast<const uint32_t*>(&ph);
            lo[k] = *reinterpret_cast<const uint32_t*>(&pl);
          }
          tmem_st_cols<KW / 2>(tmem_base + lane_off + kP_HI + (g % NB) * 32 + grp * (KW / 2), hi);
          tmem_st_cols<KW / 2>(tmem_base + lane_off + kP_LO + (g % NB) * 32 + grp * (KW / 2), lo);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g % NB]);
        if (warp == 2) ATT_TRACE(2, g, 6);
      }
      // g is now one past the item's last tile
      if (NB == 2 && T >= 2) drain(g - 2, m_hist[(g - 2) % NB], m);
      drain(g - 1, m_hist[(g - 1) % NB], m);
      // row sum and squared norm: combine the column groups in group order through the item-end exchange buffer (its
      // previous use, at the end of the last item, is separated from this one by at least one tile barrier)
      float* x = xch + 2 * NG * BM;
      x[grp * BM + r] = l;
      named_bar_sync(1 + quad, 32 * NG);
      float l_tot = 0.f;
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) l_tot += x[gg * BM + r];
      const float inv = 1.0f / (l_tot * kVPlaneScale);   // l carries the factor kPScale that o carries too
      float nsq = 0.f;
#pragma unroll
      for (int d = 0; d < KW; ++d) {
        o[d] *= inv;
        nsq = fmaf(o[d], o[d], nsq);
      }
      named_bar_sync(1 + quad, 32 * NG);               // every group has read l before the buffer is reused
      x[grp * BM + r] = nsq;
      named_bar_sync(1 + quad, 32 * NG);
      float nsq_tot = 0.f;
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) nsq_tot += x[gg * BM + r];   // dims in ascending order
      if (i < N) {
        __half* dst = a.out_f16 + b * a.bso + static_cast<long long>(i) * a.ldo + h * 64 + grp * KW;
#pragma unroll
        for (int d = 0; d < KW; d += 8) {
          __half2 h0 = __floats2half2_rn(o[d], o[d + 1]), h1 = __floats2half2_rn(o[d + 2], o[d + 3]);
          __half2 h2 = __floats2half2_rn(o[d + 4], o[d + 5]), h3 = __floats2half2_rn(o[d + 6], o[d + 7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2);
          pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(dst + d) = pk;
        }
        if (a.out_f32 != nullptr) {   // diagnostic value lane at scoring precision: the context is also kept in fp32
          float* dst32 = a.out_f32 + b * a.bso + static_cast<long long>(i) * a.ldo + h * 64 + grp * KW;
#pragma unroll
          for (int d = 0; d < KW; d += 4)
            *reinterpret_cast<float4*>(dst32 + d) = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
        }
        if (grp == 0) {
          const long long sidx = (static_cast<long long>(b) * a.H + h) * N + i;
          a.row_lse[sidx] = m + logf(l_tot * (1.0f / kPScale));
          a.out_norm[sidx] = sqrtf(nsq_tot);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
#ifdef MADTP_ATTN_TRACE
  if (warp == 0) ATT_TRACE(0, 63, 1);
#endif
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Pass 2 (persistent: every CTA walks a strided list of (sequence, 128-query tile, 128-key tile) work items)
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 consumers: TMEM lane quadrant = warp % 4 (query rows), column
//   half = (warp - 2) / 4 (64 of the tile's 128 keys).
//   The operand ring and the two S buffers are indexed by a head counter that runs ACROSS work items, so the producer
//   and the MMA warp are already several heads into the next item while the consumers reduce the current one (measured
//   on the one-item-per-CTA version: 4.4k cycles of start-up and 6.1k cycles of reduction around 15.6k cycles of MMAs).
//   The column sums over the 128 query rows are formed without shared-memory staging: a transposing butterfly over the
//   32 lanes of every warp (fixed order, deterministic) and a four-way sum over the lane quadrants.
// ------------------------------------------------------------------------------------------------
struct StatsSmem {
  static constexpr int BOX_BYTES = BM * 128;                // [128 rows x 64 halves]
  static constexpr int STAGE_BYTES = 4 * BOX_BYTES;         // Q hi, Q lo, K hi, K lo of one head
  static constexpr int STAGES = 3;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int MASK_OFF = BAR_OFF + 128;            // [2][128] additive key mask of the tile (log2 domain)
  static constexpr int PART_OFF = MASK_OFF + 2 * 128 * 4;   // [2][4][128] per-quadrant column sums
  static constexpr int TOTAL = PART_OFF + 2 * 4 * 128 * 4 + 1024;
  static constexpr int THREADS = 320;
};

// x[0..31] of the 32 lanes -> lane l returns sum over lanes of x[l] (31 shuffles; fixed summation tree)
__device__ __forceinline__ float warp_transpose_sum32(float (&x)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? x[i] : x[i + s];
      const float keep = up ? x[i + s] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return x[0];
}

__global__ void __launch_bounds__(StatsSmem::THREADS, 1)
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
  float* colmask = reinterpret_cast<float*>(smem + StatsSmem::MASK_OFF);   // [item parity][128]
  float* part = reinterpret_cast<float*>(smem + StatsSmem::PART_OFF);      // [item parity][quadrant][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.n_dev ? min(a.N, load_len(a.n_dev)) : a.N, H = a.H, HD = a.H * 64;
  if (a.n_dev) a.n_parts = (N + BM - 1) / BM;    // packed col_part [B, ceil(N/128), N] with the dynamic N
  const int NT = a.n_parts;                      // tiles per side
  const int items = NT * NT * a.B;
  constexpr float kLog2e = 1.4426950408889634f;

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
      mbar_init(&s_empty[s], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {   // warp-uniform control flow, one elected lane issues
    int gh = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int jt = item % NT, it = (item / NT) % NT, b = item / (NT * NT);
      for (int hh = 0; hh < H; ++hh, ++gh) {
        const int st = gh % StatsSmem::STAGES;
        mbar_wait(&empty[st], ((gh / StatsSmem::STAGES) & 1) ^ 1);
        uint8_t* s = smem + st * StatsSmem::STAGE_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[st], StatsSmem::STAGE_BYTES);
          tma_load_2d(&tm_hi, &full[st], s, hh * 64, b * N + it * BM);
          tma_load_2d(&tm_lo, &full[st], s + StatsSmem::BOX_BYTES, hh * 64, b * N + it * BM);
          tma_load_2d(&tm_hi, &full[st], s + 2 * StatsSmem::BOX_BYTES, HD + hh * 64, b * N + jt * BM);
          tma_load_2d(&tm_lo, &full[st], s + 3 * StatsSmem::BOX_BYTES, HD + hh * 64, b * N + jt * BM);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(0u, BM, BM);
    int gh = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      for (int hh = 0; hh < H; ++hh, ++gh) {
        const int hb = gh & 1, st = gh % StatsSmem::STAGES;
        mbar_wait(&s_empty[hb], ((gh >> 1) & 1) ^ 1);
        mbar_wait(&full[st], (gh / StatsSmem::STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t s = smem_u32(smem + st * StatsSmem::STAGE_BYTES);
        if (elect_one()) {
          issue_slice_ss(tmem_base + hb * 128, s, s + StatsSmem::BOX_BYTES, s + 2 * StatsSmem::BOX_BYTES,
                         s + 3 * StatsSmem::BOX_BYTES, idesc);
          umma_commit(&empty[st]);
          umma_commit(&s_full[hb]);
        }
        __syncwarp();
      }
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;          // row within the tile
    const int tid = threadIdx.x - 64;        // 0..255 among the consumers
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    // log P = S sc + mask - lse in the natural domain (sc = scale / 64: a power of two for the usual head dim, so the
    // product is exact and the only rounding is the final subtraction); one 2^(t log2e) per (i, j) after the head loop
    const float sc = a.scale * (1.0f / (kQkPlaneScale * kQkPlaneScale));
    // The CLS row first (it only needs the forward pass' outputs): the consumer warps would otherwise idle while the
    // operand pipeline fills. cls_attn[b,j] = sum_h P[b,h,0,j] * norm[b,h,j] / (sum_h' norm[b,h',j] + 1e-8), h ascending.
    {
      const int Tk = (N + 63) / 64;
      for (long long e = static_cast<long long>(blockIdx.x) * 256 + tid; e < static_cast<long long>(a.B) * N;
           e += static_cast<long long>(gridDim.x) * 256) {
        const int b = static_cast<int>(e / N), j = static_cast<int>(e - static_cast<long long>(b) * N);
        const long long base = static_cast<long long>(b) * H * N + j;
        float hs = 0.f;
        for (int hh = 0; hh < H; ++hh) hs += a.out_norm[base + static_cast<long long>(hh) * N];
        hs += 1e-8f;
        float acc = 0.f;
        for (int hh = 0; hh < H; ++hh) {
          const long long bh = static_cast<long long>(b) * H + hh;
          // P[b,h,0,j] = exp(logit_j - lse) = (256 p_j / 256) * exp(max of its key tile - lse of the CLS row)
          const float p = a.cls_p[bh * N + j] * (1.0f / kPScale) *
                          expf(a.cls_tile_max[bh * Tk + (j >> 6)] - a.row_lse[bh * N]);
          acc += p * (a.out_norm[base + static_cast<long long>(hh) * N] / hs);
        }
        a.cls_attn[e] = acc;
      }
    }
    int gh = 0, ip = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ip ^= 1) {
      const int jt = item % NT, it = (item / NT) % NT, b = item / (NT * NT);
      const int i = it * BM + r, j0 = jt * BM;
      const bool row_ok = (i >= 1) && (i < N);  // the CLS query row is excluded (reference vit.py:126)
      float* cm = colmask + ip * 128;
      if (tid < 128) {
        const int j = j0 + tid;
        cm[tid] = (j < N) ? (a.key_mask ? a.key_mask[static_cast<long long>(b) * N + j] : 0.f) : -INFINITY;
      }
      named_bar_sync(2, 256);   // mask of this item visible (its buffer was last read two items ago)
      float mx[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) mx[c] = -INFINITY;
      const float* lse_p = a.row_lse + static_cast<long long>(b) * H * N + (i < N ? i : 0);
      float lse_next = (i < N) ? __ldg(lse_p) : 0.f;
      for (int hh = 0; hh < H; ++hh, ++gh) {
        const int hb = gh & 1;
        const float lse = lse_next;
        if (hh + 1 < H && i < N) lse_next = __ldg(lse_p + static_cast<long long>(hh + 1) * N);   // hidden by this head
        mbar_wait(&s_full[hb], (gh >> 1) & 1);
        tcgen05_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + hb * 128 + half * 64, v0);
        tmem_ld_32x32b_x32(tmem_base + lane_off + hb * 128 + half * 64 + 32, v1);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[hb]);
        if (a.causal && j0 + BM - 1 > i) {   // warp-divergent only on the diagonal tiles
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int jj = j0 + half * 64 + k;
            const float t0 = fmaf(__uint_as_float(v0[k]), sc, cm[half * 64 + k]) - lse;
            const float t1 = fmaf(__uint_as_float(v1[k]), sc, cm[half * 64 + 32 + k]) - lse;
            mx[k] = fmaxf(mx[k], jj <= i ? t0 : -INFINITY);
            mx[32 + k] = fmaxf(mx[32 + k], jj + 32 <= i ? t1 : -INFINITY);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float t0 = fmaf(__uint_as_float(v0[k]), sc, cm[half * 64 + k]) - lse;
            const float t1 = fmaf(__uint_as_float(v1[k]), sc, cm[half * 64 + 32 + k]) - lse;
            mx[k] = fmaxf(mx[k], t0);
            mx[32 + k] = fmaxf(mx[32 + k], t1);
          }
        }
      }
      // column sums over this warp's 32 query rows (lane l ends up with columns l and 32 + l of its half), then over
      // the four lane quadrants in quadrant order
      float* pt = part + ip * 4 * 128;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float p[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) p[c] = row_ok ? ex2_approx(mx[g * 32 + c] * kLog2e) : 0.f;
        const float tot = warp_transpose_sum32(p, lane);
        pt[quad * 128 + half * 64 + g * 32 + lane] = tot;
      }
      named_bar_sync(3, 256);
      if (tid < 128) {
        const int j = j0 + tid;
        const float sum = ((pt[tid] + pt[128 + tid]) + pt[256 + tid]) + pt[384 + tid];
        if (j < N) a.col_part[(static_cast<long long>(b) * a.n_parts + it) * N + j] = sum;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int check_tc(const AttnTcArgs& a) {
  MADTP_CHECK_ARG(a.qk_hi && a.qk_lo, "attn_tc: null q/k planes");
  MADTP_CHECK_ARG(a.B >= 0 && a.H > 0 && a.N > 0, "attn_tc: bad shape B=%d H=%d N=%d", a.B, a.H, a.N);
  MADTP_CHECK_ARG(a.ld_qk >= 2LL * a.H * 64 && a.ld_qk % 8 == 0, "attn_tc: bad q/k leading dimension");
  MADTP_CHECK_ARG(a.B <= 65535 && a.H <= 65535, "attn_tc: B and H must fit the grid limits");
  return kOk;
}

template <int KW, int NB>
static int launch_fwd_kw(const CUtensorMap& tq_hi, const CUtensorMap& tq_lo, const CUtensorMap& tk_hi,
                         const CUtensorMap& tk_lo, const CUtensorMap& tv_hi, const CUtensorMap& tv_lo,
                         const AttnTcArgs& a, cudaStream_t stream) {
  using Cfg = FwdCfg<KW, NB>;
  MADTP_SMEM_ATTR_ONCE(Cfg::TOTAL, attn_fwd_tc_kernel<KW, NB>);
  const long long items = static_cast<long long>((a.N + BM - 1) / BM) * a.H * a.B;
  const long long slots = static_cast<long long>(num_sms()) * Cfg::CTAS_PER_SM;
  const int grid = static_cast<int>(items < slots ? items : slots);
  attn_fwd_tc_kernel<KW, NB><<<grid, Cfg::THREADS, Cfg::TOTAL, stream>>>(tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

int launch_attn_fwd_tc(const AttnTcArgs& a, cudaStream_t stream) {
  int st = check_tc(a);
  if (st != kOk) return st;
  MADTP_CHECK_ARG(a.vt_hi && a.vt_lo && a.ld_vt >= a.N && a.ld_vt % 8 == 0, "attn_fwd_tc: bad V^T planes");
  MADTP_CHECK_ARG(a.out_f16 && a.ldo % 8 == 0 && a.bso % 8 == 0 && a.row_lse && a.out_norm, "attn_fwd_tc: bad outputs");
  MADTP_CHECK_ARG((a.cls_p == nullptr) == (a.cls_tile_max == nullptr), "attn_fwd_tc: cls_p / cls_tile_max come in pairs");
  if (a.B == 0) return kOk;
  const long long rows = static_cast<long long>(a.B) * a.N;
  CUtensorMap tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo;
  if ((st = make_tmap(&tq_hi, a.qk_hi, false, rows, 2LL * a.H * 64, a.ld_qk, BM)) != kOk) return st;
  if ((st = make_tmap(&tq_lo, a.qk_lo, false, rows, 2LL * a.H * 64, a.ld_qk, BM)) != kOk) return st;
  if ((st = make_tmap(&tk_hi, a.qk_hi, false, rows, 2LL * a.H * 64, a.ld_qk, 64)) != kOk) return st;
  if ((st = make_tmap(&tk_lo, a.qk_lo, false, rows, 2LL * a.H * 64, a.ld_qk, 64)) != kOk) return st;
  const long long vrows = static_cast<long long>(a.B) * a.H * 64;
  if ((st = make_tmap(&tv_hi, a.vt_hi, false, vrows, a.N, a.ld_vt, 64)) != kOk) return st;
  if ((st = make_tmap(&tv_lo, a.vt_lo, false, vrows, a.N, a.ld_vt, 64)) != kOk) return st;
  // variants (development switch): 0 = 16 softmax warps x 16 keys, one CTA per SM; 1 = 8 warps x 32 keys, one CTA per
  // SM; 2 = 8 warps x 32 keys, two co-resident CTAs per SM (single-buffered P / O partial, 256 TMEM columns each)
  const char* var = getenv("MADTP_ATTN_VARIANT");
  const int v = var ? atoi(var) : 1;
  if (v == 2) return launch_fwd_kw<32, 1>(tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, a, stream);
  if (v == 1) return launch_fwd_kw<32, 2>(tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, a, stream);
  return launch_fwd_kw<16, 2>(tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, a, stream);
}

int launch_attn_stats_tc(const AttnTcArgs& a, cudaStream_t stream) {
  int st = check_tc(a);
  if (st != kOk) return st;
  MADTP_CHECK_ARG(a.row_lse && a.out_norm && a.col_part && a.cls_attn && a.cls_p && a.cls_tile_max,
                  "attn_stats_tc: null statistics buffer");
  MADTP_CHECK_ARG(a.n_parts == (a.N + BM - 1) / BM, "attn_stats_tc: n_parts must be ceil(N/128)");
  MADTP_CHECK_ARG(a.N <= 8192, "attn_stats_tc: sequence too long for the CLS-row kernel");
  if (a.B == 0) return kOk;
  const long long rows = static_cast<long long>(a.B) * a.N;
  CUtensorMap t_hi, t_lo;
  if ((st = make_tmap(&t_hi, a.qk_hi, false, rows, 2LL * a.H * 64, a.ld_qk, BM)) != kOk) return st;
  if ((st = make_tmap(&t_lo, a.qk_lo, false, rows, 2LL * a.H * 64, a.ld_qk, BM)) != kOk) return st;
  MADTP_SMEM_ATTR_ONCE(StatsSmem::TOTAL, attn_stats_tc_kernel);
  const long long items = static_cast<long long>(a.n_parts) * a.n_parts * a.B;
  const int grid = static_cast<int>(items < num_sms() ? items : num_sms());
  attn_stats_tc_kernel<<<grid, StatsSmem::THREADS, StatsSmem::TOTAL, stream>>>(t_hi, t_lo, a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

}  // namespace madtp
