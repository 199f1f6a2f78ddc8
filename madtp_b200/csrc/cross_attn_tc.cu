// Tensor-core cross-attention of a few text queries over the image tokens (value lane: fp16 operands, fp32 accumulate).
// Replaces the attention of models/nlvr_encoder.py:174-219 / models/med.py:175-217 when is_cross_attention: nothing
// downstream reads its probabilities (the pruning statistics come from the text SELF-attention), so one fp16 pass is
// enough -- the same lane as the projections on either side of it.
//
//   one CTA per (head, sequence), Lq <= 128 queries, ANY number of keys, head dim 64. The keys are walked in blocks of
//   128 with an online softmax (running maximum and sum per query row, fp32):
//   warp 0   TMA: Q [128 x 64] once, then per key block K as two [64 keys x 64] boxes and V^T as two [64 dims x 64 keys]
//            boxes through a two-deep ring
//   warp 1   S = Q K_blk^T (4 MMAs, N = 128) -> TMEM; after the softmax O_blk = P V_blk (8 MMAs, P from TMEM)
//   warps 2..9  softmax in the log2 domain, thread = (query row, half of the block's keys); two passes over S in TMEM
//            (block maximum, then exponentials relative to the new running maximum) so that S never lives in
//            registers; 256 p goes back to TMEM as packed fp16 IN PLACE over the consumed part of S; the block's partial
//            product is drained from TMEM and added to the fp32 running output (rescaled when the maximum moved).
//   256 TMEM columns per CTA (S 0..127, O 192..255): two CTAs share an SM and fill each other's pipeline bubbles.
//   Only the TMEM lane quadrants that hold real queries do any softmax work (20 text tokens -> one quadrant).
// V arrives transposed (keys contiguous), produced by running the value projection as W_v . X^T; its bias is added to
// the normalised output instead (the probabilities of a row sum to one).
// Lq and Nk may be device-resident (CrossTcArgs::lq_dev / nk_dev): the pruned lengths are then never read by the host.
#include "attention.cuh"
#include "gemm.cuh"

namespace madtp {

namespace {

struct CrossSmem {
  static constexpr int Q_BYTES = 128 * 128;
  static constexpr int KBOX = 64 * 128;
  static constexpr int KB = 128;                      // keys per block
  static constexpr int STAGE = 4 * KBOX;              // K box 0, K box 1, V^T box 0, V^T box 1
  static constexpr int STAGES = 2;
  static constexpr int K_OFF = Q_BYTES;
  static constexpr int BAR_OFF = K_OFF + STAGES * STAGE;
  static constexpr int XCH_OFF = BAR_OFF + 128;
  static constexpr int TOTAL = XCH_OFF + 3 * 2 * 128 * 4 + 1024;   // + slack for the manual 1024-byte alignment
  static constexpr int THREADS = 320;
};

}  // namespace

__global__ void __launch_bounds__(CrossSmem::THREADS, 2)
cross_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, CrossTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint8_t* kv_s = smem + CrossSmem::K_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CrossSmem::BAR_OFF);
  uint64_t* q_full = bars;          // Q landed
  uint64_t* kv_full = bars + 1;     // [2] K / V^T block landed
  uint64_t* kv_empty = bars + 3;    // [2] the block's MMAs retired
  uint64_t* s_full = bars + 5;      // S(blk) complete in TMEM
  uint64_t* p_full = bars + 6;      // P(blk) stored by every active softmax warp
  uint64_t* o_full = bars + 7;      // O(blk) complete in TMEM
  uint64_t* o_empty = bars + 8;     // O(blk) drained by every active softmax warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* xch = reinterpret_cast<float*>(smem + CrossSmem::XCH_OFF);   // [blk parity | sum][half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int Lq = a.lq_dev ? min(a.Lq, load_len(a.lq_dev)) : a.Lq;
  const int Nk = a.k_len ? min(a.Nk, __ldg(a.k_len + b)) : (a.nk_dev ? min(a.Nk, load_len(a.nk_dev)) : a.Nk);
  if (a.lq_dev) a.bso = static_cast<long long>(Lq) * a.ldo;             // packed output
  if (a.nk_dev && a.k_rows_per_batch != 0) a.k_rows_per_batch = a.vt_cols_per_batch = (Nk + 7) & ~7;
  const int NB = (Nk + CrossSmem::KB - 1) / CrossSmem::KB;              // key blocks
  const int active_warps = 2 * ((Lq + 31) / 32);                        // softmax warps that own real query rows
  constexpr uint32_t kO = 192;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, active_warps);
    mbar_init(o_full, 1);
    mbar_init(o_empty, active_warps);
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
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, CrossSmem::Q_BYTES);
      tma_load_2d(&tm_q, q_full, q_s, h * 64, b * Lq);
    }
    __syncwarp();
    const int krow0 = a.k_start ? __ldg(a.k_start + b) : b * a.k_rows_per_batch;
    const int vcol0 = a.k_start ? krow0 : b * a.vt_cols_per_batch;
    for (int blk = 0; blk < NB; ++blk) {
      const int st = blk & 1;
      mbar_wait(&kv_empty[st], ((blk >> 1) & 1) ^ 1);
      uint8_t* s = kv_s + st * CrossSmem::STAGE;
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[st], CrossSmem::STAGE);
        for (int t = 0; t < 2; ++t) {
          tma_load_2d(&tm_k, &kv_full[st], s + t * CrossSmem::KBOX, h * 64, krow0 + blk * CrossSmem::KB + t * 64);
          tma_load_2d(&tm_v, &kv_full[st], s + (2 + t) * CrossSmem::KBOX, vcol0 + blk * CrossSmem::KB + t * 64, h * 64);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc(0u, 128, CrossSmem::KB);
    constexpr uint32_t idesc_o = make_idesc(0u, 128, 64);
    mbar_wait(q_full, 0);
    const uint64_t qd = make_sw128_kmajor_desc(smem_u32(q_s));
    for (int blk = 0; blk < NB; ++blk) {
      const int st = blk & 1;
      mbar_wait(&kv_full[st], (blk >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t s = smem_u32(kv_s + st * CrossSmem::STAGE);
      // S(blk) may overwrite S / P(blk-1): every softmax warp stored P(blk-1) before P V(blk-1) was issued, and the
      // tensor pipe executes in issue order
      if (elect_one()) {
        const uint64_t kd = make_sw128_kmajor_desc(s);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_base, qd + 2 * ks, kd + 2 * ks, idesc_s, ks != 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, blk & 1);
      if (blk > 0) mbar_wait(o_empty, (blk - 1) & 1);    // O(blk-1) drained before it is overwritten
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint64_t vd = make_sw128_kmajor_desc(s + (2 + t) * CrossSmem::KBOX);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)   // 16 keys = 8 packed columns at the start of the half that owns them
            umma_f16_ts(tmem_base + kO, tmem_base + t * 64 + ks * 8, vd + 2 * ks, idesc_o, (t | ks) != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    if (quad * 32 < Lq) {                        // warp-uniform: this TMEM lane quadrant holds real queries
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      constexpr float kLog2e = 1.4426950408889634f;
      const float c1 = a.scale * kLog2e;
      const float mask_to_raw = kLog2e / c1;
      const float* mask = a.key_mask ? a.key_mask + static_cast<long long>(b) * Nk : nullptr;
      const float k0_raw = a.key0_bias ? __ldg(a.key0_bias + b) / a.scale : 0.f;   // extra logit of key 0, raw units
      const int col0 = half * 64;                // this thread's 64 keys of the block: S columns [col0, col0 + 64)
      float m = -INFINITY, l = 0.f;              // running maximum (log2 domain) and sum of 256 p
      float o[32];
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] = 0.f;
      for (int blk = 0; blk < NB; ++blk) {
        mbar_wait_spin(s_full, blk & 1);
        tcgen05_fence_after();
        const int jb = blk * CrossSmem::KB + col0;
        // one 32-column chunk of this thread's row, masked and bounded, in raw accumulator units
        auto load_chunk = [&](int c, float (&s)[32]) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_off + col0 + c * 32, v);
          tmem_ld_wait();
          const int j0 = jb + c * 32;
          if (mask != nullptr || j0 + 32 > Nk || j0 == 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int j = j0 + k;
              float x = __uint_as_float(v[k]);
              if (mask != nullptr && j < Nk) x = fmaf(__ldg(mask + j), mask_to_raw, x);
              if (j == 0) x += k0_raw;
              s[k] = (j < Nk) ? x : -INFINITY;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) s[k] = __uint_as_float(v[k]);
          }
        };
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float s[32];
          load_chunk(c, s);
#pragma unroll
          for (int k = 0; k < 32; ++k) mx = fmaxf(mx, s[k]);
        }
        mx *= c1;
        float* x = xch + (blk & 1) * 256;
        x[half * 128 + r] = mx;
        named_bar_sync(1 + quad, 64);
        const float m_new = fmaxf(m, fmaxf(mx, x[(half ^ 1) * 128 + r]));
        const float corr = ex2_approx(m - m_new);          // 2^-inf = 0 on the first block
        const float off = 8.0f - m_new;                    // p is produced as 256 * 2^(y - max)
        float ps = 0.f;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float s[32];
          load_chunk(c, s);
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float p0 = ex2_approx(fmaf(s[2 * k], c1, off)), p1 = ex2_approx(fmaf(s[2 * k + 1], c1, off));
            ps += p0 + p1;
            const __half2 ph = __floats2half2_rn(p0, p1);
            pk[k] = *reinterpret_cast<const uint32_t*>(&ph);
          }
          tmem_st_32x32b_x16(tmem_base + lane_off + col0 + c * 16, pk);   // over this thread's consumed S columns
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        l = fmaf(l, corr, ps);
        if (__any_sync(0xffffffffu, corr != 1.0f)) {
#pragma unroll
          for (int d = 0; d < 32; ++d) o[d] *= corr;
        }
        m = m_new;
        mbar_wait_spin(o_full, blk & 1);
        tcgen05_fence_after();
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_off + kO + half * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 32; ++d) o[d] += __uint_as_float(v[d]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
      float* xs = xch + 512;
      xs[half * 128 + r] = l;
      named_bar_sync(1 + quad, 64);
      const float inv = 1.0f / (xs[r] + xs[128 + r]);
      if (r < Lq) {
        const float* vb = a.v_bias ? a.v_bias + h * 64 + half * 32 : nullptr;
        __half* dst = a.out_f16 + b * a.bso + static_cast<long long>(r) * a.ldo + h * 64 + half * 32;
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = fmaf(o[d + e], inv, vb ? __ldg(vb + d + e) : 0.f);
          __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
          __half2 h2 = __floats2half2_rn(y[4], y[5]), h3 = __floats2half2_rn(y[6], y[7]);
          uint4 pk4;
          pk4.x = *reinterpret_cast<uint32_t*>(&h0);
          pk4.y = *reinterpret_cast<uint32_t*>(&h1);
          pk4.z = *reinterpret_cast<uint32_t*>(&h2);
          pk4.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(dst + d) = pk4;
        }
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

int launch_cross_attn_tc(const CrossTcArgs& a, cudaStream_t stream) {
  MADTP_CHECK_ARG(a.q && a.k && a.vt && a.out_f16, "cross_attn_tc: null pointer");
  MADTP_CHECK_ARG(a.B >= 0 && a.H > 0 && a.Lq > 0 && a.Lq <= 128 && a.Nk > 0,
                  "cross_attn_tc: needs 1 <= Lq <= 128 and Nk >= 1 (Lq=%d Nk=%d)", a.Lq, a.Nk);
  MADTP_CHECK_ARG(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ld_vt % 8 == 0 && a.ldo % 8 == 0 && a.bso % 8 == 0,
                  "cross_attn_tc: leading dimensions must be multiples of 8 halves");
  MADTP_CHECK_ARG(a.k_rows_per_batch == 0 || a.k_rows_per_batch >= a.Nk || a.k_start != nullptr,
                  "cross_attn_tc: k rows per batch is >= Nk or 0");
  // TMA box origins must be 16-byte aligned in global memory: a sequence's keys start at a multiple of 8 columns
  MADTP_CHECK_ARG(a.vt_cols_per_batch == 0 || (a.vt_cols_per_batch >= a.Nk && a.vt_cols_per_batch % 8 == 0),
                  "cross_attn_tc: V^T columns per batch must be 0 or a multiple of 8 that is >= Nk (got %d)",
                  a.vt_cols_per_batch);
  MADTP_CHECK_ARG(a.B <= 65535, "cross_attn_tc: B must fit the grid limits");
  MADTP_CHECK_ARG((a.k_start == nullptr) == (a.k_len == nullptr) && (a.key0_bias == nullptr || a.k_start != nullptr),
                  "cross_attn_tc: k_start / k_len come together (key0_bias only with them)");
  MADTP_CHECK_ARG(a.k_start == nullptr || (a.key_mask == nullptr && a.nk_dev == nullptr),
                  "cross_attn_tc: ragged keys exclude key_mask and nk_dev");
  if (a.B == 0) return kOk;
  CUtensorMap tq, tk, tv;
  int st;
  // ragged: k_rows_per_batch / vt_cols_per_batch carry the TOTAL packed rows / columns
  const long long k_rows = a.k_start ? a.k_rows_per_batch
                                     : (a.k_rows_per_batch ? static_cast<long long>(a.B) * a.k_rows_per_batch : a.Nk);
  const long long v_cols = a.k_start ? a.vt_cols_per_batch
                                     : (a.vt_cols_per_batch ? static_cast<long long>(a.B) * a.vt_cols_per_batch : a.Nk);
  if ((st = make_tmap(&tq, a.q, false, static_cast<long long>(a.B) * a.Lq, a.H * 64LL, a.ldq, 128)) != kOk) return st;
  if ((st = make_tmap(&tk, a.k, false, k_rows, a.H * 64LL, a.ldk, 64)) != kOk) return st;
  if ((st = make_tmap(&tv, a.vt, false, a.H * 64LL, v_cols, a.ld_vt, 64)) != kOk) return st;
  MADTP_SMEM_ATTR_ONCE(CrossSmem::TOTAL, cross_attn_tc_kernel);
  cross_attn_tc_kernel<<<dim3(a.H, a.B), CrossSmem::THREADS, CrossSmem::TOTAL, stream>>>(tq, tk, tv, a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

}  // namespace madtp
