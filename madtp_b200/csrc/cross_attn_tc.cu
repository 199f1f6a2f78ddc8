// Tensor-core cross-attention of a few text queries over the image tokens (value lane: fp16 operands, fp32 accumulate).
// Replaces the attention of models/nlvr_encoder.py:174-219 / models/med.py:175-217 when is_cross_attention: nothing
// downstream reads its probabilities (the pruning statistics come from the text SELF-attention), so one fp16 pass is
// enough -- the same lane as the projections on either side of it.
//
//   one CTA per (head, sequence), Lq <= 128 queries, Nk <= 256 keys, head dim 64: the whole problem is one tile.
//   warp 0   TMA: Q [128 x 64], K as up to four [64 keys x 64] boxes, V^T as up to four [64 dims x 64 keys] boxes
//   warp 1   S = Q K^T (4 MMAs, N = 64 * key tiles) -> TMEM; after the softmax O = P V (4 MMAs per key tile, P from TMEM)
//   warps 2..9  softmax in the log2 domain, thread = (query row, half of the keys); two passes over S in TMEM (maximum,
//            then exponentials) so that S never has to live in registers; 256 p goes back to TMEM as packed fp16, in
//            place over the consumed part of S (256 TMEM columns in all: two CTAs per SM).
//   Only the TMEM lane quadrants that hold real queries do any softmax work (20 text tokens -> one quadrant).
// V arrives transposed (keys contiguous), produced by running the value projection as W_v . X^T; its bias is added to
// the normalised output instead (the probabilities of a row sum to one).
#include "attention.cuh"
#include "gemm.cuh"

namespace madtp {

namespace {

struct CrossSmem {
  static constexpr int Q_BYTES = 128 * 128;
  static constexpr int KBOX = 64 * 128;
  static constexpr int K_OFF = Q_BYTES;
  static constexpr int V_OFF = K_OFF + 4 * KBOX;
  static constexpr int BAR_OFF = V_OFF + 4 * KBOX;
  static constexpr int XCH_OFF = BAR_OFF + 64;
  static constexpr int TOTAL = XCH_OFF + 2 * 2 * 128 * 4 + 1024;   // + slack for the manual 1024-byte alignment
  static constexpr int THREADS = 320;
};

}  // namespace

__global__ void __launch_bounds__(CrossSmem::THREADS, 2)
cross_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, CrossTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint8_t* k_s = smem + CrossSmem::K_OFF;
  uint8_t* v_s = smem + CrossSmem::V_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CrossSmem::BAR_OFF);
  uint64_t* full = bars;         // operands landed
  uint64_t* s_full = bars + 1;   // S complete in TMEM
  uint64_t* p_full = bars + 2;   // P stored by all eight softmax warps
  uint64_t* o_full = bars + 3;   // O complete in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* xch = reinterpret_cast<float*>(smem + CrossSmem::XCH_OFF);   // [max | sum][half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int Lq = a.Lq, Nk = a.Nk;
  const int KT = (Nk + 63) / 64;   // key tiles (1..4)
  const int NKP = KT * 64;         // padded key count = N of the first MMA
  // 256 TMEM columns, so that two CTAs share an SM: S at 0 (N = NKP <= 256 columns). The packed P of each column half
  // is written IN PLACE over the part of that half's S the writing thread has already consumed (half 0: columns
  // [0, NH/2), half 1: [NH, NH + NH/2), NH = NKP/2); O at 192..255, which the MMA only writes after every softmax warp
  // is done with S.
  constexpr uint32_t kO = 192;
  const int NH = NKP / 2;                  // keys per column half (a multiple of 32)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
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

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(full, CrossSmem::Q_BYTES + 2 * KT * CrossSmem::KBOX);
      tma_load_2d(&tm_q, full, q_s, h * 64, b * Lq);
      for (int t = 0; t < KT; ++t) {
        tma_load_2d(&tm_k, full, k_s + t * CrossSmem::KBOX, h * 64, b * a.k_rows_per_batch + t * 64);
        tma_load_2d(&tm_v, full, v_s + t * CrossSmem::KBOX, b * a.vt_cols_per_batch + t * 64, h * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    mbar_wait(full, 0);
    tcgen05_fence_after();
    const uint32_t idesc_s = make_idesc(0u, 128, static_cast<uint32_t>(NKP));
    constexpr uint32_t idesc_o = make_idesc(0u, 128, 64);
    if (elect_one()) {
      const uint64_t qd = make_sw128_kmajor_desc(smem_u32(q_s)), kd = make_sw128_kmajor_desc(smem_u32(k_s));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_base, qd + 2 * ks, kd + 2 * ks, idesc_s, ks != 0 ? 1u : 0u);
      umma_commit(s_full);
    }
    __syncwarp();
    mbar_wait(p_full, 0);
    tcgen05_fence_after();
    if (elect_one()) {
      for (int t = 0; t < KT; ++t) {
        const uint64_t vd = make_sw128_kmajor_desc(smem_u32(v_s + t * CrossSmem::KBOX));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int key0 = t * 64 + ks * 16;   // 16 keys = 8 packed columns, in the half that owns them
          const uint32_t pcol = key0 < NH ? key0 / 2 : NH + (key0 - NH) / 2;
          umma_f16_ts(tmem_base + kO, tmem_base + pcol, vd + 2 * ks, idesc_o, (t | ks) != 0 ? 1u : 0u);
        }
      }
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const bool active = quad * 32 < Lq;        // warp-uniform: this TMEM lane quadrant holds real queries
    if (!active) {
      if (lane == 0) mbar_arrive(p_full);      // rows of this quadrant are never read back
    } else {
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      constexpr float kLog2e = 1.4426950408889634f;
      const float c1 = a.scale * kLog2e;
      const float mask_to_raw = kLog2e / c1;
      const float* mask = a.key_mask ? a.key_mask + static_cast<long long>(b) * Nk : nullptr;
      const int col0 = half * NH;
      mbar_wait_spin(s_full, 0);
      tcgen05_fence_after();
      // one 32-column chunk of this thread's row, masked and bounded, in raw accumulator units
      auto load_chunk = [&](int c, float (&s)[32]) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + col0 + c * 32, v);
        tmem_ld_wait();
        const int j0 = col0 + c * 32;
        if (mask != nullptr || j0 + 32 > Nk) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int j = j0 + k;
            float x = __uint_as_float(v[k]);
            if (mask != nullptr && j < Nk) x = fmaf(__ldg(mask + j), mask_to_raw, x);
            s[k] = (j < Nk) ? x : -INFINITY;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) s[k] = __uint_as_float(v[k]);
        }
      };
      float mx = -INFINITY;
      for (int c = 0; c < NH / 32; ++c) {
        float s[32];
        load_chunk(c, s);
#pragma unroll
        for (int k = 0; k < 32; ++k) mx = fmaxf(mx, s[k]);
      }
      mx *= c1;
      xch[half * 128 + r] = mx;
      named_bar_sync(1 + quad, 64);
      mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);
      const float off = 8.0f - mx;             // p is produced as 256 * 2^(y - max)
      float l = 0.f;
      for (int c = 0; c < NH / 32; ++c) {
        float s[32];
        load_chunk(c, s);
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float p0 = ex2_approx(fmaf(s[2 * k], c1, off)), p1 = ex2_approx(fmaf(s[2 * k + 1], c1, off));
          l += p0 + p1;
          const __half2 ph = __floats2half2_rn(p0, p1);
          pk[k] = *reinterpret_cast<const uint32_t*>(&ph);
        }
        tmem_st_32x32b_x16(tmem_base + lane_off + col0 + c * 16, pk);   // over this thread's consumed S columns
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      xch[256 + half * 128 + r] = l;
      named_bar_sync(1 + quad, 64);
      const float inv = 1.0f / (xch[256 + r] + xch[256 + 128 + r]);
      mbar_wait_spin(o_full, 0);
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + lane_off + kO + half * 32, v);
      tmem_ld_wait();
      if (r < Lq) {
        const float* vb = a.v_bias ? a.v_bias + h * 64 + half * 32 : nullptr;
        __half* dst = a.out_f16 + b * a.bso + static_cast<long long>(r) * a.ldo + h * 64 + half * 32;
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf(__uint_as_float(v[d + e]), inv, vb ? __ldg(vb + d + e) : 0.f);
          __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
          __half2 h2 = __floats2half2_rn(o[4], o[5]), h3 = __floats2half2_rn(o[6], o[7]);
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
  MADTP_CHECK_ARG(a.B >= 0 && a.H > 0 && a.Lq > 0 && a.Lq <= 128 && a.Nk > 0 && a.Nk <= 256,
                  "cross_attn_tc: needs Lq <= 128 and Nk <= 256 (Lq=%d Nk=%d)", a.Lq, a.Nk);
  MADTP_CHECK_ARG(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ld_vt % 8 == 0 && a.ldo % 8 == 0 && a.bso % 8 == 0,
                  "cross_attn_tc: leading dimensions must be multiples of 8 halves");
  MADTP_CHECK_ARG(a.k_rows_per_batch == 0 || a.k_rows_per_batch >= a.Nk, "cross_attn_tc: k rows per batch is >= Nk or 0");
  // TMA box origins must be 16-byte aligned in global memory: a sequence's keys start at a multiple of 8 columns
  MADTP_CHECK_ARG(a.vt_cols_per_batch == 0 || (a.vt_cols_per_batch >= a.Nk && a.vt_cols_per_batch % 8 == 0),
                  "cross_attn_tc: V^T columns per batch must be 0 or a multiple of 8 that is >= Nk (got %d)",
                  a.vt_cols_per_batch);
  MADTP_CHECK_ARG(a.B <= 65535, "cross_attn_tc: B must fit the grid limits");
  if (a.B == 0) return kOk;
  CUtensorMap tq, tk, tv;
  int st;
  const long long k_rows = a.k_rows_per_batch ? static_cast<long long>(a.B) * a.k_rows_per_batch : a.Nk;
  const long long v_cols = a.vt_cols_per_batch ? static_cast<long long>(a.B) * a.vt_cols_per_batch : a.Nk;
  if ((st = make_tmap(&tq, a.q, false, static_cast<long long>(a.B) * a.Lq, a.H * 64LL, a.ldq, 128)) != kOk) return st;
  if ((st = make_tmap(&tk, a.k, false, k_rows, a.H * 64LL, a.ldk, 64)) != kOk) return st;
  if ((st = make_tmap(&tv, a.vt, false, a.H * 64LL, v_cols, a.ld_vt, 64)) != kOk) return st;
  MADTP_SMEM_ATTR_ONCE(CrossSmem::TOTAL, cross_attn_tc_kernel);
  cross_attn_tc_kernel<<<dim3(a.H, a.B), CrossSmem::THREADS, CrossSmem::TOTAL, stream>>>(tq, tk, tv, a);
  MADTP_LAUNCH_CHECK();
  return kOk;
}

}  // namespace madtp
