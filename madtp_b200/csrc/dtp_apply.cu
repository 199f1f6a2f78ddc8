// Fused Dynamic Token Pruning "apply" step: top-k selection, stream compaction, merged token and the LayerNorm that
// follows -- ONE kernel for what the reference does with two topk calls (one a full sort), two index-expanding gathers,
// a bmm, two cats and nn.LayerNorm (models/vit.py:148-161,202,205; models/nlvr_encoder.py:434-454; models/med.py:
// 371-390; models/utils.py:13-33), and what round 1 did with three kernels and a host read-back.
//
//   grid (S, B): S CTAs of 1024 threads per sequence (S = floor(SMs / B), 1..8: a single wave). Every CTA of a sequence recomputes the
//   selection (a few microseconds, no cross-CTA traffic), then the S * 32 warps of the sequence copy its surviving
//   rows -- one warp per row, the whole row in registers (128-bit loads and stores), so the LayerNorm of the row costs
//   no extra memory pass.
//
//   selection (n <= 1024 scores in shared memory, thread = token):
//     * order-preserving uint32 keys, 4-pass MSB-first RADIX SELECT (256-bin shared-memory histograms, warp-shuffle
//       suffix scan) finds the k-th largest key; ties are broken by the lower token index (scan of the "equal" flags),
//       exactly like the rank-count kernel it replaces;
//     * block-wide exclusive scan of the keep flags = slot of every survivor (ascending token order = stream compaction);
//     * merge weights S_j / (sum of the pruned scores + 1e-8), the sum accumulated in fp64;
//     * text encoders: the additive key mask travels with the tokens (mask_mode 1 = nlvr_encoder rank gather,
//       2 = med.py) -- for those short sequences (n <= 64) the ranks come from a count;
//     * k = *topk read on the device; k <= max_keep or n - k <= 1 means "nothing is pruned" (a plain copy + LayerNorm).
//   merged token: eight warps of the first CTA reduce the pruned rows in the same fixed order as dtp_gather_kernel
//   (partial p sums rows p, p + 8, ...; partials combined in order), so results are bit-identical to the unfused path.
//   Device-resident lengths (n_dev): packed input / output, next length and trajectory written for the next layer.
#include "dtp.cuh"

namespace madtp {

namespace {

__device__ __forceinline__ uint32_t ordered_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // larger float <=> larger key
}

// block-wide exclusive scan of a 0/1 flag over 1024 threads (thread = token); returns the exclusive prefix, *total gets
// the block total. wsum: 32 ints of shared memory.
__device__ __forceinline__ int block_excl_scan(int flag, int* wsum, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag != 0);
  const int excl = __popc(bal & ((1u << lane) - 1u));
  __syncthreads();                         // wsum may still be read from a previous scan
  if (lane == 0) wsum[warp] = __popc(bal);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 32; ++w) {
    const int c = wsum[w];
    base += (w < warp) ? c : 0;
    tot += c;
  }
  *total = tot;
  return base + excl;
}

template <int V>
__device__ __forceinline__ void store_row(const DtpApplyArgs& a, long long orow, int lane, const float4 (&v)[V]) {
  float4* o = reinterpret_cast<float4*>(a.out) + orow * (V * 32);
#pragma unroll
  for (int i = 0; i < V; ++i) o[lane + 32 * i] = v[i];
  if (a.out_f16) {
    uint2* o16 = reinterpret_cast<uint2*>(a.out_f16) + orow * (V * 32);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const __half2 h0 = __floats2half2_rn(v[i].x, v[i].y), h1 = __floats2half2_rn(v[i].z, v[i].w);
      uint2 p;
      p.x = *reinterpret_cast<const uint32_t*>(&h0);
      p.y = *reinterpret_cast<const uint32_t*>(&h1);
      o16[lane + 32 * i] = p;
    }
  }
  if (a.ln_out) {   // the same arithmetic as layernorm_kernel (rowops.cu): mean, centred variance, one warp per row
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / static_cast<float>(V * 128);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float var = warp_sum(q) / static_cast<float>(V * 128);
    const float rstd = 1.0f / sqrtf(var + a.ln_eps);
    const float4* g4 = reinterpret_cast<const float4*>(a.ln_gamma);
    const float4* b4 = reinterpret_cast<const float4*>(a.ln_beta);
    uint2* y16 = reinterpret_cast<uint2*>(a.ln_out) + orow * (V * 32);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 g = __ldg(g4 + lane + 32 * i), be = __ldg(b4 + lane + 32 * i);
      const float y0 = (v[i].x - mean) * rstd * g.x + be.x, y1 = (v[i].y - mean) * rstd * g.y + be.y;
      const float y2 = (v[i].z - mean) * rstd * g.z + be.z, y3 = (v[i].w - mean) * rstd * g.w + be.w;
      const __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
      uint2 p;
      p.x = *reinterpret_cast<const uint32_t*>(&h0);
      p.y = *reinterpret_cast<const uint32_t*>(&h1);
      y16[lane + 32 * i] = p;
    }
  }
}

}  // namespace

template <int V>   // V float4 per lane: d == 128 * V
__global__ void __launch_bounds__(1024, 1)
dtp_apply_kernel(DtpApplyArgs a) {
  __shared__ float S[kDtpMaxTokens];
  __shared__ uint32_t key[kDtpMaxTokens];
  __shared__ int slot[kDtpMaxTokens];        // survivor slot (ascending token order) or -1
  __shared__ float tw[kDtpMaxTokens];        // merge weight of a pruned token
  __shared__ int tail_list[kDtpMaxTokens];   // pruned token indices, ascending
  __shared__ int hist[256];
  __shared__ int wsum[32];
  __shared__ double red[32];
  __shared__ uint32_t sel_prefix;
  __shared__ int sel_remaining;
  extern __shared__ float4 part[];           // [8][d4] partial merged rows + [d4] the combined row

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, s = blockIdx.x, SC = gridDim.x;
  const bool dyn = a.n_dev != nullptr;
  const int n = dyn ? min(a.n, load_len(a.n_dev) - 1) : a.n;
  const int k_in = *a.topk;
  const bool identity = (k_in <= a.max_keep) || (n - k_in <= 1);    // reference early-out: nothing is pruned
  const int k = identity ? n : k_in;
  constexpr int d4 = V * 32;
  const long long in_rows = dyn ? (n + 1) : (a.bsx / (V * 128));     // rows per sequence of x / out
  const long long out_rows = dyn ? (identity ? n + 1 : k + 2) : (a.bso / (V * 128));
  const int mo_pitch = dyn ? static_cast<int>(out_rows) : a.n + 1;
  if (dyn && b == 0 && s == 0 && tid == 0) {
    if (a.n_out) *a.n_out = identity ? n + 1 : k + 2;
    if (a.k_out) *a.k_out = identity ? -1 : k;
  }

  const int j = tid;
  const bool valid = j < n;
  const float sj = valid ? a.score[static_cast<long long>(b) * n + j] : 0.f;
  S[j] = sj;
  key[j] = valid ? ordered_key(sj) : 0u;
  int keepflag = valid ? 1 : 0;
  int rank_small = -1;                      // rank in descending order, only for the masked text modes
  if (!identity) {
    // ---- radix select: the k-th largest key ----
    if (tid == 0) {
      sel_prefix = 0u;
      sel_remaining = k;
    }
    uint32_t maskbits = 0u;
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const uint32_t prefix = sel_prefix;
      const int remaining = sel_remaining;
      if (valid && (key[j] & maskbits) == prefix) atomicAdd(&hist[(key[j] >> shift) & 255u], 1);
      __syncthreads();
      if (warp == 0) {
        // lane l owns bins [8l, 8l + 8); suffix sums over lanes above, then the owning lane walks its bins from the top
        int c[8], tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          c[i] = hist[lane * 8 + i];
          tot += c[i];
        }
        // inclusive suffix sum over the 32 lanes; above = keys in the bins of higher lanes
        int run = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_down_sync(0xffffffffu, run, o);
          if (lane + o < 32) run += v;
        }
        const int above = run - tot;
        if (above < remaining && remaining <= above + tot) {
          int cum = above;
#pragma unroll
          for (int i = 7; i >= 0; --i) {
            if (cum < remaining && remaining <= cum + c[i]) {
              sel_prefix = prefix | (static_cast<uint32_t>(lane * 8 + i) << shift);
              sel_remaining = remaining - cum;
            }
            cum += c[i];
          }
        }
      }
      maskbits |= 255u << shift;
      __syncthreads();
    }
    const uint32_t kth = sel_prefix;
    const int need_eq = sel_remaining;      // keys equal to the k-th that are kept: the ones with the lowest indices
    const int eq = (valid && key[j] == kth) ? 1 : 0;
    int eq_tot;
    const int eq_rank = block_excl_scan(eq, wsum, &eq_tot);
    keepflag = (valid && (key[j] > kth || (eq && eq_rank < need_eq))) ? 1 : 0;
    if (a.mask_mode != 0 && valid) {        // short text sequences: the rank itself is needed for the mask bookkeeping
      int r = 0;
      for (int i = 0; i < n; ++i) {
        const float si = S[i];
        r += (si > sj || (si == sj && i < j)) ? 1 : 0;
      }
      rank_small = r;
    }
  }
  int kept_total;
  const int dst = block_excl_scan(keepflag, wsum, &kept_total);
  slot[j] = (valid && keepflag) ? dst : -1;
  if (valid && !keepflag) tail_list[j - dst] = j;
  // merge weights: score / (sum of the pruned scores + 1e-8), the sum in fp64 (exact for scores of similar magnitude)
  double tsum = 0.0;
  {
    double t = (valid && !keepflag) ? static_cast<double>(sj) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if (lane == 0) red[warp] = t;
    __syncthreads();
    for (int w = 0; w < 32; ++w) tsum += red[w];
  }
  const float den = static_cast<float>(tsum) + 1e-8f;
  tw[j] = (valid && !keepflag) ? __fdiv_rn(sj, den) : 0.f;

  if (s == 0) {
    if (a.keep && valid) a.keep[static_cast<long long>(b) * n + j] = static_cast<unsigned char>(keepflag);
    if (a.mask_mode != 0) {
      const float* mi = a.mask_in + static_cast<long long>(b) * (n + 1);
      float* mo = a.mask_out + static_cast<long long>(b) * mo_pitch;
      if (identity) {
        for (int i = tid; i < n + 1; i += 1024) mo[i] = mi[i];
      } else {
        if (tid == 0) mo[0] = mi[0];
        if (valid) {
          const float mj = mi[1 + j];
          if (a.mask_mode == 1) {
            if (rank_small <= k) mo[1 + rank_small] = mj;     // slot r <- mask of the r-th ranked token, r = 0..k
          } else {
            if (keepflag) mo[1 + dst] = mj;                    // the mask travels with its token
            if (rank_small == k) mo[1 + k] = mj;               // merged slot <- mask of the (k+1)-th ranked token
          }
        }
      }
    }
  }
  __syncthreads();

  const float4* xb = reinterpret_cast<const float4*>(a.x) + static_cast<long long>(b) * in_rows * d4;
  const long long ob = static_cast<long long>(b) * out_rows;

  // ---- merged token: warps 0..7 of the first CTA, same reduction order as dtp_gather_kernel ----
  if (s == 0 && !identity && warp < 8) {
    const int ntail = n - k;
    float4 acc[V];
#pragma unroll
    for (int c = 0; c < V; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = warp; i < ntail; i += 8) {
      const int jt = tail_list[i];
      const float w = tw[jt];
      const float4* src = xb + static_cast<long long>(1 + jt) * d4;
#pragma unroll
      for (int c = 0; c < V; ++c) {
        const float4 v = src[lane + 32 * c];
        acc[c].x = fmaf(w, v.x, acc[c].x);
        acc[c].y = fmaf(w, v.y, acc[c].y);
        acc[c].z = fmaf(w, v.z, acc[c].z);
        acc[c].w = fmaf(w, v.w, acc[c].w);
      }
    }
#pragma unroll
    for (int c = 0; c < V; ++c) part[warp * d4 + lane + 32 * c] = acc[c];
    named_bar_sync(1, 256);
    for (int c = tid; c < d4; c += 256) {
      float4 t = part[c];
      for (int w = 1; w < 8; ++w) {
        const float4 v = part[w * d4 + c];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      part[8 * d4 + c] = t;
    }
    named_bar_sync(1, 256);
    if (warp == 0) {
      float4 v[V];
#pragma unroll
      for (int c = 0; c < V; ++c) v[c] = part[8 * d4 + lane + 32 * c];
      store_row<V>(a, ob + 1 + k, lane, v);
    }
  }

  // ---- survivors: one warp per source row, the S * 32 warps of the sequence share the rows ----
  for (int r = s * 32 + warp; r < n + 1; r += SC * 32) {
    int dslot;
    if (r == 0) dslot = 0;
    else {
      const int sl = slot[r - 1];
      if (sl < 0) continue;
      dslot = 1 + sl;
    }
    const float4* src = xb + static_cast<long long>(r) * d4;
    float4 v[V];
#pragma unroll
    for (int c = 0; c < V; ++c) v[c] = src[lane + 32 * c];
    store_row<V>(a, ob + dslot, lane, v);
  }
}

int launch_dtp_apply(const DtpApplyArgs& a, cudaStream_t stream) {
  if (a.B == 0) return kOk;
  MADTP_CHECK_ARG(a.score && a.topk && a.x && a.out, "dtp_apply: null pointer");
  MADTP_CHECK_ARG(a.n > 0 && a.n <= kDtpMaxTokens, "dtp_apply: n=%d out of range (1..%d)", a.n, kDtpMaxTokens);
  MADTP_CHECK_ARG(a.d > 0 && a.d % 128 == 0 && a.d <= 1024, "dtp_apply: d must be a multiple of 128, <= 1024 (d=%d)", a.d);
  MADTP_CHECK_ARG(a.bsx % a.d == 0 && a.bso % a.d == 0 && a.B <= 65535, "dtp_apply: batch strides are whole rows");
  MADTP_CHECK_ARG(a.mask_mode >= 0 && a.mask_mode <= 2, "dtp_apply: mask_mode must be 0, 1 or 2");
  MADTP_CHECK_ARG(a.mask_mode == 0 || (a.mask_in && a.mask_out), "dtp_apply: mask buffers missing");
  MADTP_CHECK_ARG(a.mask_mode == 0 || a.n <= 256, "dtp_apply: the masked (text) modes are built for short sequences");
  MADTP_CHECK_ARG((a.ln_out == nullptr) || (a.ln_gamma && a.ln_beta), "dtp_apply: LayerNorm needs gamma and beta");
  MADTP_CHECK_ARG(a.n_dev == nullptr || a.n_out != nullptr, "dtp_apply: n_dev needs n_out_dev");
  // one CTA per SM (1024 threads, the row in registers): S * B <= SMs keeps the grid a single wave -- 3 x 64 CTAs on
  // 148 SMs ran as two waves, the second one 30 % full (47 us -> 3x us for block 1 of the bench)
  int S = num_sms() / a.B;
  if (S < 1) S = 1;
  if (S > 8) S = 8;
  const int d4 = a.d / 4;
  const int smem = 9 * d4 * static_cast<int>(sizeof(float4));
  dim3 grid(S, a.B);
  switch (a.d / 128) {
#define MADTP_APPLY_CASE(V)                                                   \
  case V:                                                                     \
    MADTP_SMEM_ATTR_ONCE(9 * V * 32 * 16, dtp_apply_kernel<V>);               \
    dtp_apply_kernel<V><<<grid, 1024, smem, stream>>>(a);                     \
    break;
    MADTP_APPLY_CASE(1) MADTP_APPLY_CASE(2) MADTP_APPLY_CASE(3) MADTP_APPLY_CASE(4) MADTP_APPLY_CASE(5)
    MADTP_APPLY_CASE(6) MADTP_APPLY_CASE(7) MADTP_APPLY_CASE(8)
#undef MADTP_APPLY_CASE
  }
  MADTP_LAUNCH_CHECK();
  return kOk;
}

}  // namespace madtp
