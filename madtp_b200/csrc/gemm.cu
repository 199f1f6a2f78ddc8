// GEMM kernels for the packed token stream: C[M,N] = epilogue(A[M,K] * B[N,K]^T).
//
// gemm_tcgen05_kernel — persistent, warp-specialised Blackwell kernel:
//   warp 0 (one lane)  : TMA producer   (cp.async.bulk.tensor 2-D tiles, 128-byte swizzle, mbarrier tx-count)
//   warp 1 (one lane)  : MMA issuer     (tcgen05.mma, accumulators in TMEM, tcgen05.commit -> mbarriers)
//   warps 2..5         : epilogue       (tcgen05.ld TMEM -> registers, bias / activation / residual, global stores)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Precision modes:
//   kGemmF16    fp16 operands, fp32 accumulation, one kind::f16 MMA per 16-element k-step.
//   kGemmTF32x3 error-compensated fp32: operands are pre-split into tf32 "hi" (low 13 mantissa bits zero) and the
//               exact residual "lo"; every k-step issues lo*hi + hi*lo + hi*hi (kind::tf32, fp32 accumulation).
//               This is the "scoring lane" precision: keep-masks are decided on score gaps of ~1e-7, which plain
//               tf32/fp16 products do not resolve (hard part 1).
//
// gemm_simt_kernel — plain fp32 FFMA tiling; device-side cross-check for the tensor-core path and tiny shapes.
#include "gemm.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace madtp {

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool TF32X3, bool PAIR = false>
struct GemmCfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int ROW_BYTES = 128;                       // one 128B swizzle atom of K per stage
  static constexpr int K_ELEMS = TF32X3 ? 32 : 64;            // K elements per stage
  static constexpr int UMMA_K_BYTES = 32;                     // K bytes per tcgen05.mma
  static constexpr int A_BYTES = BLOCK_M * ROW_BYTES;
  static constexpr int B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * ROW_BYTES;   // a CTA of a pair stages half of B
  static constexpr int STAGE_BYTES = (TF32X3 ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;               // double-buffered fp32 accumulator
  static constexpr int EPI_BYTES = 8 * 4096;                   // one 32x32 fp32 transpose tile per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int THREADS = 320;
};


// Epilogue for 32 consecutive accumulator columns [col0, col0+32) of one output row held by one thread:
// out = act(alpha * acc + bias) + residual, written as fp32 or fp16 with 128-bit stores when the layout allows.
__device__ __forceinline__ void store_row_chunk32(const GemmEpilogue& ep, const float (&acc)[32], long long row,
                                                  int col0, int N, bool vec_ok) {
  if (vec_ok && col0 + 32 <= N) {
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ep.bias) bv = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
      float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ep.residual) rv = __ldg(reinterpret_cast<const float4*>(ep.residual + row * ep.ldr + col0 + j));
      o[j + 0] = apply_act(ep.alpha * acc[j + 0] + bv.x, ep.act) + rv.x;
      o[j + 1] = apply_act(ep.alpha * acc[j + 1] + bv.y, ep.act) + rv.y;
      o[j + 2] = apply_act(ep.alpha * acc[j + 2] + bv.z, ep.act) + rv.z;
      o[j + 3] = apply_act(ep.alpha * acc[j + 3] + bv.w, ep.act) + rv.w;
    }
    if (ep.c_f16) {
      __half* dst = reinterpret_cast<__half*>(ep.c) + row * ep.ldc + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        __half2 h0 = __floats2half2_rn(o[j + 0], o[j + 1]);
        __half2 h1 = __floats2half2_rn(o[j + 2], o[j + 3]);
        __half2 h2 = __floats2half2_rn(o[j + 4], o[j + 5]);
        __half2 h3 = __floats2half2_rn(o[j + 6], o[j + 7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + j) = pk;
      }
    } else {
      float* dst = reinterpret_cast<float*>(ep.c) + row * ep.ldc + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = col0 + j;
      if (col < N) {
        float x = ep.alpha * acc[j];
        if (ep.bias) x += __ldg(ep.bias + col);
        x = apply_act(x, ep.act);
        if (ep.residual) x += __ldg(ep.residual + row * ep.ldr + col);
        if (ep.c_f16)
          reinterpret_cast<__half*>(ep.c)[row * ep.ldc + col] = __float2half_rn(x);
        else
          reinterpret_cast<float*>(ep.c)[row * ep.ldc + col] = x;
      }
    }
  }
}

__device__ __forceinline__ bool epilogue_vec_ok(const GemmEpilogue& ep, int N) {
  const bool al = (reinterpret_cast<uintptr_t>(ep.c) & 15) == 0 &&
                  (ep.bias == nullptr || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) &&
                  (ep.residual == nullptr || (reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0);
  return al && ((ep.ldc & 7) == 0) && ((N & 3) == 0) && (ep.residual == nullptr || (ep.ldr & 3) == 0);
}


// Coalesced epilogue for one 32 x 32 accumulator chunk owned by one warp. After tcgen05.ld every lane holds one ROW
// (32 consecutive columns); storing that directly makes each warp-wide access touch 32 different 128-byte lines.
// The chunk is therefore transposed through a 4 KB per-warp shared-memory tile (float4 granules, XOR-swizzled by
// row so both the row-wise writes and the column-wise reads are bank-conflict free) and then written with 8 lanes
// covering one 128-byte row segment, 4 rows per instruction.
//   out = act(alpha * acc + bias[col]) + residual[row, col]
// The residual sub-tile is prefetched into L2 while the warp waits for the accumulator (prefetch_residual): with only
// 8 epilogue warps per SM the memory-level parallelism of a load-then-use epilogue is far too low to cover the DRAM
// latency of the fp32 residual stream, and a register prefetch one chunk ahead spills.
__device__ __forceinline__ void load_residual(const GemmEpilogue& ep, long long row_base, int col0, int M, int N,
                                              int lane, float4 (&rv)[8]) {
  const int col = col0 + (lane & 7) * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const long long row = row_base + it * 4 + (lane >> 3);
    rv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.residual != nullptr && row < M && col < N)
      rv[it] = __ldg(reinterpret_cast<const float4*>(ep.residual + row * ep.ldr + col));
  }
}

template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  if (ACT == 1) return gelu_erf(x);
  if (ACT == 4) return gelu_tanh5(x);
  if (ACT == 2) return fmaxf(x, 0.0f);
  if (ACT == 3) return x / (1.0f + __expf(-1.702f * x));
  return x;
}

__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)
               : "memory");
  return v;
}

template <int ACT, bool RES = true, int GROUP = 8, bool SPLIT = false>
__device__ __forceinline__ void store_chunk_coalesced(const GemmEpilogue& ep, const float* acc, float* stage,
                                                      long long row_base, int col0, int M, int N, int lane,
                                                      const float4 (&rv)[8]) {
  const uint32_t sbase = smem_u32(stage);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sts_v4(sbase + static_cast<uint32_t>((lane * 8 + (j ^ (lane & 7))) * 16), acc[4 * j], acc[4 * j + 1],
           acc[4 * j + 2], acc[4 * j + 3]);
  __syncwarp();
  const int jj = lane & 7;
  const int rq = lane >> 3;
  const int col = col0 + jj * 4;
  const bool col_ok = col < N;  // N % 4 == 0 on this path, so the whole float4 is in range
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ep.bias != nullptr && col_ok) bv = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
  // GROUP rows-of-four are fetched from the tile at a time (8 = whole chunk in flight; 4 halves the live registers)
#pragma unroll
  for (int g = 0; g < 8; g += GROUP) {
    float4 a[GROUP];
#pragma unroll
    for (int i = 0; i < GROUP; ++i) {
      const int r = (g + i) * 4 + rq;
      a[i] = lds_v4(sbase + static_cast<uint32_t>((r * 8 + (jj ^ (r & 7))) * 16));
    }
#pragma unroll
    for (int i = 0; i < GROUP; ++i) {
      const int it = g + i;
      const long long row = row_base + it * 4 + rq;
      float4 o;
      o.x = act_fn<ACT>(ep.alpha * a[i].x + bv.x);
      o.y = act_fn<ACT>(ep.alpha * a[i].y + bv.y);
      o.z = act_fn<ACT>(ep.alpha * a[i].z + bv.z);
      o.w = act_fn<ACT>(ep.alpha * a[i].w + bv.w);
      if (RES) {
        o.x += rv[it].x; o.y += rv[it].y; o.z += rv[it].z; o.w += rv[it].w;
      }
      if (SPLIT) {   // fp16 hi/lo planes of kQkPlaneScale * o
        if (col_ok && row < M) {
          o.x *= kQkPlaneScale; o.y *= kQkPlaneScale; o.z *= kQkPlaneScale; o.w *= kQkPlaneScale;
          const __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
          const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
          const __half2 l0 = __floats2half2_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2half2_rn(o.z - f1.x, o.w - f1.y);
          uint2 ph, pl;
          ph.x = *reinterpret_cast<const uint32_t*>(&h0);
          ph.y = *reinterpret_cast<const uint32_t*>(&h1);
          pl.x = *reinterpret_cast<const uint32_t*>(&l0);
          pl.y = *reinterpret_cast<const uint32_t*>(&l1);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.c) + row * ep.ldc + col) = ph;
          *reinterpret_cast<uint2*>(ep.c_lo + row * ep.ldc + col) = pl;
        }
      } else if (col_ok && row < M) {
        if (ep.c_f16) {
          __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.c) + row * ep.ldc + col) = pk;
        } else {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.c) + row * ep.ldc + col) = o;
        }
      }
    }
  }
  __syncwarp();  // the tile is reused by the next chunk
}

// L2 prefetch of the residual sub-tile a warp is about to need (32 rows x NCH 128-byte lines; lane = row). Issued
// while the warp would otherwise idle on the accumulator barrier; needs no registers, unlike a register prefetch.
template <int NCH>
__device__ __forceinline__ void prefetch_residual(const GemmEpilogue& ep, long long row_base, int col_begin, int M,
                                                  int N, int lane) {
  if (ep.residual == nullptr) return;
  const long long row = row_base + lane;
  if (row >= M) return;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = col_begin + c * 32;
    if (col < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.residual + row * ep.ldr + col));
  }
}

// Epilogue of NCH consecutive 32-column chunks starting at TMEM address taddr / output column col_begin, for the 32
// rows [row_base, row_base+32) owned by this warp.
template <int ACT, int NCH>
__device__ __forceinline__ void epilogue_chunks_tmem(const GemmEpilogue& ep, uint32_t taddr, float* stage,
                                                     long long row_base, int col_begin, int M, int N, int lane) {
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    const int col0 = col_begin + c * 32;
    if (col0 >= N) break;
    uint32_t v[32];
    tmem_ld_32x32b_x32(taddr + c * 32, v);
    float4 rv[8];
    load_residual(ep, row_base, col0, M, N, lane, rv);   // L2 hits (prefetched); overlaps the TMEM load + transpose
    tmem_ld_wait();
    float acc32[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc32[j] = __uint_as_float(v[j]);
    store_chunk_coalesced<ACT>(ep, acc32, stage, row_base, col0, M, N, lane, rv);
  }
}

// PAIR = true: clusters of two CTAs form a cta_group::2 pair working on two vertically adjacent output tiles (same
// n-tile): ONE MMA of M = 256 spans both SMs, every CTA stages its own 128 rows of A and only half of B's rows, the
// leader (cluster rank 0) issues all MMAs and owns the operand-full barriers (see gemm_split3_kernel).
template <int BLOCK_N, bool TF32X3, bool PAIR = false>
__global__ void __launch_bounds__(320, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a_lo,
                    const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_b_lo,
                    GemmEpilogue ep, int M_cap, int N_cap, int K) {
  using Cfg = GemmCfg<BLOCK_N, TF32X3, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  static_assert(!PAIR || !TF32X3, "the CTA-pair variant is fp16 only");
  const int M = gemm_dyn_m(ep, M_cap), N = gemm_dyn_n(ep, N_cap);   // device-resident extents (gemm.cuh)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  constexpr int CL = PAIR ? 2 : 1;
  const int num_tiles = ((m_tiles + CL - 1) / CL) * n_tiles;   // work items (a pair owns CL vertically adjacent tiles)
  const int num_kb = (K + Cfg::K_ELEMS - 1) / Cfg::K_ELEMS;
  const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_item = blockIdx.x / CL, item_stride = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (TF32X3) {
      tma_prefetch_desc(&tm_a_lo);
      tma_prefetch_desc(&tm_b_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], PAIR ? 16 : 8);   // pair: the epilogue warps of both CTAs release the leader's buffer
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    if (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();   // the peer's barriers are initialised before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps run their loops warp-uniformly and let ONE elected lane issue (elect.sync): the TMA / MMA
  // operands then live in uniform registers. Guarding the loops with `lane == 0` makes every tcgen05.mma and TMA a
  // ~100-cycle R2UR + waterfall sequence (measured on the attention kernel: -23% time from this change alone).
  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_item; tile < num_tiles; tile += item_stride) {
        const int m0 = ((tile / n_tiles) * CL + rank) * Cfg::BLOCK_M;
        const int n0 = (tile % n_tiles) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* s = smem + stage * Cfg::STAGE_BYTES;
          const int k0 = kb * Cfg::K_ELEMS;
          if (PAIR) {
            // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the whole pair
            if (elect_one()) {
              const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
              tma_load_2d_pair(&tm_a, bar, s, k0, m0);
              tma_load_2d_pair(&tm_b, bar, s + Cfg::A_BYTES, k0, n0 + rank * (BLOCK_N / 2));
            }
          } else if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(&tm_a, &full_bar[stage], s, k0, m0);
            tma_load_2d(&tm_b, &full_bar[stage], s + Cfg::A_BYTES, k0, n0);
            if (TF32X3) {
              tma_load_2d(&tm_a_lo, &full_bar[stage], s + Cfg::A_BYTES + Cfg::B_BYTES, k0, m0);
              tma_load_2d(&tm_b_lo, &full_bar[stage], s + 2 * Cfg::A_BYTES + Cfg::B_BYTES, k0, n0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (!PAIR || rank == 0) {   // pair: only the leader issues MMAs
      constexpr uint32_t idesc = make_idesc(TF32X3 ? 2u : 0u, PAIR ? 2 * Cfg::BLOCK_M : Cfg::BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_item; tile < num_tiles; tile += item_stride) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t s = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t a_hi = make_sw128_kmajor_desc(s);
          const uint64_t b_hi = make_sw128_kmajor_desc(s + Cfg::A_BYTES);
          const uint64_t a_lo = make_sw128_kmajor_desc(s + Cfg::A_BYTES + Cfg::B_BYTES);
          const uint64_t b_lo = make_sw128_kmajor_desc(s + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < Cfg::ROW_BYTES / Cfg::UMMA_K_BYTES; ++k) {
              const uint64_t koff = static_cast<uint64_t>((k * Cfg::UMMA_K_BYTES) >> 4);
              const uint32_t first = (kb | k) != 0 ? 1u : 0u;
              if (TF32X3) {
                // small terms first, then the dominant hi*hi product
                umma_tf32(d_tmem, a_lo + koff, b_hi + koff, idesc, first);
                umma_tf32(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
                umma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
              } else if (PAIR) {
                umma_f16_pair(d_tmem, a_hi + koff, b_hi + koff, idesc, first);
              } else {
                umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc, first);
              }
            }
            if (PAIR) {
              umma_commit_pair(&empty_bar[stage], 3);                          // frees the stage in both CTAs
              if (kb == num_kb - 1) umma_commit_pair(&tmem_full[acc], 3);      // accumulators complete in both CTAs
            } else {
              umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs retire
              if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..9) ------------------------------
    // warp % 4 = TMEM lane quadrant (a warp may only touch lanes [32*(warp%4), +32)); (warp-2)/4 = column half.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int NCH = BLOCK_N / 64;   // 32-column chunks per warp
    float* stage = epi_stage + (warp - 2) * 1024;
    const bool vec_ok = epilogue_vec_ok(ep, N);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_item; tile < num_tiles; tile += item_stride) {
      const int m0 = ((tile / n_tiles) * CL + rank) * Cfg::BLOCK_M;
      const int n0 = (tile % n_tiles) * BLOCK_N;
      const long long row_base = m0 + quad * 32;
      const int col_begin = n0 + half * (BLOCK_N / 2);
      if (vec_ok) prefetch_residual<NCH>(ep, row_base, col_begin, M, N, lane);   // lands in L2 while the MMAs finish
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * BLOCK_N + half * (BLOCK_N / 2));
      if (vec_ok) {
        switch (ep.act) {
          case 1: epilogue_chunks_tmem<1, NCH>(ep, taddr, stage, row_base, col_begin, M, N, lane); break;
          case 2: epilogue_chunks_tmem<2, NCH>(ep, taddr, stage, row_base, col_begin, M, N, lane); break;
          case 3: epilogue_chunks_tmem<3, NCH>(ep, taddr, stage, row_base, col_begin, M, N, lane); break;
          case 4: epilogue_chunks_tmem<4, NCH>(ep, taddr, stage, row_base, col_begin, M, N, lane); break;
          default: epilogue_chunks_tmem<0, NCH>(ep, taddr, stage, row_base, col_begin, M, N, lane); break;
        }
      } else {
        const long long row = row_base + lane;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const int col0 = col_begin + c * 32;
          if (col0 >= N) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (row < M) {
            float acc32[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc32[j] = __uint_as_float(v[j]);
            store_row_chunk32(ep, acc32, row, col0, N, false);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
        else mbar_arrive(&tmem_empty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();   // no CTA leaves while its peer may still arrive on its barriers or read its operands
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------
// TF32x3 kernel with chunk-drained accumulation.
//
// Measured on B200: the tensor core's fp32 accumulator truncates on every accumulate, so a K=768 TF32x3 GEMM that
// accumulates all of K in TMEM comes out ~5e-6 relative (and ~2e-5 at K=3072), 50x worse than an fp32 FFMA GEMM --
// too coarse for the scoring lane. Here every 32-element K chunk (one pipeline stage: 4 k-steps x 3 MMAs) is
// accumulated into its own TMEM buffer (two buffers, ping-pong) and the eight epilogue warps drain each chunk with
// tcgen05.ld and add it into fp32 registers with round-to-nearest. The tensor core therefore only ever sums 32
// products plus the small cross terms; the long K reduction is IEEE fp32.
//   warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: drain + epilogue (warp%4 = TMEM lane quadrant,
//   (warp-2)/4 = column half of the tile).
// ------------------------------------------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ void umma_x(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  if (F16) umma_f16(d_tmem, a, b, idesc, accumulate);
  else umma_tf32(d_tmem, a, b, idesc, accumulate);
}

template <int BLOCK_N, bool PAIR = false>
struct Tf32Cfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int K_ELEMS = 32;
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * 128;   // a CTA of a pair stages half of B's rows
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int EPI_BYTES = 8 * 4096;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
  static constexpr int THREADS = 320;
  static constexpr int COLS_PER_WARP = BLOCK_N / 2;
};

// CL = 1: independent CTAs. CL = 2: clusters of two CTAs that work on vertically adjacent output tiles (same n-tile)
// and share the B operand: each CTA fetches half of the B tile and TMA-multicasts it to both, which halves the L2 ->
// shared-memory traffic of the (weight) operand -- the K = 768 projections are L2-bandwidth bound otherwise.
// F16 = true: the same pipeline with fp16 hi/lo planes (kind::f16): a 128-byte operand row then holds 64 k-elements
// instead of 32 and every MMA covers 16 of them, so a k-block costs the same bytes and MMA slots but twice the K.
// PAIR = true (with CL = 2, F16): the two CTAs form a cta_group::2 pair -- ONE MMA of M = 256 spans both SMs, every
// CTA stages its own 128 rows of A and only HALF of B's rows (64 KB per k-block instead of 96 KB: this kernel is bound
// by the L2 -> shared-memory operand traffic), the leader issues all MMAs and owns the operand-full barriers.
template <int BLOCK_N, int CL, bool F16 = false, bool PAIR = false>
__global__ void __launch_bounds__(320, 1)
gemm_split3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_b_lo,
                   GemmEpilogue ep, int M_cap, int N_cap, int K) {
  using Cfg = Tf32Cfg<BLOCK_N, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CW = Cfg::COLS_PER_WARP;
  static_assert(!PAIR || (CL == 2 && F16), "the CTA-pair variant is the fp16-plane kernel on clusters of two");
  const int M = gemm_dyn_m(ep, M_cap), N = gemm_dyn_n(ep, N_cap);   // device-resident extents (gemm.cuh)
  const int n_tok = (ep.mode == 1 && ep.m_dev) ? load_len(ep.m_dev) : ep.n_tok;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = ((m_tiles + CL - 1) / CL) * n_tiles;   // work items per cluster: CL vertically adjacent tiles
  constexpr int K_ELEMS = F16 ? 64 : Cfg::K_ELEMS;
  const int num_kb = (K + K_ELEMS - 1) / K_ELEMS;
  const int chunk = ep.chunk_kb > 0 ? ep.chunk_kb : 1;   // k-blocks accumulated in TMEM between two drains
  const int rank = (CL > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_item = blockIdx.x / CL, item_stride = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], PAIR ? 1 : CL);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], PAIR ? 16 : 8);   // pair: the epilogue warps of BOTH CTAs release the leader's buffer
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    if (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();   // the peer's barriers are initialised before anything is multicast to them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {   // warp-uniform loop, one elected lane issues (see gemm_tcgen05_kernel)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_item; tile < num_tiles; tile += item_stride) {
        const int m0 = ((tile / n_tiles) * CL + rank) * Cfg::BLOCK_M;
        const int n0 = (tile % n_tiles) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);   // CL > 1: BOTH CTAs have retired their MMAs on this stage
          uint8_t* s = smem + stage * Cfg::STAGE_BYTES;
          const int k0 = kb * K_ELEMS;
          if (PAIR) {
            // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the whole pair
            if (elect_one()) {
              const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
              const int nr = n0 + rank * (BLOCK_N / 2);
              tma_load_2d_pair(&tm_a, bar, s, k0, m0);
              tma_load_2d_pair(&tm_b, bar, s + Cfg::A_BYTES, k0, nr);
              tma_load_2d_pair(&tm_a_lo, bar, s + Cfg::A_BYTES + Cfg::B_BYTES, k0, m0);
              tma_load_2d_pair(&tm_b_lo, bar, s + 2 * Cfg::A_BYTES + Cfg::B_BYTES, k0, nr);
            }
          } else if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(&tm_a, &full_bar[stage], s, k0, m0);
            tma_load_2d(&tm_a_lo, &full_bar[stage], s + Cfg::A_BYTES + Cfg::B_BYTES, k0, m0);
            if (CL == 1) {
              tma_load_2d(&tm_b, &full_bar[stage], s + Cfg::A_BYTES, k0, n0);
              tma_load_2d(&tm_b_lo, &full_bar[stage], s + 2 * Cfg::A_BYTES + Cfg::B_BYTES, k0, n0);
            } else {   // this CTA's half of the B tile, delivered to both CTAs of the cluster
              constexpr int HB = Cfg::B_BYTES / CL;
              const int nr = n0 + rank * (BLOCK_N / CL);
              tma_load_2d_multicast(&tm_b, &full_bar[stage], s + Cfg::A_BYTES + rank * HB, k0, nr, (1u << CL) - 1);
              tma_load_2d_multicast(&tm_b_lo, &full_bar[stage], s + 2 * Cfg::A_BYTES + Cfg::B_BYTES + rank * HB, k0,
                                    nr, (1u << CL) - 1);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!PAIR || rank == 0) {   // pair: only the leader issues MMAs
      constexpr uint32_t idesc = make_idesc(F16 ? 0u : 2u, PAIR ? 2 * Cfg::BLOCK_M : Cfg::BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t buf_phase = 0;
      for (int tile = first_item; tile < num_tiles; tile += item_stride) {
        for (int kb = 0; kb < num_kb; ++kb) {
          const bool chunk_first = (kb % chunk) == 0;                       // first k-block of a drained chunk
          const bool chunk_last = ((kb + 1) % chunk) == 0 || kb + 1 == num_kb;
          if (chunk_first) mbar_wait(&tmem_empty[buf], buf_phase ^ 1u);   // (polling here measured 6 % slower)
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BLOCK_N);
          const uint32_t s = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t a_hi = make_sw128_kmajor_desc(s);
          const uint64_t b_hi = make_sw128_kmajor_desc(s + Cfg::A_BYTES);
          const uint64_t a_lo = make_sw128_kmajor_desc(s + Cfg::A_BYTES + Cfg::B_BYTES);
          const uint64_t b_lo = make_sw128_kmajor_desc(s + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
          // cross terms first (tiny partial sums), then the dominant hi*hi products: only the last four
          // accumulates truncate at the full partial-sum magnitude
          if (PAIR) {
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_pair(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, (k != 0 || !chunk_first) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_pair(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_pair(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
              umma_commit_pair(&empty_bar[stage], 3);   // frees the stage and publishes the chunk in both CTAs
              if (chunk_last) umma_commit_pair(&tmem_full[buf], 3);
            }
          } else if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_x<F16>(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, (k != 0 || !chunk_first) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_x<F16>(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_x<F16>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
            if (CL == 1) umma_commit(&empty_bar[stage]);
            else umma_commit_multicast(&empty_bar[stage], (1u << CL) - 1);   // frees the stage in both CTAs
            if (chunk_last) umma_commit(&tmem_full[buf]);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
          if (chunk_last) {
            buf ^= 1;
            if (buf == 0) buf_phase ^= 1u;
          }
        }
      }
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    int buf = 0;
    uint32_t buf_phase = 0;
    const bool vec_ok = epilogue_vec_ok(ep, N);
    const int num_chunks = (num_kb + chunk - 1) / chunk;
    for (int tile = first_item; tile < num_tiles; tile += item_stride) {
      const int m0 = ((tile / n_tiles) * CL + rank) * Cfg::BLOCK_M;
      const int n0 = (tile % n_tiles) * BLOCK_N;
      float sum[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) sum[j] = 0.f;
      for (int ch = 0; ch < num_chunks; ++ch) {
        mbar_wait(&tmem_full[buf], buf_phase);
        tcgen05_fence_after();
        const uint32_t tbase =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(buf * BLOCK_N + half * CW);
#pragma unroll
        for (int c = 0; c < CW / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tbase + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c * 32 + j] += __uint_as_float(v[j]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[buf]), 0));
          else mbar_arrive(&tmem_empty[buf]);
        }
        buf ^= 1;
        if (buf == 0) buf_phase ^= 1u;
      }
      const long long row = m0 + quad * 32 + lane;
#pragma unroll
      for (int c = 0; c < CW / 32; ++c) {
        const int col0 = n0 + half * CW + c * 32;
        if (col0 < N) {
          if (ep.mode == 1) {
            if (col0 < ep.qk_cols) {
              float4 rv[8];
              store_chunk_coalesced<0, false, (CW > 64 ? 4 : 8), true>(ep, &sum[c * 32], epi_stage + (warp - 2) * 1024,
                                                                       m0 + quad * 32, col0, M, N, lane, rv);
            } else if (row < M) {
              // value projection: lane = token, so consecutive lanes write consecutive keys of one V^T row
              const int b = static_cast<int>(row / n_tok);
              const int i = static_cast<int>(row - static_cast<long long>(b) * n_tok);
              const int vc = col0 - ep.qk_cols;
              const long long off =
                  (static_cast<long long>(b * ep.heads + (vc >> 6)) * 64 + (vc & 63)) * ep.ld_vt + i;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float o =
                    kVPlaneScale * fmaf(ep.alpha, sum[c * 32 + j], ep.bias ? __ldg(ep.bias + col0 + j) : 0.f);
                const __half h = __float2half_rn(o);
                ep.vt_hi[off + j * ep.ld_vt] = h;
                ep.vt_lo[off + j * ep.ld_vt] = __float2half_rn(o - __half2float(h));
              }
            }
          } else if (vec_ok && ep.act == 0) {
            float* stg = epi_stage + (warp - 2) * 1024;
            float4 rv[8];
            if (ep.residual != nullptr) {
              load_residual(ep, m0 + quad * 32, col0, M, N, lane, rv);
              store_chunk_coalesced<0, true, (CW > 64 ? 4 : 8)>(ep, &sum[c * 32], stg, m0 + quad * 32, col0, M, N, lane, rv);
            } else {
              store_chunk_coalesced<0, false, (CW > 64 ? 4 : 8)>(ep, &sum[c * 32], stg, m0 + quad * 32, col0, M, N, lane, rv);
            }
          } else if (row < M) {
            float acc32[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc32[j] = sum[c * 32 + j];
            store_row_chunk32(ep, acc32, row, col0, N, false);
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();   // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT fp32 kernel: 64x64 tile, 16-deep k slices, 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
                 GemmEpilogue ep, int M_cap, int N_cap, int K) {
  const int M = gemm_dyn_m(ep, M_cap), N = gemm_dyn_n(ep, N_cap);
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + kk;
      As[kk][r] = (gm < M && gk < K) ? A[gm * lda + gk] : 0.f;
      Bs[kk][r] = (gn < N && gk < K) ? B[gn * ldb + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      float x = ep.alpha * acc[i][j];
      if (ep.bias) x += ep.bias[col];
      x = apply_act(x, ep.act);
      if (ep.residual) x += ep.residual[row * ep.ldr + col];
      if (ep.c_f16)
        reinterpret_cast<__half*>(ep.c)[row * ep.ldc + col] = __float2half_rn(x);
      else
        reinterpret_cast<float*>(ep.c)[row * ep.ldc + col] = x;
    }
  }
}

// fp32 dot-product kernel for a handful of rows (M <= 64: cls_head / itm_head): one warp per output element, so even
// a [32 x 2] output spreads over the machine instead of running as one 48-step dependent loop in a single CTA.
__global__ void __launch_bounds__(256)
gemm_rowdot_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
                   GemmEpilogue ep, int M_cap, int N_cap, int K) {
  const int M = gemm_dyn_m(ep, M_cap), N = gemm_dyn_n(ep, N_cap);
  const long long o = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= static_cast<long long>(M) * N) return;
  const int row = static_cast<int>(o / N), col = static_cast<int>(o % N);
  const float4* a4 = reinterpret_cast<const float4*>(A + row * lda);
  const float4* b4 = reinterpret_cast<const float4*>(B + col * ldb);
  float acc = 0.f;
  for (int k4 = lane; k4 < K / 4; k4 += 32) {
    const float4 x = __ldg(a4 + k4), y = __ldg(b4 + k4);
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
    acc = fmaf(x.z, y.z, acc);
    acc = fmaf(x.w, y.w, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float x = ep.alpha * acc;
    if (ep.bias) x += ep.bias[col];
    x = apply_act(x, ep.act);
    if (ep.residual) x += ep.residual[row * ep.ldr + col];
    if (ep.c_f16)
      reinterpret_cast<__half*>(ep.c)[row * ep.ldc + col] = __float2half_rn(x);
    else
      reinterpret_cast<float*>(ep.c)[row * ep.ldc + col] = x;
  }
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor maps + dispatch
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// Row-major [rows, cols] matrix with leading dimension ld (elements); box = box_rows x (128 bytes of columns).
int make_tmap(CUtensorMap* map, const void* ptr, bool f32, long long rows, long long cols, long long ld,
                     int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return kCudaError;
  }
  const int es = f32 ? 4 : 2;
  MADTP_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand pointer must be 16-byte aligned");
  MADTP_CHECK_ARG((ld * es) % 16 == 0, "GEMM operand row pitch must be a multiple of 16 bytes (ld=%lld)", ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / es), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows, cols, ld);
    return kCudaError;
  }
  return kOk;
}

// Row-major fp16 [rows, cols] matrix (cols a multiple of 64) seen as [cols / 64 column groups][rows][64 columns]:
// one box = `box_groups` column groups x `box_rows` rows x 128 bytes, written to shared memory group after group, each
// group a [box_rows x 128 bytes] tile with the 128-byte swizzle -- an MN-major tensor-core operand in ONE copy.
int make_tmap_colgroups(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                        int box_groups) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return kCudaError;
  }
  MADTP_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "operand pointer must be 16-byte aligned");
  MADTP_CHECK_ARG((ld * 2) % 16 == 0 && cols % 64 == 0, "operand row pitch / width unsupported (ld=%lld cols=%lld)", ld, cols);
  cuuint64_t gdim[3] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(cols / 64)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * 2, 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_groups)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (column groups) failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows,
              cols, ld);
    return kCudaError;
  }
  return kOk;
}

template <int BLOCK_N, bool TF32X3, bool PAIR = false>
static int launch_tc(const void* a, const void* a_lo, long long lda, const void* b, const void* b_lo, long long ldb,
                     const GemmEpilogue& ep, int M, int N, int K, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, TF32X3, PAIR>;
  CUtensorMap ta, tal, tb, tbl;
  int st;
  if ((st = make_tmap(&ta, a, TF32X3, M, K, lda, Cfg::BLOCK_M)) != kOk) return st;
  if ((st = make_tmap(&tb, b, TF32X3, N, K, ldb, PAIR ? BLOCK_N / 2 : BLOCK_N)) != kOk) return st;
  if (TF32X3) {
    if ((st = make_tmap(&tal, a_lo, true, M, K, lda, Cfg::BLOCK_M)) != kOk) return st;
    if ((st = make_tmap(&tbl, b_lo, true, N, K, ldb, BLOCK_N)) != kOk) return st;
  } else {
    tal = ta;
    tbl = tb;
  }
  MADTP_SMEM_ATTR_ONCE(Cfg::SMEM_BYTES, gemm_tcgen05_kernel<BLOCK_N, TF32X3, PAIR>);
  constexpr int CL = PAIR ? 2 : 1;
  const int m_tiles = (M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int items = ((m_tiles + CL - 1) / CL) * n_tiles;
  const int max_clusters = num_sms() / CL;
  const int clusters = items < max_clusters ? items : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MADTP_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BLOCK_N, TF32X3, PAIR>, ta, tal, tb, tbl, ep, M, N, K));
  return kOk;
}


template <int BLOCK_N, int CL, bool F16 = false, bool PAIR = false>
static int launch_tf32_cl(const void* a, const void* a_lo, long long lda, const void* b, const void* b_lo,
                          long long ldb, const GemmEpilogue& ep, int M, int N, int K, cudaStream_t stream) {
  using Cfg = Tf32Cfg<BLOCK_N, PAIR>;
  CUtensorMap ta, tal, tb, tbl;
  int st;
  if ((st = make_tmap(&ta, a, !F16, M, K, lda, Cfg::BLOCK_M)) != kOk) return st;
  if ((st = make_tmap(&tb, b, !F16, N, K, ldb, BLOCK_N / CL)) != kOk) return st;
  if ((st = make_tmap(&tal, a_lo, !F16, M, K, lda, Cfg::BLOCK_M)) != kOk) return st;
  if ((st = make_tmap(&tbl, b_lo, !F16, N, K, ldb, BLOCK_N / CL)) != kOk) return st;
  MADTP_SMEM_ATTR_ONCE(Cfg::SMEM_BYTES, gemm_split3_kernel<BLOCK_N, CL, F16, PAIR>);
  GemmEpilogue epc = ep;
  if (epc.chunk_kb <= 0) {
    static const int env_chunk = getenv("MADTP_CHUNK_KB") ? atoi(getenv("MADTP_CHUNK_KB")) : 0;
    epc.chunk_kb = env_chunk > 0 ? env_chunk : (F16 ? kF16ChunkKb : 1);
  }
  const int m_tiles = (M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int items = ((m_tiles + CL - 1) / CL) * ((N + BLOCK_N - 1) / BLOCK_N);
  const int max_clusters = num_sms() / CL;
  const int clusters = items < max_clusters ? items : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MADTP_CUDA(cudaLaunchKernelEx(&cfg, gemm_split3_kernel<BLOCK_N, CL, F16, PAIR>, ta, tal, tb, tbl, epc, M, N, K));
  return kOk;
}

// Clusters of two with a TMA-multicast B operand are implemented and verified but OFF by default: measured on B200
// (M=36928, N=2304, K=768) 585 us with clusters vs 560 us without -- this kernel is bound by its two-deep operand
// pipeline (TMA latency), not by L2 bandwidth, and pairing CTAs adds lock-step stalls. MADTP_CLUSTER=1 enables it.
template <int BLOCK_N, bool F16 = false>
static int launch_tf32(const void* a, const void* a_lo, long long lda, const void* b, const void* b_lo, long long ldb,
                       const GemmEpilogue& ep, int M, int N, int K, cudaStream_t stream) {
  if (F16) {
    // CTA pairs (cta_group::2) are implemented and verified but OFF by default (MADTP_PAIR=1 enables them). Measured on
    // B200, M=36928 N=2304 K=768: 360 us against 350 us with one drained chunk per k-block (the leader's accumulator
    // hand-off then crosses the cluster 12 times per tile), 323 against 332 us with two k-blocks per chunk, 272 against
    // 309 us with a single drain per tile -- the pair pays off only where the chunked accumulation is not needed.
    const bool pair = BLOCK_N == 256 && getenv("MADTP_PAIR") != nullptr &&
                      static_cast<long long>((M + 127) / 128) * ((N + BLOCK_N - 1) / BLOCK_N) >= 2LL * num_sms();
    if (pair) return launch_tf32_cl<256, 2, true, true>(a, a_lo, lda, b, b_lo, ldb, ep, M, N, K, stream);
    return launch_tf32_cl<BLOCK_N, 1, true>(a, a_lo, lda, b, b_lo, ldb, ep, M, N, K, stream);
  }
  const int m_tiles = (M + 127) / 128;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  static const bool use_cluster = getenv("MADTP_CLUSTER") != nullptr;
  if (use_cluster && m_tiles >= 2 && static_cast<long long>(m_tiles) * n_tiles >= 2LL * num_sms())
    return launch_tf32_cl<BLOCK_N, 2>(a, a_lo, lda, b, b_lo, ldb, ep, M, N, K, stream);
  return launch_tf32_cl<BLOCK_N, 1>(a, a_lo, lda, b, b_lo, ldb, ep, M, N, K, stream);
}

int launch_gemm_qkv(const void* a_hi, const void* a_lo, long long lda, const void* w_hi, const void* w_lo,
                    long long ldb, const float* bias, float alpha, int M, int K, int n_tok, int heads, void* qk_hi,
                    void* qk_lo, long long ld_qk, void* vt_hi, void* vt_lo, long long ld_vt, const int* n_dev,
                    cudaStream_t stream) {
  MADTP_CHECK_ARG(a_hi && a_lo && w_hi && w_lo && qk_hi && qk_lo && vt_hi && vt_lo, "gemm_qkv: null pointer");
  MADTP_CHECK_ARG(M >= 0 && K > 0 && n_tok > 0 && heads > 0 && M % n_tok == 0, "gemm_qkv: bad shape M=%d n_tok=%d", M,
                  n_tok);
  MADTP_CHECK_ARG(ld_qk % 8 == 0 && ld_qk >= 2LL * heads * 64 && ld_vt >= n_tok && ld_vt % 8 == 0,
                  "gemm_qkv: bad leading dimensions");
  MADTP_CHECK_ARG((reinterpret_cast<uintptr_t>(qk_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(qk_lo) & 15) == 0 &&
                      (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
                  "gemm_qkv: outputs and bias must be 16-byte aligned");
  if (M == 0) return kOk;
  GemmEpilogue ep = {};
  ep.c = qk_hi;
  ep.c_lo = static_cast<__half*>(qk_lo);
  ep.ldc = ld_qk;
  ep.bias = bias;
  ep.alpha = alpha;
  ep.mode = 1;
  ep.vt_hi = static_cast<__half*>(vt_hi);
  ep.vt_lo = static_cast<__half*>(vt_lo);
  ep.ld_vt = ld_vt;
  ep.n_tok = n_tok;
  ep.heads = heads;
  ep.qk_cols = 2 * heads * 64;
  ep.m_dev = n_dev;              // dynamic tokens per sequence: M = *n_dev * (number of sequences)
  ep.m_mult = M / n_tok;
  return launch_tf32<256, true>(a_hi, a_lo, lda, w_hi, w_lo, ldb, ep, M, 3 * heads * 64, K, stream);
}

// Pick the N tile that wastes the fewest SM-waves (persistent grid of num_sms CTAs).
static int pick_block_n(int M, int N) {
  const int sms = num_sms();
  const int m_tiles = (M + 127) / 128;
  auto cost = [&](int bn) {
    const long long tiles = 1LL * m_tiles * ((N + bn - 1) / bn);
    const long long waves = (tiles + sms - 1) / sms;
    return waves * bn;  // time ~ waves x tile width
  };
  // a 128-wide tile re-reads A twice as often and pays the per-tile epilogue twice per output column: measured on
  // B200 (M = 22208, fc1 + fc2) 245 us against 184 us for the 256-wide tile at equal wave counts, so the narrow tile
  // must save more than a quarter of the waves to win
  return 4 * cost(256) <= 5 * cost(128) ? 256 : 128;
}

int launch_gemm(int precision, const void* a, const void* a_lo, long long lda, const void* b, const void* b_lo,
                long long ldb, const GemmEpilogue& ep, int M, int N, int K, cudaStream_t stream) {
  MADTP_CHECK_ARG(M >= 0 && N > 0 && K > 0, "bad GEMM shape M=%d N=%d K=%d", M, N, K);
  MADTP_CHECK_ARG(a && b && ep.c, "null GEMM operand");
  if (M == 0) return kOk;
  if (precision == kGemmSimtF32 && M <= 64 && (K & 3) == 0 && (lda & 3) == 0 && (ldb & 3) == 0 &&
      ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
    // a handful of rows (classification heads): one warp per output element, 128-bit loads
    const long long outs = static_cast<long long>(M) * N;
    gemm_rowdot_kernel<<<static_cast<int>((outs + 7) / 8), 256, 0, stream>>>(
        static_cast<const float*>(a), lda, static_cast<const float*>(b), ldb, ep, M, N, K);
    MADTP_LAUNCH_CHECK();
    return kOk;
  }
  if (precision == kGemmSimtF32) {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(static_cast<const float*>(a), lda, static_cast<const float*>(b), ldb,
                                               ep, M, N, K);
    MADTP_LAUNCH_CHECK();
    return kOk;
  }
  const int bn = pick_block_n(M, N);
  if (precision == kGemmTF32x3) {
    MADTP_CHECK_ARG(a_lo && b_lo, "TF32x3 GEMM needs the lo halves of both operands");
    return bn == 256 ? launch_tf32<256>(a, a_lo, lda, b, b_lo, ldb, ep, M, N, K, stream)
                     : launch_tf32<128>(a, a_lo, lda, b, b_lo, ldb, ep, M, N, K, stream);
  }
  if (precision == kGemmF16x3) {
    MADTP_CHECK_ARG(a_lo && b_lo, "F16x3 GEMM needs the lo halves of both operands");
    return bn == 256 ? launch_tf32<256, true>(a, a_lo, lda, b, b_lo, ldb, ep, M, N, K, stream)
                     : launch_tf32<128, true>(a, a_lo, lda, b, b_lo, ldb, ep, M, N, K, stream);
  }
  if (precision == kGemmF16) {
    // CTA pairs (cta_group::2, 256 x 256 per pair) are implemented and verified but OFF by default (MADTP_PAIR=1):
    // measured on B200 they gain 3-7 % on warm, isolated GEMMs (fc2 at M = 36928: 154 us against 165 us) and nothing
    // inside the forward, where these GEMMs start on cold operands.
    static const bool use_pair = getenv("MADTP_PAIR") != nullptr;
    const long long tiles256 = static_cast<long long>((M + 127) / 128) * ((N + 255) / 256);
    if (use_pair && bn == 256 && tiles256 >= 2LL * num_sms())
      return launch_tc<256, false, true>(a, nullptr, lda, b, nullptr, ldb, ep, M, N, K, stream);
    // few rows (the text encoders: M = 640 and shrinking): a launch is bound by how fast each SM can pull its own
    // operands through the pipeline, so 64-wide tiles on twice as many SMs finish sooner
    const long long m_tiles = (M + 127) / 128;
    const long long tiles128 = m_tiles * ((N + 127) / 128), tiles64 = m_tiles * ((N + 63) / 64);
    static const bool no_bn64 = getenv("MADTP_NO_BN64") != nullptr;
    if (!no_bn64 && bn == 128 && 2 * tiles128 <= num_sms() && tiles64 <= num_sms())
      return launch_tc<64, false>(a, nullptr, lda, b, nullptr, ldb, ep, M, N, K, stream);
    return bn == 256 ? launch_tc<256, false>(a, nullptr, lda, b, nullptr, ldb, ep, M, N, K, stream)
                     : launch_tc<128, false>(a, nullptr, lda, b, nullptr, ldb, ep, M, N, K, stream);
  }
  set_error("unknown GEMM precision %d", precision);
  return kInvalidArgument;
}

}  // namespace madtp
