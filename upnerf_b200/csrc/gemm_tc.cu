// Dense layer on the 5th-generation tensor cores (tcgen05) -- upnerf_gemm_bf16.
//
//   C[M,N] (bf16) = epilogue( A[M,K] (bf16, K-major) * B[N,K]^T (bf16, K-major) )
//
// Replaces the nn.Linear (+activation) calls of NeRF.forward (reference models/nerf.py:84-123)
// and, with transposed weights, their data-gradient in backward.
//
// Structure (one persistent CTA per SM, 18 warps, warp-specialised):
//   warp 0       TMA producer: A tile 128x64 and B tile Nx64 per stage, 128-byte swizzle
//   warp 1       TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x N x 16)
//   warps 2..17  epilogue, four groups of four warps (a group covers the four TMEM lane
//                quadrants; thread = output row).  Output boxes of 128 rows x 64 columns are
//                dealt round-robin to the groups; each group tcgen05.ld's its box, applies
//                bias / per-ray bias / rank-1 / aux-add / ReLU / ReLU-mask / row-dot heads,
//                packs bf16 into its swizzled smem box and TMA-stores it.
// Two 256-column TMEM accumulators ping-pong so the epilogue of tile i overlaps the MMAs of
// tile i+1.  The epilogue is the instruction-bound part of a 256-wide layer (one warp per
// SM sub-partition cannot hide its own latencies), hence four warps per sub-partition.
// The aux operand (residual or ReLU mask) is TMA-loaded into the very smem box the result
// is later stored from, one box ahead of its use.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "ptx_sm100.cuh"

namespace upnerf {
namespace {

using namespace ptx;

constexpr int kMaxStages = 8;      // A ring depth = min(kMaxStages, what the B ring leaves of kRingBytes): 5 at N = 256, 7 at 128, 8 at 64
constexpr int kBStages = 4;        // B (the layer's weights) comes from L2: 64 KB of stages (2 at N = 256, 4 below) cover its latency
constexpr int kBM = 128;          // rows per tile (UMMA M)
constexpr int kBK = 64;           // K elements per stage = one 128-byte swizzle span
constexpr int kABytes = kBM * 128;
constexpr int kBBytesMax = 256 * 128;
constexpr int kRingBytes = 3 * (kBM * 128 + kBBytesMax);   // operand area: [B ring: 2 x (N x 64)] [A ring: n x (128 x 64)]
constexpr int kCBytes = kBM * 128;  // one 128 x 64 bf16 output box
constexpr int kGroups = 4;          // epilogue warp groups (one C box each)
constexpr int kEpiThreads = kGroups * 128;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kMaxN = 256;
constexpr int kMaxHeads = 3;

constexpr int kOffA = 0;
constexpr int kOffC = kOffA + kRingBytes;
constexpr int kOffVec = kOffC + kGroups * kCBytes;
constexpr int kVecFloats = kMaxN * (2 + kMaxHeads);
constexpr int kOffHead = kOffVec + kVecFloats * 4;
constexpr int kHeadFloats = 2 * kGroups * kBM * kMaxHeads;   // [acc][group][row][head]
constexpr int kOffBar = kOffHead + kHeadFloats * 4;
constexpr int kNumBars = 2 * kMaxStages + 2 * kBStages + 4 + kGroups;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmem + 16 + 1024;  // + slack for manual 1024-byte alignment
static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");

struct GemmArgs {
  int64_t M;
  int N;
  int K;
  int kb1;        // K-blocks (of 64) taken from the first A operand; the rest come from the second (gemm2)
  int num_tiles;
  upnerf_epilogue ep;
  // lsu_store: finished 128 x 64 output boxes are copied out by the group's own threads (coalesced
  // 16-byte st.global read back from the swizzled box) instead of a TMA store, which takes the
  // output stream off the SM's TMA unit (~27 B/clk, shared with the operand loads)
  int lsu_store;
  __nv_bfloat16* C;
  int64_t ldc;
};

__device__ __forceinline__ float softplus_ref(float x) {
  // torch.nn.Softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.f / (1.f + expf(-x)); }

// kRB: the epilogue adds a per-ray bias; kHD: it computes row-dot heads.  Compile-time so that the plain layers
// (every data-gradient GEMM) keep the short epilogue: with both as run-time branches inside the unrolled column
// loop the plain N = K = 256 layer went from 0.159 to 0.181 ms.
template <bool kRB, bool kHD>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmAux, const __grid_constant__ GemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem + kOffA;
  uint8_t* sC = smem + kOffC;
  float* sVec = reinterpret_cast<float*>(smem + kOffVec);
  float* sHead = reinterpret_cast<float*>(smem + kOffHead);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_full = bars;                       // A stage landed
  uint64_t* bar_empty = bars + kMaxStages;         // A stage consumed
  uint64_t* bar_bfull = bars + 2 * kMaxStages;     // B stage landed
  uint64_t* bar_bempty = bar_bfull + kBStages;     // B stage consumed
  uint64_t* bar_tfull = bar_bempty + kBStages;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint64_t* bar_aux = bar_tempty + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N = args.N;
  const int kblocks = args.K / kBK;
  const int nchunks = N / 64;
  const bool has_aux = args.ep.aux_mode != 0;
  // Ring geometry of THIS layer.  A (activations) streams from HBM, B (the weights, re-read for every tile) from
  // L2: they have rings of their own so that the HBM stream can run several stages ahead of the short B ring --
  // with one joint ring a 256-wide layer kept only 3 x 16 KB of A in flight per SM, too little under the step's
  // HBM contention (the same effect as in wgrad_tc.cu).
  const int kBStageBytes = N * 128;
  const int nB = 65536 / kBStageBytes < kBStages ? 65536 / kBStageBytes : kBStages;   // B ring depth of this layer
  uint8_t* sB = sA;                                           // B ring first
  uint8_t* sAr = sA + nB * kBStageBytes;                      // then the A ring
  const int a_room = (kRingBytes - nB * kBStageBytes) / kABytes;
  const int kStages = a_room < kMaxStages ? a_room : kMaxStages;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    if (args.kb1 < args.K / kBK) prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    if (has_aux) prefetch_tmap(&tmAux);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&bar_bfull[i], 1);
      mbar_init(&bar_bempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], kEpiThreads);
    }
    for (int i = 0; i < kGroups; ++i) mbar_init(&bar_aux[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  if (warp >= 2) {
    // epilogue vectors -> smem (read with broadcast in the inner loops)
    const int t = threadIdx.x - 64;
    for (int i = t; i < N; i += kEpiThreads) {
      sVec[i] = args.ep.bias ? args.ep.bias[i] : 0.f;
      sVec[kMaxN + i] = args.ep.rank1_row ? args.ep.rank1_col[i] : 0.f;
      for (int h = 0; h < args.ep.n_heads; ++h)
        sVec[(2 + h) * kMaxN + i] = args.ep.head_w[h * N + i];
    }
    // two-contributor head exchange (below): [3 buffers][row][3 sums + arrival count], zero between uses
    for (int i = t; i < 3 * kBM * 4; i += kEpiThreads) sHead[i] = 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // one thread issues both streams in order; the A load of K-block j is issued `skew` blocks before the B
      // load of the same block, so A runs ahead by up to its whole ring while B stays two stages deep
      int my_tiles = 0;
      for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) ++my_tiles;
      const int total = my_tiles * kblocks;
      const int skew = kStages - nB > 0 ? kStages - nB : 0;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int a_tile = blockIdx.x, a_kb = 0, b_kb = 0;
      for (int j = 0; j < total + skew; ++j) {
        if (j < total) {
          mbar_wait(&bar_empty[sa], pha ^ 1);
          mbar_arrive_expect_tx(&bar_full[sa], kABytes);
          if (a_kb < args.kb1) tma_load_2d(sAr + sa * kABytes, &tmA, &bar_full[sa], a_kb * kBK, a_tile * kBM);
          else tma_load_2d(sAr + sa * kABytes, &tmA2, &bar_full[sa], (a_kb - args.kb1) * kBK, a_tile * kBM);
          if (++a_kb == kblocks) {
            a_kb = 0;
            a_tile += gridDim.x;
          }
          if (++sa == kStages) {
            sa = 0;
            pha ^= 1;
          }
        }
        if (j >= skew) {
          mbar_wait(&bar_bempty[sb], phb ^ 1);
          mbar_arrive_expect_tx(&bar_bfull[sb], kBStageBytes);
          tma_load_2d(sB + sb * kBStageBytes, &tmB, &bar_bfull[sb], b_kb * kBK, 0);
          if (++b_kb == kblocks) b_kb = 0;
          if (++sb == nB) {
            sb = 0;
            phb ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kBM, N, 0, 0);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int t = 0;
      for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x, ++t) {
        const int acc = t & 1;
        mbar_wait(&bar_tempty[acc], ((t >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&bar_full[sa], pha);
          mbar_wait(&bar_bfull[sb], phb);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sAr + sa * kABytes);
          const uint32_t b_addr = smem_u32(sB + sb * kBStageBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_desc(a_addr + k * 32, 16, 1024, kLayoutSw128);
            const uint64_t db = umma_desc(b_addr + k * 32, 16, 1024, kLayoutSw128);
            mma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          mma_commit(&bar_empty[sa]);    // both stages are free when these MMAs retire
          mma_commit(&bar_bempty[sb]);
          if (++sa == kStages) {
            sa = 0;
            pha ^= 1;
          }
          if (++sb == nB) {
            sb = 0;
            phb ^= 1;
          }
        }
        mma_commit(&bar_tfull[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int grp = ew >> 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    const bool leader = ((ew & 3) == 0) && lane == 0;
    const uint32_t bar_id = 1 + grp;  // named barrier of this group (128 threads)
    const upnerf_epilogue& ep = args.ep;
    const int nh = kHD ? ep.n_heads : 0;
    uint8_t* cbuf = sC + grp * kCBytes;
    uint8_t* crow_ptr = cbuf + row_in_tile * 128;
    uint64_t* my_aux = &bar_aux[grp];

    int my_tiles = 0;
    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) ++my_tiles;
    const int total_boxes = my_tiles * nchunks;
    // first box of this group at or after box index b0 (box = local_tile * nchunks + chunk)
    auto next_box = [&](int b0) -> int {
      const int b = b0 + ((grp - b0) & (kGroups - 1));
      return b < total_boxes ? b : -1;
    };
    auto issue_aux = [&](int box) {
      const int lt = box / nchunks;
      const int ch = box - lt * nchunks;
      const int tile = blockIdx.x + lt * gridDim.x;
      mbar_arrive_expect_tx(my_aux, kCBytes);
      tma_load_2d(cbuf, &tmAux, my_aux, ch * 64, tile * kBM);
    };
    if (has_aux && leader) {
      const int b = next_box(0);
      if (b >= 0) issue_aux(b);
    }
    uint32_t qg = 0;  // boxes this group has processed (aux barrier phase)
    // Per-ray bias rows are read with broadcast loads inside the box loop; their first touch
    // would expose a full L2 round trip per 8 columns, so the lines a warp needs for local tile
    // `lt` are prefetched into L1 one tile ahead (lanes 0/1 cover the first row's ray, lanes
    // 30/31 the last row's: a warp of 32 consecutive rows spans at most two rays for S >= 32).
    auto prefetch_ray_bias = [&](int lt) {
      if (!kRB || !ep.ray_bias || ep.rows_per_ray % 32 == 0 || (lane > 1 && lane < 30)) return;
      const int tile = blockIdx.x + lt * gridDim.x;
      if (tile >= args.num_tiles) return;
      int64_t row = static_cast<int64_t>(tile) * kBM + row_in_tile;
      if (row >= args.M) row = args.M - 1;
      const float* rbp = ep.ray_bias + (row / ep.rows_per_ray) * N + (lane & 1) * 32;
      for (int ch = (grp - lt * nchunks) & (kGroups - 1); ch < nchunks; ch += kGroups)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rbp + ch * 64));
    };
    prefetch_ray_bias(0);

    int t = 0;
    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      const int64_t grow = static_cast<int64_t>(tile) * kBM + row_in_tile;
      const bool row_ok = grow < args.M;
      const int64_t crow = row_ok ? grow : (args.M - 1);
      const float r1 = ep.rank1_row ? ep.rank1_row[crow] : 0.f;
      const float* rb = (kRB && ep.ray_bias) ? ep.ray_bias + (crow / ep.rows_per_ray) * N : nullptr;
      const bool rb_uniform = kRB && rb && (ep.rows_per_ray % 32 == 0);
      // row-dot heads accumulate as packed pairs (even / odd columns): fma.rn.f32x2 halves the instruction count and
      // the dependent-chain length of the 64 x n_heads multiply-adds per box and row
      float2 hacc[kMaxHeads] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
      prefetch_ray_bias(t + 1);

      // chunks of this tile owned by this group: first, first + 4, ...
      const int first = (grp - t * nchunks) & (kGroups - 1);
      int last = -1;
      for (int ch = first; ch < nchunks; ch += kGroups) last = ch;

      mbar_wait(&bar_tfull[acc], (t >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256;
      if (last < 0) {
        tc_fence_before_sync();
        mbar_arrive(&bar_tempty[acc]);
      }

      for (int ch = first; ch < nchunks; ch += kGroups) {
        const int nh_here = ch * 64 >= ep.head_col_begin ? nh : 0;   // (warp-uniform) heads skip the columns before
        // Per-ray bias of this box.  A warp's 32 rows start at a multiple of 32, so with rows_per_ray % 32 == 0
        // they all belong to ONE ray: each lane fetches two of the box's 64 columns now (one coalesced 256-byte
        // read, in flight across the barrier and the TMEM load below) and the values reach the rows by shuffle.
        // (Sixteen dependent 16-byte loads per thread and box -- L1 hits or not -- were 42 % of this kernel's
        // stall samples on the stacked candidate|rgb layer.)
        float2 rbv = make_float2(0.f, 0.f);
        if (kRB && rb_uniform) rbv = __ldg(reinterpret_cast<const float2*>(rb + ch * 64) + lane);
        // the smem box is free once the previous store of this group has been read out (the
        // leader waited for that before issuing the aux load / arriving at the barrier)
        if (has_aux) {
          mbar_wait(my_aux, qg & 1);
        } else {
          named_bar_sync(bar_id, 128);
        }
        // accumulator columns per TMEM load: 32, or 16 in the variant that carries both a per-ray bias and row-dot
        // heads (its epilogue spilled registers inside this loop at the 96 the 18-warp block allows: 5 warps per
        // scheduler x 32 x 96 is all of a scheduler's register file)
        constexpr int kLd = (kRB && kHD) ? 16 : 32;
#pragma unroll
        for (int part = 0; part < 64 / kLd; ++part) {
          uint32_t acc_r[kLd];
          if constexpr (kLd == 32) tmem_ld_32x32(taddr + ch * 64 + part * 32, acc_r);
          else tmem_ld_32x16(taddr + ch * 64 + part * 16, acc_r);
          tmem_ld_wait();
          if (part == 64 / kLd - 1 && ch == last) {
            // accumulator fully read by this thread: hand it back to the MMA warp
            tc_fence_before_sync();
            mbar_arrive(&bar_tempty[acc]);
          }
#pragma unroll
          for (int c4 = 0; c4 < kLd / 8; ++c4) {
            const int c8 = part * (kLd / 8) + c4;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(acc_r[c4 * 8 + e]);
            const int col = ch * 64 + c8 * 8;
            if (!kRB || ep.bias) {   // (a layer whose bias rides in its per-ray bias passes none: no loads at all)
              const float4 b0 = *reinterpret_cast<const float4*>(&sVec[col]);
              const float4 b1 = *reinterpret_cast<const float4*>(&sVec[col + 4]);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (!kRB) {
            } else if (rb_uniform) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[2 * e] += __shfl_sync(0xffffffffu, rbv.x, c8 * 4 + e);
                v[2 * e + 1] += __shfl_sync(0xffffffffu, rbv.y, c8 * 4 + e);
              }
            } else if (rb) {
              const float4 p0 = __ldg(reinterpret_cast<const float4*>(rb + col));
              const float4 p1 = __ldg(reinterpret_cast<const float4*>(rb + col + 4));
              v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
              v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
            }
            if (ep.rank1_row) {
              const float4 c0 = *reinterpret_cast<const float4*>(&sVec[kMaxN + col]);
              const float4 c1 = *reinterpret_cast<const float4*>(&sVec[kMaxN + col + 4]);
              v[0] += r1 * c0.x; v[1] += r1 * c0.y; v[2] += r1 * c0.z; v[3] += r1 * c0.w;
              v[4] += r1 * c1.x; v[5] += r1 * c1.y; v[6] += r1 * c1.z; v[7] += r1 * c1.w;
            }
            uint4* slot = reinterpret_cast<uint4*>(crow_ptr + ((c8 ^ (row_in_tile & 7)) << 4));
            float a[8];
            if (has_aux) {
              const uint4 au = *slot;
              const __nv_bfloat162* ab = reinterpret_cast<const __nv_bfloat162*>(&au);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(ab[e]);
                a[2 * e] = f.x;
                a[2 * e + 1] = f.y;
              }
              if (ep.aux_mode == 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += a[e];
              }
            }
            if (ep.act == 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            if (ep.aux_mode == 2) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = a[e] > 0.f ? v[e] : 0.f;
            }
            // (fully unrolled with a predicate so that hacc[] stays in registers)
#pragma unroll
            for (int h = 0; h < kMaxHeads; ++h) {
              if (kHD && h < nh_here) {
                const float4 w0 = *reinterpret_cast<const float4*>(&sVec[(2 + h) * kMaxN + col]);
                const float4 w1 = *reinterpret_cast<const float4*>(&sVec[(2 + h) * kMaxN + col + 4]);
                hacc[h] = __ffma2_rn(make_float2(v[0], v[1]), make_float2(w0.x, w0.y), hacc[h]);
                hacc[h] = __ffma2_rn(make_float2(v[2], v[3]), make_float2(w0.z, w0.w), hacc[h]);
                hacc[h] = __ffma2_rn(make_float2(v[4], v[5]), make_float2(w1.x, w1.y), hacc[h]);
                hacc[h] = __ffma2_rn(make_float2(v[6], v[7]), make_float2(w1.z, w1.w), hacc[h]);
              }
            }
            uint4 out;
            __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
            for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            *slot = out;
          }
        }
        if (args.lsu_store) {
          named_bar_sync(bar_id, 128);
          const int wq = ew & 3;
          const uint32_t cb = smem_u32(cbuf);
          __nv_bfloat16* obase = args.C + ch * 64 + (lane & 7) * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr2 = wq * 32 + i * 4 + (lane >> 3);
            const float4 vv = lds128(cb + rr2 * 128 + (((lane & 7) ^ (rr2 & 7)) << 4));
            const int64_t gr = static_cast<int64_t>(tile) * kBM + rr2;
            if (gr < args.M) __stcs(reinterpret_cast<float4*>(obase + gr * args.ldc), vv);
          }
          if (has_aux) {
            named_bar_sync(bar_id, 128);   // every thread has read its rows: the box may take the next aux tile
            if (leader) {
              const int nb = next_box(t * nchunks + ch + 1);
              if (nb >= 0) issue_aux(nb);
            }
          }
        } else {
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (leader) {
            tma_store_2d(&tmC, cbuf, ch * 64, tile * kBM);
            tma_store_commit();
            tma_store_wait_read<0>();  // box reusable; overlaps the other groups' work
            if (has_aux) {
              const int nb = next_box(t * nchunks + ch + 1);
              if (nb >= 0) issue_aux(nb);
            }
          }
        }
        ++qg;
      }

      if (kHD && nh > 0) {
        // Row-dot heads: every group that owned a chunk at or after head_col_begin holds a partial sum per row;
        // one group combines them.  Where possible only the contributing groups meet at the named barrier (a
        // barrier over all 512 epilogue threads per tile keeps the groups in lock step and cost the
        // 128-wide layers 20 % of their time); with 2 chunks per tile the set alternates between {0,1} and {2,3}
        // -- one barrier id per parity.
        const int c0 = ep.head_col_begin >> 6;
        const int ncontrib = nchunks - c0;
        // (With 4 chunks per tile everybody meets and group 0 -- whose own box carries no head columns in the
        //  stacked layer -- combines: making a contributing group combine as well measured 27 us SLOWER there,
        //  it becomes the critical group.  With 2 chunks per tile the two idle groups skip the barrier: -16 us.)
        const bool regular = nchunks == 2 || (ncontrib == 1 && nchunks <= 2);
        const bool mine = regular ? (first >= c0 && first < nchunks) : true;
        const int comb_grp = regular ? ((c0 + t * nchunks) & (kGroups - 1)) : 0;
        // exchange buffer: a slot is rewritten two uses of the same group set later, i.e. after a barrier the
        // combiner of the earlier tile has passed too (sets alternate tile by tile when nchunks == 2)
        const int hb = nchunks == 2 ? ((t >> 1) & 1) : (t & 1);
        const bool mine2 = first >= c0 && first < nchunks;   // my box carries head columns
        if (ncontrib == 2) {
          // Exactly two groups hold a partial sum per row (the stacked candidate|rgb layer: boxes 2 and 3; 128-wide
          // layers: both boxes).  No barrier: each adds its partial into the row's slot with shared-memory atomics
          // and counts its arrival; whoever arrives second finalises the row and clears the slot.  a + b is
          // commutative, so the result does not depend on who is first.  Three buffers: a group cannot be three
          // tiles ahead of another (it would need the accumulator the other one has not yet handed back).
          if (mine2) {
            float* acc3 = sHead + ((t % 3) * kBM + row_in_tile) * 4;
#pragma unroll
            for (int h = 0; h < kMaxHeads; ++h)
              if (h < nh) atomicAdd(&acc3[h], hacc[h].x + hacc[h].y);
            __threadfence_block();
            const int prev = atomicAdd(reinterpret_cast<int*>(&acc3[3]), 1);
            if (prev == 1) {
              __threadfence_block();
              volatile float* va = acc3;
              for (int h = 0; h < nh; ++h) {
                float x = ep.head_b[h] + va[h];
                va[h] = 0.f;
                if (ep.head_act == 1) x = softplus_ref(x);
                else if (ep.head_act == 2) x = sigmoid_ref(x);
                if (row_ok) ep.head_out[grow * nh + h] = x;
              }
              *reinterpret_cast<volatile int*>(&acc3[3]) = 0;
            }
          }
        } else if (mine) {
          float* slot = sHead + ((hb * kGroups + grp) * kBM + row_in_tile) * kMaxHeads;
#pragma unroll
          for (int h = 0; h < kMaxHeads; ++h)
            if (h < nh) slot[h] = hacc[h].x + hacc[h].y;
          if (!regular) named_bar_sync(6, kEpiThreads);
          else if (ncontrib > 1) named_bar_sync(6 + (nchunks == 2 ? (t & 1) : 0), 128 * ncontrib);
          if (grp == comb_grp && row_ok) {
            for (int h = 0; h < nh; ++h) {
              float x = ep.head_b[h];
              if (regular) {
                for (int j = 0; j < ncontrib; ++j)
                  x += sHead[((hb * kGroups + ((c0 + j + t * nchunks) & (kGroups - 1))) * kBM + row_in_tile) * kMaxHeads + h];
              } else {
#pragma unroll
                for (int g = 0; g < kGroups; ++g)
                  x += sHead[((hb * kGroups + g) * kBM + row_in_tile) * kMaxHeads + h];
              }
              if (ep.head_act == 1) x = softplus_ref(x);
              else if (ep.head_act == 2) x = sigmoid_ref(x);
              ep.head_out[grow * nh + h] = x;
            }
          }
        }
      }
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace
}  // namespace upnerf

namespace upnerf {
static int gemm_launch(const void* A, int64_t lda, int K1, const void* A2, int64_t lda2, const void* B, int64_t ldb,
                       void* C, int64_t ldc, int64_t M, int N, int K, const upnerf_epilogue* ep, void* stream);
}  // namespace upnerf

extern "C" int upnerf_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                                int64_t ldc, int64_t M, int N, int K, const upnerf_epilogue* ep,
                                void* stream) {
  return upnerf::gemm_launch(A, lda, K, nullptr, 0, B, ldb, C, ldc, M, N, K, ep, stream);
}

extern "C" int upnerf_gemm2_bf16(const void* A1, int64_t lda1, int K1, const void* A2, int64_t lda2, int K2,
                                 const void* B, int64_t ldb, void* C, int64_t ldc, int64_t M, int N,
                                 const upnerf_epilogue* ep, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(A2 && K1 >= 64 && K1 % 64 == 0 && K2 >= 64 && K2 % 64 == 0, UPNERF_ERR_BAD_SHAPE,
                 "gemm2_bf16: K1=%d K2=%d must be positive multiples of 64", K1, K2);
  return gemm_launch(A1, lda1, K1, A2, lda2, B, ldb, C, ldc, M, N, K1 + K2, ep, stream);
}

static int upnerf::gemm_launch(const void* A, int64_t lda, int K1, const void* A2, int64_t lda2, const void* B,
                               int64_t ldb, void* C, int64_t ldc, int64_t M, int N, int K, const upnerf_epilogue* ep,
                               void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(M > 0, UPNERF_ERR_BAD_SHAPE, "gemm_bf16: M=%lld", (long long)M);
  UPNERF_REQUIRE(N >= 64 && N <= kMaxN && N % 64 == 0, UPNERF_ERR_BAD_SHAPE,
                 "gemm_bf16: N=%d must be a multiple of 64 in [64,256]", N);
  UPNERF_REQUIRE(K >= 64 && K % 64 == 0, UPNERF_ERR_BAD_SHAPE,
                 "gemm_bf16: K=%d must be a positive multiple of 64", K);
  GemmArgs args;
  memset(&args, 0, sizeof(args));
  args.M = M;
  args.N = N;
  args.K = K;
  const int64_t tiles = ceil_div64(M, kBM);
  UPNERF_REQUIRE(tiles < (1ll << 28), UPNERF_ERR_BAD_SHAPE, "gemm_bf16: M too large");
  args.num_tiles = static_cast<int>(tiles);
  if (ep) args.ep = *ep;
  UPNERF_REQUIRE(args.ep.n_heads >= 0 && args.ep.n_heads <= kMaxHeads, UPNERF_ERR_BAD_SHAPE,
                 "gemm_bf16: n_heads=%d", args.ep.n_heads);
  UPNERF_REQUIRE(args.ep.head_col_begin >= 0 && args.ep.head_col_begin % 64 == 0, UPNERF_ERR_BAD_SHAPE,
                 "gemm_bf16: head_col_begin=%d must be a multiple of 64", args.ep.head_col_begin);
  UPNERF_REQUIRE(args.ep.aux_mode == 0 || args.ep.aux != nullptr, UPNERF_ERR_BAD_SHAPE,
                 "gemm_bf16: aux_mode set without aux");
  UPNERF_REQUIRE(!args.ep.ray_bias || args.ep.rows_per_ray > 0, UPNERF_ERR_BAD_SHAPE,
                 "gemm_bf16: ray_bias without rows_per_ray");

  {
    // opt-in: measured slower here (N = K = 256, M = 786k: 0.154 -> 0.196 ms) -- this kernel's TMA
    // unit is not saturated (~19 B/clk per SM), so the copy-out only lengthens the epilogue
    const char* e = getenv("UPNERF_GEMM_LSU_STORE");
    args.lsu_store = (e && e[0] == '1') ? 1 : 0;
    args.C = static_cast<__nv_bfloat16*>(C);
    args.ldc = ldc;
    if ((reinterpret_cast<uintptr_t>(C) & 15) != 0 || (ldc & 7) != 0) args.lsu_store = 0;
  }
  args.kb1 = K1 / kBK;
  CUtensorMap tmA, tmA2, tmB, tmC, tmAux;
  UPNERF_TRY(make_tmap_bf16_2d(&tmA, A, M, K1, lda, kBM, kBK));
  if (A2) UPNERF_TRY(make_tmap_bf16_2d(&tmA2, A2, M, K - K1, lda2, kBM, kBK));
  else tmA2 = tmA;
  UPNERF_TRY(make_tmap_bf16_2d(&tmB, B, N, K, ldb, N, kBK));
  UPNERF_TRY(make_tmap_bf16_2d(&tmC, C, M, N, ldc, kBM, 64));
  if (args.ep.aux_mode != 0) {
    UPNERF_TRY(make_tmap_bf16_2d(&tmAux, args.ep.aux, M, N, args.ep.ldaux, kBM, 64));
  } else {
    tmAux = tmC;
  }
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                            const CUtensorMap, const GemmArgs);
  static const KernelFn kernels[4] = {gemm_tc_kernel<false, false>, gemm_tc_kernel<true, false>,
                                      gemm_tc_kernel<false, true>, gemm_tc_kernel<true, true>};
  static bool attr_set = false;
  if (!attr_set) {
    for (KernelFn f : kernels)
      UPNERF_CHECK_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const KernelFn fn = kernels[(args.ep.ray_bias ? 1 : 0) + (args.ep.n_heads > 0 ? 2 : 0)];
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  LaunchScope scope(kCatGemmTc, as_stream(stream), 2.0 * M * N * K,
                    2.0 * M * (K + N * (args.ep.aux_mode ? 2 : 1)) + 2.0 * N * K);
  fn<<<grid, kThreads, kSmemBytes, as_stream(stream)>>>(tmA, tmA2, tmB, tmC, tmAux, args);
  UPNERF_CHECK_LAUNCH("gemm_tc_kernel");
  return UPNERF_OK;
}
