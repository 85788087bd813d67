// Weight gradient on tcgen05 -- upnerf_wgrad_bf16.
//
//   dW[n, colmap(k)] += sum_m dY[m,n] * X[m,k]      db[n] += sum_m dY[m,n]
//
// This is the autograd backward of every nn.Linear on the path (reference
// models/nerf.py:38-78) with respect to weight and bias.  The reduction runs over the
// SAMPLE axis, so both operands are read "MN-major": dY[m, n] supplies A[n_out, k=m] and
// X[m, k] supplies B[k_in, k=m] without any transpose pass -- the TMA boxes land in the
// canonical MN-major 128-byte-swizzle layout and the UMMA descriptors say so.
//
// Grid: (N/128 output-row chunks) x (splits over the sample axis).  Each CTA accumulates
// a 128 x K fp32 tile in TMEM over its sample range.  The bias gradient is one extra N=16 MMA
// per K-step against an all-ones B tile.
//
// Split reduction, two modes:
//   * partial != nullptr (the render path): every CTA writes its tile with plain coalesced
//     stores to partial[split][chunk][k][row] and ONE wgrad_reduce_kernel per network pass sums
//     the splits of all layers in a fixed order into the parameter gradients.  Deterministic,
//     and it removes the ~25 us per launch that 4.8 M contended fp32 L2 atomics cost (measured:
//     0.068 ms at M = 262144 where the operand stream alone takes 0.038 ms).
//   * partial == nullptr (stand-alone upnerf_wgrad_bf16): fp32 atomics straight into dW.
#include <cuda_bf16.h>
#include <string.h>

#include "common.h"
#include "internal.h"
#include "ptx_sm100.cuh"

namespace upnerf {
namespace {

using namespace ptx;

constexpr int kMaxStages = 8;            // ring depth = min(kMaxStages, kStageBoxes / boxes per stage): 4 at K = 256
constexpr int kStageBoxes = 27;          // shared memory given to the operand ring, in 8 KB boxes
constexpr int kBS = 64;                 // samples per stage
constexpr int kBoxBytes = kBS * 128;    // one 64-sample x 64-column bf16 box
constexpr int kABoxes = 2;              // 128 output rows
constexpr int kMaxBBoxes = 5;           // K <= 320
constexpr int kOnesBytes = 1024;
constexpr int kMaxK = 320;
constexpr int kBiasCol = 384;           // TMEM column of the bias-gradient accumulator
constexpr int kThreads = 192;

constexpr int kOffStage = 0;
constexpr int kOffOnes = kOffStage + kStageBoxes * kBoxBytes;
constexpr int kOffMap = kOffOnes + kOnesBytes;
constexpr int kOffBar = kOffMap + kMaxK * 4;
constexpr int kNumBars = 2 * kMaxStages + 1;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmem + 16 + 1024;

constexpr int kMaxSeg = 4;

struct WgradArgs {
  float* dW;
  int64_t lddw;
  float* dW_hi;      // optional: destination of output rows 128..255 (two stacked layers)
  int64_t lddw_hi;
  float* db;
  float* partial;    // optional: [splits][chunks][K + 1][128] fp32 split partials (column K = bias)
  int64_t M;
  int N;
  int K;
  int splits;
  int chunks;              // N / 128
  int64_t rows_per_split;  // multiple of kBS
  int n_seg;
  int seg_src[kMaxSeg], seg_len[kMaxSeg], seg_dst[kMaxSeg];
};

// A launch covers a GROUP of independent weight gradients (all layers of one network pass): the linear block
// index selects (problem, 128-row chunk, split).  Every problem gets a share of the grid proportional to the
// bytes it streams, so the whole group is ONE wave of ~sm_count CTAs that finish together -- one ramp-up, one
// tail and ~8 split partials per layer instead of one launch, one tail and ~74 partials per layer.
struct WgradGroupParams {
  CUtensorMap tmY[kMaxWgradSlots];
  CUtensorMap tmX[kMaxWgradSlots];
  WgradArgs p[kMaxWgradSlots];
  int block_begin[kMaxWgradSlots + 1];
  int n;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WgradGroupParams grp) {
  int pi = 0;
  while (pi + 1 < grp.n && static_cast<int>(blockIdx.x) >= grp.block_begin[pi + 1]) ++pi;
  const WgradArgs& args = grp.p[pi];
  const CUtensorMap* const tmY = &grp.tmY[pi];
  const CUtensorMap* const tmX = &grp.tmX[pi];
  const int local = static_cast<int>(blockIdx.x) - grp.block_begin[pi];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sStage = smem + kOffStage;
  uint8_t* sOnes = smem + kOffOnes;
  int* sMap = reinterpret_cast<int*>(smem + kOffMap);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + kMaxStages;
  uint64_t* bar_done = bars + 2 * kMaxStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunk = local % args.chunks;  // which 128 output rows
  const int split = local / args.chunks;
  const int K = args.K;
  const int bboxes = K / 64;
  // ring geometry of THIS problem: a stage holds 64 samples of both operands
  const int kStageBytes = (kABoxes + bboxes) * kBoxBytes;
  const int kStages = kStageBoxes / (kABoxes + bboxes) < kMaxStages ? kStageBoxes / (kABoxes + bboxes) : kMaxStages;

  const int64_t row_begin = static_cast<int64_t>(split) * args.rows_per_split;
  int64_t row_end = row_begin + args.rows_per_split;
  if (row_end > args.M) row_end = args.M;
  const int nsteps = row_end > row_begin ? static_cast<int>((row_end - row_begin + kBS - 1) / kBS) : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(tmY);
    prefetch_tmap(tmX);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    // all-ones bf16 tile (B operand of the bias-gradient MMA)
    for (int i = t; i < kOnesBytes / 4; i += 128)
      reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
    // packed column -> parameter column (or -1 for padding)
    for (int c = t; c < kMaxK; c += 128) {
      int d = -1;
      for (int s = 0; s < args.n_seg; ++s)
        if (c >= args.seg_src[s] && c < args.seg_src[s] + args.seg_len[s])
          d = args.seg_dst[s] + (c - args.seg_src[s]);
      sMap[c] = d;
    }
    fence_proxy_async_smem();  // generic-proxy writes to sOnes -> visible to the MMA
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (nsteps > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = (kABoxes + bboxes) * kBoxBytes;
        for (int s = 0; s < nsteps; ++s) {
          const int r0 = static_cast<int>(row_begin + static_cast<int64_t>(s) * kBS);
          mbar_wait(&bar_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bar_full[stage], tx);
          uint8_t* base = sStage + stage * kStageBytes;
          for (int j = 0; j < kABoxes; ++j)
            tma_load_2d(base + j * kBoxBytes, tmY, &bar_full[stage], chunk * 128 + j * 64, r0);
          for (int j = 0; j < bboxes; ++j)
            tma_load_2d(base + (kABoxes + j) * kBoxBytes, tmX, &bar_full[stage], j * 64, r0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const int n1 = K < 256 ? K : 256;
        const int n2 = K - n1;
        const uint32_t idesc1 = umma_idesc_bf16(128, n1, 1, 1);
        const uint32_t idesc2 = umma_idesc_bf16(128, n2 > 0 ? n2 : 16, 1, 1);
        const uint32_t idesc_ones = umma_idesc_bf16(128, 16, 1, 0);
        const uint64_t d_ones = umma_desc(smem_u32(sOnes), 128, 256, kLayoutNone);
        int stage = 0;
        uint32_t phase = 0;
        for (int s = 0; s < nsteps; ++s) {
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sStage + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kABoxes * kBoxBytes;
#pragma unroll
          for (int ks = 0; ks < kBS / 16; ++ks) {
            // 16 samples = two 8-row swizzle groups (SBO = 1024); 64-column blocks are one
            // TMA box apart (LBO = kBoxBytes).
            const uint64_t da = umma_desc(a_addr + ks * 2048, kBoxBytes, 1024, kLayoutSw128);
            const uint64_t db = umma_desc(b_addr + ks * 2048, kBoxBytes, 1024, kLayoutSw128);
            const uint32_t accum = (s | ks) != 0;
            mma_bf16_ss(tmem_base, da, db, idesc1, accum);
            if (n2 > 0) {
              const uint64_t db2 =
                  umma_desc(b_addr + 4 * kBoxBytes + ks * 2048, kBoxBytes, 1024, kLayoutSw128);
              mma_bf16_ss(tmem_base + 256, da, db2, idesc2, accum);
            }
            mma_bf16_ss(tmem_base + kBiasCol, da, d_ones, idesc_ones, accum);
          }
          mma_commit(&bar_empty[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        mma_commit(bar_done);
      }
    } else {
      const int quad = warp & 3;
      const int n = chunk * 128 + quad * 32 + lane;  // output row (always < N: N % 128 == 0)
      mbar_wait(bar_done, 0);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      if (args.partial) {
        // thread = output row: lanes hold consecutive rows, so column-major partials are
        // written as full 128-byte lines straight from the TMEM registers
        float* dst = args.partial +
                     (static_cast<int64_t>(split) * args.chunks + chunk) * (K + 1) * 128 + quad * 32 + lane;
        for (int g = 0; g < K / 32; ++g) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + g * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) dst[(g * 32 + i) * 128] = __uint_as_float(v[i]);
        }
        if (args.db) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + kBiasCol, v);
          tmem_ld_wait();
          dst[K * 128] = __uint_as_float(v[0]);
        }
      } else {
      // Transpose the 128 x K fp32 tile through shared memory (the operand stages are idle now:
      // every TMA load has landed and every MMA has retired) so that the reduction into dW is
      // COALESCED: one warp per output row, lanes on consecutive columns.  Thread-per-row
      // atomics straight from TMEM would be 32 scattered 4-byte transactions per instruction.
      float* sT = reinterpret_cast<float*>(sStage);
      const int ldt = K + 1;  // +1: the row-strided writes below hit 32 different banks
      for (int g = 0; g < K / 32; ++g) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + g * 32, v);
        tmem_ld_wait();
        float* dst = sT + (quad * 32 + lane) * ldt + g * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) dst[i] = __uint_as_float(v[i]);
      }
      named_bar_sync(1, 128);
      for (int r = warp - 2; r < 128; r += 4) {
        const float* src = sT + r * ldt;
        float* wrow = (chunk == 1 && args.dW_hi) ? args.dW_hi + static_cast<int64_t>(r) * args.lddw_hi
                                                  : args.dW + static_cast<int64_t>(chunk * 128 + r) * args.lddw;
        for (int c = lane; c < K; c += 32) {
          const int d = sMap[c];
          if (d >= 0) atomicAdd(wrow + d, src[c]);
        }
      }
      if (args.db) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + kBiasCol, v);
        tmem_ld_wait();
        atomicAdd(args.db + n, __uint_as_float(v[0]));
      }
      }
    }
  } else if (args.partial && warp >= 2) {
    // a split without rows still owns a slot of the partial buffer
    float* dst = args.partial +
                 (static_cast<int64_t>(split) * args.chunks + chunk) * (K + 1) * 128 + (threadIdx.x - 64);
    for (int c = 0; c <= K; ++c) dst[c * 128] = 0.f;
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// One block per (slot, 128-row chunk, tile of 8 packed columns).  Phase 1: thread = output row sums the splits
// of its 8 columns in a fixed order (bit-reproducible; the partials are column-major, so these reads are
// full lines).  Phase 2: the 128 x 32 tile goes through shared memory so that the read-modify-write of the
// parameter gradient is row-major -- a warp covers 8 consecutive columns (one full 32-byte sector) of 4 rows;
// the strided version touched a sector per ELEMENT and took longer than reading the partials.
constexpr int kRedCols = 8;
__global__ void __launch_bounds__(128)
wgrad_reduce_kernel(const __grid_constant__ WgradReduceList list) {
  __shared__ float tile[kRedCols][128 + 1];
  int si = 0;
  while (si + 1 < list.n && static_cast<int>(blockIdx.x) >= list.s[si + 1].block_begin) ++si;
  const WgradReduceSlot& s = list.s[si];
  const int ctiles = (s.K + 1 + kRedCols - 1) / kRedCols;
  const int local = blockIdx.x - s.block_begin;
  const int chunk = local / ctiles;
  const int col0 = (local - chunk * ctiles) * kRedCols;
  const int row = threadIdx.x;
  const int ncols = s.K + 1 - col0 < kRedCols ? s.K + 1 - col0 : kRedCols;
  const int64_t stride = static_cast<int64_t>(s.chunks) * (s.K + 1) * 128;
  const float* p0 = s.partial + (static_cast<int64_t>(chunk) * (s.K + 1) + col0) * 128 + row;
  for (int c = 0; c < ncols; ++c) {
    const float* p = p0 + c * 128;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int i = 0;
    for (; i + 8 <= s.splits; i += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldcs(p + (i + j) * stride);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j & 3] += v[j];
    }
    for (; i < s.splits; ++i) acc[i & 3] += __ldcs(p + i * stride);
    tile[c][row] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
  __syncthreads();
  const int cidx = threadIdx.x & (kRedCols - 1), r0 = threadIdx.x / kRedCols;   // 16 rows per sweep
  // destination column of this thread's packed column (or -1: padding), and the bias column K
  const int col = col0 + cidx;
  int d = -1;
  if (cidx < ncols && col < s.K)
    for (int i = 0; i < s.n_seg; ++i)
      if (col >= s.seg_src[i] && col < s.seg_src[i] + s.seg_len[i]) d = s.seg_dst[i] + (col - s.seg_src[i]);
  const bool is_bias = cidx < ncols && col == s.K;
  for (int r = r0; r < 128; r += 128 / kRedCols) {
    const float v = tile[cidx][r];
    if (d >= 0) {
      float* dst = (chunk == 1 && s.dW_hi) ? s.dW_hi + static_cast<int64_t>(r) * s.lddw_hi + d
                                           : s.dW + static_cast<int64_t>(chunk * 128 + r) * s.lddw + d;
      *dst += v;
    } else if (is_bias && s.db) {
      s.db[chunk * 128 + r] += v;
    }
  }
}

}  // namespace
}  // namespace upnerf

extern "C" int upnerf_wgrad_bf16(const void* dY, int64_t lddy, const void* X, int64_t ldx,
                                 float* dW, int64_t lddw, float* db, int64_t M, int N, int K,
                                 int n_seg, const int* seg_src_host, const int* seg_len_host,
                                 const int* seg_dst_host, void* stream) {
  return upnerf::wgrad_launch(dY, lddy, X, ldx, dW, lddw, nullptr, 0, db, M, N, K, n_seg, seg_src_host,
                              seg_len_host, seg_dst_host, nullptr, stream);
}
extern "C" int upnerf_wgrad2_bf16(const void* dY, int64_t lddy, const void* X, int64_t ldx,
                                  float* dW_lo, int64_t lddw_lo, float* dW_hi, int64_t lddw_hi,
                                  int64_t M, int K, int n_seg, const int* seg_src_host,
                                  const int* seg_len_host, const int* seg_dst_host, void* stream) {
  UPNERF_REQUIRE(dW_lo && dW_hi, UPNERF_ERR_BAD_SHAPE, "wgrad2_bf16: both destinations are required");
  return upnerf::wgrad_launch(dY, lddy, X, ldx, dW_lo, lddw_lo, dW_hi, lddw_hi, nullptr, M, 256, K, n_seg,
                              seg_src_host, seg_len_host, seg_dst_host, nullptr, stream);
}
namespace upnerf {
namespace {

int fill_args(WgradArgs* args, const WgradProblem& q) {
  UPNERF_REQUIRE(q.M > 0, UPNERF_ERR_BAD_SHAPE, "wgrad_bf16: M=%lld", (long long)q.M);
  UPNERF_REQUIRE(q.N >= 128 && q.N <= 256 && q.N % 128 == 0, UPNERF_ERR_BAD_SHAPE,
                 "wgrad_bf16: N=%d must be 128 or 256", q.N);
  UPNERF_REQUIRE(q.K >= 64 && q.K <= kMaxK && q.K % 64 == 0, UPNERF_ERR_BAD_SHAPE,
                 "wgrad_bf16: K=%d must be a multiple of 64 in [64,320]", q.K);
  UPNERF_REQUIRE(q.n_seg >= 1 && q.n_seg <= kMaxSeg, UPNERF_ERR_BAD_SHAPE, "wgrad_bf16: n_seg=%d", q.n_seg);
  memset(args, 0, sizeof(*args));
  args->dW = q.dW;
  args->lddw = q.lddw;
  args->dW_hi = q.dW_hi;
  args->lddw_hi = q.lddw_hi;
  args->db = q.db;
  args->M = q.M;
  args->N = q.N;
  args->K = q.K;
  args->chunks = q.N / 128;
  args->n_seg = q.n_seg;
  for (int i = 0; i < q.n_seg; ++i) {
    args->seg_src[i] = q.seg_src[i];
    args->seg_len[i] = q.seg_len[i];
    args->seg_dst[i] = q.seg_dst[i];
    UPNERF_REQUIRE(q.seg_src[i] >= 0 && q.seg_src[i] + q.seg_len[i] <= q.K, UPNERF_ERR_BAD_SHAPE,
                   "wgrad_bf16: segment %d out of range", i);
  }
  return UPNERF_OK;
}

// One launch for a group of problems.  `batch` (optional) supplies the pool of split partials and collects the
// reduce slots; without it every CTA adds its tile to dW with atomics.
int launch_group(const WgradProblem* probs, int n, WgradBatch* batch, cudaStream_t st) {
  UPNERF_REQUIRE(n >= 1 && n <= kMaxWgradSlots, UPNERF_ERR_BAD_SHAPE, "wgrad group of %d problems", n);
  static WgradGroupParams* g = nullptr;     // ~6.5 KB of kernel parameters, built in place
  if (!g) g = static_cast<WgradGroupParams*>(malloc(sizeof(WgradGroupParams)));
  UPNERF_REQUIRE(g, UPNERF_ERR_WORKSPACE, "wgrad: out of host memory");
  memset(g, 0, sizeof(*g));
  g->n = n;
  // a CTA of problem p streams (128 + K) columns of bf16 per sample row: split the grid in proportion
  double cost[kMaxWgradSlots], total = 0.0;
  int64_t steps[kMaxWgradSlots];
  int splits[kMaxWgradSlots];
  for (int i = 0; i < n; ++i) {
    UPNERF_TRY(fill_args(&g->p[i], probs[i]));
    steps[i] = ceil_div64(probs[i].M, kBS);
    cost[i] = static_cast<double>(g->p[i].chunks) * (128 + probs[i].K) * static_cast<double>(steps[i]);
    total += cost[i];
  }
  const int T = sm_count();
  int used = 0;
  for (int i = 0; i < n; ++i) {
    int64_t sp = static_cast<int64_t>(T * cost[i] / total / g->p[i].chunks);
    if (sp > steps[i]) sp = steps[i];
    if (sp < 1) sp = 1;
    splits[i] = static_cast<int>(sp);
    used += splits[i] * g->p[i].chunks;
  }
  for (;;) {       // hand the CTAs left over by the rounding to the problems with the longest per-CTA stream
    int best = -1;
    double load = 0.0;
    for (int i = 0; i < n; ++i) {
      if (splits[i] >= steps[i] || used + g->p[i].chunks > T) continue;
      const double l = cost[i] / (static_cast<double>(g->p[i].chunks) * splits[i]);
      if (l > load) { load = l; best = i; }
    }
    if (best < 0) break;
    ++splits[best];
    used += g->p[best].chunks;
  }
  double flop = 0.0, bytes = 0.0;
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    WgradArgs& a = g->p[i];
    const WgradProblem& q = probs[i];
    a.splits = splits[i];
    a.rows_per_split = ceil_div64(steps[i], splits[i]) * kBS;
    g->block_begin[i] = blocks;
    blocks += splits[i] * a.chunks;
    if (batch && batch->pool && batch->list.n < kMaxWgradSlots) {
      const uint64_t need = static_cast<uint64_t>(splits[i]) * a.chunks * (q.K + 1) * 128;
      if (batch->used + need <= batch->pool_floats) {
        a.partial = batch->pool + batch->used;
        batch->used += need;
        WgradReduceSlot& sl = batch->list.s[batch->list.n++];
        memset(&sl, 0, sizeof(sl));
        sl.partial = a.partial;
        sl.dW = q.dW; sl.lddw = q.lddw; sl.dW_hi = q.dW_hi; sl.lddw_hi = q.lddw_hi; sl.db = q.db;
        sl.splits = splits[i]; sl.chunks = a.chunks; sl.K = q.K; sl.n_seg = q.n_seg;
        for (int j = 0; j < q.n_seg; ++j) {
          sl.seg_src[j] = q.seg_src[j]; sl.seg_len[j] = q.seg_len[j]; sl.seg_dst[j] = q.seg_dst[j];
        }
        sl.block_begin = batch->blocks;
        batch->blocks += a.chunks * ((q.K + 1 + kRedCols - 1) / kRedCols);
      }
    }
    UPNERF_TRY(make_tmap_bf16_2d(&g->tmY[i], q.dY, q.M, q.N, q.lddy, kBS, 64));
    UPNERF_TRY(make_tmap_bf16_2d(&g->tmX[i], q.X, q.M, q.K, q.ldx, kBS, 64));
    flop += 2.0 * q.M * q.N * q.K;
    bytes += 2.0 * q.M * (q.N + q.K) + 4.0 * q.N * q.K;
  }
  g->block_begin[n] = blocks;
  static bool attr_set = false;
  if (!attr_set) {
    UPNERF_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  LaunchScope scope(kCatWgradTc, st, flop, bytes);
  wgrad_tc_kernel<<<blocks, kThreads, kSmemBytes, st>>>(*g);
  UPNERF_CHECK_LAUNCH("wgrad_tc_kernel");
  return UPNERF_OK;
}

}  // namespace
}  // namespace upnerf

int upnerf::wgrad_launch(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw,
                         float* dW_hi, int64_t lddw_hi, float* db, int64_t M, int N, int K, int n_seg,
                         const int* seg_src_host, const int* seg_len_host, const int* seg_dst_host,
                         WgradBatch* batch, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n_seg >= 1 && n_seg <= kMaxSeg, UPNERF_ERR_BAD_SHAPE, "wgrad_bf16: n_seg=%d", n_seg);
  WgradProblem q;
  memset(&q, 0, sizeof(q));
  q.dY = dY; q.lddy = lddy; q.X = X; q.ldx = ldx;
  q.dW = dW; q.lddw = lddw; q.dW_hi = dW_hi; q.lddw_hi = lddw_hi; q.db = db;
  q.M = M; q.N = N; q.K = K; q.n_seg = n_seg;
  for (int i = 0; i < n_seg; ++i) {
    q.seg_src[i] = seg_src_host[i]; q.seg_len[i] = seg_len_host[i]; q.seg_dst[i] = seg_dst_host[i];
  }
  if (batch && batch->pool && batch->defer) {
    // grouped mode: the problem is queued; wgrad_reduce() launches the whole group, then the reduction
    WgradArgs probe;
    UPNERF_TRY(fill_args(&probe, q));       // validate now, where the caller can be named
    if (batch->n_pending == kMaxWgradSlots) UPNERF_TRY(wgrad_flush(batch, as_stream(stream)));
    batch->pending[batch->n_pending++] = q;
    return UPNERF_OK;
  }
  return launch_group(&q, 1, batch, as_stream(stream));
}

// Launches the queued problems of a deferring batch as one group (their partials stay parked in the pool).
int upnerf::wgrad_flush(WgradBatch* batch, cudaStream_t st) {
  if (!batch || batch->n_pending == 0) return UPNERF_OK;
  const int n = batch->n_pending;
  batch->n_pending = 0;
  return launch_group(batch->pending, n, batch, st);
}

uint64_t upnerf::wgrad_pool_floats() {
  // per slot at most sm_count CTAs x (K + 1) x 128 floats; one pass of the D=8, W=256 network runs
  // wgrads with K = 128, 9 x 256, 320 and 64 (see render.cu:pass_bwd) -- rounded up generously
  return static_cast<uint64_t>(sm_count()) * 128 * 3072;
}

int upnerf::wgrad_reduce(WgradBatch* batch, cudaStream_t st) {
  if (!batch) return UPNERF_OK;
  UPNERF_TRY(wgrad_flush(batch, st));
  if (batch->list.n == 0) return UPNERF_OK;
  {
    LaunchScope scope(kCatWgradReduce, st, 0.0, 0.0);
    wgrad_reduce_kernel<<<batch->blocks, 128, 0, st>>>(batch->list);
    UPNERF_CHECK_LAUNCH("wgrad_reduce_kernel");
  }
  batch->list.n = 0;
  batch->used = 0;
  batch->blocks = 0;
  return UPNERF_OK;
}

extern "C" int upnerf_wgrad_bf16_det(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW,
                                     int64_t lddw, float* db, int64_t M, int N, int K, int n_seg,
                                     const int* seg_src_host, const int* seg_len_host,
                                     const int* seg_dst_host, void* workspace, uint64_t workspace_bytes,
                                     void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(workspace && workspace_bytes >= upnerf_wgrad_det_workspace_bytes(N, K), UPNERF_ERR_WORKSPACE,
                 "wgrad_bf16_det: workspace too small");
  WgradBatch b;
  memset(&b, 0, sizeof(b));
  b.pool = static_cast<float*>(workspace);
  b.pool_floats = workspace_bytes / sizeof(float);
  UPNERF_TRY(wgrad_launch(dY, lddy, X, ldx, dW, lddw, nullptr, 0, db, M, N, K, n_seg, seg_src_host, seg_len_host,
                          seg_dst_host, &b, stream));
  return wgrad_reduce(&b, as_stream(stream));
}

extern "C" uint64_t upnerf_wgrad_det_workspace_bytes(int N, int K) {
  (void)N;
  return static_cast<uint64_t>(upnerf::sm_count()) * 128 * (K + 1) * sizeof(float);
}
