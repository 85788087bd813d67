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

constexpr int kStages = 3;
constexpr int kBS = 64;                 // samples per stage
constexpr int kBoxBytes = kBS * 128;    // one 64-sample x 64-column bf16 box
constexpr int kABoxes = 2;              // 128 output rows
constexpr int kMaxBBoxes = 5;           // K <= 320
constexpr int kStageBytes = (kABoxes + kMaxBBoxes) * kBoxBytes;
constexpr int kOnesBytes = 1024;
constexpr int kMaxK = 320;
constexpr int kBiasCol = 384;           // TMEM column of the bias-gradient accumulator
constexpr int kThreads = 192;

constexpr int kOffStage = 0;
constexpr int kOffOnes = kOffStage + kStages * kStageBytes;
constexpr int kOffMap = kOffOnes + kOnesBytes;
constexpr int kOffBar = kOffMap + kMaxK * 4;
constexpr int kNumBars = 2 * kStages + 1;
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
  int64_t rows_per_split;  // multiple of kBS
  int n_seg;
  int seg_src[kMaxSeg], seg_len[kMaxSeg], seg_dst[kMaxSeg];
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ WgradArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sStage = smem + kOffStage;
  uint8_t* sOnes = smem + kOffOnes;
  int* sMap = reinterpret_cast<int*>(smem + kOffMap);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + kStages;
  uint64_t* bar_done = bars + 2 * kStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x;  // which 128 output rows
  const int split = blockIdx.y;
  const int K = args.K;
  const int bboxes = K / 64;

  const int64_t row_begin = static_cast<int64_t>(split) * args.rows_per_split;
  int64_t row_end = row_begin + args.rows_per_split;
  if (row_end > args.M) row_end = args.M;
  const int nsteps = row_end > row_begin ? static_cast<int>((row_end - row_begin + kBS - 1) / kBS) : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmY);
    prefetch_tmap(&tmX);
    for (int i = 0; i < kStages; ++i) {
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
            tma_load_2d(base + j * kBoxBytes, &tmY, &bar_full[stage], chunk * 128 + j * 64, r0);
          for (int j = 0; j < bboxes; ++j)
            tma_load_2d(base + (kABoxes + j) * kBoxBytes, &tmX, &bar_full[stage], j * 64, r0);
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
                     (static_cast<int64_t>(split) * gridDim.x + chunk) * (K + 1) * 128 + quad * 32 + lane;
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
                 (static_cast<int64_t>(split) * gridDim.x + chunk) * (K + 1) * 128 + (threadIdx.x - 64);
    for (int c = 0; c <= K; ++c) dst[c * 128] = 0.f;
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// One block per (slot, 128-row chunk, packed column): thread = output row.  Sums the splits in a
// fixed order (bit-reproducible) and adds the result to the parameter gradient.
__global__ void __launch_bounds__(128)
wgrad_reduce_kernel(const __grid_constant__ WgradReduceList list) {
  int si = 0;
  while (si + 1 < list.n && static_cast<int>(blockIdx.x) >= list.s[si + 1].block_begin) ++si;
  const WgradReduceSlot& s = list.s[si];
  const int local = blockIdx.x - s.block_begin;
  const int chunk = local / (s.K + 1);
  const int col = local - chunk * (s.K + 1);
  const int row = threadIdx.x;
  float* dst = nullptr;
  if (col == s.K) {
    if (s.db) dst = s.db + chunk * 128 + row;
  } else {
    int d = -1;
    for (int i = 0; i < s.n_seg; ++i)
      if (col >= s.seg_src[i] && col < s.seg_src[i] + s.seg_len[i]) d = s.seg_dst[i] + (col - s.seg_src[i]);
    if (d >= 0)
      dst = (chunk == 1 && s.dW_hi) ? s.dW_hi + static_cast<int64_t>(row) * s.lddw_hi + d
                                    : s.dW + static_cast<int64_t>(chunk * 128 + row) * s.lddw + d;
  }
  if (!dst) return;
  const int64_t stride = static_cast<int64_t>(s.chunks) * (s.K + 1) * 128;
  const float* p = s.partial + (static_cast<int64_t>(chunk) * (s.K + 1) + col) * 128 + row;
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
  *dst += (acc[0] + acc[1]) + (acc[2] + acc[3]);
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
int upnerf::wgrad_launch(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw,
                         float* dW_hi, int64_t lddw_hi, float* db, int64_t M, int N, int K, int n_seg,
                         const int* seg_src_host, const int* seg_len_host, const int* seg_dst_host,
                         WgradBatch* batch, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(M > 0, UPNERF_ERR_BAD_SHAPE, "wgrad_bf16: M=%lld", (long long)M);
  UPNERF_REQUIRE(N >= 128 && N <= 256 && N % 128 == 0, UPNERF_ERR_BAD_SHAPE,
                 "wgrad_bf16: N=%d must be 128 or 256", N);
  UPNERF_REQUIRE(K >= 64 && K <= kMaxK && K % 64 == 0, UPNERF_ERR_BAD_SHAPE,
                 "wgrad_bf16: K=%d must be a multiple of 64 in [64,320]", K);
  UPNERF_REQUIRE(n_seg >= 1 && n_seg <= kMaxSeg, UPNERF_ERR_BAD_SHAPE, "wgrad_bf16: n_seg=%d",
                 n_seg);
  WgradArgs args;
  memset(&args, 0, sizeof(args));
  args.dW = dW;
  args.lddw = lddw;
  args.dW_hi = dW_hi;
  args.lddw_hi = lddw_hi;
  args.db = db;
  args.M = M;
  args.N = N;
  args.K = K;
  args.n_seg = n_seg;
  for (int i = 0; i < n_seg; ++i) {
    args.seg_src[i] = seg_src_host[i];
    args.seg_len[i] = seg_len_host[i];
    args.seg_dst[i] = seg_dst_host[i];
    UPNERF_REQUIRE(seg_src_host[i] >= 0 && seg_src_host[i] + seg_len_host[i] <= K,
                   UPNERF_ERR_BAD_SHAPE, "wgrad_bf16: segment %d out of range", i);
  }
  const int chunks = N / 128;
  int splits = sm_count() / chunks;
  const int64_t steps = ceil_div64(M, kBS);
  if (splits > steps) splits = static_cast<int>(steps);
  if (splits < 1) splits = 1;
  args.rows_per_split = ceil_div64(steps, splits) * kBS;
  args.splits = splits;
  if (batch && batch->pool && batch->list.n < kMaxWgradSlots) {
    const uint64_t need = static_cast<uint64_t>(splits) * chunks * (K + 1) * 128;
    if (batch->used + need <= batch->pool_floats) {
      args.partial = batch->pool + batch->used;
      batch->used += need;
      WgradReduceSlot& sl = batch->list.s[batch->list.n++];
      memset(&sl, 0, sizeof(sl));
      sl.partial = args.partial;
      sl.dW = dW; sl.lddw = lddw; sl.dW_hi = dW_hi; sl.lddw_hi = lddw_hi; sl.db = db;
      sl.splits = splits; sl.chunks = chunks; sl.K = K; sl.n_seg = n_seg;
      for (int i = 0; i < n_seg; ++i) {
        sl.seg_src[i] = seg_src_host[i]; sl.seg_len[i] = seg_len_host[i]; sl.seg_dst[i] = seg_dst_host[i];
      }
      sl.block_begin = batch->blocks;
      batch->blocks += chunks * (K + 1);
    }
  }

  CUtensorMap tmY, tmX;
  UPNERF_TRY(make_tmap_bf16_2d(&tmY, dY, M, N, lddy, kBS, 64));
  UPNERF_TRY(make_tmap_bf16_2d(&tmX, X, M, K, ldx, kBS, 64));
  static bool attr_set = false;
  if (!attr_set) {
    UPNERF_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmemBytes));
    attr_set = true;
  }
  dim3 grid(chunks, splits);
  LaunchScope scope(kCatWgradTc, as_stream(stream), 2.0 * M * N * K, 2.0 * M * (N + K) + 4.0 * N * K);
  wgrad_tc_kernel<<<grid, kThreads, kSmemBytes, as_stream(stream)>>>(tmY, tmX, args);
  UPNERF_CHECK_LAUNCH("wgrad_tc_kernel");
  return UPNERF_OK;
}

uint64_t upnerf::wgrad_pool_floats() {
  // per slot at most sm_count CTAs x (K + 1) x 128 floats; one pass of the D=8, W=256 network runs
  // wgrads with K = 128, 9 x 256, 320 and 64 (see render.cu:pass_bwd) -- rounded up generously
  return static_cast<uint64_t>(sm_count()) * 128 * 3072;
}

int upnerf::wgrad_reduce(WgradBatch* batch, cudaStream_t st) {
  if (!batch || batch->list.n == 0) return UPNERF_OK;
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
