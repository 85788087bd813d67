// Small fp32-in / fp32-out products on tcgen05 (kind::tf32) -- upnerf_gemm_tf32.
//
//   C[m,n] (+)= sum_k A[m,k] B[n,k]  (+ bias[n] + rank1_row[m] rank1_col[n])     arbitrary element strides
//
// The production (bf16) path has a few dozen products per step whose operands are PER-RAY or
// parameter-space fp32 tensors, not per-sample activations: the per-ray bias of the head layers
// (W[:, cols] . [PE(dir) | a_emb], models/nerf.py:97-113 with the embeddings of
// models/rendering.py:255-258), the 384-d feature projections applied after compositing
// (feat_share_layer / feat_candidate_layer, models/nerf.py:53,76) with their data and weight
// gradients, and the chain rule through the folded matrix W_rgb0[:, :F] W_sf.  They used to run on the
// strided fp32 FMA kernel (gemm_simt.cu: 36 launches, ~1.1 ms of SM time per step beside the
// tensor-core chain); here they run on the tensor cores with tf32 operands (10-bit mantissa, rounded
// to nearest when the tile is staged -- finer than the bf16 activations around them) and fp32
// accumulation.  The fp32 validation mode keeps the FMA kernel.
//
// One CTA = one 128 x 128 output tile over one split of K.  The operands have arbitrary strides
// (transposed views, column slices of parameter matrices), so producer warps stage them themselves:
// global -> registers (coalesced along whichever axis is contiguous) -> shared memory in the canonical
// K-major 128-byte-swizzle layout (32 tf32 per row); four producer groups each own one stage and every
// fourth K-chunk; a 17th warp issues 4 x tcgen05.mma 128x128x8 per chunk, in order, into a 128-column
// TMEM accumulator.  Epilogue: TMEM -> registers -> (per-warp shared-memory transpose when
// the output is row-major) -> plain or atomic (split-K) stores.
#include <string.h>

#include "common.h"
#include "ptx_sm100.cuh"

namespace upnerf {
namespace {

using namespace ptx;

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int kGroups = 3;                        // producer groups = shared-memory stages
constexpr int kGroupThreads = 256;
constexpr int kLoadThreads = kGroups * kGroupThreads;
constexpr int kThreads = kLoadThreads + 32;       // + the MMA-issuing warp
constexpr int kTileBytes = BM * BK * 4;           // 16 KB: 128 rows x 128 bytes
constexpr int kStageBytes = 2 * kTileBytes;       // A tile | B tile
constexpr int kOffBar = kGroups * kStageBytes;
constexpr int kOffTmem = kOffBar + (2 * kGroups + 1) * 8;
constexpr int kSmemBytes = kOffTmem + 16 + 1024;  // + alignment slack
constexpr int kPer = BM * BK / kGroupThreads;     // 16 elements per thread and operand

struct Tf32Args {
  const float* A;
  int64_t sam, sak;
  const float* B;
  int64_t sbn, sbk;
  float* C;
  int64_t scm, scn;
  int64_t M, N, K;
  int accumulate;
  int atomic;
  int64_t k_per_split;
  const float* bias;
  const float* rank1_row;
  const float* rank1_col;
};

// Instruction descriptor for kind::tf32: D fp32, A/B tf32 (format 2), both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// A group's 256 threads move one 128 x 32 operand tile, 16 elements per thread.  Element i of thread t is
//   KFAST (operand contiguous along k):  k = t & 31, row = (t >> 5) + 8 i   -- a warp reads 128 contiguous bytes
//   else  (contiguous along the row):    row = t & 127, k = (t >> 7) + 2 i  -- a warp reads 32 consecutive rows
// so global reads coalesce either way, every address is one constant stride after the previous one, and the
// shared-memory offsets (canonical K-major SWIZZLE_128B: row r at r*128 bytes, 16-byte chunk j at j ^ (r & 7))
// are compile-time functions of i apart from one per-thread term.
template <bool KFAST>
__device__ __forceinline__ void load_tile(const float* __restrict__ base, int64_t srow, int64_t sk, int64_t rows_left,
                                          int64_t k_left, int t, float (&r)[kPer]) {
  if (KFAST) {
    const int k = t & 31, row0 = t >> 5;
    const float* p = base + row0 * srow + k;
    const bool kin = k < k_left;
    const int64_t step = 8 * srow;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      r[i] = (kin && row0 + 8 * i < rows_left) ? __ldg(p) : 0.f;
      p += step;
    }
  } else {
    const int row = t & 127, kb = t >> 7;
    const float* p = base + row * srow + kb * sk;
    const bool rin = row < rows_left;
    const int64_t step = 2 * sk;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      r[i] = (rin && kb + 2 * i < k_left) ? __ldg(p) : 0.f;
      p += step;
    }
  }
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <bool KFAST>
__device__ __forceinline__ void store_tile(uint32_t tile, int t, const float (&r)[kPer]) {   // tile: shared address
  if (KFAST) {
    const int k = t & 31, row0 = t >> 5;       // row & 7 == row0 for every i
    const uint32_t p = tile + row0 * 128 + ((((k >> 2) ^ row0) << 4) | ((k & 3) << 2));
#pragma unroll
    for (int i = 0; i < kPer; ++i) sts32(p + i * 1024, to_tf32(r[i]));
  } else {
    const int row = t & 127, kb = t >> 7, r7 = row & 7;
    const uint32_t p = tile + row * 128 + (kb << 2);
#pragma unroll
    for (int i = 0; i < kPer; ++i)     // k = kb + 2 i:  k >> 2 = i >> 1,  k & 3 = kb + 2 (i & 1)
      sts32(p + ((((i >> 1) ^ r7) << 4) | ((i & 1) << 3)), to_tf32(r[i]));
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ Tf32Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kOffBar);   // [kGroups]: the group staged its chunk
  uint64_t* bar_free = bar_full + kGroups;                            // [kGroups]: the MMAs reading it retired
  uint64_t* bar_done = bar_free + kGroups;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kOffTmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM;
  const int64_t n0 = static_cast<int64_t>(blockIdx.y) * BN;
  const int64_t kbeg = static_cast<int64_t>(blockIdx.z) * a.k_per_split;
  int64_t kend = kbeg + a.k_per_split;
  if (kend > a.K) kend = a.K;
  const int nchunks = kend > kbeg ? static_cast<int>((kend - kbeg + BK - 1) / BK) : 0;

  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) {
      mbar_init(&bar_full[i], kGroupThreads);
      mbar_init(&bar_free[i], 1);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == kLoadThreads / 32) tmem_alloc<128>(tmem_holder);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  // These products are LATENCY-bound (a few dozen CTAs, 1..12 K-chunks each).  Four producer groups of four
  // warps each own one shared-memory stage and every fourth K-chunk: a group issues all 64 loads of its chunk
  // per thread at once, stages them, fences and arrives -- while it waits for its loads the other three groups
  // are at other points of the same cycle.  (A proxy fence drains the issuing thread's outstanding loads, so
  // prefetching ACROSS a fence inside one thread does not overlap anything; splitting the chunks over groups does.)
  if (tid < kLoadThreads) {
    const int g = tid / kGroupThreads, t = tid % kGroupThreads;
    const uint32_t sA = smem_u32(smem + g * kStageBytes);
    const uint32_t sB = sA + kTileBytes;
    const bool a_kfast = (a.sak == 1), b_kfast = (a.sbk == 1);
    for (int c = g; c < nchunks; c += kGroups) {
      const int64_t k0 = kbeg + static_cast<int64_t>(c) * BK;
      float ra[kPer], rb[kPer];
      const float* pa = a.A + m0 * a.sam + k0 * a.sak;
      const float* pb = a.B + n0 * a.sbn + k0 * a.sbk;
      if (a_kfast) load_tile<true>(pa, a.sam, a.sak, a.M - m0, kend - k0, t, ra);
      else load_tile<false>(pa, a.sam, a.sak, a.M - m0, kend - k0, t, ra);
      if (b_kfast) load_tile<true>(pb, a.sbn, a.sbk, a.N - n0, kend - k0, t, rb);
      else load_tile<false>(pb, a.sbn, a.sbk, a.N - n0, kend - k0, t, rb);
      if (c >= kGroups) mbar_wait(&bar_free[g], ((c / kGroups) - 1) & 1);   // the MMAs that read this stage retired
      if (a_kfast) store_tile<true>(sA, t, ra); else store_tile<false>(sA, t, ra);
      if (b_kfast) store_tile<true>(sB, t, rb); else store_tile<false>(sB, t, rb);
      fence_proxy_async_smem();        // generic-proxy stores -> visible to the tensor core's async proxy
      mbar_arrive(&bar_full[g]);
    }
  } else if (lane == 0) {
    // ---- MMA issuer: chunks in order (the first one overwrites the accumulator)
    const uint32_t idesc = umma_idesc_tf32(BM, BN);
    for (int c = 0; c < nchunks; ++c) {
      const int g = c % kGroups;
      mbar_wait(&bar_full[g], (c / kGroups) & 1);
      tc_fence_after_sync();
      const uint32_t a_addr = smem_u32(smem + g * kStageBytes);
      const uint32_t b_addr = a_addr + kTileBytes;
#pragma unroll
      for (int k = 0; k < BK / 8; ++k)
        mma_tf32_ss(tmem_base, umma_desc(a_addr + k * 32, 16, 1024, kLayoutSw128),
                    umma_desc(b_addr + k * 32, 16, 1024, kLayoutSw128), idesc, (c | k) != 0);
      mma_commit(&bar_free[g]);
    }
    mma_commit(bar_done);
  }

  // ---- epilogue: warp w owns TMEM lanes (w & 3) * 32 .. +31 (hardware rule) and columns (w >> 2) * 32 .. +31
  if (nchunks > 0 && warp < 16) {
    mbar_wait(bar_done, 0);
    tc_fence_after_sync();
    const int quad = warp & 3, cbase = (warp >> 2) * 32;
    const int64_t m = m0 + quad * 32 + lane;
    // the operand stages are idle now: 32 x 33 floats per warp for the transpose
    float* sT = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    const bool row_major = a.scn == 1 && a.scm != 1;
    if (n0 + cbase < a.N && m0 + quad * 32 < a.M) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + cbase, v);
      tmem_ld_wait();
      const bool atomic = a.atomic != 0, acc = a.accumulate != 0;
      if (row_major) {
        // lanes across columns: transpose the warp's 32 x 32 block through shared memory
#pragma unroll
        for (int i = 0; i < 32; ++i) sT[lane * 33 + i] = __uint_as_float(v[i]);
        __syncwarp();
        const int64_t nn = n0 + cbase + lane;
        const bool nok = nn < a.N;
        const float bn = (nok && a.bias) ? __ldg(a.bias + nn) : 0.f;
        const float cn = (nok && a.rank1_row) ? __ldg(a.rank1_col + nn) : 0.f;
        const int64_t mrow = m0 + quad * 32;
        const int rows = a.M - mrow < 32 ? static_cast<int>(a.M - mrow) : 32;
        float* cptr = a.C + mrow * a.scm + nn;
        if (nok) {
#pragma unroll 8
          for (int r = 0; r < rows; ++r, cptr += a.scm) {
            float x = sT[r * 33 + lane];
            if (atomic) { atomicAdd(cptr, x); continue; }
            x += bn;
            if (a.rank1_row) x += __ldg(a.rank1_row + mrow + r) * cn;
            if (acc) x += *cptr;
            *cptr = x;
          }
        }
      } else if (m < a.M) {
        // lanes across rows (column-major or strided output): stores straight from the TMEM registers
        const float rm = a.rank1_row ? __ldg(a.rank1_row + m) : 0.f;
        float* cptr = a.C + m * a.scm + (n0 + cbase) * a.scn;
        const int cols = a.N - (n0 + cbase) < 32 ? static_cast<int>(a.N - (n0 + cbase)) : 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < cols) {
            float x = __uint_as_float(v[i]);
            float* cp = cptr + i * a.scn;
            if (atomic) {
              atomicAdd(cp, x);
            } else {
              if (a.bias) x += __ldg(a.bias + n0 + cbase + i);
              if (a.rank1_row) x += rm * __ldg(a.rank1_col + n0 + cbase + i);
              if (acc) x += *cp;
              *cp = x;
            }
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kLoadThreads / 32) {
    tc_fence_after_sync();
    tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace
}  // namespace upnerf

extern "C" int upnerf_gemm_tf32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk,
                                float* C, int64_t scm, int64_t scn, int64_t M, int64_t N, int64_t K,
                                const upnerf_epilogue* ep, int accumulate, int split_k, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(upnerf_device_ok(), UPNERF_ERR_CUDA, "gemm_tf32 needs an sm_100a GPU (no fallback)");
  UPNERF_REQUIRE(M > 0 && N > 0 && K > 0, UPNERF_ERR_BAD_SHAPE, "gemm_tf32: M=%lld N=%lld K=%lld", (long long)M,
                 (long long)N, (long long)K);
  UPNERF_REQUIRE(A && B && C, UPNERF_ERR_BAD_SHAPE, "gemm_tf32: missing operand");
  Tf32Args a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.sam = sam; a.sak = sak;
  a.B = B; a.sbn = sbn; a.sbk = sbk;
  a.C = C; a.scm = scm; a.scn = scn;
  a.M = M; a.N = N; a.K = K;
  a.accumulate = accumulate;
  if (split_k < 1) split_k = 1;
  a.atomic = split_k > 1;
  // k ranges are multiples of BK so tiles of different splits never overlap
  const int64_t kps = ceil_div64(ceil_div64(K, split_k), BK) * BK;
  split_k = static_cast<int>(ceil_div64(K, kps));
  a.k_per_split = kps;
  if (ep) {
    UPNERF_REQUIRE(ep->n_heads == 0 && ep->aux_mode == 0 && ep->act == 0 && !ep->ray_bias, UPNERF_ERR_BAD_CONFIG,
                   "gemm_tf32: only bias and rank-1 epilogues are supported");
    UPNERF_REQUIRE(!a.atomic || (!ep->bias && !ep->rank1_row), UPNERF_ERR_BAD_CONFIG,
                   "gemm_tf32: split-K accumulates with atomics and takes no epilogue");
    UPNERF_REQUIRE(!ep->rank1_row || ep->rank1_col, UPNERF_ERR_BAD_SHAPE, "gemm_tf32: rank1_col missing");
    a.bias = ep->bias;
    a.rank1_row = ep->rank1_row;
    a.rank1_col = ep->rank1_col;
  }
  const int64_t mt = ceil_div64(M, BM), nt = ceil_div64(N, BN);
  UPNERF_REQUIRE(mt < (1ll << 31) && nt < 65536 && split_k < 65536, UPNERF_ERR_BAD_SHAPE, "gemm_tf32: grid too large");
  static bool attr_set = false;
  if (!attr_set) {
    UPNERF_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  dim3 grid(static_cast<unsigned>(mt), static_cast<unsigned>(nt), static_cast<unsigned>(split_k));
  LaunchScope scope(kCatGemmTf32, as_stream(stream), 2.0 * M * N * K);
  gemm_tf32_kernel<<<grid, kThreads, kSmemBytes, as_stream(stream)>>>(a);
  UPNERF_CHECK_LAUNCH("gemm_tf32_kernel");
  return UPNERF_OK;
}
