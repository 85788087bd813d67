// fp32 SIMT GEMM with arbitrary element strides -- upnerf_gemm_f32.
//
// Two jobs: (1) the fp32 "validation mode" of the MLP (reference arithmetic is fp32,
// models/nerf.py:84-123), where it stands in for every tcgen05 GEMM so that outputs can be
// compared with the oracle at 1e-4; (2) the small per-ray products of the production path
// (per-ray embedding biases, the 384-d feature projection applied AFTER compositing).
// Generic strides let the same kernel serve forward (A.W^T), data gradient (dY.W) and
// weight gradient (dY^T.X, split over the sample axis with atomics) without transposes.
#include <string.h>

#include "common.h"

namespace upnerf {
namespace {

constexpr int TM = 64, TN = 64, TK = 32;

struct SimtArgs {
  const float* A;
  int64_t sam, sak;
  const float* B;
  int64_t sbn, sbk;
  float* C;
  int64_t scm, scn;
  int64_t M, N, K;
  int accumulate;
  int split_k;
  int atomic;  // accumulate with atomicAdd (split-k requested), epilogue skipped
  int64_t k_per_split;
  upnerf_epilogue ep;
};

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const SimtArgs a) {
  // double-buffered tiles: the global loads of tile i+1 are in flight while tile i is consumed
  __shared__ float As[2][TK][TM + 4];
  __shared__ float Bs[2][TK][TN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * TM;
  const int64_t n0 = static_cast<int64_t>(blockIdx.y) * TN;
  const int64_t kbeg = static_cast<int64_t>(blockIdx.z) * a.k_per_split;
  int64_t kend = kbeg + a.k_per_split;
  if (kend > a.K) kend = a.K;

  const int tx = tid & 15;  // n direction
  const int ty = tid >> 4;  // m direction
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kfast = (a.sak == 1);
  const bool b_kfast = (a.sbk == 1);
  constexpr int kPerThread = TM * TK / 256;
  float ra[kPerThread], rb[kPerThread];

  auto load_tile = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int e = tid + i * 256;
      int kk, mm;
      if (a_kfast) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      const int64_t gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < a.M && gk < kend) ? __ldg(a.A + gm * a.sam + gk * a.sak) : 0.f;
      int kb, nn;
      if (b_kfast) { kb = e % TK; nn = e / TK; } else { nn = e % TN; kb = e / TN; }
      const int64_t gn = n0 + nn, gk2 = k0 + kb;
      rb[i] = (gn < a.N && gk2 < kend) ? __ldg(a.B + gn * a.sbn + gk2 * a.sbk) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int e = tid + i * 256;
      int kk, mm;
      if (a_kfast) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      As[buf][kk][mm] = ra[i];
      int kb, nn;
      if (b_kfast) { kb = e % TK; nn = e / TK; } else { nn = e % TN; kb = e / TN; }
      Bs[buf][kb][nn] = rb[i];
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tile(kbeg);
    store_tile(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
    const bool more = k0 + TK < kend;
    if (more) load_tile(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[buf][kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[buf][kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) store_tile(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  const upnerf_epilogue& ep = a.ep;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j];
      float* c = a.C + m * a.scm + n * a.scn;
      if (a.atomic) {
        atomicAdd(c, v);
        continue;
      }
      if (ep.bias) v += ep.bias[n];
      if (ep.ray_bias) v += ep.ray_bias[(m / ep.rows_per_ray) * a.N + n];
      if (ep.rank1_row) v += ep.rank1_row[m] * ep.rank1_col[n];
      float aux = 0.f;
      if (ep.aux_mode != 0) aux = static_cast<const float*>(ep.aux)[m * ep.ldaux + n];
      if (ep.aux_mode == 1) v += aux;
      if (ep.act == 1) v = fmaxf(v, 0.f);
      if (ep.aux_mode == 2) v = aux > 0.f ? v : 0.f;
      if (a.accumulate) v += *c;
      *c = v;
    }
  }
}

}  // namespace
}  // namespace upnerf

extern "C" int upnerf_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B,
                               int64_t sbn, int64_t sbk, float* C, int64_t scm, int64_t scn,
                               int64_t M, int64_t N, int64_t K, const upnerf_epilogue* ep,
                               int accumulate, int split_k, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(M > 0 && N > 0 && K > 0, UPNERF_ERR_BAD_SHAPE, "gemm_f32: M=%lld N=%lld K=%lld",
                 (long long)M, (long long)N, (long long)K);
  SimtArgs a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.sam = sam; a.sak = sak;
  a.B = B; a.sbn = sbn; a.sbk = sbk;
  a.C = C; a.scm = scm; a.scn = scn;
  a.M = M; a.N = N; a.K = K;
  a.accumulate = accumulate;
  if (split_k < 1) split_k = 1;
  a.atomic = split_k > 1;
  // k ranges are multiples of TK so tiles of different splits never overlap
  int64_t kps = ceil_div64(ceil_div64(K, split_k), TK) * TK;
  split_k = static_cast<int>(ceil_div64(K, kps));
  a.split_k = split_k;
  a.k_per_split = kps;
  if (ep) a.ep = *ep;
  UPNERF_REQUIRE(a.ep.n_heads == 0, UPNERF_ERR_BAD_CONFIG, "gemm_f32: heads are not supported");
  UPNERF_REQUIRE(!(a.ep.aux_mode != 0) || a.ep.aux != nullptr, UPNERF_ERR_BAD_SHAPE,
                 "gemm_f32: aux_mode set without aux");
  const int64_t mt = ceil_div64(M, TM), nt = ceil_div64(N, TN);
  UPNERF_REQUIRE(mt < (1ll << 31) && nt < 65536 && split_k < 65536, UPNERF_ERR_BAD_SHAPE,
                 "gemm_f32: grid too large");
  dim3 grid(static_cast<unsigned>(mt), static_cast<unsigned>(nt), static_cast<unsigned>(split_k));
  LaunchScope scope(kCatGemmSimt, as_stream(stream), 2.0 * M * N * K);
  gemm_simt_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
  UPNERF_CHECK_LAUNCH("gemm_simt_kernel");
  return UPNERF_OK;
}
