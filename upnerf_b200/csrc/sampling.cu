// Depth sampling along rays (subsystem (d)): stratified coarse samples, inverse-CDF
// hierarchical resampling and the sort-merge that builds the fine sample set.
//
//   upnerf_stratified_z      models/rendering.py:231-249
//   upnerf_sample_pdf        models/rendering.py:7-50   (searchsorted(right=True) contract)
//   upnerf_searchsorted_right  torch.searchsorted(cdf, u, right=True) on caller-provided CDFs
//   upnerf_resample_merge    models/rendering.py:262-307 (1 or 2 sample_pdf draws + sort)
//
// One warp per ray.  The CDF is accumulated sequentially in fp32 (like torch.cumsum on the
// CPU) by lane 0 and shared through shared memory; every lane then binary-searches its own
// uniforms.  Bin indices are bit-exact for identical CDFs and uniforms.
#include "common.h"

namespace upnerf {
namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxBins = 256;   // S - 1 <= 255
constexpr int kMaxFine = 512;   // S + N_importance <= 512

__global__ void stratified_z_kernel(const float* __restrict__ rays, const float* __restrict__ prand,
                                    float perturb, int use_disp, int64_t R, int S,
                                    float* __restrict__ z) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= R * S) return;
  const int64_t r = i / S;
  const int j = static_cast<int>(i - r * S);
  const float near = rays[r * 8 + 6], far = rays[r * 8 + 7];
  // torch.linspace(0, 1, S): step = 1/(S-1); the upper half is computed from the end
  auto lin = [S](int k) -> float {
    if (S == 1) return 0.f;
    const float step = 1.f / static_cast<float>(S - 1);
    return k < S / 2 ? __fmul_rn(step, k) : 1.f - __fmul_rn(step, S - 1 - k);
  };
  auto zval = [&](int k) -> float {
    const float s = lin(k);
    return use_disp ? 1.f / __fadd_rn(__fmul_rn(1.f / near, 1.f - s), __fmul_rn(1.f / far, s))
                    : __fadd_rn(__fmul_rn(near, 1.f - s), __fmul_rn(far, s));
  };
  float v = zval(j);
  if (perturb > 0.f) {
    const float lo = j == 0 ? v : 0.5f * (zval(j - 1) + v);
    const float hi = j == S - 1 ? v : 0.5f * (v + zval(j + 1));
    v = __fadd_rn(lo, __fmul_rn(hi - lo, __fmul_rn(perturb, prand[i])));
  }
  z[i] = v;
}

// upper_bound: first index i in [0, n] with cdf[i] > u   (searchsorted right=True)
__device__ __forceinline__ int upper_bound(const float* cdf, int n, float u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Builds cdf[0..nw] (nw+1 entries, cdf[0] = 0) from weights w[0..nw) in shared memory.
__device__ __forceinline__ void build_cdf(const float* __restrict__ w, int nw, float eps,
                                          float* cdf, int lane) {
  // sum of (w + eps): warp tree reduction (the reference's reduction order is unspecified)
  float part = 0.f;
  for (int i = lane; i < nw; i += 32) part += w[i] + eps;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  for (int i = lane; i < nw; i += 32) cdf[i + 1] = (w[i] + eps) / part;   // pdf
  __syncwarp();
  if (lane == 0) {
    cdf[0] = 0.f;
    float run = 0.f;
    for (int i = 1; i <= nw; ++i) {
      run += cdf[i];
      cdf[i] = run;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int nw, float u,
                                            float eps, int* ind_out) {
  const int ind = upper_bound(cdf, nw + 1, u);
  if (ind_out) *ind_out = ind;
  const int below = ind - 1 < 0 ? 0 : ind - 1;
  const int above = ind > nw ? nw : ind;
  const float c0 = cdf[below], c1 = cdf[above];
  float denom = c1 - c0;
  if (denom < eps) denom = 1.f;
  const float b0 = bins[below], b1 = bins[above];
  // explicit rounding steps (no FMA contraction) so results match torch's fp32 ops bit for bit
  return __fadd_rn(b0, __fmul_rn(__fdiv_rn(u - c0, denom), b1 - b0));
}

__device__ __forceinline__ float det_u(int j, int n) {
  if (n == 1) return 0.f;
  const float step = 1.f / static_cast<float>(n - 1);
  return j < n / 2 ? __fmul_rn(step, j) : 1.f - __fmul_rn(step, n - 1 - j);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_pdf_kernel(const float* __restrict__ bins, int64_t ld_bins, const float* __restrict__ weights,
                  int64_t ld_w, const float* __restrict__ u, int64_t R, int nw, int N, float eps,
                  float* __restrict__ samples, int64_t* __restrict__ inds, float* __restrict__ cdf_out) {
  __shared__ float s_cdf[kWarpsPerBlock][kMaxBins + 1];
  __shared__ float s_bins[kWarpsPerBlock][kMaxBins + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(kWarpsPerBlock) + warp;
  if (r >= R) return;
  float* cdf = s_cdf[warp];
  float* sb = s_bins[warp];
  for (int i = lane; i <= nw; i += 32) sb[i] = bins[r * ld_bins + i];
  build_cdf(weights + r * ld_w, nw, eps, cdf, lane);
  if (cdf_out)
    for (int i = lane; i <= nw; i += 32) cdf_out[r * (nw + 1) + i] = cdf[i];
  for (int j = lane; j < N; j += 32) {
    const float uu = u ? u[r * N + j] : det_u(j, N);
    int ind;
    samples[r * N + j] = invert_cdf(cdf, sb, nw, uu, eps, &ind);
    if (inds) inds[r * N + j] = ind;
  }
}

__global__ void searchsorted_right_kernel(const float* __restrict__ cdf, int ncdf,
                                          const float* __restrict__ u, int N, int64_t R,
                                          int64_t* __restrict__ inds) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= R * N) return;
  const int64_t r = i / N;
  inds[i] = upper_bound(cdf + r * ncdf, ncdf, u[i]);
}

// In-shared-memory bitonic sort of n (power of two) floats by one warp.
__device__ __forceinline__ void warp_bitonic_sort(float* v, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        const int p = i ^ j;
        if (p > i) {
          const float a = v[i], b = v[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            v[i] = b;
            v[p] = a;
          }
        }
      }
      __syncwarp();
    }
  }
}

struct ResampleArgs {
  const float* z;        // [R,S] coarse depths
  const float* w0;       // [R, ld_w] weights of draw 0 (pointer already offset to column 1)
  const float* w1;       // draw 1 or nullptr
  int64_t ld_w;
  const float* u0;       // [R,n0] or nullptr (deterministic)
  const float* u1;       // [R,n1] or nullptr
  int n0, n1;
  int64_t R;
  int S;
  float eps;
  float* z_fine;         // [R, S+n0+n1]
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32) resample_merge_kernel(ResampleArgs a) {
  __shared__ float s_cdf[kWarpsPerBlock][kMaxBins + 1];
  __shared__ float s_bins[kWarpsPerBlock][kMaxBins + 1];
  __shared__ float s_all[kWarpsPerBlock][kMaxFine];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(kWarpsPerBlock) + warp;
  if (r >= a.R) return;
  const int S = a.S, nb = S - 1, nw = S - 2;
  float* cdf = s_cdf[warp];
  float* sb = s_bins[warp];
  float* all = s_all[warp];
  const float* zr = a.z + r * S;
  for (int i = lane; i < S; i += 32) all[i] = zr[i];
  for (int i = lane; i < nb; i += 32) sb[i] = 0.5f * (zr[i] + zr[i + 1]);   // rendering.py:264-266
  __syncwarp();
  int filled = S;
  for (int draw = 0; draw < 2; ++draw) {
    const float* w = draw == 0 ? a.w0 : a.w1;
    const float* u = draw == 0 ? a.u0 : a.u1;
    const int n = draw == 0 ? a.n0 : a.n1;
    if (!w || n <= 0) continue;
    build_cdf(w + r * a.ld_w, nw, a.eps, cdf, lane);
    for (int j = lane; j < n; j += 32) {
      const float uu = u ? u[r * n + j] : det_u(j, n);
      all[filled + j] = invert_cdf(cdf, sb, nw, uu, a.eps, nullptr);
    }
    filled += n;
    __syncwarp();
  }
  int n2 = 1;
  while (n2 < filled) n2 <<= 1;
  for (int i = filled + lane; i < n2; i += 32) all[i] = __int_as_float(0x7f800000);  // +inf pad
  __syncwarp();
  warp_bitonic_sort(all, n2, lane);
  for (int i = lane; i < filled; i += 32) a.z_fine[r * filled + i] = all[i];
}

// ---------------------------------------------------------------------------------------
// Fast path of resample_merge for the shapes training uses (S and N_importance multiples of 32
// with S == N_importance: 64+64, 128+128).  Still one warp per ray, but nothing is sorted in
// shared memory: the CDF is a warp scan, the new samples are bitonic-sorted in REGISTERS
// (shuffles for strides < 32, register swaps above), and the sorted union with the (already
// sorted) coarse depths is a rank merge -- each element finds its output slot with one binary
// search in the other list.  ~6x fewer instructions per ray than the generic kernel.
template <int E>   // E = S / 32 = N_new / 32 elements per lane
__global__ void __launch_bounds__(kWarpsPerBlock * 32) resample_merge_fast_kernel(ResampleArgs a) {
  constexpr int S = 32 * E, NN = 32 * E, NW = S - 2;
  __shared__ float s_z[kWarpsPerBlock][S];
  __shared__ float s_bins[kWarpsPerBlock][S];
  __shared__ float s_cdf[kWarpsPerBlock][2][S];
  __shared__ float s_new[kWarpsPerBlock][NN];
  __shared__ float s_out[kWarpsPerBlock][S + NN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(kWarpsPerBlock) + warp;
  if (r >= a.R) return;
  float* zs = s_z[warp];
  float* sb = s_bins[warp];
  float* snew = s_new[warp];
  float* sout = s_out[warp];
  const float* zr = a.z + r * S;
  float zreg[E];
#pragma unroll
  for (int i = 0; i < E; ++i) {
    zreg[i] = zr[lane + 32 * i];
    zs[lane + 32 * i] = zreg[i];
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int k = lane + 32 * i;
    if (k < S - 1) sb[k] = 0.5f * (zs[k] + zs[k + 1]);   // rendering.py:264-266
  }
  // ---- CDFs of the one or two draws: pdf -> per-lane contiguous partial sums -> warp scan
#pragma unroll
  for (int draw = 0; draw < 2; ++draw) {
    const float* w = draw == 0 ? a.w0 : a.w1;
    const int n = draw == 0 ? a.n0 : a.n1;
    if (!w || n <= 0) continue;
    float* cdf = s_cdf[warp][draw];
    const float* wr = w + r * a.ld_w;
    float wv[E], part = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int k = lane + 32 * i;
      wv[i] = k < NW ? wr[k] + a.eps : 0.f;
      part += wv[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int k = lane + 32 * i;
      if (k < NW) cdf[k + 1] = wv[i] / part;   // pdf, staged so each lane can take a contiguous run
    }
    __syncwarp();
    // lane owns pdf[lane*E .. lane*E+E) (cdf index +1)
    float run[E], tot = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int k = lane * E + i;
      tot += k < NW ? cdf[k + 1] : 0.f;
      run[i] = tot;
    }
    float incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const float base = incl - tot;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int k = lane * E + i;
      if (k < NW) cdf[k + 1] = base + run[i];
    }
    if (lane == 0) cdf[0] = 0.f;
  }
  __syncwarp();
  // ---- inverse-CDF samples, E per lane (element j = lane + 32 i of the concatenated draws)
  float v[E];
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int j = lane + 32 * i;
    const bool first = j < a.n0;
    const float* cdf = s_cdf[warp][first ? 0 : 1];
    const float* u = first ? a.u0 : a.u1;
    const int n = first ? a.n0 : a.n1;
    const int jj = first ? j : j - a.n0;
    const float uu = u ? u[r * n + jj] : det_u(jj, n);
    v[i] = invert_cdf(cdf, sb, NW, uu, a.eps, nullptr);
  }
  // ---- bitonic sort of the NN new samples; element index e = i*32 + lane
#pragma unroll
  for (int k = 2; k <= NN; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int dj = j >> 5;
#pragma unroll
        for (int i = 0; i < E; ++i) {
          const int p = i ^ dj;
          if (p > i) {
            const bool up = (((i * 32) & k) == 0);
            const float lo = fminf(v[i], v[p]), hi = fmaxf(v[i], v[p]);
            v[i] = up ? lo : hi;
            v[p] = up ? hi : lo;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < E; ++i) {
          const float o = __shfl_xor_sync(0xffffffffu, v[i], j);
          const int e = i * 32 + lane;
          const bool up = (e & k) == 0;
          const bool lower = (lane & j) == 0;   // I hold the smaller index of the pair
          v[i] = (lower == up) ? fminf(v[i], o) : fmaxf(v[i], o);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < E; ++i) snew[i * 32 + lane] = v[i];
  __syncwarp();
  // ---- rank merge: coarse z first on ties
#pragma unroll
  for (int i = 0; i < E; ++i) {
    {  // z element: slot = own index + #new strictly smaller
      const int k = lane + 32 * i;
      const float x = zreg[i];
      int lo = 0, hi = NN;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (snew[mid] < x) lo = mid + 1; else hi = mid;
      }
      sout[k + lo] = x;
    }
    {  // new element: slot = own index + #z smaller or equal
      const int k = i * 32 + lane;
      const float x = v[i];
      int lo = 0, hi = S;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (zs[mid] <= x) lo = mid + 1; else hi = mid;
      }
      sout[k + lo] = x;
    }
  }
  __syncwarp();
  float* out = a.z_fine + r * (S + NN);
#pragma unroll
  for (int i = 0; i < 2 * E; ++i) out[lane + 32 * i] = sout[lane + 32 * i];
}

}  // namespace
}  // namespace upnerf

extern "C" {

int upnerf_stratified_z(const float* rays, const float* perturb_rand, float perturb, int use_disp,
                        int64_t n_rays, int n_samples, float* z, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n_rays > 0 && n_samples > 0, UPNERF_ERR_BAD_SHAPE, "stratified_z: bad sizes");
  UPNERF_REQUIRE(!(perturb > 0.f) || perturb_rand, UPNERF_ERR_BAD_SHAPE,
                 "stratified_z: perturb > 0 needs perturb_rand");
  const int64_t n = n_rays * n_samples;
  LaunchScope scope(kCatSampling, as_stream(stream));
  stratified_z_kernel<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, as_stream(stream)>>>(
      rays, perturb_rand, perturb, use_disp, n_rays, n_samples, z);
  UPNERF_CHECK_LAUNCH("stratified_z_kernel");
  return UPNERF_OK;
}

int upnerf_sample_pdf(const float* bins, int64_t ld_bins, const float* weights, int64_t ld_weights,
                      const float* u, int64_t n_rays, int n_weights, int n_importance, float eps,
                      float* samples, int64_t* inds, float* cdf_out, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n_rays > 0 && n_importance > 0, UPNERF_ERR_BAD_SHAPE, "sample_pdf: bad sizes");
  UPNERF_REQUIRE(n_weights >= 1 && n_weights + 1 <= kMaxBins, UPNERF_ERR_BAD_SHAPE,
                 "sample_pdf: n_weights=%d unsupported (max %d)", n_weights, kMaxBins - 1);
  LaunchScope scope(kCatSampling, as_stream(stream));
  sample_pdf_kernel<<<static_cast<unsigned>(ceil_div64(n_rays, kWarpsPerBlock)), kWarpsPerBlock * 32,
                      0, as_stream(stream)>>>(bins, ld_bins, weights, ld_weights, u, n_rays,
                                               n_weights, n_importance, eps, samples, inds, cdf_out);
  UPNERF_CHECK_LAUNCH("sample_pdf_kernel");
  return UPNERF_OK;
}

int upnerf_searchsorted_right(const float* cdf, int n_cdf, const float* u, int n_u, int64_t n_rays,
                              int64_t* inds, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n_rays > 0 && n_cdf > 0 && n_u > 0, UPNERF_ERR_BAD_SHAPE, "searchsorted: bad sizes");
  const int64_t n = n_rays * n_u;
  LaunchScope scope(kCatSampling, as_stream(stream));
  searchsorted_right_kernel<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, as_stream(stream)>>>(
      cdf, n_cdf, u, n_u, n_rays, inds);
  UPNERF_CHECK_LAUNCH("searchsorted_right_kernel");
  return UPNERF_OK;
}

int upnerf_resample_merge(const float* z, const float* w0, const float* w1, int64_t ld_w,
                          const float* u0, const float* u1, int n0, int n1, int64_t n_rays,
                          int n_samples, float eps, float* z_fine, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n_rays > 0 && n_samples >= 3, UPNERF_ERR_BAD_SHAPE, "resample_merge: bad sizes");
  UPNERF_REQUIRE(n_samples - 1 <= kMaxBins && n_samples + n0 + n1 <= kMaxFine, UPNERF_ERR_BAD_SHAPE,
                 "resample_merge: S=%d n0=%d n1=%d exceeds limits", n_samples, n0, n1);
  ResampleArgs a{z, w0, w1, ld_w, u0, u1, n0, n1, n_rays, n_samples, eps, z_fine};
  // algorithmic traffic: z, the weight rows, the uniforms in; the merged depths out
  const double bytes = 4.0 * n_rays * (n_samples + (n_samples - 2) * (w1 && n1 > 0 ? 2 : 1) +
                                       (u0 ? n0 : 0) + (u1 ? n1 : 0) + n_samples + n0 + n1);
  LaunchScope scope(kCatSampling, as_stream(stream), 0.0, bytes);
  const unsigned blocks = static_cast<unsigned>(ceil_div64(n_rays, kWarpsPerBlock));
  if (n0 + n1 == n_samples && (n_samples == 64 || n_samples == 128)) {
    if (n_samples == 64) resample_merge_fast_kernel<2><<<blocks, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(a);
    else resample_merge_fast_kernel<4><<<blocks, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(a);
    UPNERF_CHECK_LAUNCH("resample_merge_fast_kernel");
    return UPNERF_OK;
  }
  resample_merge_kernel<<<static_cast<unsigned>(ceil_div64(n_rays, kWarpsPerBlock)),
                          kWarpsPerBlock * 32, 0, as_stream(stream)>>>(a);
  UPNERF_CHECK_LAUNCH("resample_merge_kernel");
  return UPNERF_OK;
}

}  // extern "C"
