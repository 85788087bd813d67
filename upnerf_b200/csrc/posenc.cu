// Coarse-to-fine positional encoding (subsystem (b), encoding part), forward and backward.
//
//   upnerf_c2f_weights        models/nerf.py:137-142   w_k = (1 - cos(pi clamp(alpha-k,0,1)))/2
//   upnerf_posenc_fwd         models/nerf.py:126-147   standalone encoding of (M,3) inputs
//   upnerf_points_posenc_fwd  models/rendering.py:251,308 + nerf.py:126-147 fused:
//                             x = o + d z is formed in registers and never written
//   upnerf_points_posenc_bwd  d(encoding) -> d(rays_o), d(rays_d): the pose-gradient path
//
// Output row layout (width 3+6L, padded with zeros to `ld_out`):
//   [x0 x1 x2 | for c in 0..2: w_k sin(x_c f_k) k<L, w_k cos(x_c f_k) k<L],  f_k = fp32(pi) 2^k.
// sin/cos arguments reach ~1e4, so the full-accuracy sincosf is used (no fast-math here) -- in fp32
// (validation) mode for every band.  In bf16 mode the kernels that feed / read bf16 tensors take ONE
// full-accuracy sincosf per coordinate (argument pi x, |x| of a few units) and derive the other bands
// by angle doubling, sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a: the error doubles per band
// (~5e-5 absolute at band 9, below the 1.5e-4 the reference's own fp32 argument rounding moves the
// top band, and 80x below bf16 resolution) and the instruction count drops ~6x -- these kernels
// were bound by 60 sincosf calls per sample, not by HBM.
#include <cuda_bf16.h>

#include "common.h"

namespace upnerf {
namespace {

constexpr int kMaxL = 16;

__device__ __forceinline__ void store_val(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_val(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ float load_val(const float* p) { return *p; }
__device__ __forceinline__ float load_val(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__global__ void c2f_weights_kernel(const float* __restrict__ progress, float start, float end,
                                   int use_c2f, int L, float* __restrict__ w) {
  const int k = threadIdx.x;
  if (k >= L) return;
  if (!use_c2f) {
    w[k] = 1.f;
    return;
  }
  const float kPi = 3.14159265358979323846f;
  // same fp32 operation order as the reference: ((p - start) / (end - start)) * L
  const float alpha = __fmul_rn(__fdiv_rn(progress[0] - start, end - start), static_cast<float>(L));
  float t = alpha - static_cast<float>(k);
  t = fminf(fmaxf(t, 0.f), 1.f);
  w[k] = (1.f - cosf(__fmul_rn(t, kPi))) / 2.f;
}

// next band by angle doubling (bf16 paths only, see the file comment)
__device__ __forceinline__ void double_angle(float& sn, float& cs) {
  const float s2 = 2.f * sn * cs;
  cs = fmaf(-2.f * sn, sn, 1.f);
  sn = s2;
}

template <typename T, bool kFast = false>
__device__ __forceinline__ void encode_row(const float x[3], int L, const float* __restrict__ bw,
                                           T* __restrict__ out, int ld_out) {
  const float kPi = 3.14159265358979323846f;
  store_val(out + 0, x[0]);
  store_val(out + 1, x[1]);
  store_val(out + 2, x[2]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float f = kPi;
    T* o = out + 3 + c * 2 * L;
    float sn, cs;
    if (kFast) sincosf(__fmul_rn(x[c], kPi), &sn, &cs);
    for (int k = 0; k < L; ++k) {
      if (!kFast) sincosf(__fmul_rn(x[c], f), &sn, &cs);
      else if (k > 0) double_angle(sn, cs);
      const float wk = bw[k];
      store_val(o + k, sn * wk);
      store_val(o + L + k, cs * wk);
      f *= 2.f;
    }
  }
  for (int i = 3 + 6 * L; i < ld_out; ++i) store_val(out + i, 0.f);
}

template <typename T>
__global__ void posenc_fwd_kernel(const float* __restrict__ x, int64_t ld_x, int64_t M, int L,
                                  const float* __restrict__ band_w, T* __restrict__ out,
                                  int64_t ld_row, int width) {
  __shared__ float bw[kMaxL];
  if (threadIdx.x < L) bw[threadIdx.x] = band_w[threadIdx.x];
  __syncthreads();
  const int64_t m = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (m >= M) return;
  const float v[3] = {x[m * ld_x], x[m * ld_x + 1], x[m * ld_x + 2]};
  encode_row<T>(v, L, bw, out + m * ld_row, width);
}

template <typename T>
__global__ void points_posenc_fwd_kernel(const float* __restrict__ rays, const float* __restrict__ z,
                                         int64_t M, int S, int L, const float* __restrict__ band_w,
                                         T* __restrict__ out, int64_t ld_row, int width) {
  __shared__ float bw[kMaxL];
  if (threadIdx.x < L) bw[threadIdx.x] = band_w[threadIdx.x];
  __syncthreads();
  const int64_t m = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (m >= M) return;
  const int64_t r = m / S;
  const float* ray = rays + r * 8;
  const float zz = z[m];
  float v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = __fadd_rn(ray[c], __fmul_rn(ray[3 + c], zz));
  encode_row<T>(v, L, bw, out + m * ld_row, width);
}


// Tiled variant for width == 64 (the xyz encoding feeding layer 1 / the skip buffer): each thread
// encodes one row into shared memory, then the block writes the 128 rows with 16-byte stores
// (a warp covers four 128-byte rows per instruction instead of 32 scattered 2-byte stores).
template <typename T>
__global__ void __launch_bounds__(128)
points_posenc_fwd_tiled_kernel(const float* __restrict__ rays, const float* __restrict__ z, int64_t M, int S,
                               int L, const float* __restrict__ band_w, T* __restrict__ out, int64_t ld_row) {
  __shared__ float bw[kMaxL];
  // staged in the OUTPUT type (bf16 rows are 132 B: twice the resident blocks per SM of an fp32 tile);
  // odd word stride -> the row-per-thread writes are bank-conflict free
  constexpr int kStride = sizeof(T) == 2 ? 66 : 65;
  __shared__ T tile[128][kStride];
  if (threadIdx.x < L) bw[threadIdx.x] = band_w[threadIdx.x];
  __syncthreads();
  const int64_t m0 = blockIdx.x * 128ll;
  const int64_t m = m0 + threadIdx.x;
  if (m < M) {
    const int64_t r = m / S;
    const float* ray = rays + r * 8;
    const float zz = z[m];
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __fadd_rn(ray[c], __fmul_rn(ray[3 + c], zz));
    encode_row<T, sizeof(T) == 2>(v, L, bw, &tile[threadIdx.x][0], 64);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 128 * 8; idx += 128) {
    const int row = idx >> 3, c8 = idx & 7;
    if (m0 + row >= M) continue;
    const T* src = &tile[row][c8 * 8];
    T* dst = out + (m0 + row) * ld_row + c8 * 8;
    if constexpr (sizeof(T) == 2) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(src);   // rows are 4-byte aligned (stride 132 B)
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      float4* d4 = reinterpret_cast<float4*>(dst);
      d4[0] = make_float4(src[0], src[1], src[2], src[3]);
      d4[1] = make_float4(src[4], src[5], src[6], src[7]);
    }
  }
}

// One warp per ray: each lane walks samples lane, lane+32, ...; dx is reduced over the ray.
template <typename T>
__global__ void __launch_bounds__(128)
points_posenc_bwd_kernel(const T* __restrict__ d_pe, int64_t ld_row, const float* __restrict__ rays,
                         const float* __restrict__ z, int64_t R, int S, int L,
                         const float* __restrict__ band_w, float* __restrict__ d_rays) {
  __shared__ float bw[kMaxL];
  if (threadIdx.x < L) bw[threadIdx.x] = band_w[threadIdx.x];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * 4ll + warp;
  if (r >= R) return;
  const float kPi = 3.14159265358979323846f;
  const float* ray = rays + r * 8;
  float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
  for (int s = lane; s < S; s += 32) {
    const int64_t m = r * S + s;
    const float zz = z[m];
    const T* g = d_pe + m * ld_row;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xc = __fadd_rn(ray[c], __fmul_rn(ray[3 + c], zz));
      float dx = load_val(g + c);
      float f = kPi;
      const T* gc = g + 3 + c * 2 * L;
      constexpr bool kFast = sizeof(T) == 2;
      float sn, cs;
      if (kFast) sincosf(__fmul_rn(xc, kPi), &sn, &cs);
      for (int k = 0; k < L; ++k) {
        if (!kFast) sincosf(__fmul_rn(xc, f), &sn, &cs);
        else if (k > 0) double_angle(sn, cs);
        dx += bw[k] * f * (cs * load_val(gc + k) - sn * load_val(gc + L + k));
        f *= 2.f;
      }
      go[c] += dx;
      gd[c] += dx * zz;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      go[c] += __shfl_xor_sync(0xffffffffu, go[c], o);
      gd[c] += __shfl_xor_sync(0xffffffffu, gd[c], o);
    }
  }
  if (lane == 0) {
    float* out = d_rays + r * 8;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      out[c] += go[c];
      out[3 + c] += gd[c];
    }
  }
}

// Same, for rows that are exactly 64 elements wide and 16-byte aligned (the skip-buffer layout,
// 3 + 6L <= 64): a lane fetches its whole gradient row with 16-byte loads before the trigonometry
// instead of 63 scattered element loads.  L is a template parameter so the row stays in registers.
template <typename T, int L>
__global__ void __launch_bounds__(128)
points_posenc_bwd_row_kernel(const T* __restrict__ d_pe, int64_t ld_row, const float* __restrict__ rays,
                             const float* __restrict__ z, int64_t R, int S,
                             const float* __restrict__ band_w, float* __restrict__ d_rays) {
  static_assert(3 + 6 * L <= 64, "row does not fit the 64-wide layout");
  __shared__ float bw[kMaxL];
  if (threadIdx.x < L) bw[threadIdx.x] = band_w[threadIdx.x];
  __syncthreads();
  // one warp per (ray, 32-sample chunk): four times the warps of a warp-per-ray mapping, so a
  // few thousand rays still fill the machine evenly; the per-ray sum finishes with atomics
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpr = (S + 31) >> 5;
  const int64_t q = blockIdx.x * 4ll + warp;
  if (q >= R * cpr) return;
  const int64_t r = q / cpr;
  const int s = static_cast<int>(q - r * cpr) * 32 + lane;
  const float kPi = 3.14159265358979323846f;
  const float* ray = rays + r * 8;
  float o3[3], d3[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o3[c] = ray[c];
    d3[c] = ray[3 + c];
  }
  float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
  // bf16 rows (128 bytes): the warp fetches its 32 rows with eight COALESCED 16-byte loads per lane (one
  // instruction covers four whole rows) and hands each lane its row through shared memory.  A lane reading its
  // own row straight from global memory touches 32 different lines per instruction: 256 LSU wavefronts per 4 KB
  // instead of 32 + 32 + 32, which capped the kernel at ~3.7 TB/s.
  __shared__ uint4 stage[sizeof(T) == 2 ? 4 : 1][sizeof(T) == 2 ? 32 : 1][sizeof(T) == 2 ? 9 : 1];
  if constexpr (sizeof(T) == 2) {
    const int s0 = s - lane;
    const T* base = d_pe + (r * S + s0) * ld_row;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int idx = v * 32 + lane, row = idx >> 3, c = idx & 7;
      uint4 t = make_uint4(0u, 0u, 0u, 0u);
      if (s0 + row < S) t = __ldg(reinterpret_cast<const uint4*>(base + row * ld_row) + c);
      stage[warp][row][c] = t;
    }
    __syncwarp();
  }
  if (s < S) {
    const int64_t m = r * S + s;
    const float zz = z[m];
    float g[64];
    if constexpr (sizeof(T) == 2) {
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const uint4 t = stage[warp][lane][v];
        const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(b[i]);
          g[v * 8 + 2 * i] = f.x;
          g[v * 8 + 2 * i + 1] = f.y;
        }
      }
    } else {
      const float4* src = reinterpret_cast<const float4*>(d_pe + m * ld_row);
#pragma unroll
      for (int v = 0; v < 16; ++v) {
        const float4 t = __ldg(src + v);
        g[v * 4] = t.x; g[v * 4 + 1] = t.y; g[v * 4 + 2] = t.z; g[v * 4 + 3] = t.w;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xc = __fadd_rn(o3[c], __fmul_rn(d3[c], zz));
      float dx = g[c];
      float f = kPi;
      constexpr bool kFast = sizeof(T) == 2;
      float sn, cs;
      if (kFast) sincosf(__fmul_rn(xc, kPi), &sn, &cs);
#pragma unroll
      for (int k = 0; k < L; ++k) {
        if (!kFast) sincosf(__fmul_rn(xc, f), &sn, &cs);
        else if (k > 0) double_angle(sn, cs);
        dx += bw[k] * f * (cs * g[3 + c * 2 * L + k] - sn * g[3 + c * 2 * L + L + k]);
        f *= 2.f;
      }
      go[c] += dx;
      gd[c] += dx * zz;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      go[c] += __shfl_xor_sync(0xffffffffu, go[c], o);
      gd[c] += __shfl_xor_sync(0xffffffffu, gd[c], o);
    }
  }
  if (lane == 0) {
    float* out = d_rays + r * 8;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicAdd(out + c, go[c]);
      atomicAdd(out + 3 + c, gd[c]);
    }
  }
}

}  // namespace
}  // namespace upnerf

extern "C" {

int upnerf_c2f_weights(const float* progress_dev, float start, float end, int use_c2f, int L,
                       float* band_w, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(L >= 1 && L <= kMaxL, UPNERF_ERR_BAD_SHAPE, "c2f_weights: L=%d", L);
  UPNERF_REQUIRE(!use_c2f || progress_dev, UPNERF_ERR_BAD_SHAPE, "c2f_weights: progress missing");
  LaunchScope scope(kCatPosenc, as_stream(stream));
  c2f_weights_kernel<<<1, 32, 0, as_stream(stream)>>>(progress_dev, start, end, use_c2f, L, band_w);
  UPNERF_CHECK_LAUNCH("c2f_weights_kernel");
  return UPNERF_OK;
}

int upnerf_posenc_fwd(const float* x, int64_t ld_x, int64_t M, int L, const float* band_w, void* out,
                      int64_t ld_out, int width, int dtype, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(M > 0 && L >= 1 && L <= kMaxL && width >= 3 + 6 * L && ld_out >= width,
                 UPNERF_ERR_BAD_SHAPE, "posenc_fwd: M=%lld L=%d width=%d ld=%lld", (long long)M, L,
                 width, (long long)ld_out);
  const unsigned grid = static_cast<unsigned>(ceil_div64(M, 128));
  LaunchScope scope(kCatPosenc, as_stream(stream));
  if (dtype == UPNERF_BF16)
    posenc_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, as_stream(stream)>>>(
        x, ld_x, M, L, band_w, static_cast<__nv_bfloat16*>(out), ld_out, width);
  else
    posenc_fwd_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(x, ld_x, M, L, band_w,
                                                                   static_cast<float*>(out), ld_out, width);
  UPNERF_CHECK_LAUNCH("posenc_fwd_kernel");
  return UPNERF_OK;
}

int upnerf_points_posenc_fwd(const float* rays, const float* z, int64_t n_rays, int n_samples, int L,
                             const float* band_w, void* out, int64_t ld_out, int width, int dtype,
                             void* stream) {
  using namespace upnerf;
  const int64_t M = n_rays * n_samples;
  UPNERF_REQUIRE(M > 0 && L >= 1 && L <= kMaxL && width >= 3 + 6 * L && ld_out >= width,
                 UPNERF_ERR_BAD_SHAPE, "points_posenc_fwd: bad sizes");
  const unsigned grid = static_cast<unsigned>(ceil_div64(M, 128));
  LaunchScope scope(kCatPosenc, as_stream(stream));
  const size_t es = dtype == UPNERF_BF16 ? 2 : 4;
  if (width == 64 && (ld_out * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    if (dtype == UPNERF_BF16)
      points_posenc_fwd_tiled_kernel<__nv_bfloat16><<<grid, 128, 0, as_stream(stream)>>>(
          rays, z, M, n_samples, L, band_w, static_cast<__nv_bfloat16*>(out), ld_out);
    else
      points_posenc_fwd_tiled_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(
          rays, z, M, n_samples, L, band_w, static_cast<float*>(out), ld_out);
    UPNERF_CHECK_LAUNCH("points_posenc_fwd_tiled_kernel");
    return UPNERF_OK;
  }
  if (dtype == UPNERF_BF16)
    points_posenc_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, as_stream(stream)>>>(
        rays, z, M, n_samples, L, band_w, static_cast<__nv_bfloat16*>(out), ld_out, width);
  else
    points_posenc_fwd_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(
        rays, z, M, n_samples, L, band_w, static_cast<float*>(out), ld_out, width);
  UPNERF_CHECK_LAUNCH("points_posenc_fwd_kernel");
  return UPNERF_OK;
}

int upnerf_points_posenc_bwd(const void* d_pe, int64_t ld_pe, const float* rays, const float* z,
                             int64_t n_rays, int n_samples, int L, const float* band_w,
                             float* d_rays, int dtype, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n_rays > 0 && n_samples > 0 && L >= 1 && L <= kMaxL, UPNERF_ERR_BAD_SHAPE,
                 "points_posenc_bwd: bad sizes");
  const unsigned grid = static_cast<unsigned>(ceil_div64(n_rays, 4));
  LaunchScope scope(kCatPosenc, as_stream(stream));
  const size_t es = dtype == UPNERF_BF16 ? 2 : 4;
  if (L == 10 && ld_pe >= 64 && (ld_pe * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(d_pe) & 15) == 0) {
    const unsigned grid = static_cast<unsigned>(ceil_div64(n_rays * ((n_samples + 31) / 32), 4));
    if (dtype == UPNERF_BF16)
      points_posenc_bwd_row_kernel<__nv_bfloat16, 10><<<grid, 128, 0, as_stream(stream)>>>(
          static_cast<const __nv_bfloat16*>(d_pe), ld_pe, rays, z, n_rays, n_samples, band_w, d_rays);
    else
      points_posenc_bwd_row_kernel<float, 10><<<grid, 128, 0, as_stream(stream)>>>(
          static_cast<const float*>(d_pe), ld_pe, rays, z, n_rays, n_samples, band_w, d_rays);
    UPNERF_CHECK_LAUNCH("points_posenc_bwd_row_kernel");
    return UPNERF_OK;
  }
  if (dtype == UPNERF_BF16)
    points_posenc_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(d_pe), ld_pe, rays, z, n_rays, n_samples, L, band_w, d_rays);
  else
    points_posenc_bwd_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(
        static_cast<const float*>(d_pe), ld_pe, rays, z, n_rays, n_samples, L, band_w, d_rays);
  UPNERF_CHECK_LAUNCH("points_posenc_bwd_kernel");
  return UPNERF_OK;
}

}  // extern "C"
