// Small kernels around the dense layers: per-ray embedding gathers/scatters
// (models/rendering.py:255-258,309-312), the rgb output layer's backward
// (models/nerf.py:59-64), per-ray reductions of per-sample gradients, the fp32-mode
// row-dot heads (share_sigma / candidate_sigma / rgb out, models/nerf.py:51,62,74) and
// weight packing (fp32 master parameters -> bf16/fp32 GEMM operands, padded / permuted /
// transposed as the kernels want them).
#include <cuda_bf16.h>

#include "common.h"
#include "internal.h"

namespace upnerf {
namespace {

__device__ __forceinline__ float ldv(const float* p) { return *p; }
__device__ __forceinline__ float ldv(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stv(float* p, float v) { *p = v; }
__device__ __forceinline__ void stv(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// four consecutive elements (16-byte / 8-byte aligned) as one memory transaction
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 t;
  *reinterpret_cast<__nv_bfloat162*>(&t.x) = __floats2bfloat162_rn(v[0], v[1]);
  *reinterpret_cast<__nv_bfloat162*>(&t.y) = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ float softplus_ref(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void gather_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx,
                                   int64_t R, int dim, float* __restrict__ out, int64_t ld_out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= R * dim) return;
  const int64_t r = i / dim;
  const int c = static_cast<int>(i - r * dim);
  out[r * ld_out + c] = table[idx[r] * dim + c];
}

__global__ void scatter_add_rows_kernel(const float* __restrict__ src, int64_t ld_src,
                                        const int64_t* __restrict__ idx, int64_t R, int dim,
                                        float* __restrict__ table) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= R * dim) return;
  const int64_t r = i / dim;
  const int c = static_cast<int>(i - r * dim);
  atomicAdd(table + idx[r] * dim + c, src[r * ld_src + c]);
}

// 16 bytes = kVec elements of T per lane
template <typename T> struct Ld16;
template <> struct Ld16<float> {
  static constexpr int kVec = 4;
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct Ld16<__nv_bfloat16> {
  static constexpr int kVec = 8;
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(b[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};

// out[r, :] = sum over the S samples of ray r of X[r*S + s, :]   (N = 128, one warp per ray).
// A row is 128/kVec lanes of 16 bytes, so a warp covers kRows = 32*kVec/128 rows per load and
// keeps four such loads in flight.
template <typename T>
__global__ void __launch_bounds__(128)
ray_sum128_kernel(const T* __restrict__ X, int64_t ld, int64_t R, int S, float* __restrict__ out) {
  constexpr int kVec = Ld16<T>::kVec;
  constexpr int kLanesPerRow = 128 / kVec;       // 16 (bf16) or 32 (fp32)
  constexpr int kRows = 32 / kLanesPerRow;       // rows per warp-wide load
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * 4ll + warp;
  if (r >= R) return;
  const int sub = lane / kLanesPerRow, cl = lane % kLanesPerRow;
  float acc[kVec];
#pragma unroll
  for (int e = 0; e < kVec; ++e) acc[e] = 0.f;
  const T* base = X + r * S * ld + cl * kVec;
  for (int s0 = 0; s0 < S; s0 += 4 * kRows) {
    float v[4][kVec];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int srow = s0 + j * kRows + sub;
      Ld16<T>::ld(base + static_cast<int64_t>(srow < S ? srow : S - 1) * ld, v[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (s0 + j * kRows + sub < S) {
#pragma unroll
        for (int e = 0; e < kVec; ++e) acc[e] += v[j][e];
      }
    }
  }
  if (kRows == 2) {
#pragma unroll
    for (int e = 0; e < kVec; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
  }
  if (sub == 0) {
    float* o = out + r * 128 + cl * kVec;
#pragma unroll
    for (int e = 0; e < kVec; e += 4)
      *reinterpret_cast<float4*>(o + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
  }
}

// Backward of  rgb = sigmoid(W2 q + b2)  with q = relu(...) (N = 128 hidden units):
//   d q_pre = (W2^T (d_rgb * rgb (1-rgb))) * (q > 0);  dW2, db2 accumulated;  per-ray sum of
//   d q_pre -> d_raybias (the per-ray bias of the hidden layer carries direction/appearance).
template <typename T>
__global__ void __launch_bounds__(128, 7)
rgb_head_bwd_kernel(const T* __restrict__ Q, int64_t ldq, const float* __restrict__ rgb,
                    const float* __restrict__ d_rgb, const float* __restrict__ W2, int64_t R, int S,
                    T* __restrict__ dQ, int64_t lddq, float* __restrict__ d_raybias,
                    float* __restrict__ dW2, float* __restrict__ db2) {
  __shared__ float red[4][3 * 128 + 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float w2[3][4];
#pragma unroll
  for (int h = 0; h < 3; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) w2[h][e] = W2[h * 128 + lane * 4 + e];
  float aw[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float ab[3] = {0.f, 0.f, 0.f};
  constexpr int kB = 4;  // samples whose loads are in flight together (8 was slower: 130 vs 92 us per launch)
  for (int64_t r = blockIdx.x * 4ll + warp; r < R; r += gridDim.x * 4ll) {
    float rb[4] = {0.f, 0.f, 0.f, 0.f};
    for (int sb = 0; sb < S; sb += 32) {
      // g = d_rgb * rgb (1 - rgb) of 32 samples at once: lane j owns sample sb + j (two coalesced 384-byte reads
      // instead of six broadcast loads per sample and lane: the kernel was LSU-issue bound), shuffled to the
      // warp sample by sample below
      float gl[3] = {0.f, 0.f, 0.f};
      if (sb + lane < S) {
        const int64_t m = r * S + sb + lane;
#pragma unroll
        for (int h = 0; h < 3; ++h) {
          const float y = __ldg(rgb + m * 3 + h);
          gl[h] = __ldg(d_rgb + m * 3 + h) * y * (1.f - y);
        }
      }
#pragma unroll
      for (int h = 0; h < 3; ++h) ab[h] += gl[h];       // (summed over lanes at the end)
      const int nb = S - sb < 32 ? S - sb : 32;
      for (int s0 = 0; s0 < nb; s0 += kB) {
        float qv[kB][4];
#pragma unroll
        for (int j = 0; j < kB; ++j) {
          const int64_t m = r * S + sb + min(s0 + j, nb - 1);
          ld4(Q + m * ldq + lane * 4, qv[j]);
        }
#pragma unroll
        for (int j = 0; j < kB; ++j) {
          if (s0 + j >= nb) break;
          const int64_t m = r * S + sb + s0 + j;
          const float g0 = __shfl_sync(0xffffffffu, gl[0], s0 + j);
          const float g1 = __shfl_sync(0xffffffffu, gl[1], s0 + j);
          const float g2 = __shfl_sync(0xffffffffu, gl[2], s0 + j);
          float d[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            d[e] = qv[j][e] > 0.f ? (g0 * w2[0][e] + g1 * w2[1][e] + g2 * w2[2][e]) : 0.f;
            rb[e] += d[e];
            aw[0][e] += g0 * qv[j][e];
            aw[1][e] += g1 * qv[j][e];
            aw[2][e] += g2 * qv[j][e];
          }
          st4(dQ + m * lddq + lane * 4, d);
        }
      }
    }
    float* o = d_raybias + r * 128 + lane * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = rb[e];
  }
#pragma unroll
  for (int h = 0; h < 3; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) red[warp][h * 128 + lane * 4 + e] = aw[h][e];
#pragma unroll
  for (int h = 0; h < 3; ++h)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ab[h] += __shfl_xor_sync(0xffffffffu, ab[h], o);
  if (lane == 0) {
    red[warp][384] = ab[0];
    red[warp][385] = ab[1];
    red[warp][386] = ab[2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 387; i += 128) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) t += red[w][i];
    if (i < 384) atomicAdd(dW2 + i, t); else atomicAdd(db2 + (i - 384), t);
  }
}

// out[n] += sum_m s[m] X[m,n];  out_s += sum_m s[m]   (s == nullptr -> ones).  N <= 256, even.
template <typename T>
__global__ void __launch_bounds__(256)
rowscale_colsum_kernel(const T* __restrict__ X, int64_t ld, const float* __restrict__ s, int64_t M,
                       int N, float* __restrict__ out, float* __restrict__ out_s) {
  __shared__ float red[256][2];
  const int half = N / 2;
  const int groups = 256 / half;            // rows handled concurrently by one block
  const int g = threadIdx.x / half;
  const int c2 = threadIdx.x % half;
  float a0 = 0.f, a1 = 0.f, as = 0.f;
  if (g < groups) {
    for (int64_t m = blockIdx.x * static_cast<int64_t>(groups) + g; m < M;
         m += static_cast<int64_t>(gridDim.x) * groups) {
      const float sv = s ? s[m] : 1.f;
      a0 = fmaf(sv, ldv(X + m * ld + 2 * c2), a0);
      a1 = fmaf(sv, ldv(X + m * ld + 2 * c2 + 1), a1);
      as += sv;
    }
  }
  red[threadIdx.x][0] = a0;
  red[threadIdx.x][1] = a1;
  __syncthreads();
  if (threadIdx.x < half) {
    float t0 = 0.f, t1 = 0.f;
    for (int k = 0; k < groups; ++k) {
      t0 += red[k * half + threadIdx.x][0];
      t1 += red[k * half + threadIdx.x][1];
    }
    atomicAdd(out + 2 * threadIdx.x, t0);
    atomicAdd(out + 2 * threadIdx.x + 1, t1);
  }
  if (out_s && c2 == 0 && g < groups) atomicAdd(out_s, as);
}

// Same reduction for rows that are whole 16-byte vectors (N * sizeof(T) % 16 == 0, aligned): a lane
// owns one 16-byte column chunk, 256 / (lanes per row) rows are read per block-wide load and four
// such loads are in flight.
template <typename T>
__global__ void __launch_bounds__(256)
rowscale_colsum_vec_kernel(const T* __restrict__ X, int64_t ld, const float* __restrict__ s, int64_t M,
                           int N, float* __restrict__ out, float* __restrict__ out_s) {
  constexpr int kVec = Ld16<T>::kVec;
  __shared__ float red[256][kVec + 1];
  const int lpr = N / kVec;                 // lanes per row (power of two, <= 64)
  const int rpb = 256 / lpr;                // rows per block-wide load
  const int g = threadIdx.x / lpr, cl = threadIdx.x % lpr;
  float acc[kVec];
#pragma unroll
  for (int e = 0; e < kVec; ++e) acc[e] = 0.f;
  float as = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * rpb;
  for (int64_t m0 = blockIdx.x * static_cast<int64_t>(rpb) + g; m0 < M; m0 += 4 * stride) {
    float v[4][kVec], sv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t m = m0 + j * stride;
      const int64_t mc = m < M ? m : M - 1;
      Ld16<T>::ld(X + mc * ld + cl * kVec, v[j]);
      sv[j] = m < M ? (s ? __ldg(s + mc) : 1.f) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < kVec; ++e) acc[e] = fmaf(sv[j], v[j][e], acc[e]);
      as += sv[j];
    }
  }
#pragma unroll
  for (int e = 0; e < kVec; ++e) red[threadIdx.x][e] = acc[e];
  red[threadIdx.x][kVec] = as;
  __syncthreads();
  if (threadIdx.x < N) {
    const int c = threadIdx.x / kVec, e = threadIdx.x % kVec;
    float t = 0.f;
    for (int k = 0; k < rpb; ++k) t += red[k * lpr + c][e];
    atomicAdd(out + threadIdx.x, t);
  }
  if (out_s && threadIdx.x == 255) {
    float t = 0.f;
    for (int k = 0; k < rpb; ++k) t += red[k * lpr][kVec];
    atomicAdd(out_s, t);
  }
}

// fp32-mode row-dot heads: out[m,h] = act(X[m,:] . w[h,:] + b[h]); one warp per row.
__global__ void __launch_bounds__(128)
rowdot_head_kernel(const float* __restrict__ X, int64_t ld, int64_t M, int N, int nh,
                   const float* __restrict__ w, const float* __restrict__ b, int act,
                   float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = blockIdx.x * 4ll + warp;
  if (m >= M) return;
  for (int h = 0; h < nh; ++h) {
    float acc = 0.f;
    for (int n = lane; n < N; n += 32) acc = fmaf(X[m * ld + n], w[h * N + n], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float x = acc + (b ? b[h] : 0.f);
      if (act == 1) x = softplus_ref(x);
      else if (act == 2) x = sigmoid_ref(x);
      out[m * nh + h] = x;
    }
  }
}

template <typename T>
__global__ void pack_kernel(const PackList list) {
  const PackOp& op = list.ops[blockIdx.y];
  const int64_t n = static_cast<int64_t>(op.rows) * op.cols;
  T* dst = static_cast<T*>(op.dst);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / op.cols), c = static_cast<int>(i % op.cols);
    const float v = op.src[static_cast<int64_t>(r) * op.ld_src + c];
    if (op.transpose) stv(dst + static_cast<int64_t>(c) * op.ld_dst + r, v);
    else stv(dst + static_cast<int64_t>(r) * op.ld_dst + c, v);
  }
}

}  // namespace

int gather_rows(const float* table, const int64_t* idx, int64_t R, int dim, float* out, int64_t ld_out,
                cudaStream_t st) {
  const int64_t n = R * dim;
  LaunchScope scope(kCatHeads, st);
  gather_rows_kernel<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, st>>>(table, idx, R, dim, out, ld_out);
  UPNERF_CHECK_LAUNCH("gather_rows_kernel");
  return UPNERF_OK;
}

int scatter_add_rows(const float* src, int64_t ld_src, const int64_t* idx, int64_t R, int dim,
                     float* table, cudaStream_t st) {
  const int64_t n = R * dim;
  LaunchScope scope(kCatHeads, st);
  scatter_add_rows_kernel<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, st>>>(src, ld_src, idx, R, dim, table);
  UPNERF_CHECK_LAUNCH("scatter_add_rows_kernel");
  return UPNERF_OK;
}

int ray_sum128(const void* X, int64_t ld, int64_t R, int S, float* out, int dtype, cudaStream_t st) {
  const unsigned grid = static_cast<unsigned>(ceil_div64(R, 4));
  LaunchScope scope(kCatHeads, st);
  if (dtype == UPNERF_BF16)
    ray_sum128_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(X), ld, R, S, out);
  else
    ray_sum128_kernel<float><<<grid, 128, 0, st>>>(static_cast<const float*>(X), ld, R, S, out);
  UPNERF_CHECK_LAUNCH("ray_sum128_kernel");
  return UPNERF_OK;
}

int rgb_head_bwd(const void* Q, int64_t ldq, const float* rgb, const float* d_rgb, const float* W2,
                 int64_t R, int S, void* dQ, int64_t lddq, float* d_raybias, float* dW2, float* db2,
                 int dtype, cudaStream_t st) {
  int64_t blocks = ceil_div64(R, 4);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 7;
  if (blocks > cap) blocks = cap;
  LaunchScope scope(kCatHeads, st);
  if (dtype == UPNERF_BF16)
    rgb_head_bwd_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 128, 0, st>>>(
        static_cast<const __nv_bfloat16*>(Q), ldq, rgb, d_rgb, W2, R, S, static_cast<__nv_bfloat16*>(dQ),
        lddq, d_raybias, dW2, db2);
  else
    rgb_head_bwd_kernel<float><<<static_cast<unsigned>(blocks), 128, 0, st>>>(
        static_cast<const float*>(Q), ldq, rgb, d_rgb, W2, R, S, static_cast<float*>(dQ), lddq,
        d_raybias, dW2, db2);
  UPNERF_CHECK_LAUNCH("rgb_head_bwd_kernel");
  return UPNERF_OK;
}

int rowscale_colsum(const void* X, int64_t ld, const float* s, int64_t M, int N, float* out,
                    float* out_s, int dtype, cudaStream_t st) {
  // wide inputs are processed in column tiles of 128
  if (N > 256) {
    UPNERF_REQUIRE(N % 128 == 0, UPNERF_ERR_BAD_SHAPE, "rowscale_colsum: N=%d", N);
    const size_t es = dtype == UPNERF_BF16 ? 2 : 4;
    for (int c = 0; c < N; c += 128)
      UPNERF_TRY(rowscale_colsum(static_cast<const uint8_t*>(X) + c * es, ld, s, M, 128, out + c,
                                 c == 0 ? out_s : nullptr, dtype, st));
    return UPNERF_OK;
  }
  UPNERF_REQUIRE(N >= 2 && N <= 256 && N % 2 == 0 && 256 % (N / 2) == 0, UPNERF_ERR_BAD_SHAPE,
                 "rowscale_colsum: N=%d", N);
  {
    const size_t es = dtype == UPNERF_BF16 ? 2 : 4;
    const int kvec = static_cast<int>(16 / es);
    const int lpr = N / kvec;
    if (N % kvec == 0 && lpr >= 1 && lpr <= 64 && (lpr & (lpr - 1)) == 0 && (ld * es) % 16 == 0 &&
        reinterpret_cast<uintptr_t>(X) % 16 == 0) {
      const int rpb = 256 / lpr;
      int64_t blocks = ceil_div64(M, static_cast<int64_t>(rpb) * 4);
      const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
      if (blocks > cap) blocks = cap;
      if (blocks < 1) blocks = 1;
      LaunchScope scope(kCatHeads, st);
      if (dtype == UPNERF_BF16)
        rowscale_colsum_vec_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(X), ld, s, M, N, out, out_s);
      else
        rowscale_colsum_vec_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            static_cast<const float*>(X), ld, s, M, N, out, out_s);
      UPNERF_CHECK_LAUNCH("rowscale_colsum_vec_kernel");
      return UPNERF_OK;
    }
  }
  const int groups = 256 / (N / 2);
  int64_t blocks = ceil_div64(M, groups * 8);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  LaunchScope scope(kCatHeads, st);
  if (dtype == UPNERF_BF16)
    rowscale_colsum_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(X), ld, s, M, N, out, out_s);
  else
    rowscale_colsum_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        static_cast<const float*>(X), ld, s, M, N, out, out_s);
  UPNERF_CHECK_LAUNCH("rowscale_colsum_kernel");
  return UPNERF_OK;
}

int rowdot_head(const float* X, int64_t ld, int64_t M, int N, int nh, const float* w, const float* b,
                int act, float* out, cudaStream_t st) {
  LaunchScope scope(kCatHeads, st);
  rowdot_head_kernel<<<static_cast<unsigned>(ceil_div64(M, 4)), 128, 0, st>>>(X, ld, M, N, nh, w, b, act, out);
  UPNERF_CHECK_LAUNCH("rowdot_head_kernel");
  return UPNERF_OK;
}

int run_pack(const PackList& list, int dtype, cudaStream_t st) {
  if (list.n == 0) return UPNERF_OK;
  dim3 grid(64, list.n);
  LaunchScope scope(kCatPack, st);
  if (dtype == UPNERF_BF16) pack_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(list);
  else pack_kernel<float><<<grid, 256, 0, st>>>(list);
  UPNERF_CHECK_LAUNCH("pack_kernel");
  return UPNERF_OK;
}

}  // namespace upnerf
