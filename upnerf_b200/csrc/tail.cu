// Per-ray tail of the train step in one launch -- upnerf_tail_loss.
//
// Replaces, per batch of R rays: the monocular-depth affine correction of
// NeRFSystem.training_step (reference models/nerf_system.py:169-177), UPNeRFLoss.forward
// (losses.py:21-64) and its autograd backward, and the psnr of models/nerf_system.py:202-207.
// The loss is a sum of means of per-ray terms, so every gradient with respect to the render /
// TransientNet outputs is an elementwise expression of the same inputs: the kernel emits the loss
// terms AND those gradients (for an upstream gradient of 1, which is what manual_backward(loss)
// feeds), and scatters d(loss)/d(depth_scale) into the embedding gradient.  ~120 elementwise /
// reduction / sort launches of the eager formulation become one.
//
// One warp per ray; lanes stride over the feature vector with 16-byte accesses.  Loss terms are
// reduced warp -> block -> per-block partials in global memory; the last block to finish (ticket)
// sums the partials in block order, so the reported scalars are bit-reproducible.
#include <math.h>
#include <string.h>

#include "common.h"
#include "internal.h"

namespace upnerf {
namespace {

constexpr int kTerms = 9;   // l_depth_c, l_feat_c, l_rgb_c, l_depth_f, l_feat_f, l_rgb_f, l_beta, l_alpha, mse
constexpr int kWarps = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) - (x < 0.f); }

__global__ void __launch_bounds__(kWarps * 32)
tail_loss_kernel(const __grid_constant__ upnerf_tail_args a, float* __restrict__ partials,
                 unsigned int* __restrict__ ticket) {
  __shared__ float s_part[kWarps][kTerms];
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t R = a.n_rays;
  const int F = a.feat_dim;
  const float m = a.sched_mult_dev ? __ldg(a.sched_mult_dev) : a.sched_mult;
  const bool lo = m < 1.f, hi = m > 0.f;
  const bool fine = a.has_fine != 0;
  const float invR = 1.f / static_cast<float>(R);
  const float w_depth = a.depth_mult * (1.f - m) * invR;
  const float w_feat = (1.f - m) / (static_cast<float>(R) * static_cast<float>(F));
  const float w_rgb = invR / 3.f;

  float acc[kTerms];
#pragma unroll
  for (int i = 0; i < kTerms; ++i) acc[i] = 0.f;

  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kWarps + warp; r < R;
       r += static_cast<int64_t>(gridDim.x) * kWarps) {
    if (lo) {
      // ---- monocular depth target (models/nerf_system.py:169-177) and the two L1 depth terms
      float depth = 0.f, dd_dscale = 0.f, dd_dshift = 0.f, g_depth = 0.f;
      if (lane == 0) {
        const int64_t im = a.img_idx[r];
        const float es = expf(a.depth_scale[im * 2 + 0]);
        const float shift = a.depth_scale[im * 2 + 1];
        const float id = a.inv_depths[r];
        float inv = id * es + shift;
        const bool c1 = inv < 1.f / a.far_;
        if (c1) inv = 1.f / a.far_;
        depth = 1.f / inv;
        const bool c2 = depth < a.near_;
        if (c2) depth = a.near_;
        const float dd_dinv = (c1 || c2) ? 0.f : -1.f / (inv * inv);
        dd_dscale = dd_dinv * id * es;
        dd_dshift = dd_dinv;
        {
          const float diff = a.s_depth_c[r] - depth;
          const float tw = a.t_weight_c ? 1.f - a.t_weight_c[r] : 1.f;
          acc[0] += fabsf(diff) * tw;
          const float g = sgn(diff) * tw * w_depth;
          if (a.g_s_depth_c) a.g_s_depth_c[r] = g;
          g_depth -= g;
        }
        if (fine) {
          const float diff = a.s_depth_f[r] - depth;
          const float tw = a.t_weight_f ? 1.f - a.t_weight_f[r] : 1.f;
          acc[3] += fabsf(diff) * tw;
          const float g = sgn(diff) * tw * w_depth;
          if (a.g_s_depth_f) a.g_s_depth_f[r] = g;
          g_depth -= g;
        }
        if (a.d_depth_scale && g_depth != 0.f) {
          if (dd_dscale != 0.f) atomicAdd(a.d_depth_scale + im * 2 + 0, g_depth * dd_dscale);
          if (dd_dshift != 0.f) atomicAdd(a.d_depth_scale + im * 2 + 1, g_depth * dd_dshift);
        }
      }
      // ---- feature L2 terms (losses.py:33-35,52-54)
      const float* ft = a.feats + r * F;
      const float* fc = a.feat_c + r * F;
      const float* ff = fine ? a.feat_f + r * F : nullptr;
      float* gc = a.g_feat_c ? a.g_feat_c + r * F : nullptr;
      float* gf = (fine && a.g_feat_f) ? a.g_feat_f + r * F : nullptr;
      for (int c = lane * 4; c < F; c += 128) {
        const float4 t = *reinterpret_cast<const float4*>(ft + c);
        {
          const float4 v = *reinterpret_cast<const float4*>(fc + c);
          const float d0 = v.x - t.x, d1 = v.y - t.y, d2 = v.z - t.z, d3 = v.w - t.w;
          acc[1] += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          if (gc) {
            const float k = 2.f * w_feat;
            *reinterpret_cast<float4*>(gc + c) = make_float4(k * d0, k * d1, k * d2, k * d3);
          }
        }
        if (ff) {
          const float4 v = *reinterpret_cast<const float4*>(ff + c);
          const float d0 = v.x - t.x, d1 = v.y - t.y, d2 = v.z - t.z, d3 = v.w - t.w;
          acc[4] += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          if (gf) {
            const float k = 2.f * w_feat;
            *reinterpret_cast<float4*>(gf + c) = make_float4(k * d0, k * d1, k * d2, k * d3);
          }
        }
      }
    }
    // ---- colour terms, beta / alpha regularisers (losses.py:39-41,60-64) and the psnr mse
    if (lane < 3) {
      const float t = a.rgbs[r * 3 + lane];
      if (hi) {
        const float d = a.s_rgb_c[r * 3 + lane] - t;
        acc[2] += d * d;
        if (a.g_s_rgb_c) a.g_s_rgb_c[r * 3 + lane] = d * m * w_rgb;   // 2 d * (m/2) / (3R)
      }
      const float* sf = fine ? a.s_rgb_f : a.s_rgb_c;
      float sq = 0.f;
      if (sf) {
        const float d = sf[r * 3 + lane] - t;
        sq = d * d;
        acc[8] += sq;
        if (hi && fine) {
          if (a.t_beta) {
            const float b = a.t_beta[r];
            const float ib2 = 1.f / (2.f * b * b);
            acc[5] += sq * ib2;
            if (a.g_s_rgb_f) a.g_s_rgb_f[r * 3 + lane] = 2.f * d * ib2 * m * w_rgb;
          } else {
            acc[5] += sq;
            if (a.g_s_rgb_f) a.g_s_rgb_f[r * 3 + lane] = 2.f * d * m * w_rgb;
          }
        }
      }
      // d/d beta of sum_c sq_c / (2 beta^2): -sum_c sq_c / beta^3
      float s3 = sq + __shfl_down_sync(0x7u, sq, 1) + __shfl_down_sync(0x7u, sq, 2);
      if (lane == 0 && hi && fine && a.t_beta) {
        const float b = a.t_beta[r];
        acc[6] += logf(b);
        acc[7] += a.t_alpha[r];
        if (a.g_t_beta) a.g_t_beta[r] = -s3 / (b * b * b) * m * w_rgb + m * invR / b;
        if (a.g_t_alpha) a.g_t_alpha[r] = a.alpha_reg * m * invR;
      }
    }
  }

#pragma unroll
  for (int i = 0; i < kTerms; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < kTerms; ++i) s_part[warp][i] = acc[i];
  __syncthreads();
  if (threadIdx.x < kTerms) {
    float s = 0.f;
    for (int w = 0; w < kWarps; ++w) s += s_part[w][threadIdx.x];
    partials[blockIdx.x * kTerms + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < kTerms) {
    float s = 0.f;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(partials + b * kTerms + threadIdx.x);
    s_part[0][threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float* S = s_part[0];
    float* L = a.losses;
    const float n3 = 3.f * static_cast<float>(R);
    for (int i = 0; i < UPNERF_TAIL_LOSS_SLOTS; ++i) L[i] = 0.f;
    float total = 0.f;
    // same term order as the reference's dict (losses.py:21-64): the total is summed in that order
    if (lo) {
      L[0] = S[0] * invR * a.depth_mult * (1.f - m); total += L[0];
      L[1] = S[1] / (static_cast<float>(R) * static_cast<float>(F)) * (1.f - m); total += L[1];
    }
    if (hi) { L[2] = S[2] / n3 * m / 2.f; total += L[2]; }
    if (fine) {
      if (lo) {
        L[3] = S[3] * invR * a.depth_mult * (1.f - m); total += L[3];
        L[4] = S[4] / (static_cast<float>(R) * static_cast<float>(F)) * (1.f - m); total += L[4];
      }
      if (hi) {
        L[5] = S[5] / n3 * m; total += L[5];
        if (a.t_beta) {
          L[6] = S[6] * invR * m; total += L[6];
          L[7] = S[7] * invR * a.alpha_reg * m; total += L[7];
        }
      }
    }
    L[8] = total;
    L[9] = (a.s_rgb_c || a.s_rgb_f) ? -10.f * log10f(S[8] / n3) : 0.f;   // psnr of s_rgb_{fine|coarse}
    *ticket = 0;   // re-armed for the next launch
  }
}

}  // namespace
}  // namespace upnerf

extern "C" uint64_t upnerf_tail_workspace_bytes(void) {
  return (static_cast<uint64_t>(upnerf::sm_count()) * 4 * upnerf::kTerms + 64) * sizeof(float);
}

extern "C" int upnerf_tail_loss(const upnerf_tail_args* a, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(a && a->n_rays > 0 && a->losses && a->workspace, UPNERF_ERR_BAD_SHAPE, "tail_loss: missing arguments");
  UPNERF_REQUIRE(a->workspace_bytes >= upnerf_tail_workspace_bytes(), UPNERF_ERR_WORKSPACE, "tail_loss: workspace too small");
  UPNERF_REQUIRE(a->rgbs, UPNERF_ERR_BAD_SHAPE, "tail_loss: rgbs missing");
  const bool lo = a->sched_mult < 1.f, hi = a->sched_mult > 0.f;
  if (lo) {
    UPNERF_REQUIRE(a->feat_dim > 0 && a->feat_dim % 4 == 0, UPNERF_ERR_BAD_CONFIG,
                   "tail_loss: feat_dim=%d (the encode_feat=False colour-candidate loss is not implemented)", a->feat_dim);
    UPNERF_REQUIRE(a->img_idx && a->inv_depths && a->depth_scale && a->feats && a->s_depth_c && a->feat_c,
                   UPNERF_ERR_BAD_SHAPE, "tail_loss: depth / feature inputs missing for sched_mult < 1");
    UPNERF_REQUIRE(!a->has_fine || (a->s_depth_f && a->feat_f), UPNERF_ERR_BAD_SHAPE, "tail_loss: fine inputs missing");
  }
  if (hi) {
    UPNERF_REQUIRE(a->s_rgb_c && (!a->has_fine || a->s_rgb_f), UPNERF_ERR_BAD_SHAPE, "tail_loss: s_rgb missing");
    UPNERF_REQUIRE(!a->t_beta || a->t_alpha, UPNERF_ERR_BAD_SHAPE, "tail_loss: t_alpha missing");
  }
  cudaStream_t st = as_stream(stream);
  float* partials = static_cast<float*>(a->workspace) + 64;
  unsigned int* ticket = static_cast<unsigned int*>(a->workspace);   // zero on first use (caller zero-fills once)
  int grid = sm_count() * 4;
  const int64_t need = ceil_div64(a->n_rays, kWarps);
  if (grid > need) grid = static_cast<int>(need);
  LaunchScope scope(kCatHeads, st, 0.0,
                    4.0 * a->n_rays * ((lo ? (a->has_fine ? 5.0 : 3.0) * a->feat_dim : 0.0) + 24.0));
  tail_loss_kernel<<<grid, kWarps * 32, 0, st>>>(*a, partials, ticket);
  UPNERF_CHECK_LAUNCH("tail_loss_kernel");
  return UPNERF_OK;
}
