// Fused Adam over one flat fp32 parameter buffer -- upnerf_adam_step.
//
// Replaces torch.optim.Adam.step for the reference's two optimisers (models/nerf_system.py:41-73,
// 188-195; utils/optim.py:20-44: Adam(eps=1e-8), no weight decay, no amsgrad).  The reference
// optimiser holds ~70 parameter tensors and skips every tensor whose .grad is None in the current
// schedule phase (no update, no step increment, moments untouched).  Here all tensors are views of
// one flat buffer, so the buffer is described as consecutive SEGMENTS, each with its own liveness
// and bias corrections (the host keeps one step counter per segment class): a dead segment is not
// read or written at all, a live one gets exactly torch's single-tensor update
//     m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g g
//     p = p - step_size * m / (sqrt(v) / bc2_sqrt + eps),
//     step_size = lr / (1 - b1^t),  bc2_sqrt = sqrt(1 - b2^t)     (scalars computed in double on the host).
// HBM-bound: 16 B read + 12 B written per live element.
#include <string.h>

#include "common.h"

namespace upnerf {
namespace {

__device__ __forceinline__ int find_segment(const upnerf_adam_args& a, int64_t i) {
  int lo = 0, hi = a.n_segments - 1;  // first segment with seg_end > i
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a.seg_end[mid] > i) hi = mid; else lo = mid + 1;
  }
  return lo;
}

struct Betas {
  float b2, om1, om2, eps;   // beta2, fp32(1 - beta1), fp32(1 - beta2) (rounded from double like torch), eps
  float decay;               // AdamW: param.mul_(1 - lr * weight_decay) first; 1 = plain Adam
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const Betas& k,
                                         float step_size, float bc2_sqrt) {
  const float eps = k.eps;
  p = p * k.decay;
  m = m + (g - m) * k.om1;                            // exp_avg.lerp_(grad, 1 - beta1)
  v = __fmaf_rn(k.om2, g * g, v * k.b2);              // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
  const float denom = __fdiv_rn(__fsqrt_rn(v), bc2_sqrt) + eps;
  p = p - step_size * __fdiv_rn(m, denom);            // param.addcdiv_(exp_avg, denom, -step_size)
}

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ upnerf_adam_args a, const Betas k0) {
  const int64_t i0 = (blockIdx.x * 256ll + threadIdx.x) * 4;
  if (i0 >= a.n) return;
  // per-step scalars: by value, or (CUDA-graph replay) from the device array the host rewrites every step
  const float* ds = a.dev_scalars;
  Betas k = k0;
  if (ds) {
    const float d = __ldg(ds + 2 * UPNERF_ADAM_MAX_SEGMENTS);
    k.decay = (d == 0.f) ? 1.f : d;
  }
  auto step_size = [&](int s) { return ds ? __ldg(ds + s) : a.seg_step_size[s]; };
  auto bc2_sqrt = [&](int s) { return ds ? __ldg(ds + UPNERF_ADAM_MAX_SEGMENTS + s) : a.seg_bc2_sqrt[s]; };
  const int s0 = find_segment(a, i0);
  if (i0 + 4 <= a.seg_end[s0] && i0 + 4 <= a.n) {
    if (!a.seg_live[s0]) return;
    const float ss = step_size(s0), bc = bc2_sqrt(s0);
    float4 p = *reinterpret_cast<float4*>(a.params + i0);
    const float4 g = *reinterpret_cast<const float4*>(a.grads + i0);
    float4 m = *reinterpret_cast<float4*>(a.exp_avg + i0);
    float4 v = *reinterpret_cast<float4*>(a.exp_avg_sq + i0);
    adam_one(p.x, g.x, m.x, v.x, k, ss, bc);
    adam_one(p.y, g.y, m.y, v.y, k, ss, bc);
    adam_one(p.z, g.z, m.z, v.z, k, ss, bc);
    adam_one(p.w, g.w, m.w, v.w, k, ss, bc);
    *reinterpret_cast<float4*>(a.params + i0) = p;
    *reinterpret_cast<float4*>(a.exp_avg + i0) = m;
    *reinterpret_cast<float4*>(a.exp_avg_sq + i0) = v;
    return;
  }
  int s = s0;
  for (int64_t i = i0; i < i0 + 4 && i < a.n; ++i) {   // a group of four straddling a boundary
    while (a.seg_end[s] <= i) ++s;
    if (!a.seg_live[s]) continue;
    adam_one(a.params[i], a.grads[i], a.exp_avg[i], a.exp_avg_sq[i], k, step_size(s), bc2_sqrt(s));
  }
}

}  // namespace
}  // namespace upnerf

extern "C" int upnerf_adam_step(const upnerf_adam_args* a, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(a && a->params && a->grads && a->exp_avg && a->exp_avg_sq && a->n > 0, UPNERF_ERR_BAD_SHAPE,
                 "adam_step: missing buffers");
  UPNERF_REQUIRE(a->n_segments >= 1 && a->n_segments <= UPNERF_ADAM_MAX_SEGMENTS, UPNERF_ERR_BAD_SHAPE,
                 "adam_step: n_segments=%d", a->n_segments);
  int64_t prev = 0;
  for (int s = 0; s < a->n_segments; ++s) {
    UPNERF_REQUIRE(a->seg_end[s] > prev, UPNERF_ERR_BAD_SHAPE, "adam_step: segment ends must ascend");
    prev = a->seg_end[s];
  }
  UPNERF_REQUIRE(prev == a->n, UPNERF_ERR_BAD_SHAPE, "adam_step: last segment must end at n");
  UPNERF_REQUIRE((reinterpret_cast<uintptr_t>(a->params) | reinterpret_cast<uintptr_t>(a->grads) |
                  reinterpret_cast<uintptr_t>(a->exp_avg) | reinterpret_cast<uintptr_t>(a->exp_avg_sq)) % 16 == 0,
                 UPNERF_ERR_BAD_SHAPE, "adam_step: buffers must be 16-byte aligned");
  const unsigned grid = static_cast<unsigned>(ceil_div64(ceil_div64(a->n, 4), 256));
  LaunchScope scope(kCatHeads, as_stream(stream), 0.0, 28.0 * a->n);
  const Betas k{static_cast<float>(a->beta2), static_cast<float>(1.0 - a->beta1), static_cast<float>(1.0 - a->beta2),
                static_cast<float>(a->eps),
                (a->decay_mul == 0.0) ? 1.f : static_cast<float>(a->decay_mul)};
  adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(*a, k);
  UPNERF_CHECK_LAUNCH("adam_kernel");
  return UPNERF_OK;
}
