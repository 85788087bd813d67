// GPU-resident training-ray batcher: PhototourismDataset.__getitem__
// for split "train" (datasets/phototourism.py:420-454) over a whole batch + default_collate, as one
// gather launch over tables that stay in HBM.
//
// One warp per ray.  The 4-tap feature interpolation streams four feat_dim-long pixel rows
// (contiguous, 16-byte vector loads, all four taps in flight before the first use) and writes one
// row; the small per-ray fields are copied by the first lanes.  Arithmetic is the reference's,
// operation by operation, with explicitly rounded multiplies and adds (no FMA contraction), so
// the result is bit-identical to the CPU path.
#include "common.h"

namespace upnerf {
namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float tap4(float w11, float p11, float w12, float p12, float w21, float p21,
                                      float w22, float p22) {
  // ((w11*p11 + w12*p12) + w21*p21) + w22*p22 with every product and sum rounded to fp32
  float r = __fadd_rn(__fmul_rn(w11, p11), __fmul_rn(w12, p12));
  r = __fadd_rn(r, __fmul_rn(w21, p21));
  return __fadd_rn(r, __fmul_rn(w22, p22));
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) ray_batch_gather_kernel(upnerf_ray_batch_args a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kWarpsPerBlock;
  for (int64_t r = warp0; r < a.n_rays; r += nwarps) {
    const int64_t i = a.idx[r];
    bool ok = i >= 0 && i < a.n_total;
    int64_t img = 0;
    if (ok) {
      img = static_cast<int64_t>(a.ray_infos[i * 3 + 2]);   // .long(): truncation toward zero
      ok = img >= 0 && img < a.n_images;
    }
    if (!ok) {
      if (lane == 0 && a.status) atomicOr(a.status, 1);
      continue;
    }
    // small fields: lanes 0..1 ray_infos, 2..4 directions, 5..7 rgbs, 8 inv_depth, 9 img_idx, 12..23 pose
    if (lane < 2) a.out_ray_infos[r * 2 + lane] = a.ray_infos[i * 3 + lane];
    else if (lane < 5) a.out_directions[r * 3 + lane - 2] = a.directions[i * 3 + lane - 2];
    else if (lane < 8) a.out_rgbs[r * 3 + lane - 5] = a.rgbs[i * 3 + lane - 5];
    else if (lane == 8) { if (a.inv_depths && a.out_inv_depths) a.out_inv_depths[r] = a.inv_depths[i]; }
    else if (lane == 9) a.out_img_idx[r] = img;
    else if (lane >= 12 && lane < 24) a.out_c2w[r * 12 + lane - 12] = a.poses[img * 12 + lane - 12];
    if (!a.feat_maps || !a.out_feats) continue;

    // datasets/phototourism.py:430-450
    const int h = a.feat_h;
    const float hm1 = static_cast<float>(h - 1);
    const float y = __fmul_rn(a.pxl_coords[i * 2 + 0], hm1);
    const float x = __fmul_rn(a.pxl_coords[i * 2 + 1], hm1);
    const long long y1 = static_cast<long long>(floorf(y));
    const long long x1 = static_cast<long long>(floorf(x));
    const long long y2 = y1 + 1 < h - 1 ? y1 + 1 : h - 1;
    const long long x2 = x1 + 1 < h - 1 ? x1 + 1 : h - 1;    // the reference clamps x with h as well (h == w)
    if (y1 < 0 || x1 < 0 || y1 >= a.feat_h || x1 >= a.feat_w) {
      if (lane == 0 && a.status) atomicOr(a.status, 1);
      continue;
    }
    const float fy1 = static_cast<float>(y1), fy2 = static_cast<float>(y2);
    const float fx1 = static_cast<float>(x1), fx2 = static_cast<float>(x2);
    const float w11 = __fmul_rn(__fsub_rn(fy2, y), __fsub_rn(fx2, x));
    const float w12 = __fmul_rn(__fsub_rn(fy2, y), __fsub_rn(x, fx1));
    const float w21 = __fmul_rn(__fsub_rn(y, fy1), __fsub_rn(fx2, x));
    const float w22 = __fmul_rn(__fsub_rn(y, fy1), __fsub_rn(x, fx1));
    const int F = a.feat_dim;
    const float* base = a.feat_maps + img * static_cast<int64_t>(a.feat_h) * a.feat_w * F;
    const float* p11 = base + (y1 * a.feat_w + x1) * F;
    const float* p12 = base + (y1 * a.feat_w + x2) * F;
    const float* p21 = base + (y2 * a.feat_w + x1) * F;
    const float* p22 = base + (y2 * a.feat_w + x2) * F;
    float* out = a.out_feats + r * F;
    if ((F & 3) == 0) {
      for (int c = lane * 4; c < F; c += 128) {
        const float4 a11 = ldg4(p11 + c), a12 = ldg4(p12 + c), a21 = ldg4(p21 + c), a22 = ldg4(p22 + c);
        float4 o;
        o.x = tap4(w11, a11.x, w12, a12.x, w21, a21.x, w22, a22.x);
        o.y = tap4(w11, a11.y, w12, a12.y, w21, a21.y, w22, a22.y);
        o.z = tap4(w11, a11.z, w12, a12.z, w21, a21.z, w22, a22.z);
        o.w = tap4(w11, a11.w, w12, a12.w, w21, a21.w, w22, a22.w);
        __stcs(reinterpret_cast<float4*>(out + c), o);
      }
    } else {
      for (int c = lane; c < F; c += 32)
        out[c] = tap4(w11, __ldg(p11 + c), w12, __ldg(p12 + c), w21, __ldg(p21 + c), w22, __ldg(p22 + c));
    }
  }
}

}  // namespace
}  // namespace upnerf

extern "C" int upnerf_ray_batch_gather(const upnerf_ray_batch_args* a, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(a != nullptr, UPNERF_ERR_BAD_SHAPE, "ray_batch: args missing");
  UPNERF_REQUIRE(upnerf_device_ok(), UPNERF_ERR_CUDA,
                 "upnerf_b200 needs a compute-capability 10.x GPU (sm_100a); there is no fallback");
  UPNERF_REQUIRE(a->n_rays >= 0 && a->n_total > 0 && a->n_images > 0, UPNERF_ERR_BAD_SHAPE,
                 "ray_batch: n_rays=%lld n_total=%lld n_images=%d", (long long)a->n_rays, (long long)a->n_total,
                 a->n_images);
  if (a->n_rays == 0) return UPNERF_OK;
  UPNERF_REQUIRE(a->idx && a->ray_infos && a->directions && a->rgbs && a->poses, UPNERF_ERR_BAD_SHAPE,
                 "ray_batch: idx / ray_infos / directions / rgbs / poses are required");
  UPNERF_REQUIRE(a->out_ray_infos && a->out_directions && a->out_img_idx && a->out_c2w && a->out_rgbs,
                 UPNERF_ERR_BAD_SHAPE, "ray_batch: an output buffer is missing");
  if (a->feat_maps) {
    UPNERF_REQUIRE(a->pxl_coords && a->out_feats, UPNERF_ERR_BAD_SHAPE,
                   "ray_batch: feat_maps without pxl_coords / out_feats");
    UPNERF_REQUIRE(a->feat_h >= 2 && a->feat_h == a->feat_w && a->feat_dim > 0, UPNERF_ERR_BAD_SHAPE,
                   "ray_batch: feature maps must be square (the reference asserts h == w), got %d x %d x %d",
                   a->feat_h, a->feat_w, a->feat_dim);
    UPNERF_REQUIRE((a->feat_dim & 3) != 0 ||
                       ((reinterpret_cast<uintptr_t>(a->feat_maps) | reinterpret_cast<uintptr_t>(a->out_feats)) & 15) == 0,
                   UPNERF_ERR_BAD_SHAPE, "ray_batch: feat_maps / out_feats must be 16-byte aligned");
  }
  const int64_t want = ceil_div64(a->n_rays, kWarpsPerBlock);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;     // 8 resident blocks of 8 warps per SM
  const int grid = static_cast<int>(want < cap ? want : cap);
  const double bytes = static_cast<double>(a->n_rays) *
                       ((a->feat_maps ? 5.0 * a->feat_dim * 4 + 8 : 0.0) + 8 + 2 * (12 + 12 + 8 + 48) + 12 + 8 + 8);
  LaunchScope scope(kCatSampling, as_stream(stream), 0.0, bytes);
  ray_batch_gather_kernel<<<grid, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(*a);
  UPNERF_CHECK_LAUNCH("ray_batch_gather_kernel");
  return UPNERF_OK;
}
