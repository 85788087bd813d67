// SE(3) pose refinement + ray casting, forward and backward (subsystem (a)).
//
//   forward : wu = se3_table[img_idx]  ->  [R_ref | t_ref] = exp(wu)      (utils/camera.py:87-98)
//             [R | t] = c2w o refine    (R = R_c R_ref, t = R_c t_ref + t_c) (utils/camera.py:51-58)
//             d = normalize(R dir), o = t                                 (utils/ray.py:44-56 / 57-65)
//             rays = [o, d, near, far]                                    (models/nerf_system.py:166)
//   backward: d(rays_o, rays_d) -> d se3_table (N_img, 6), atomically summed per image.
//
// The reference evaluates sin(t)/t, (1-cos t)/t^2, (t-sin t)/t^3 as order-10 Taylor sums
// (utils/camera.py:126-152); they are even in t, so here they are polynomials in s = |w|^2,
// which makes the gradient at w = 0 finite without special-casing.  One thread per ray.
#include "common.h"

namespace upnerf {
namespace {

struct Taylor {
  float A, B, C;     // values
  float dA, dB, dC;  // derivatives with respect to s = theta^2
};

__device__ __forceinline__ Taylor taylor_abc(float s) {
  // denominators (2i+1)!, (2i+2)!, (2i+3)! built the way the reference does (running product)
  Taylor t;
  t.A = t.B = t.C = 0.f;
  t.dA = t.dB = t.dC = 0.f;
  float dena = 1.f, denb = 1.f, denc = 1.f;
  float p = 1.f;       // s^i
  float pm1 = 0.f;     // s^(i-1) (0 for i = 0)
  float sign = 1.f;
#pragma unroll
  for (int i = 0; i <= 10; ++i) {
    if (i > 0) dena *= static_cast<float>((2 * i) * (2 * i + 1));
    denb *= static_cast<float>((2 * i + 1) * (2 * i + 2));
    denc *= static_cast<float>((2 * i + 2) * (2 * i + 3));
    t.A += sign * p / dena;
    t.B += sign * p / denb;
    t.C += sign * p / denc;
    const float fi = static_cast<float>(i);
    t.dA += sign * fi * pm1 / dena;
    t.dB += sign * fi * pm1 / denb;
    t.dC += sign * fi * pm1 / denc;
    pm1 = p;
    p *= s;
    sign = -sign;
  }
  return t;
}

struct Mat3 {
  float m[3][3];
};

__device__ __forceinline__ Mat3 matmul(const Mat3& a, const Mat3& b) {
  Mat3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return c;
}

__device__ __forceinline__ void exp_map(const float w[3], const float u[3], const Taylor& t,
                                        Mat3& K, Mat3& K2, Mat3& Rm, Mat3& V, float tref[3]) {
  K.m[0][0] = 0.f;   K.m[0][1] = -w[2]; K.m[0][2] = w[1];
  K.m[1][0] = w[2];  K.m[1][1] = 0.f;   K.m[1][2] = -w[0];
  K.m[2][0] = -w[1]; K.m[2][1] = w[0];  K.m[2][2] = 0.f;
  K2 = matmul(K, K);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float e = (i == j) ? 1.f : 0.f;
      Rm.m[i][j] = e + t.A * K.m[i][j] + t.B * K2.m[i][j];
      V.m[i][j] = e + t.B * K.m[i][j] + t.C * K2.m[i][j];
    }
#pragma unroll
  for (int i = 0; i < 3; ++i) tref[i] = V.m[i][0] * u[0] + V.m[i][1] * u[1] + V.m[i][2] * u[2];
}

struct PoseArgs {
  const float* table;     // [n_img, 6] or nullptr (no refinement)
  const int64_t* img_idx; // [R]
  const float* c2w;       // [R,3,4] or [3,4] (c2w_stride = 0)
  int64_t c2w_stride;     // 12 or 0
  const float* dirs;      // [R,3]
  const float* near_far;  // [R,2] or nullptr
  int64_t R;
};

__device__ __forceinline__ void load_pose(const PoseArgs& a, int64_t r, float w[3], float u[3],
                                          Mat3& Rc, float tc[3]) {
  if (a.table) {
    const float* row = a.table + a.img_idx[r] * 6;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      w[i] = row[i];
      u[i] = row[3 + i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = u[i] = 0.f;
  }
  const float* c = a.c2w + r * a.c2w_stride;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) Rc.m[i][j] = c[i * 4 + j];
    tc[i] = c[i * 4 + 3];
  }
}

__global__ void pose_rays_fwd_kernel(PoseArgs a, float* __restrict__ rays, float* __restrict__ pose_out) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= a.R) return;
  float w[3], u[3], tc[3];
  Mat3 Rc;
  load_pose(a, r, w, u, Rc, tc);
  Mat3 Rf;
  float tf[3];
  if (a.table) {
    const float s = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const Taylor t = taylor_abc(s);
    Mat3 K, K2, Rr, V;
    float tref[3];
    exp_map(w, u, t, K, K2, Rr, V, tref);
    Rf = matmul(Rc, Rr);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      tf[i] = Rc.m[i][0] * tref[0] + Rc.m[i][1] * tref[1] + Rc.m[i][2] * tref[2] + tc[i];
  } else {
    Rf = Rc;
#pragma unroll
    for (int i = 0; i < 3; ++i) tf[i] = tc[i];
  }
  const float dx = a.dirs[r * 3 + 0], dy = a.dirs[r * 3 + 1], dz = a.dirs[r * 3 + 2];
  float d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = Rf.m[i][0] * dx + Rf.m[i][1] * dy + Rf.m[i][2] * dz;
  const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  float* out = rays + r * 8;
  out[0] = tf[0]; out[1] = tf[1]; out[2] = tf[2];
  out[3] = d[0] / nrm; out[4] = d[1] / nrm; out[5] = d[2] / nrm;
  if (a.near_far) {
    out[6] = a.near_far[r * 2];
    out[7] = a.near_far[r * 2 + 1];
  }
  if (pose_out) {
    float* p = pose_out + r * 12;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) p[i * 4 + j] = Rf.m[i][j];
      p[i * 4 + 3] = tf[i];
    }
  }
}

__global__ void pose_rays_bwd_kernel(PoseArgs a, const float* __restrict__ d_rays,
                                     float* __restrict__ d_table) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= a.R) return;
  float w[3], u[3], tc[3];
  Mat3 Rc;
  load_pose(a, r, w, u, Rc, tc);
  const float s = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const Taylor t = taylor_abc(s);
  Mat3 K, K2, Rr, V;
  float tref[3];
  exp_map(w, u, t, K, K2, Rr, V, tref);
  const Mat3 Rf = matmul(Rc, Rr);
  const float dir[3] = {a.dirs[r * 3 + 0], a.dirs[r * 3 + 1], a.dirs[r * 3 + 2]};
  float draw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) draw[i] = Rf.m[i][0] * dir[0] + Rf.m[i][1] * dir[1] + Rf.m[i][2] * dir[2];
  const float nrm = sqrtf(draw[0] * draw[0] + draw[1] * draw[1] + draw[2] * draw[2]);
  const float dn[3] = {draw[0] / nrm, draw[1] / nrm, draw[2] / nrm};
  const float* g = d_rays + r * 8;
  const float go[3] = {g[0], g[1], g[2]};
  const float gd[3] = {g[3], g[4], g[5]};
  // through the normalisation
  const float dot = dn[0] * gd[0] + dn[1] * gd[1] + dn[2] * gd[2];
  float graw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) graw[i] = (gd[i] - dn[i] * dot) / nrm;
  // d/dR_ref = R_c^T (graw (x) dir),  d/dt_ref = R_c^T go
  float gq[3], gt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gq[i] = Rc.m[0][i] * graw[0] + Rc.m[1][i] * graw[1] + Rc.m[2][i] * graw[2];
    gt[i] = Rc.m[0][i] * go[0] + Rc.m[1][i] * go[1] + Rc.m[2][i] * go[2];
  }
  Mat3 gR, gV;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      gR.m[i][j] = gq[i] * dir[j];
      gV.m[i][j] = gt[i] * u[j];
    }
  float gu[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) gu[j] = V.m[0][j] * gt[0] + V.m[1][j] * gt[1] + V.m[2][j] * gt[2];
  // scalar coefficient gradients and dK
  float gA = 0.f, gB = 0.f, gC = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      gA += gR.m[i][j] * K.m[i][j];
      gB += gR.m[i][j] * K2.m[i][j] + gV.m[i][j] * K.m[i][j];
      gC += gV.m[i][j] * K2.m[i][j];
    }
  // G2 = dL/dK2 = B gR + C gV ;  dL/dK = A gR + B gV + G2 K^T + K^T G2
  Mat3 G2, gK;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) G2.m[i][j] = t.B * gR.m[i][j] + t.C * gV.m[i][j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float acc = t.A * gR.m[i][j] + t.B * gV.m[i][j];
#pragma unroll
      for (int k = 0; k < 3; ++k) acc += G2.m[i][k] * K.m[j][k] + K.m[k][i] * G2.m[k][j];
      gK.m[i][j] = acc;
    }
  const float gs = gA * t.dA + gB * t.dB + gC * t.dC;
  float gw[3];
  gw[0] = 2.f * gs * w[0] + gK.m[2][1] - gK.m[1][2];
  gw[1] = 2.f * gs * w[1] + gK.m[0][2] - gK.m[2][0];
  gw[2] = 2.f * gs * w[2] + gK.m[1][0] - gK.m[0][1];
  float* out = d_table + a.img_idx[r] * 6;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    atomicAdd(out + i, gw[i]);
    atomicAdd(out + 3 + i, gu[i]);
  }
}


// ---- stand-alone pieces with the reference's granularity (drop-in for lie.se3_to_SE3,
//      pose.compose_pair and get_rays on poses that carry their own gradient) ----

__global__ void se3_exp_fwd_kernel(const float* __restrict__ wu, int64_t N, float* __restrict__ out) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= N) return;
  const float w[3] = {wu[r * 6], wu[r * 6 + 1], wu[r * 6 + 2]};
  const float u[3] = {wu[r * 6 + 3], wu[r * 6 + 4], wu[r * 6 + 5]};
  const Taylor t = taylor_abc(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  Mat3 K, K2, Rr, V;
  float tref[3];
  exp_map(w, u, t, K, K2, Rr, V, tref);
  float* p = out + r * 12;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) p[i * 4 + j] = Rr.m[i][j];
    p[i * 4 + 3] = tref[i];
  }
}

// gradient of exp(wu) given dL/d[R|t] (3x4)
__device__ __forceinline__ void exp_map_bwd(const float w[3], const float u[3], const Mat3& gR,
                                            const float gt[3], float gw[3], float gu[3]) {
  const float s = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const Taylor t = taylor_abc(s);
  Mat3 K, K2, Rr, V;
  float tref[3];
  exp_map(w, u, t, K, K2, Rr, V, tref);
  Mat3 gV;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) gV.m[i][j] = gt[i] * u[j];
#pragma unroll
  for (int j = 0; j < 3; ++j) gu[j] = V.m[0][j] * gt[0] + V.m[1][j] * gt[1] + V.m[2][j] * gt[2];
  float gA = 0.f, gB = 0.f, gC = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      gA += gR.m[i][j] * K.m[i][j];
      gB += gR.m[i][j] * K2.m[i][j] + gV.m[i][j] * K.m[i][j];
      gC += gV.m[i][j] * K2.m[i][j];
    }
  Mat3 G2, gK;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) G2.m[i][j] = t.B * gR.m[i][j] + t.C * gV.m[i][j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float acc = t.A * gR.m[i][j] + t.B * gV.m[i][j];
#pragma unroll
      for (int k = 0; k < 3; ++k) acc += G2.m[i][k] * K.m[j][k] + K.m[k][i] * G2.m[k][j];
      gK.m[i][j] = acc;
    }
  const float gs = gA * t.dA + gB * t.dB + gC * t.dC;
  gw[0] = 2.f * gs * w[0] + gK.m[2][1] - gK.m[1][2];
  gw[1] = 2.f * gs * w[1] + gK.m[0][2] - gK.m[2][0];
  gw[2] = 2.f * gs * w[2] + gK.m[1][0] - gK.m[0][1];
}

__global__ void se3_exp_bwd_kernel(const float* __restrict__ wu, const float* __restrict__ d_pose,
                                   int64_t N, float* __restrict__ d_wu) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= N) return;
  const float w[3] = {wu[r * 6], wu[r * 6 + 1], wu[r * 6 + 2]};
  const float u[3] = {wu[r * 6 + 3], wu[r * 6 + 4], wu[r * 6 + 5]};
  Mat3 gR;
  float gt[3], gw[3], gu[3];
  const float* g = d_pose + r * 12;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) gR.m[i][j] = g[i * 4 + j];
    gt[i] = g[i * 4 + 3];
  }
  exp_map_bwd(w, u, gR, gt, gw, gu);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d_wu[r * 6 + i] = gw[i];
    d_wu[r * 6 + 3 + i] = gu[i];
  }
}

// out = b o a :  R = R_b R_a, t = R_b t_a + t_b.  stride 0 broadcasts a single (3,4) pose.
__global__ void pose_compose_fwd_kernel(const float* __restrict__ a, int64_t sa, const float* __restrict__ b,
                                        int64_t sb, int64_t N, float* __restrict__ out) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= N) return;
  const float* pa = a + r * sa;
  const float* pb = b + r * sb;
  float* o = out + r * 12;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = pb[i * 4 + 0] * pa[0 * 4 + j] + pb[i * 4 + 1] * pa[1 * 4 + j] + pb[i * 4 + 2] * pa[2 * 4 + j];
      if (j == 3) acc += pb[i * 4 + 3];
      o[i * 4 + j] = acc;
    }
  }
}

__global__ void pose_compose_bwd_kernel(const float* __restrict__ a, int64_t sa, const float* __restrict__ b,
                                        int64_t sb, const float* __restrict__ g, int64_t N,
                                        float* __restrict__ da, float* __restrict__ db) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= N) return;
  const float* pa = a + r * sa;
  const float* pb = b + r * sb;
  const float* go = g + r * 12;
  if (da) {  // d a[k][j] = sum_i R_b[i][k] g[i][j]
    float* o = da + r * 12;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[k * 4 + j] = pb[0 * 4 + k] * go[0 * 4 + j] + pb[1 * 4 + k] * go[1 * 4 + j] + pb[2 * 4 + k] * go[2 * 4 + j];
  }
  if (db) {  // d R_b[i][k] = sum_j g[i][j] a[k][j] (j over 4 incl. translation), d t_b = g[:,3]
    float* o = db + r * 12;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        o[i * 4 + k] = go[i * 4 + 0] * pa[k * 4 + 0] + go[i * 4 + 1] * pa[k * 4 + 1] +
                       go[i * 4 + 2] * pa[k * 4 + 2] + go[i * 4 + 3] * pa[k * 4 + 3];
      o[i * 4 + 3] = go[i * 4 + 3];
    }
  }
}

// gradient of get_rays with respect to the pose(s): d c2w[:, :3] = graw (x) dir, d c2w[:, 3] = d o
__global__ void get_rays_bwd_kernel(PoseArgs a, const float* __restrict__ d_rays, float* __restrict__ d_c2w) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= a.R) return;
  const float* c = a.c2w + r * a.c2w_stride;
  const float dir[3] = {a.dirs[r * 3 + 0], a.dirs[r * 3 + 1], a.dirs[r * 3 + 2]};
  float draw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) draw[i] = c[i * 4] * dir[0] + c[i * 4 + 1] * dir[1] + c[i * 4 + 2] * dir[2];
  const float nrm = sqrtf(draw[0] * draw[0] + draw[1] * draw[1] + draw[2] * draw[2]);
  const float dn[3] = {draw[0] / nrm, draw[1] / nrm, draw[2] / nrm};
  const float* g = d_rays + r * 8;
  const float dot = dn[0] * g[3] + dn[1] * g[4] + dn[2] * g[5];
  float* o = d_c2w + (a.c2w_stride ? r * 12 : 0);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float graw = (g[3 + i] - dn[i] * dot) / nrm;
    if (a.c2w_stride) {
      o[i * 4 + 0] = graw * dir[0];
      o[i * 4 + 1] = graw * dir[1];
      o[i * 4 + 2] = graw * dir[2];
      o[i * 4 + 3] = g[i];
    } else {  // one shared pose: reduce over rays
      atomicAdd(o + i * 4 + 0, graw * dir[0]);
      atomicAdd(o + i * 4 + 1, graw * dir[1]);
      atomicAdd(o + i * 4 + 2, graw * dir[2]);
      atomicAdd(o + i * 4 + 3, g[i]);
    }
  }
}

}  // namespace
}  // namespace upnerf

extern "C" {

int upnerf_pose_rays_fwd(const float* se3_table, const int64_t* img_idx, const float* c2w,
                         int c2w_is_single, const float* directions, const float* near_far,
                         int64_t n_rays, float* rays, float* pose_out, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n_rays > 0, UPNERF_ERR_BAD_SHAPE, "pose_rays_fwd: n_rays=%lld", (long long)n_rays);
  UPNERF_REQUIRE(!se3_table || img_idx, UPNERF_ERR_BAD_SHAPE, "pose_rays_fwd: table without img_idx");
  PoseArgs a{se3_table, img_idx, c2w, c2w_is_single ? 0 : 12, directions, near_far, n_rays};
  const int threads = 128;
  LaunchScope scope(kCatPoseRays, as_stream(stream));
  pose_rays_fwd_kernel<<<static_cast<unsigned>(ceil_div64(n_rays, threads)), threads, 0,
                         as_stream(stream)>>>(a, rays, pose_out);
  UPNERF_CHECK_LAUNCH("pose_rays_fwd_kernel");
  return UPNERF_OK;
}

int upnerf_pose_rays_bwd(const float* se3_table, const int64_t* img_idx, const float* c2w,
                         int c2w_is_single, const float* directions, int64_t n_rays,
                         const float* d_rays, float* d_se3_table, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n_rays > 0 && se3_table && img_idx, UPNERF_ERR_BAD_SHAPE, "pose_rays_bwd: bad args");
  PoseArgs a{se3_table, img_idx, c2w, c2w_is_single ? 0 : 12, directions, nullptr, n_rays};
  const int threads = 128;
  LaunchScope scope(kCatPoseRays, as_stream(stream));
  pose_rays_bwd_kernel<<<static_cast<unsigned>(ceil_div64(n_rays, threads)), threads, 0,
                         as_stream(stream)>>>(a, d_rays, d_se3_table);
  UPNERF_CHECK_LAUNCH("pose_rays_bwd_kernel");
  return UPNERF_OK;
}

int upnerf_se3_exp_fwd(const float* wu, int64_t n, float* pose_out, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n > 0, UPNERF_ERR_BAD_SHAPE, "se3_exp_fwd: n=%lld", (long long)n);
  LaunchScope scope(kCatPoseRays, as_stream(stream));
  se3_exp_fwd_kernel<<<static_cast<unsigned>(ceil_div64(n, 128)), 128, 0, as_stream(stream)>>>(wu, n, pose_out);
  UPNERF_CHECK_LAUNCH("se3_exp_fwd_kernel");
  return UPNERF_OK;
}

int upnerf_se3_exp_bwd(const float* wu, const float* d_pose, int64_t n, float* d_wu, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n > 0, UPNERF_ERR_BAD_SHAPE, "se3_exp_bwd: n=%lld", (long long)n);
  LaunchScope scope(kCatPoseRays, as_stream(stream));
  se3_exp_bwd_kernel<<<static_cast<unsigned>(ceil_div64(n, 128)), 128, 0, as_stream(stream)>>>(wu, d_pose, n, d_wu);
  UPNERF_CHECK_LAUNCH("se3_exp_bwd_kernel");
  return UPNERF_OK;
}

int upnerf_pose_compose_fwd(const float* pose_a, int a_is_single, const float* pose_b, int b_is_single,
                            int64_t n, float* out, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n > 0, UPNERF_ERR_BAD_SHAPE, "pose_compose_fwd: n=%lld", (long long)n);
  LaunchScope scope(kCatPoseRays, as_stream(stream));
  pose_compose_fwd_kernel<<<static_cast<unsigned>(ceil_div64(n, 128)), 128, 0, as_stream(stream)>>>(
      pose_a, a_is_single ? 0 : 12, pose_b, b_is_single ? 0 : 12, n, out);
  UPNERF_CHECK_LAUNCH("pose_compose_fwd_kernel");
  return UPNERF_OK;
}

int upnerf_pose_compose_bwd(const float* pose_a, int a_is_single, const float* pose_b, int b_is_single,
                            const float* d_out, int64_t n, float* d_a, float* d_b, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n > 0, UPNERF_ERR_BAD_SHAPE, "pose_compose_bwd: n=%lld", (long long)n);
  LaunchScope scope(kCatPoseRays, as_stream(stream));
  pose_compose_bwd_kernel<<<static_cast<unsigned>(ceil_div64(n, 128)), 128, 0, as_stream(stream)>>>(
      pose_a, a_is_single ? 0 : 12, pose_b, b_is_single ? 0 : 12, d_out, n, d_a, d_b);
  UPNERF_CHECK_LAUNCH("pose_compose_bwd_kernel");
  return UPNERF_OK;
}

int upnerf_get_rays_bwd(const float* c2w, int c2w_is_single, const float* directions, int64_t n_rays,
                        const float* d_rays, float* d_c2w, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(n_rays > 0, UPNERF_ERR_BAD_SHAPE, "get_rays_bwd: n_rays=%lld", (long long)n_rays);
  PoseArgs a{nullptr, nullptr, c2w, c2w_is_single ? 0 : 12, directions, nullptr, n_rays};
  LaunchScope scope(kCatPoseRays, as_stream(stream));
  get_rays_bwd_kernel<<<static_cast<unsigned>(ceil_div64(n_rays, 128)), 128, 0, as_stream(stream)>>>(a, d_rays, d_c2w);
  UPNERF_CHECK_LAUNCH("get_rays_bwd_kernel");
  return UPNERF_OK;
}

}  // extern "C"
