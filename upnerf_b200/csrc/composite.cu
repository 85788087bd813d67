// Alpha compositing of static + candidate density along rays, forward and backward
// (subsystem (c); reference models/rendering.py:124-219).  One warp per ray.
//
//   delta_i = z_{i+1} - z_i (last = 1e2);  a^s = 1-exp(-delta s_sigma), a^c likewise,
//   a = 1-exp(-delta (s_sigma + c_sigma));  T = excl. cumprod(1-a),  T^s = excl. cumprod(1-a^s)
//   candidate pass (sched_mult < 1, candidate head on):
//       c_weights = a T, c_depth = sum a T z, t_weight = sum a^c T,
//       feat      = sum a^s T s_feat + sum a^c T c_feat
//   static pass: s_weights = a^s T^s, s_rgb = sum s_weights s_rgb, s_depth = sum s_weights z
//
// The 384-d feature heads are linear (models/nerf.py:53,76), so instead of compositing
// per-sample 384-d features this kernel composites the hidden vectors feeding them
// (hF: 256-d, g2: 128-d) and the per-ray weight sums; the projection is applied once per
// ray afterwards:  feat = W_sf (sum w hF) + b_sf sum w + W_cf (sum w' g2) + b_cf sum w'.
//
// Backward is the division-free suffix-sum form: with e_k = sum of (weight_k * dL/dweight_k)
// over the weight families sharing a transmittance and E_i = sum_{k>i} e_k,
//   dL/dsigma_i = delta_i * ( (1-alpha_i) T_i dL/dalpha-terms - E_i ),
// so alpha == 1 (which the 1e2 last delta produces on every ray) needs no special case.
#include <cuda_bf16.h>

#include "common.h"

namespace upnerf {
namespace {

constexpr int kWarps = 4;
constexpr int kMaxS = 256;
constexpr int kHF = 256;   // hidden width W
constexpr int kG2 = 128;   // W / 2
constexpr int kBatch = 4;  // rows whose loads are in flight together per warp

using CompArgs = upnerf_composite_args;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inclusive product scan over the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
// inclusive suffix sum over the warp (lane i gets sum of lanes >= i)
__device__ __forceinline__ float warp_suffix_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += n;
  }
  return v;
}

template <typename T> struct Vec;
template <> struct Vec<float> {
  // lane reads NC consecutive fp32 columns
  template <int NC> static __device__ __forceinline__ void load(const float* row, int lane, float (&v)[NC]) {
    const float4* p = reinterpret_cast<const float4*>(row + lane * NC);
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) {
      const float4 q = __ldg(p + i);
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  }
  template <int NC> static __device__ __forceinline__ void store(float* row, int lane, const float (&v)[NC]) {
    float4* p = reinterpret_cast<float4*>(row + lane * NC);
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  template <int NC> static __device__ __forceinline__ void load(const __nv_bfloat16* row, int lane, float (&v)[NC]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(row + lane * NC);
    if constexpr (NC == 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
      const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(b[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    } else {
      const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
      const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int i = 0; i < 2; ++i) { const float2 f = __bfloat1622float2(b[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
  }
  template <int NC> static __device__ __forceinline__ void store(__nv_bfloat16* row, int lane, const float (&v)[NC]) {
    if constexpr (NC == 8) {
      uint4 q;
      __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      *reinterpret_cast<uint4*>(row + lane * NC) = q;
    } else {
      uint2 q;
      __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
      for (int i = 0; i < 2; ++i) b[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      *reinterpret_cast<uint2*>(row + lane * NC) = q;
    }
  }
};

// Per-sample quantities of one 32-sample chunk, computed identically in fwd and bwd.
struct Sample {
  float z, delta, as, ac, a;  // alphas
};

__device__ __forceinline__ Sample load_sample(const CompArgs& p, int64_t r, int i, bool valid) {
  Sample s;
  s.z = 0.f; s.delta = 0.f; s.as = 0.f; s.ac = 0.f; s.a = 0.f;
  if (!valid) return s;
  const int64_t m = r * p.S + i;
  s.z = p.z[m];
  s.delta = (i == p.S - 1) ? 1e2f : p.z[m + 1] - s.z;
  const float ss = p.s_sigma[m];
  s.as = 1.f - expf(__fmul_rn(-s.delta, ss));
  if (p.cand) {
    const float cs = p.c_sigma[m];
    s.ac = 1.f - expf(__fmul_rn(-s.delta, cs));
    s.a = 1.f - expf(__fmul_rn(-s.delta, __fadd_rn(ss, cs)));
  }
  return s;
}

template <typename T>
// (7 blocks per SM: 4096 rays = 1024 blocks fit ONE wave of 148 x 7; at 6 blocks per SM the 136 blocks of a second
//  wave doubled the kernel's time at the training batch size)
__global__ void __launch_bounds__(kWarps * 32, 7) composite_fwd_kernel(const CompArgs p) {
  __shared__ float s_ws[kWarps][kMaxS];
  __shared__ float s_wc[kWarps][kMaxS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(kWarps) + warp;
  if (r >= p.R) return;
  const int S = p.S;
  float carryT = 1.f, carryTs = 1.f;
  float acc_cdepth = 0.f, acc_tw = 0.f, acc_sdepth = 0.f, acc_ws = 0.f;
  float acc_rgb[3] = {0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < S; c0 += 32) {
    const int i = c0 + lane;
    const bool valid = i < S;
    const Sample s = load_sample(p, r, i, valid);
    // static transmittance
    const float incs = warp_scan_mul(1.f - s.as, lane);
    float exs = __shfl_up_sync(0xffffffffu, incs, 1);
    if (lane == 0) exs = 1.f;
    const float Ts = carryTs * exs;
    carryTs *= __shfl_sync(0xffffffffu, incs, 31);
    const float wstat = s.as * Ts;
    float ws = wstat, wc = 0.f;
    if (p.cand) {
      const float inc = warp_scan_mul(1.f - s.a, lane);
      float ex = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) ex = 1.f;
      const float Tc = carryT * ex;
      carryT *= __shfl_sync(0xffffffffu, inc, 31);
      const float wcomb = s.a * Tc;
      ws = s.as * Tc;
      wc = s.ac * Tc;
      if (valid) {
        p.c_weights[r * S + i] = wcomb;
        acc_cdepth += wcomb * s.z;
        acc_tw += wc;
      }
    }
    if (valid) {
      if (p.s_weights) p.s_weights[r * S + i] = wstat;
      acc_sdepth += wstat * s.z;
      if (p.stat_rgb) {
        const float* c = p.rgb + (r * S + i) * 3;
        acc_rgb[0] += wstat * c[0];
        acc_rgb[1] += wstat * c[1];
        acc_rgb[2] += wstat * c[2];
      }
      acc_ws += ws;
      s_ws[warp][i] = ws;
      s_wc[warp][i] = wc;
    }
  }
  acc_sdepth = warp_sum(acc_sdepth);
  if (lane == 0) p.s_depth[r] = acc_sdepth;
  if (p.cand) {
    acc_cdepth = warp_sum(acc_cdepth);
    acc_tw = warp_sum(acc_tw);
    if (lane == 0) {
      p.c_depth[r] = acc_cdepth;
      p.t_weight[r] = acc_tw;
    }
  }
  if (p.stat_rgb) {
#pragma unroll
    for (int c = 0; c < 3; ++c) acc_rgb[c] = warp_sum(acc_rgb[c]);
    if (lane == 0) {
      p.s_rgb[r * 3] = acc_rgb[0];
      p.s_rgb[r * 3 + 1] = acc_rgb[1];
      p.s_rgb[r * 3 + 2] = acc_rgb[2];
    }
  }
  if (p.feat_mode == 0) return;
  __syncwarp();
  acc_ws = warp_sum(acc_ws);
  if (lane == 0) {
    p.ws_sum[r] = acc_ws;
    if (p.cand) p.wc_sum[r] = acc_tw;
  }
  float ah[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float ag[4] = {0.f, 0.f, 0.f, 0.f};
  const T* hf = static_cast<const T*>(p.hf) + r * S * p.ld_hf;
  const T* g2 = p.cand ? static_cast<const T*>(p.g2) + r * S * p.ld_g2 : nullptr;
  // kBatch rows per step, every load issued before the first use: at a few thousand rays the
  // kernel is bound by bytes in flight per warp, not by bandwidth
  for (int i0 = 0; i0 < S; i0 += kBatch) {
    float v[kBatch][8], g[kBatch][4];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int i = min(i0 + j, S - 1);
      Vec<T>::template load<8>(hf + i * p.ld_hf, lane, v[j]);
      if (p.cand) Vec<T>::template load<4>(g2 + i * p.ld_g2, lane, g[j]);
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const bool ok = i0 + j < S;
      const float ws = ok ? s_ws[warp][i0 + j] : 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) ah[e] = fmaf(ws, v[j][e], ah[e]);
      if (p.cand) {
        const float wc = ok ? s_wc[warp][i0 + j] : 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) ag[e] = fmaf(wc, g[j][e], ag[e]);
      }
    }
  }
  Vec<float>::store<8>(p.hf_ray + r * kHF, lane, ah);
  if (p.cand) Vec<float>::store<4>(p.g2_ray + r * kG2, lane, ag);
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32, 7) composite_bwd_kernel(const CompArgs p) {
  // per-warp scratch: weights, per-sample dots, candidate pre-activation gradient
  __shared__ float s_ws[kWarps][kMaxS];
  __shared__ float s_wc[kWarps][kMaxS];
  __shared__ float s_dh[kWarps][kMaxS];
  __shared__ float s_dg[kWarps][kMaxS];
  __shared__ float s_T[kWarps][kMaxS];
  __shared__ float s_Ts[kWarps][kMaxS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(kWarps) + warp;
  if (r >= p.R) return;
  const int S = p.S;
  const bool feat = p.feat_mode != 0;
  const T* hf = feat ? static_cast<const T*>(p.hf) + r * S * p.ld_hf : nullptr;
  const T* g2 = (feat && p.cand) ? static_cast<const T*>(p.g2) + r * S * p.ld_g2 : nullptr;

  // ---- pass 1: per-sample dots  hF_i . g_hf_ray  and  g2_i . g_g2_ray
  float gh[8], gg[4];
  if (feat) {
    Vec<float>::load<8>(p.g_hf_ray + r * kHF, lane, gh);
    if (p.cand) Vec<float>::load<4>(p.g_g2_ray + r * kG2, lane, gg);
    for (int i0 = 0; i0 < S; i0 += kBatch) {
      float v[kBatch][8], g[kBatch][4];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int i = min(i0 + j, S - 1);
        Vec<T>::template load<8>(hf + i * p.ld_hf, lane, v[j]);
        if (p.cand) Vec<T>::template load<4>(g2 + i * p.ld_g2, lane, g[j]);
      }
      float d[kBatch], dg[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        d[j] = 0.f;
        dg[j] = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) d[j] = fmaf(v[j][e], gh[e], d[j]);
        if (p.cand) {
#pragma unroll
          for (int e = 0; e < 4; ++e) dg[j] = fmaf(g[j][e], gg[e], dg[j]);
        }
      }
      // kBatch independent butterfly reductions (interleaved by the unroll)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
          if (p.cand) dg[j] += __shfl_xor_sync(0xffffffffu, dg[j], o);
        }
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        if (lane == j && i0 + j < S) {
          s_dh[warp][i0 + j] = d[j];
          s_dg[warp][i0 + j] = dg[j];
        }
      }
    }
    __syncwarp();
  }

  // ---- pass 2: transmittances forward (kept in registers per chunk), then gradients
  // chunks are revisited in reverse for the suffix sums; per-sample T values go to smem.
  {
    float carryT = 1.f, carryTs = 1.f;
    for (int c0 = 0; c0 < S; c0 += 32) {
      const int i = c0 + lane;
      const Sample s = load_sample(p, r, i, i < S);
      const float incs = warp_scan_mul(1.f - s.as, lane);
      float exs = __shfl_up_sync(0xffffffffu, incs, 1);
      if (lane == 0) exs = 1.f;
      const float ts_val = carryTs * exs;
      carryTs *= __shfl_sync(0xffffffffu, incs, 31);
      float t_val = 1.f;
      if (p.cand) {
        const float inc = warp_scan_mul(1.f - s.a, lane);
        float ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = 1.f;
        t_val = carryT * ex;
        carryT *= __shfl_sync(0xffffffffu, inc, 31);
      }
      if (i < S) {
        s_T[warp][i] = t_val;
        s_Ts[warp][i] = ts_val;
      }
    }
  }
  const float g_cdepth = (p.cand && p.g_c_depth) ? p.g_c_depth[r] : 0.f;
  const float g_tw = (p.cand && p.g_t_weight) ? p.g_t_weight[r] : 0.f;
  const float g_sdepth = p.g_s_depth ? p.g_s_depth[r] : 0.f;
  const float g_ws = (feat && p.g_ws_sum) ? p.g_ws_sum[r] : 0.f;
  const float g_wc = (feat && p.cand && p.g_wc_sum) ? p.g_wc_sum[r] : 0.f;
  float g_rgb[3] = {0.f, 0.f, 0.f};
  if (p.stat_rgb && p.g_s_rgb) {
    g_rgb[0] = p.g_s_rgb[r * 3];
    g_rgb[1] = p.g_s_rgb[r * 3 + 1];
    g_rgb[2] = p.g_s_rgb[r * 3 + 2];
  }
  float carryE = 0.f, carryEs = 0.f;  // suffix sums over later chunks
  const int nchunks = (S + 31) / 32;
  for (int c = nchunks - 1; c >= 0; --c) {
    const int i = c * 32 + lane;
    const bool valid = i < S;
    const Sample s = load_sample(p, r, i, valid);
    const int64_t m = r * S + i;
    const float Ts = valid ? s_Ts[warp][i] : 0.f, Tc = valid ? s_T[warp][i] : 0.f;
    const float wstat = s.as * Ts;
    // gradient w.r.t. the static-pass weight
    float H = g_sdepth * s.z;
    float rgbv[3] = {0.f, 0.f, 0.f};
    if (valid) {
      if (p.g_s_weights) H += p.g_s_weights[m];
      if (p.stat_rgb) {
        rgbv[0] = p.rgb[m * 3]; rgbv[1] = p.rgb[m * 3 + 1]; rgbv[2] = p.rgb[m * 3 + 2];
        H += rgbv[0] * g_rgb[0] + rgbv[1] * g_rgb[1] + rgbv[2] * g_rgb[2];
      }
      if (p.feat_mode == 1) H += s_dh[warp][i] + g_ws;
    }
    float es = valid ? wstat * H : 0.f;
    const float incs = warp_suffix_sum(es, lane);
    const float Es = carryEs + incs - es;          // exclusive: later samples only
    carryEs += __shfl_sync(0xffffffffu, incs, 0);
    float d_ss = s.delta * ((1.f - s.as) * Ts * H - Es);
    float d_cs = 0.f;
    float ws = wstat, wc = 0.f;
    if (p.cand) {
      float G1 = g_cdepth * s.z, G2v = 0.f, G3 = g_tw;
      if (valid) {
        if (p.g_c_weights) G1 += p.g_c_weights[m];
        if (p.feat_mode == 2) {
          G2v = s_dh[warp][i] + g_ws;
          G3 += s_dg[warp][i] + g_wc;
        }
      }
      ws = s.as * Tc;
      wc = s.ac * Tc;
      const float e = valid ? (s.a * Tc * G1 + ws * G2v + wc * G3) : 0.f;
      const float inc = warp_suffix_sum(e, lane);
      const float E = carryE + inc - e;
      carryE += __shfl_sync(0xffffffffu, inc, 0);
      const float common = (1.f - s.a) * Tc * G1 - E;
      d_ss += s.delta * (common + (1.f - s.as) * Tc * G2v);
      d_cs = s.delta * (common + (1.f - s.ac) * Tc * G3);
    }
    if (valid) {
      // softplus'(x) = 1 - exp(-softplus(x))
      const float ssig = p.s_sigma[m];
      p.d_ssig_pre[m] = d_ss * (1.f - expf(-ssig));
      float dcp = 0.f;
      if (p.cand) {
        const float csig = p.c_sigma[m];
        dcp = d_cs * (1.f - expf(-csig));
        p.d_csig_pre[m] = dcp;
      }
      if (p.stat_rgb) {
        p.d_rgb[m * 3] = wstat * g_rgb[0];
        p.d_rgb[m * 3 + 1] = wstat * g_rgb[1];
        p.d_rgb[m * 3 + 2] = wstat * g_rgb[2];
      }
      s_ws[warp][i] = ws;
      s_wc[warp][i] = wc;
      s_dg[warp][i] = dcp;   // reuse: candidate pre-activation gradient for pass 3
    }
  }
  if (!feat) return;
  __syncwarp();

  // ---- pass 3: per-sample gradients of the composited hidden vectors
  T* dhf = static_cast<T*>(p.d_hf) + r * S * p.ld_dhf;
  T* dg2 = p.cand ? static_cast<T*>(p.d_g2pre) + r * S * p.ld_dg2 : nullptr;
  float wcs[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.cand) {  // parameter pointer: only 4-byte aligned inside the flat buffer
#pragma unroll
    for (int e = 0; e < 4; ++e) wcs[e] = __ldg(p.w_csigma + lane * 4 + e);
  }
  for (int i0 = 0; i0 < S; i0 += kBatch) {
    float g[kBatch][4];
    if (p.cand) {
#pragma unroll
      for (int j = 0; j < kBatch; ++j)
        Vec<T>::template load<4>(g2 + min(i0 + j, S - 1) * p.ld_g2, lane, g[j]);
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int i = i0 + j;
      if (i >= S) break;
      const float ws = s_ws[warp][i];
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = ws * gh[e];
      Vec<T>::template store<8>(dhf + i * p.ld_dhf, lane, o);
      if (p.cand) {
        const float wc = s_wc[warp][i];
        const float dcp = s_dg[warp][i];
        float q[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) q[e] = g[j][e] > 0.f ? (wc * gg[e] + dcp * wcs[e]) : 0.f;
        Vec<T>::template store<4>(dg2 + i * p.ld_dg2, lane, q);
      }
    }
  }
}

int check_args(const CompArgs& p, bool bwd) {
  UPNERF_REQUIRE(p.R > 0 && p.S >= 1 && p.S <= kMaxS, UPNERF_ERR_BAD_SHAPE,
                 "composite: R=%lld S=%d (S must be <= %d)", (long long)p.R, p.S, kMaxS);
  UPNERF_REQUIRE(p.z && p.s_sigma, UPNERF_ERR_BAD_SHAPE, "composite: z/s_sigma missing");
  UPNERF_REQUIRE(!p.cand || p.c_sigma, UPNERF_ERR_BAD_SHAPE, "composite: c_sigma missing");
  UPNERF_REQUIRE(!p.stat_rgb || p.rgb, UPNERF_ERR_BAD_SHAPE, "composite: rgb missing");
  UPNERF_REQUIRE(p.feat_mode >= 0 && p.feat_mode <= 2 && (p.feat_mode != 2 || p.cand) &&
                     (p.feat_mode != 1 || !p.cand),
                 UPNERF_ERR_BAD_CONFIG, "composite: feat_mode=%d cand=%d", p.feat_mode, p.cand);
  UPNERF_REQUIRE(p.feat_mode == 0 || p.hf, UPNERF_ERR_BAD_SHAPE, "composite: hf missing");
  UPNERF_REQUIRE(p.feat_mode != 2 || p.g2, UPNERF_ERR_BAD_SHAPE, "composite: g2 missing");
  if (!bwd) {
    UPNERF_REQUIRE(p.s_depth, UPNERF_ERR_BAD_SHAPE, "composite_fwd: s_depth missing");
    UPNERF_REQUIRE(!p.cand || (p.c_weights && p.c_depth && p.t_weight), UPNERF_ERR_BAD_SHAPE,
                   "composite_fwd: candidate outputs missing");
    UPNERF_REQUIRE(!p.stat_rgb || p.s_rgb, UPNERF_ERR_BAD_SHAPE, "composite_fwd: s_rgb missing");
    UPNERF_REQUIRE(p.feat_mode == 0 || (p.hf_ray && p.ws_sum), UPNERF_ERR_BAD_SHAPE,
                   "composite_fwd: feature outputs missing");
    UPNERF_REQUIRE(p.feat_mode != 2 || (p.g2_ray && p.wc_sum), UPNERF_ERR_BAD_SHAPE,
                   "composite_fwd: candidate feature outputs missing");
  } else {
    UPNERF_REQUIRE(p.d_ssig_pre, UPNERF_ERR_BAD_SHAPE, "composite_bwd: d_ssig_pre missing");
    UPNERF_REQUIRE(!p.cand || p.d_csig_pre, UPNERF_ERR_BAD_SHAPE, "composite_bwd: d_csig_pre missing");
    UPNERF_REQUIRE(!p.stat_rgb || p.d_rgb, UPNERF_ERR_BAD_SHAPE, "composite_bwd: d_rgb missing");
    UPNERF_REQUIRE(p.feat_mode == 0 || (p.g_hf_ray && p.d_hf), UPNERF_ERR_BAD_SHAPE,
                   "composite_bwd: hidden-vector gradients missing");
    UPNERF_REQUIRE(p.feat_mode != 2 || (p.g_g2_ray && p.d_g2pre && p.w_csigma), UPNERF_ERR_BAD_SHAPE,
                   "composite_bwd: candidate gradients missing");
  }
  return UPNERF_OK;
}

}  // namespace

int composite_fwd_impl(const upnerf_composite_args* a, void* stream) {
  const CompArgs& p = *a;
  UPNERF_TRY(check_args(p, false));
  const unsigned grid = static_cast<unsigned>(ceil_div64(p.R, kWarps));
  LaunchScope scope(kCatComposite, as_stream(stream));
  if (a->dtype == UPNERF_BF16)
    composite_fwd_kernel<__nv_bfloat16><<<grid, kWarps * 32, 0, as_stream(stream)>>>(p);
  else
    composite_fwd_kernel<float><<<grid, kWarps * 32, 0, as_stream(stream)>>>(p);
  UPNERF_CHECK_LAUNCH("composite_fwd_kernel");
  return UPNERF_OK;
}

int composite_bwd_impl(const upnerf_composite_args* a, void* stream) {
  const CompArgs& p = *a;
  UPNERF_TRY(check_args(p, true));
  const unsigned grid = static_cast<unsigned>(ceil_div64(p.R, kWarps));
  LaunchScope scope(kCatComposite, as_stream(stream));
  if (a->dtype == UPNERF_BF16)
    composite_bwd_kernel<__nv_bfloat16><<<grid, kWarps * 32, 0, as_stream(stream)>>>(p);
  else
    composite_bwd_kernel<float><<<grid, kWarps * 32, 0, as_stream(stream)>>>(p);
  UPNERF_CHECK_LAUNCH("composite_bwd_kernel");
  return UPNERF_OK;
}

}  // namespace upnerf

extern "C" {
int upnerf_composite_fwd(const upnerf_composite_args* a, void* stream) {
  return upnerf::composite_fwd_impl(a, stream);
}
int upnerf_composite_bwd(const upnerf_composite_args* a, void* stream) {
  return upnerf::composite_bwd_impl(a, stream);
}
}
