// TransientNet on the tensor cores -- upnerf_tnet_fwd / upnerf_tnet_bwd.
//
// The reference's per-ray transient MLP (models/transient_net.py:6-38): on the 384-d image feature of a ray
//   h    = 4 x (Linear 256 + ReLU)                        feat_encoder
//   t    = ReLU(Linear_{384->128}([final_encoder(h) | embedding_t[img_idx]]))
//   alpha = sigmoid(alpha_layer(h)),  rgb = sigmoid(rgb_layer(t)),
//   beta  = softplus(beta_layer(t)) * alpha + beta_min
// and its autograd backward.  Six dense layers on [R, 384/256/128] rows: <0.5 % of the step's flops, but as
// fp32 cuBLAS / cutlass SIMT sgemm calls from torch eager they cost ~1.0 ms per step on a side stream (40
// launches) and shared the SMs with the 148-CTA tcgen05 grids.  Here every contraction is one
// upnerf_gemm_bf16 / wgrad launch (tcgen05, fp32 accumulate), the five N <= 3 heads and the activation
// algebra are two warp-per-ray kernels, and the weight gradients accumulate straight into the caller's
// gradient buffers (one split reduction for all of them).
#include <cuda_bf16.h>
#include <string.h>

#include "common.h"
#include "internal.h"

namespace upnerf {
namespace {

constexpr int HW = 256, TD = 128;   // hidden width, transient width (= t_encoder's output)

enum P {   // parameter slots in state_dict order (models/transient_net.py:9-25)
  pEmb = 0, pW0, pB0, pW1, pB1, pW2, pB2, pW3, pB3, pWf, pBf, pWt, pBt, pWa, pBa, pWb, pBb, pWr, pBr
};

__device__ __forceinline__ float softplus_ref(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fp32 rows -> bf16 rows (the feature batch), 8 elements per thread
__global__ void cvt_rows_kernel(const float* __restrict__ src, int64_t ld_src, __nv_bfloat16* __restrict__ dst,
                                int64_t ld_dst, int64_t R, int cols) {
  const int per_row = cols / 8;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= R * per_row) return;
  const int64_t r = i / per_row;
  const int c = static_cast<int>(i - r * per_row) * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src + r * ld_src + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(src + r * ld_src + c + 4));
  uint4 o;
  *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(a.x, a.y);
  *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(a.z, a.w);
  *reinterpret_cast<__nv_bfloat162*>(&o.z) = __floats2bfloat162_rn(b.x, b.y);
  *reinterpret_cast<__nv_bfloat162*>(&o.w) = __floats2bfloat162_rn(b.z, b.w);
  *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = o;
}

// embedding_t[img_idx] -> bf16 columns of the concat buffer
__global__ void gather_rows_bf16_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int64_t R,
                                        int dim, __nv_bfloat16* __restrict__ out, int64_t ld_out) {
  const int per_row = dim / 4;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= R * per_row) return;
  const int64_t r = i / per_row;
  const int c = static_cast<int>(i - r * per_row) * 4;
  const float4 a = __ldg(reinterpret_cast<const float4*>(table + idx[r] * dim + c));
  uint2 o;
  *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(a.x, a.y);
  *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(a.z, a.w);
  *reinterpret_cast<uint2*>(out + r * ld_out + c) = o;
}

// bf16 gradient rows -> fp32 atomics into the embedding table's gradient
__global__ void scatter_add_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, int64_t ld_src,
                                             const int64_t* __restrict__ idx, int64_t R, int dim,
                                             float* __restrict__ table) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= R * dim) return;
  const int64_t r = i / dim;
  const int c = static_cast<int>(i - r * dim);
  atomicAdd(table + idx[r] * dim + c, __bfloat162float(src[r * ld_src + c]));
}

struct HeadArgs {
  const __nv_bfloat16* H4;   // [R, 256]
  const __nv_bfloat16* T;    // [R, 128]
  const float *wa, *ba, *wb, *bb, *wr, *br;
  float beta_min;
  int64_t R;
  // forward outputs
  float *alpha, *beta, *rgb;
  // backward
  const float *g_alpha, *g_beta, *g_rgb;
  __nv_bfloat16* dT;         // [R, 128] gradient of t_encoder's pre-activation
  float* dA;                 // [R] gradient of alpha_layer's pre-activation
  float* dBR;                // [4][R]: gradient of beta_layer's pre-activation, then the three of rgb_layer
};

// One warp per ray: the row-dots of the three heads.  lane holds 8 columns of h (256) and 4 of t (128).
template <bool kBackward>
__global__ void __launch_bounds__(128) tnet_heads_kernel(HeadArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (r >= a.R) return;
  float h[8], t[4];
  {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.H4 + r * HW + lane * 8));
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(b[i]);
      h[2 * i] = f.x;
      h[2 * i + 1] = f.y;
    }
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(a.T + r * TD + lane * 4));
    const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
    const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
    t[0] = f0.x; t[1] = f0.y; t[2] = f1.x; t[3] = f1.y;
  }
  float sa = 0.f, sb = 0.f, sr[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) sa += h[i] * __ldg(a.wa + lane * 8 + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sb += t[i] * __ldg(a.wb + lane * 4 + i);
#pragma unroll
    for (int c = 0; c < 3; ++c) sr[c] += t[i] * __ldg(a.wr + c * TD + lane * 4 + i);
  }
  sa = warp_sum(sa) + __ldg(a.ba);
  sb = warp_sum(sb) + __ldg(a.bb);
#pragma unroll
  for (int c = 0; c < 3; ++c) sr[c] = warp_sum(sr[c]) + __ldg(a.br + c);
  const float alpha = sigmoid_ref(sa);
  const float sp = softplus_ref(sb);
  if (!kBackward) {
    if (lane == 0) {
      a.alpha[r] = alpha;
      a.beta[r] = sp * alpha + a.beta_min;
#pragma unroll
      for (int c = 0; c < 3; ++c) a.rgb[r * 3 + c] = sigmoid_ref(sr[c]);
    }
    return;
  }
  const float gb = a.g_beta ? a.g_beta[r] : 0.f;
  const float ga = (a.g_alpha ? a.g_alpha[r] : 0.f) + gb * sp;      // alpha feeds beta too
  // d softplus(x) = sigmoid(x) below the threshold, 1 above (torch.nn.Softplus(beta=1, threshold=20))
  const float d_b = gb * alpha * (sb > 20.f ? 1.f : sigmoid_ref(sb));
  const float d_a = ga * alpha * (1.f - alpha);
  float d_r[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = sigmoid_ref(sr[c]);
    d_r[c] = a.g_rgb ? a.g_rgb[r * 3 + c] * s * (1.f - s) : 0.f;
  }
  if (lane == 0) {
    a.dA[r] = d_a;
    a.dBR[r] = d_b;
#pragma unroll
    for (int c = 0; c < 3; ++c) a.dBR[(c + 1) * a.R + r] = d_r[c];
  }
  float o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = d_b * __ldg(a.wb + lane * 4 + i);
#pragma unroll
    for (int c = 0; c < 3; ++c) v += d_r[c] * __ldg(a.wr + c * TD + lane * 4 + i);
    o[i] = t[i] > 0.f ? v : 0.f;          // through t_encoder's ReLU
  }
  uint2 w;
  *reinterpret_cast<__nv_bfloat162*>(&w.x) = __floats2bfloat162_rn(o[0], o[1]);
  *reinterpret_cast<__nv_bfloat162*>(&w.y) = __floats2bfloat162_rn(o[2], o[3]);
  *reinterpret_cast<uint2*>(a.dT + r * TD + lane * 4) = w;
}

struct Bump {
  uint8_t* base;
  uint64_t off;
  template <typename U> U* take(uint64_t count) {
    off = (off + 255) & ~uint64_t(255);
    U* p = base ? reinterpret_cast<U*>(base + off) : nullptr;
    off += count * sizeof(U);
    return p;
  }
};

struct Plan {
  __nv_bfloat16 *X0, *H[4], *CAT, *T;                       // kept by forward
  __nv_bfloat16 *W[4], *Wf, *Wt, *WT[4], *WfT, *WtT;        // packed operands (WT[0] unused)
  __nv_bfloat16 *dT, *dFIN, *dEMB, *dH[4];                  // backward scratch
  float *dA, *dBR, *pool;
  uint64_t pool_floats, bytes;
  void* region;
  uint64_t region_bytes;
};

int make_plan(const upnerf_tnet_args& a, void* base, Plan* p) {
  UPNERF_REQUIRE(a.n_rays > 0 && a.feat_dim > 0 && a.feat_dim % 64 == 0 && a.feat_dim <= 512, UPNERF_ERR_BAD_SHAPE,
                 "tnet: n_rays=%lld feat_dim=%d (a multiple of 64, <= 512)", (long long)a.n_rays, a.feat_dim);
  UPNERF_REQUIRE(a.transient_dim == TD && a.hidden == HW, UPNERF_ERR_BAD_CONFIG,
                 "tnet: only the shipped widths (hidden 256, transient 128) are implemented");
  const int64_t R = a.n_rays;
  const int F = a.feat_dim;
  Bump b{static_cast<uint8_t*>(base), 0};
  p->X0 = b.take<__nv_bfloat16>(R * F);
  for (int i = 0; i < 4; ++i) p->H[i] = b.take<__nv_bfloat16>(R * HW);
  p->CAT = b.take<__nv_bfloat16>(R * (HW + TD));
  p->T = b.take<__nv_bfloat16>(R * TD);
  b.off = (b.off + 255) & ~uint64_t(255);
  const uint64_t start = b.off;
  p->region = base ? static_cast<uint8_t*>(base) + start : nullptr;
  p->W[0] = b.take<__nv_bfloat16>(HW * F);
  for (int i = 1; i < 4; ++i) p->W[i] = b.take<__nv_bfloat16>(HW * HW);
  p->Wf = b.take<__nv_bfloat16>(HW * HW);
  p->Wt = b.take<__nv_bfloat16>(TD * (HW + TD));
  p->WT[0] = nullptr;
  for (int i = 1; i < 4; ++i) p->WT[i] = b.take<__nv_bfloat16>(HW * HW);
  p->WfT = b.take<__nv_bfloat16>(HW * HW);
  p->WtT = b.take<__nv_bfloat16>((HW + TD) * TD);
  p->region_bytes = b.off - start;
  p->dT = b.take<__nv_bfloat16>(R * TD);
  p->dFIN = b.take<__nv_bfloat16>(R * HW);
  p->dEMB = b.take<__nv_bfloat16>(R * TD);
  for (int i = 0; i < 4; ++i) p->dH[i] = b.take<__nv_bfloat16>(R * HW);
  p->dA = b.take<float>(R);
  p->dBR = b.take<float>(4 * R);
  // split partials of the weight-gradient launches of upnerf_tnet_bwd (same split rule as wgrad_launch)
  auto need = [&](int N, int K) -> uint64_t {
    const int chunks = N / 128;
    int64_t splits = sm_count() / chunks;
    const int64_t steps = ceil_div64(R, 64);
    if (splits > steps) splits = steps;
    if (splits < 1) splits = 1;
    return static_cast<uint64_t>(splits) * chunks * (K + 1) * 128;
  };
  p->pool_floats = need(TD, HW) + need(TD, TD) + 4 * need(HW, HW);
  for (int c = 0; c < F; c += HW) p->pool_floats += need(HW, F - c < HW ? F - c : HW);
  p->pool = b.take<float>(p->pool_floats);
  p->bytes = b.off + 256;
  return UPNERF_OK;
}

upnerf_epilogue ep_none() {
  upnerf_epilogue e;
  memset(&e, 0, sizeof(e));
  return e;
}

int check(const upnerf_tnet_args& a, const Plan& p) {
  UPNERF_REQUIRE(upnerf_device_ok(), UPNERF_ERR_CUDA,
                 "upnerf_b200 needs a compute-capability 10.x GPU (sm_100a); there is no fallback");
  UPNERF_REQUIRE(a.feats && a.img_idx, UPNERF_ERR_BAD_SHAPE, "tnet: feats / img_idx missing");
  for (int i = 0; i < UPNERF_TNET_PARAMS; ++i)
    UPNERF_REQUIRE(a.params[i], UPNERF_ERR_BAD_SHAPE, "tnet: parameter %d missing", i);
  UPNERF_REQUIRE(a.workspace && a.workspace_bytes >= p.bytes, UPNERF_ERR_WORKSPACE,
                 "tnet: workspace too small (%llu < %llu bytes)", (unsigned long long)a.workspace_bytes,
                 (unsigned long long)p.bytes);
  return UPNERF_OK;
}

}  // namespace
}  // namespace upnerf

extern "C" {

uint64_t upnerf_tnet_workspace_bytes(const upnerf_tnet_args* a) {
  upnerf::Plan p;
  if (upnerf::make_plan(*a, nullptr, &p) != 0) return 0;
  return p.bytes;
}

int upnerf_tnet_fwd(const upnerf_tnet_args* a, void* stream) {
  using namespace upnerf;
  Plan p;
  UPNERF_TRY(make_plan(*a, a->workspace, &p));
  UPNERF_TRY(check(*a, p));
  UPNERF_REQUIRE(a->alpha && a->beta && a->rgb, UPNERF_ERR_BAD_SHAPE, "tnet_fwd: outputs missing");
  cudaStream_t st = as_stream(stream);
  const int64_t R = a->n_rays;
  const int F = a->feat_dim, CW = HW + TD;
  // operands: fp32 master weights -> bf16 (K-major as stored) and the transposes the data gradients need
  PackList pl;
  pl.n = 0;
  auto add = [&](const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int rows, int cols, int tr) {
    pl.ops[pl.n++] = PackOp{src, ld_src, dst, ld_dst, rows, cols, tr};
  };
  const int wslot[4] = {pW0, pW1, pW2, pW3};
  add(a->params[pW0], F, p.W[0], F, HW, F, 0);
  for (int i = 1; i < 4; ++i) {
    add(a->params[wslot[i]], HW, p.W[i], HW, HW, HW, 0);
    add(a->params[wslot[i]], HW, p.WT[i], HW, HW, HW, 1);
  }
  add(a->params[pWf], HW, p.Wf, HW, HW, HW, 0);
  add(a->params[pWf], HW, p.WfT, HW, HW, HW, 1);
  add(a->params[pWt], CW, p.Wt, CW, TD, CW, 0);
  add(a->params[pWt], CW, p.WtT, TD, TD, CW, 1);
  UPNERF_TRY(run_pack(pl, UPNERF_BF16, st));
  {
    const int64_t n = R * (F / 8);
    LaunchScope scope(kCatTnet, st);
    cvt_rows_kernel<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, st>>>(a->feats, F, p.X0, F, R, F);
    UPNERF_CHECK_LAUNCH("cvt_rows_kernel");
  }
  {
    const int64_t n = R * (TD / 4);
    LaunchScope scope(kCatTnet, st);
    gather_rows_bf16_kernel<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, st>>>(a->params[pEmb], a->img_idx, R,
                                                                                     TD, p.CAT + HW, CW);
    UPNERF_CHECK_LAUNCH("gather_rows_bf16_kernel");
  }
  const int bslot[4] = {pB0, pB1, pB2, pB3};
  for (int i = 0; i < 4; ++i) {
    upnerf_epilogue e = ep_none();
    e.bias = a->params[bslot[i]];
    e.act = 1;
    const void* in = i == 0 ? p.X0 : p.H[i - 1];
    const int K = i == 0 ? F : HW;
    UPNERF_TRY(upnerf_gemm_bf16(in, K, p.W[i], K, p.H[i], HW, R, HW, K, &e, stream));
  }
  upnerf_epilogue e = ep_none();
  e.bias = a->params[pBf];
  UPNERF_TRY(upnerf_gemm_bf16(p.H[3], HW, p.Wf, HW, p.CAT, CW, R, HW, HW, &e, stream));     // final_encoder -> [FIN | .]
  e = ep_none();
  e.bias = a->params[pBt];
  e.act = 1;
  UPNERF_TRY(upnerf_gemm_bf16(p.CAT, CW, p.Wt, CW, p.T, TD, R, TD, CW, &e, stream));        // t_encoder
  HeadArgs h;
  memset(&h, 0, sizeof(h));
  h.H4 = p.H[3]; h.T = p.T;
  h.wa = a->params[pWa]; h.ba = a->params[pBa]; h.wb = a->params[pWb]; h.bb = a->params[pBb];
  h.wr = a->params[pWr]; h.br = a->params[pBr];
  h.beta_min = a->beta_min; h.R = R;
  h.alpha = a->alpha; h.beta = a->beta; h.rgb = a->rgb;
  {
    LaunchScope scope(kCatTnet, st);
    tnet_heads_kernel<false><<<static_cast<unsigned>(ceil_div64(R, 4)), 128, 0, st>>>(h);
    UPNERF_CHECK_LAUNCH("tnet_heads_kernel<fwd>");
  }
  return UPNERF_OK;
}

int upnerf_tnet_bwd(const upnerf_tnet_args* a, void* stream) {
  using namespace upnerf;
  Plan p;
  UPNERF_TRY(make_plan(*a, a->workspace, &p));
  UPNERF_TRY(check(*a, p));
  cudaStream_t st = as_stream(stream);
  const int64_t R = a->n_rays;
  const int F = a->feat_dim, CW = HW + TD;
  float* const* g = a->d_params;
  // 1. heads: pre-activation gradients of alpha / beta / rgb, dT through t_encoder's ReLU
  HeadArgs h;
  memset(&h, 0, sizeof(h));
  h.H4 = p.H[3]; h.T = p.T;
  h.wa = a->params[pWa]; h.ba = a->params[pBa]; h.wb = a->params[pWb]; h.bb = a->params[pBb];
  h.wr = a->params[pWr]; h.br = a->params[pBr];
  h.beta_min = a->beta_min; h.R = R;
  h.g_alpha = a->g_alpha; h.g_beta = a->g_beta; h.g_rgb = a->g_rgb;
  h.dT = p.dT; h.dA = p.dA; h.dBR = p.dBR;
  {
    LaunchScope scope(kCatTnet, st);
    tnet_heads_kernel<true><<<static_cast<unsigned>(ceil_div64(R, 4)), 128, 0, st>>>(h);
    UPNERF_CHECK_LAUNCH("tnet_heads_kernel<bwd>");
  }
  if (g[pWa]) UPNERF_TRY(rowscale_colsum(p.H[3], HW, p.dA, R, HW, g[pWa], g[pBa], UPNERF_BF16, st));
  if (g[pWb]) UPNERF_TRY(rowscale_colsum(p.T, TD, p.dBR, R, TD, g[pWb], g[pBb], UPNERF_BF16, st));
  if (g[pWr] && a->g_rgb)
    for (int c = 0; c < 3; ++c)
      UPNERF_TRY(rowscale_colsum(p.T, TD, p.dBR + (c + 1) * R, R, TD, g[pWr] + c * TD, g[pBr] ? g[pBr] + c : nullptr,
                                 UPNERF_BF16, st));
  // 2. data gradients down the chain
  upnerf_epilogue e = ep_none();
  UPNERF_TRY(upnerf_gemm_bf16(p.dT, TD, p.WtT, TD, p.dFIN, HW, R, HW, TD, &e, stream));                  // d final_encoder out
  if (g[pEmb]) {
    UPNERF_TRY(upnerf_gemm_bf16(p.dT, TD, p.WtT + static_cast<int64_t>(HW) * TD, TD, p.dEMB, TD, R, TD, TD, &e, stream));
    LaunchScope scope(kCatTnet, st);
    scatter_add_rows_bf16_kernel<<<static_cast<unsigned>(ceil_div64(R * TD, 256)), 256, 0, st>>>(p.dEMB, TD, a->img_idx, R,
                                                                                                 TD, g[pEmb]);
    UPNERF_CHECK_LAUNCH("scatter_add_rows_bf16_kernel");
  }
  // dH4_pre = (dFIN W_final + dA (x) w_alpha) * [h4 > 0]
  e = ep_none();
  e.rank1_row = p.dA; e.rank1_col = a->params[pWa];
  e.aux = p.H[3]; e.ldaux = HW; e.aux_mode = 2;
  UPNERF_TRY(upnerf_gemm_bf16(p.dFIN, HW, p.WfT, HW, p.dH[3], HW, R, HW, HW, &e, stream));
  for (int i = 3; i >= 1; --i) {
    e = ep_none();
    e.aux = p.H[i - 1]; e.ldaux = HW; e.aux_mode = 2;
    UPNERF_TRY(upnerf_gemm_bf16(p.dH[i], HW, p.WT[i], HW, p.dH[i - 1], HW, R, HW, HW, &e, stream));
  }
  // 3. weight gradients: dW += dY^T X, db += colsum(dY); split partials parked, ONE reduction at the end
  WgradBatch wb;
  memset(&wb, 0, sizeof(wb));
  wb.pool = p.pool;
  wb.pool_floats = p.pool_floats;
  wb.defer = 1;   // one grouped launch for all of them
  auto wg = [&](const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw, float* db, int N,
                int K, int dst0) -> int {
    const int src = 0, len = K, dst = dst0;
    return wgrad_launch(dY, lddy, X, ldx, dW, lddw, nullptr, 0, db, R, N, K, 1, &src, &len, &dst, &wb, stream);
  };
  if (g[pWt]) {      // K = 384 as 256 (final_encoder part) + 128 (embedding part)
    UPNERF_TRY(wg(p.dT, TD, p.CAT, CW, g[pWt], CW, g[pBt], TD, HW, 0));
    UPNERF_TRY(wg(p.dT, TD, p.CAT + HW, CW, g[pWt], CW, nullptr, TD, TD, HW));
  }
  if (g[pWf]) UPNERF_TRY(wg(p.dFIN, HW, p.H[3], HW, g[pWf], HW, g[pBf], HW, HW, 0));
  const int wslot[4] = {pW0, pW1, pW2, pW3}, bslot[4] = {pB0, pB1, pB2, pB3};
  for (int i = 3; i >= 1; --i)
    if (g[wslot[i]]) UPNERF_TRY(wg(p.dH[i], HW, p.H[i - 1], HW, g[wslot[i]], HW, g[bslot[i]], HW, HW, 0));
  if (g[pW0]) {      // K = feat_dim in column blocks of <= 256
    for (int c = 0; c < F; c += HW) {
      const int k = F - c < HW ? F - c : HW;
      UPNERF_TRY(wg(p.dH[0], HW, p.X0 + c, F, g[pW0], F, c == 0 ? g[pB0] : nullptr, HW, k, c));
    }
  }
  return wgrad_reduce(&wb, st);
}

}  // extern "C"
