// Host orchestration of the whole path: render_rays (reference models/rendering.py:53-314)
// driving NeRF.forward (models/nerf.py:80-124) for the coarse and fine networks, forward and
// backward, as a fixed sequence of kernel launches on one stream.
//
// Data flow of one network pass over M = R*S samples (T = bf16 or fp32):
//   PE(x) --L1--> H1 .. H4 | skip [H4|PE] --L5--> H5 .. H8 --(row-dot)--> s_sigma
//   H8 --final--> HF --(+per-ray bias)--> G1 -> G2 --(row-dot)--> c_sigma      (candidate head)
//                  HF --(W_rgb0[:, :F] W_sf folded, +per-ray bias)--> Q --(row-dots)--> s_rgb
//   compositing reduces {s_sigma, c_sigma, s_rgb, HF, G2} per ray; the two linear 384-d
//   feature projections run once per RAY on the composited hidden vectors.
// Per-ray inputs (appearance / candidate embeddings, direction encoding) enter their layers
// as a per-ray bias, so nothing per-ray is ever repeated per sample.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "internal.h"

namespace upnerf {
namespace {

constexpr int W = 256, H = 128, PEW = 64, X4W = 320;

// ------------------------------------------------------------------ parameter layout
struct NetLayout {
  int in_xyz, in_dir, rgb_in, F, ad, cd;
  int64_t progress, Wl[8], bl[8], Wf, bf, Ws, bs, Wsf, bsf, Wr0, br0, Wr2, br2;
  int64_t Wc0, bc0, Wc2, bc2, Wcs, bcs, Wcf, bcf;
  int64_t total;
};

int make_layout(const upnerf_net_config& c, NetLayout* L) {
  UPNERF_REQUIRE(c.D == 8 && c.W == 256, UPNERF_ERR_BAD_CONFIG,
                 "only the D=8, W=256, skips=[4] trunk is implemented (got D=%d W=%d)", c.D, c.W);
  UPNERF_REQUIRE(c.xyz_L >= 1 && c.xyz_L <= 10 && c.dir_L >= 1 && c.dir_L <= 16, UPNERF_ERR_BAD_CONFIG,
                 "xyz_L=%d (<=10) / dir_L=%d (<=16) unsupported", c.xyz_L, c.dir_L);
  UPNERF_REQUIRE(!c.encode_feat || c.feat_dim > 0, UPNERF_ERR_BAD_CONFIG, "encode_feat needs feat_dim");
  UPNERF_REQUIRE(c.encode_feat || c.candidate_dim == 0, UPNERF_ERR_BAD_CONFIG,
                 "encode_feat=False with a candidate head (c_rgb path) is not implemented");
  memset(L, 0, sizeof(*L));
  L->in_xyz = 6 * c.xyz_L + 3;
  L->in_dir = 6 * c.dir_L + 3;
  L->F = c.encode_feat ? c.feat_dim : 0;
  L->ad = c.appearance_dim;
  L->cd = c.candidate_dim;
  L->rgb_in = (c.encode_feat ? c.feat_dim : W) + L->in_dir + L->ad;
  int64_t o = 0;
  L->progress = o; o += 1;
  for (int i = 0; i < 8; ++i) {
    const int k = i == 0 ? L->in_xyz : (i == 4 ? W + L->in_xyz : W);
    L->Wl[i] = o; o += static_cast<int64_t>(W) * k;
    L->bl[i] = o; o += W;
  }
  L->Wf = o; o += W * W;
  L->bf = o; o += W;
  L->Ws = o; o += W;
  L->bs = o; o += 1;
  if (c.encode_feat) {
    L->Wsf = o; o += static_cast<int64_t>(L->F) * W;
    L->bsf = o; o += L->F;
  }
  L->Wr0 = o; o += static_cast<int64_t>(H) * L->rgb_in;
  L->br0 = o; o += H;
  L->Wr2 = o; o += 3 * H;
  L->br2 = o; o += 3;
  if (c.candidate_dim > 0) {
    L->Wc0 = o; o += static_cast<int64_t>(H) * (W + L->cd);
    L->bc0 = o; o += H;
    L->Wc2 = o; o += H * H;
    L->bc2 = o; o += H;
    L->Wcs = o; o += H;
    L->bcs = o; o += 1;
    L->Wcf = o; o += static_cast<int64_t>(L->F) * H;
    L->bcf = o; o += L->F;
  }
  L->total = o;
  return UPNERF_OK;
}

// ------------------------------------------------------------------ workspace carving
struct Bump {
  uint8_t* base;
  uint64_t off;
  template <typename U> U* take(uint64_t count) {
    off = (off + 255) & ~uint64_t(255);
    U* p = base ? reinterpret_cast<U*>(base + off) : nullptr;
    off += count * sizeof(U);
    return p;
  }
  void* take_bytes(uint64_t bytes) { return take<uint8_t>(bytes); }
};

constexpr int WCAT = UPNERF_TRUNK_WCAT_COLS;    // trunk forward weights, concatenated along K
constexpr int WCATT = UPNERF_TRUNK_WCATT_COLS;  // transposed trunk weights in backward-chain order
struct Packed {  // GEMM operands derived from the fp32 parameters (element type T)
  void* Wcat;                                           // [256, WCAT]: W1|W2|W3|W4|W5=[h|PE]|W6|W7|W8|WF
  void *W1, *Wk[8], *W5, *WF;                           // views into Wcat (row stride WCAT)
  void* Wcq;                                            // [256, 256]: [Wc1; Wq] stacked along N (views below)
  void *Wc1, *Wq;                                       // [128, 256] K-major, rows 0..127 / 128..255 of Wcq
  void* Wc2;                                            // [128, 128]
  void* WcatT;                                          // [256, WCATT]: WF^T|W8^T|W7^T|W6^T|W5h^T|W4^T|W3^T|W2^T
  void *WkT[8], *W5T, *WFT;                             // views into WcatT (row stride WCATT); W5T = h part
  void *W1T, *W5peT;                                    // [64, 256]: rows = PE columns (for dPE)
  void* WcqT;                                           // [256, 256]: [Wc1^T | Wq^T] side by side along K
  void *Wc1T, *WqT;                                     // [256, 128] views (row stride 256)
  void* Wc2T;
  float* hw3;                                           // [3, 256] fp32: rgb_share_layer.2 weights on columns 128..255
  float* Wq32;      // [128,256] fp32 folded rgb weight
  float* bq_const;  // [128]
  float* band_xyz;  // [16]
  float* band_dir;  // [16]
  void* region;     // start of the packed T region (zeroed before packing)
  uint64_t region_bytes;
};

struct PassBufs {
  int64_t R, M;
  int S;
  // saved by forward
  void *X4, *Hs[9], *HF, *G1, *G2, *Q;  // Hs[1..8]; Hs[4] aliases X4 (ld 320)
  float *ssig, *csig, *rgb, *z;
  uint32_t* relu_mask;  // bit masks of H1..H8 (bf16 mode: written by the fused forward)
  float *HFr, *G2r, *Wsum, *Wcsum, *P, *Crows, *Bq, *Bc;
  Packed pk;
};

struct Scratch {
  void *dY[8], *dHF, *dG2p, *dG1p, *dQp, *dPE;  // dY[j] = gradient of layer (8-j)'s pre-activation
  float *dssig, *dcsig, *drgb;
  float *gHFr, *gG2r, *gWs, *gWc, *dBq, *dBc, *dP, *dCrows;
  float *dWq2[2], *dbq2[2];  // per network pass (fine, coarse): read by side-stream work that outlives the pass
  float* gdump;              // sink of the tiny side-effect gradients when the weights are frozen
  float* wg_pool;            // split partials of the tcgen05 weight gradients (bf16 mode)
  uint64_t wg_pool_floats;
};

void carve_pass(Bump& b, int64_t R, int S, size_t es, const NetLayout& L, bool keep, PassBufs* p) {
  const int64_t M = R * S;
  p->R = R; p->M = M; p->S = S;
  p->X4 = b.take_bytes(M * X4W * es);
  // forward-only in bf16 mode: the fused trunk keeps H1..H8 on chip, only the PE columns of X4 exist
  const bool on_chip = !keep && es == 2;
  for (int i = 1; i <= 8; ++i) p->Hs[i] = (i == 4) ? p->X4 : (on_chip ? nullptr : b.take_bytes(M * W * es));
  p->Hs[0] = nullptr;
  p->HF = b.take_bytes(M * W * es);
  // candidate and rgb hidden layers side by side in one [M, 256] buffer: G1 = columns 0..127,
  // Q = columns 128..255 (both with row stride 256), so the two layers can run as one stacked GEMM
  p->G1 = b.take_bytes(M * 2 * H * es);
  p->Q = p->G1 ? static_cast<uint8_t*>(p->G1) + H * es : nullptr;
  p->G2 = b.take_bytes(M * H * es);
  p->ssig = b.take<float>(M);
  p->csig = b.take<float>(M);
  p->rgb = b.take<float>(M * 3);
  p->z = b.take<float>(M);
  p->relu_mask = (es == 2 && keep) ? b.take<uint32_t>(upnerf_trunk_mask_words(M)) : nullptr;
  p->HFr = b.take<float>(R * W);
  p->G2r = b.take<float>(R * H);
  p->Wsum = b.take<float>(R);
  p->Wcsum = b.take<float>(R);
  p->P = b.take<float>(R * (L.in_dir + L.ad));
  p->Crows = b.take<float>(R * (L.cd > 0 ? L.cd : 1));
  // per-ray biases: stacked mode [R, 256] rows [Bc | Bq]; otherwise two contiguous [R, 128] blocks
  p->Bc = b.take<float>(R * 2 * H);
  p->Bq = p->Bc ? p->Bc + R * H : nullptr;
  Packed& k = p->pk;
  k.Wq32 = b.take<float>(H * W);
  k.bq_const = b.take<float>(H);
  k.band_xyz = b.take<float>(16);
  k.band_dir = b.take<float>(16);
  b.off = (b.off + 255) & ~uint64_t(255);
  const uint64_t start = b.off;
  k.region = b.base ? b.base + start : nullptr;
  k.Wcat = b.take_bytes(static_cast<uint64_t>(W) * WCAT * es);
  {
    int64_t c = 0;
    auto view = [&](int width) {
      void* p = k.Wcat ? static_cast<uint8_t*>(k.Wcat) + c * es : nullptr;
      c += width;
      return p;
    };
    k.W1 = view(PEW);
    for (int i = 0; i < 8; ++i) {
      k.Wk[i] = nullptr;
      if (i == 0) continue;
      if (i == 4) k.W5 = view(X4W);
      else k.Wk[i] = view(W);
    }
    k.WF = view(W);
  }
  k.WcatT = b.take_bytes(static_cast<uint64_t>(W) * WCATT * es);
  {
    int64_t c = 0;
    auto view = [&]() {
      void* p = k.WcatT ? static_cast<uint8_t*>(k.WcatT) + c * es : nullptr;
      c += W;
      return p;
    };
    k.WFT = view();
    for (int i = 7; i >= 1; --i) {
      if (i == 4) k.W5T = view();
      else k.WkT[i] = view();
    }
    k.WkT[0] = k.WkT[4] = nullptr;
  }
  // [W5pe^T | W1^T] side by side along K ([64, 512], row stride 2W): dPE is ONE GEMM over [dY5 | dY1]
  k.W5peT = b.take_bytes(PEW * 2 * W * es);
  k.W1T = k.W5peT ? static_cast<uint8_t*>(k.W5peT) + W * es : nullptr;
  k.Wcq = b.take_bytes(2 * H * W * es);
  k.Wc1 = k.Wcq;
  k.Wq = k.Wcq ? static_cast<uint8_t*>(k.Wcq) + static_cast<uint64_t>(H) * W * es : nullptr;
  k.WcqT = b.take_bytes(W * 2 * H * es);
  k.Wc1T = k.WcqT;
  k.WqT = k.WcqT ? static_cast<uint8_t*>(k.WcqT) + H * es : nullptr;
  k.Wc2 = b.take_bytes(H * H * es);
  k.Wc2T = b.take_bytes(H * H * es);
  k.hw3 = b.take<float>(3 * 2 * H);
  k.region_bytes = b.off - start;
}

void carve_scratch(Bump& b, int64_t R, int S, size_t es, const NetLayout& L, Scratch* s) {
  const int64_t M = R * S;
  // bf16 mode keeps all eight (the fused chain writes them, the weight gradients read them);
  // fp32 mode ping-pongs two
  for (int j = 0; j < 8; ++j) s->dY[j] = (es == 2 || j < 2) ? b.take_bytes(M * W * es) : nullptr;
  s->dHF = b.take_bytes(M * W * es);
  s->dG2p = b.take_bytes(M * H * es);
  s->dG1p = b.take_bytes(M * 2 * H * es);   // [dG1p | dQp], row stride 256
  s->dQp = s->dG1p ? static_cast<uint8_t*>(s->dG1p) + H * es : nullptr;
  s->dPE = b.take_bytes(M * PEW * es);
  s->dssig = b.take<float>(M);
  s->dcsig = b.take<float>(M);
  s->drgb = b.take<float>(M * 3);
  s->gHFr = b.take<float>(R * W);
  s->gG2r = b.take<float>(R * H);
  s->gWs = b.take<float>(R);
  s->gWc = b.take<float>(R);
  s->dBq = b.take<float>(R * H);
  s->dBc = b.take<float>(R * H);
  s->dP = b.take<float>(R * (L.in_dir + L.ad));
  s->dCrows = b.take<float>(R * (L.cd > 0 ? L.cd : 1));
  for (int i = 0; i < 2; ++i) {
    s->dWq2[i] = b.take<float>(H * W);
    s->dbq2[i] = b.take<float>(H);
  }
  s->gdump = b.take<float>(4 * H);
  s->wg_pool_floats = es == 2 ? wgrad_pool_floats() : 0;
  s->wg_pool = s->wg_pool_floats ? b.take<float>(s->wg_pool_floats) : nullptr;
}

struct Plan {
  NetLayout L;
  PassBufs coarse, fine;
  Scratch scratch;
  uint64_t bytes;
  size_t es;
  int S_f;
};

int make_plan(const upnerf_render_args& a, void* base, Plan* pl) {
  UPNERF_TRY(make_layout(a.cfg, &pl->L));
  UPNERF_REQUIRE(a.dtype == UPNERF_F32 || a.dtype == UPNERF_BF16, UPNERF_ERR_BAD_CONFIG, "dtype=%d", a.dtype);
  UPNERF_REQUIRE(a.n_rays > 0 && a.n_samples >= 3 && a.n_samples <= 256 && a.n_importance >= 0 &&
                     a.n_samples + a.n_importance <= 256,
                 UPNERF_ERR_BAD_SHAPE, "n_rays=%lld n_samples=%d n_importance=%d unsupported",
                 (long long)a.n_rays, a.n_samples, a.n_importance);
  pl->es = a.dtype == UPNERF_BF16 ? 2 : 4;
  pl->S_f = a.n_samples + a.n_importance;
  UPNERF_REQUIRE((a.n_rays * a.n_samples) % 8 == 0, UPNERF_ERR_BAD_SHAPE,
                 "n_rays*n_samples must be a multiple of 8");
  Bump b{static_cast<uint8_t*>(base), 0};
  const bool keep = !a.no_grad;
  carve_pass(b, a.n_rays, a.n_samples, pl->es, pl->L, keep, &pl->coarse);
  if (a.n_importance > 0) carve_pass(b, a.n_rays, pl->S_f, pl->es, pl->L, keep, &pl->fine);
  if (keep) carve_scratch(b, a.n_rays, a.n_importance > 0 ? pl->S_f : a.n_samples, pl->es, pl->L, &pl->scratch);
  pl->bytes = b.off + 256;
  return UPNERF_OK;
}

// ------------------------------------------------------------------ phase flags
struct Phase {
  bool cand, rgb, feat;
  int feat_mode;
};
Phase make_phase(const upnerf_net_config& c, float m) {
  Phase p;
  p.cand = (m < 1.f) && c.encode_candidate && c.candidate_dim > 0 && c.encode_feat;
  p.rgb = c.encode_feat ? (m > 0.f) : true;
  p.feat = c.encode_feat && (m < 1.f);
  p.feat_mode = !p.feat ? 0 : (p.cand ? 2 : 1);
  return p;
}

// ------------------------------------------------------------------ GEMM dispatch by precision
struct Ctx {
  int dtype;
  size_t es;
  cudaStream_t st;
  WgradBatch* wb;  // bf16 backward: weight gradients park split partials here (one reduce per pass)
  // Leaf work (per-ray products, parameter-space chain rules, bias/embedding gradients: small,
  // latency-bound launches nothing downstream in the pass waits for) goes to a second in-order stream
  // so it runs beside the tensor-core chain instead of inside it.  side == st when disabled.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_pro[2] = {nullptr, nullptr};   // side stream: head operands + per-ray biases of pass 0 / 1 are ready
  Ctx leaf() const { Ctx c = *this; c.st = side; c.wb = nullptr; return c; }
  bool has_side() const { return side != st; }
};

// One side stream and one fork/join event pair per device, created on first use and kept for the
// life of the process (UPNERF_SIDE_STREAM=0 disables the overlap: everything stays on one stream).
struct SideRes { cudaStream_t st; cudaEvent_t fork_ev, join_ev, pro_ev[2]; int state; };
SideRes* side_res() {
  static SideRes res[32];
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("UPNERF_SIDE_STREAM");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
  SideRes& r = res[dev];
  if (r.state == 0) {
    r.state = -1;
    if (cudaStreamCreateWithFlags(&r.st, cudaStreamNonBlocking) == cudaSuccess &&
        cudaEventCreateWithFlags(&r.fork_ev, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&r.join_ev, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&r.pro_ev[0], cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&r.pro_ev[1], cudaEventDisableTiming) == cudaSuccess)
      r.state = 1;
  }
  return r.state == 1 ? &r : nullptr;
}
Ctx make_ctx(int dtype, size_t es, cudaStream_t st, WgradBatch* wb) {
  Ctx c;
  c.dtype = dtype; c.es = es; c.st = st; c.wb = wb;
  c.side = st;
  if (SideRes* r = side_res()) {
    c.side = r->st; c.ev_fork = r->fork_ev; c.ev_join = r->join_ev;
    c.ev_pro[0] = r->pro_ev[0]; c.ev_pro[1] = r->pro_ev[1];
  }
  return c;
}
// side stream picks up everything launched on the main stream so far
int fork_side(const Ctx& c) {
  if (!c.has_side()) return UPNERF_OK;
  UPNERF_CHECK_CUDA(cudaEventRecord(c.ev_fork, c.st));
  UPNERF_CHECK_CUDA(cudaStreamWaitEvent(c.side, c.ev_fork, 0));
  return UPNERF_OK;
}
// main stream waits for everything launched on the side stream so far
int join_side(const Ctx& c) {
  if (!c.has_side()) return UPNERF_OK;
  UPNERF_CHECK_CUDA(cudaEventRecord(c.ev_join, c.side));
  UPNERF_CHECK_CUDA(cudaStreamWaitEvent(c.st, c.ev_join, 0));
  return UPNERF_OK;
}

inline void* col(void* p, int64_t c, size_t es) { return static_cast<uint8_t*>(p) + c * es; }
inline const void* col(const void* p, int64_t c, size_t es) { return static_cast<const uint8_t*>(p) + c * es; }

// C[M,N] = epi(A[M,K] B[N,K]^T) with T operands
int linear(const Ctx& c, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
           int64_t M, int N, int K, upnerf_epilogue ep) {
  if (c.dtype == UPNERF_BF16) return upnerf_gemm_bf16(A, lda, B, ldb, C, ldc, M, N, K, &ep, c.st);
  upnerf_epilogue e2 = ep;
  e2.n_heads = 0;
  UPNERF_TRY(upnerf_gemm_f32(static_cast<const float*>(A), lda, 1, static_cast<const float*>(B), ldb, 1,
                             static_cast<float*>(C), ldc, 1, M, N, K, &e2, 0, 1, c.st));
  if (ep.n_heads > 0)
    UPNERF_TRY(rowdot_head(static_cast<const float*>(C), ldc, M, N, ep.n_heads, ep.head_w, ep.head_b,
                           ep.head_act, ep.head_out, c.st));
  return UPNERF_OK;
}

struct Seg { int src, len, dst; };

// dW[n, map(k)] += dY^T X ; db[n] += colsum(dY)
int wgrad(const Ctx& c, const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw,
          float* db, int64_t M, int N, int K, const Seg* segs, int nseg) {
  if (c.dtype == UPNERF_BF16) {
    int src[4], len[4], dst[4];
    for (int i = 0; i < nseg; ++i) { src[i] = segs[i].src; len[i] = segs[i].len; dst[i] = segs[i].dst; }
    return wgrad_launch(dY, lddy, X, ldx, dW, lddw, nullptr, 0, db, M, N, K, nseg, src, len, dst, c.wb, c.st);
  }
  const float* y = static_cast<const float*>(dY);
  const float* x = static_cast<const float*>(X);
  for (int i = 0; i < nseg; ++i) {
    const int64_t tiles = ceil_div64(N, 64) * ceil_div64(segs[i].len, 64);
    int split = static_cast<int>(ceil_div64(4 * sm_count(), tiles));
    const int64_t maxsplit = ceil_div64(M, 256);
    if (split > maxsplit) split = static_cast<int>(maxsplit);
    if (split < 2) split = 2;  // split_k >= 2 selects the atomic accumulate path
    UPNERF_TRY(upnerf_gemm_f32(y, 1, lddy, x + segs[i].src, 1, ldx, dW + segs[i].dst, lddw, 1, N,
                               segs[i].len, M, nullptr, 0, split, c.st));
  }
  if (db) UPNERF_TRY(rowscale_colsum(dY, lddy, nullptr, M, N, db, nullptr, c.dtype, c.st));
  return UPNERF_OK;
}

// small fp32 product with arbitrary strides (per-ray work and parameter-space chain rules)
int mm(const Ctx& c, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk,
       float* C, int64_t scm, int64_t scn, int64_t M, int64_t N, int64_t K, const upnerf_epilogue* ep,
       int accumulate, int split_k = 1) {
  // production mode: tensor cores with tf32 operands; validation mode: fp32 FMA
  if (c.dtype == UPNERF_BF16) {
    // an accumulating product without an epilogue may as well be split over K (atomics into the existing
    // values): the parameter-space chain rules have 1..3 output tiles and would otherwise run 8 K-chunks in a row
    if (accumulate && !ep && split_k == 1 && K > 32) split_k = static_cast<int>(K / 32 < 8 ? K / 32 : 8);
    // one CTA per SM: keep tiles x splits within ONE wave (6 tiles x 32 splits = 192 CTAs ran as two)
    if (split_k > 2) {
      const int64_t tiles = ceil_div64(M, 128) * ceil_div64(N, 128);
      const int64_t cap = sm_count() / tiles;
      if (cap >= 2 && split_k > cap) split_k = static_cast<int>(cap);
    }
    return upnerf_gemm_tf32(A, sam, sak, B, sbn, sbk, C, scm, scn, M, N, K, ep, accumulate, split_k, c.st);
  }
  return upnerf_gemm_f32(A, sam, sak, B, sbn, sbk, C, scm, scn, M, N, K, ep, accumulate, split_k, c.st);
}
// out[n] += sum_m s[m] X[m,n] for fp32 X of any width (falls back to the GEMM for odd widths)
int colsum_any(const float* X, int64_t ld, const float* sc, int64_t M, int N, float* out, cudaStream_t st) {
  if (N % 128 == 0 || (N <= 256 && N % 2 == 0 && 256 % (N / 2) == 0))
    return rowscale_colsum(X, ld, sc, M, N, out, nullptr, UPNERF_F32, st);
  return upnerf_gemm_f32(sc, 0, 1, X, 1, ld, out, 0, 1, 1, N, M, nullptr, 0, 64, st);
}
int split_for(int64_t K) {
  int64_t s = ceil_div64(K, 128);
  if (s < 2) s = 2;
  if (s > 256) s = 256;
  return static_cast<int>(s);
}

upnerf_epilogue ep_none() {
  upnerf_epilogue e;
  memset(&e, 0, sizeof(e));
  return e;
}

// ------------------------------------------------------------------ weight packing
// The trunk operands are packed on the main stream (the fused trunk needs them first); the head
// operands -- behind the folded-matrix product Wq = W_rgb0[:, :F] W_sf -- on the side stream `cl`,
// which the caller joins before the first head layer.
int pack_weights(const Ctx& c, const Ctx& cl, const upnerf_net_config& cfg, const NetLayout& L, const Phase& ph,
                 const float* prm, Packed& k, bool reuse) {
  if (reuse) return fork_side(c);   // inference chunk after the first: everything below is still in the workspace
  UPNERF_CHECK_CUDA(cudaMemsetAsync(k.region, 0, k.region_bytes, c.st));
  UPNERF_TRY(upnerf_c2f_weights(prm + L.progress, cfg.c2f_start, cfg.c2f_end, cfg.use_c2f, cfg.xyz_L,
                                k.band_xyz, c.st));
  UPNERF_TRY(upnerf_c2f_weights(prm + L.progress, cfg.c2f_start, cfg.c2f_end, cfg.use_c2f, cfg.dir_L,
                                k.band_dir, c.st));
  PackList pl;
  pl.n = 0;
  auto add = [&](const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int rows, int cols, int tr) {
    pl.ops[pl.n++] = PackOp{src, ld_src, dst, ld_dst, rows, cols, tr};
  };
  const int ix = L.in_xyz;
  // layer 1: [256, in_xyz] -> [256, 64] (zero padded) and its transpose [64, 256]
  add(prm + L.Wl[0], ix, k.W1, WCAT, W, ix, 0);
  add(prm + L.Wl[0], ix, k.W1T, 2 * W, W, ix, 1);
  for (int i = 1; i < 8; ++i) {
    if (i == 4) continue;
    add(prm + L.Wl[i], W, k.Wk[i], WCAT, W, W, 0);
    add(prm + L.Wl[i], W, k.WkT[i], WCATT, W, W, 1);
  }
  // skip layer: reference input is [PE | h]; the packed input is [h | PE | 0]
  add(prm + L.Wl[4] + ix, W + ix, k.W5, WCAT, W, W, 0);
  add(prm + L.Wl[4], W + ix, col(k.W5, W, c.es), WCAT, W, ix, 0);
  add(prm + L.Wl[4] + ix, W + ix, k.W5T, WCATT, W, W, 1);
  add(prm + L.Wl[4], W + ix, k.W5peT, 2 * W, W, ix, 1);
  add(prm + L.Wf, W, k.WF, WCAT, W, W, 0);
  add(prm + L.Wf, W, k.WFT, WCATT, W, W, 1);
  UPNERF_TRY(run_pack(pl, c.dtype, c.st));
  UPNERF_TRY(fork_side(c));
  pl.n = 0;
  if (ph.cand) {
    add(prm + L.Wc0, W + L.cd, k.Wc1, W, H, W, 0);
    add(prm + L.Wc0, W + L.cd, k.Wc1T, 2 * H, H, W, 1);
    add(prm + L.Wc2, H, k.Wc2, H, H, H, 0);
    add(prm + L.Wc2, H, k.Wc2T, H, H, H, 1);
  }
  if (ph.rgb) {
    if (cfg.encode_feat) {
      // Wq = W_rgb0[:, :F] W_sf  (128 x 256);  bq_const = W_rgb0[:, :F] b_sf + b_rgb0
      // (not split over K: atomics would make the folded weights -- and with them every forward output --
      //  differ in the last bits from call to call)
      UPNERF_TRY(mm(cl, prm + L.Wr0, L.rgb_in, 1, prm + L.Wsf, 1, W, k.Wq32, W, 1, H, W, L.F, nullptr, 0));
      upnerf_epilogue e = ep_none();
      e.bias = prm + L.br0;
      UPNERF_TRY(mm(cl, prm + L.bsf, 0, 1, prm + L.Wr0, L.rgb_in, 1, k.bq_const, 0, 1, 1, H, L.F, &e, 0));
      add(k.Wq32, W, k.Wq, W, H, W, 0);
      add(k.Wq32, W, k.WqT, 2 * H, H, W, 1);
    } else {
      UPNERF_CHECK_CUDA(cudaMemcpyAsync(k.bq_const, prm + L.br0, H * sizeof(float),
                                        cudaMemcpyDeviceToDevice, cl.st));
      add(prm + L.Wr0, L.rgb_in, k.Wq, W, H, W, 0);
      add(prm + L.Wr0, L.rgb_in, k.WqT, 2 * H, H, W, 1);
    }
  }
  return pl.n ? run_pack(pl, c.dtype, cl.st) : UPNERF_OK;
}

// ------------------------------------------------------------------ one network, forward
// Everything of a pass that depends on parameters, rays and embeddings only -- not on the sample depths: packed
// GEMM operands (trunk on the main stream, heads on the side stream) and the per-ray biases of the head layers.
// The render forward issues the prologues of BOTH passes up front: the fine pass's side-stream chain (~100 us of
// small kernels that cannot share an SM with the tensor-core kernels) then runs in the TMEM-free window between
// the coarse head layers and the fine trunk instead of holding up the fine head layers.
int pass_prologue(const Ctx& c, const upnerf_render_args& a, const NetLayout& L, const Phase& ph,
                  const upnerf_pass_io& io, PassBufs& p, int which) {
  const upnerf_net_config& cfg = a.cfg;
  const float* prm = io.params;
  const int64_t R = p.R;
  Packed& k = p.pk;
  const Ctx cl = c.leaf();
  const bool reuse = a.no_grad && a.reuse_packed;
  UPNERF_TRY(pack_weights(c, cl, cfg, L, ph, prm, k, reuse));   // forks the side stream

  // Per-ray inputs of the head layers (embeddings, direction encoding) enter as a per-ray bias
  // W[:, cols] e_ray: computed on the side stream while the trunk runs.
  const int H2 = 2 * H;
  const bool stack = ph.cand && ph.rgb && c.dtype == UPNERF_BF16;
  float* Bc = p.Bc;
  float* Bq = stack ? p.Bc + H : p.Bq;
  const int64_t ldbias = stack ? H2 : H;
  upnerf_epilogue e = ep_none();
  if (ph.cand) {
    // per-ray bias of candidate_encoding.0: W[:, 256:] c_emb + b
    UPNERF_TRY(gather_rows(io.emb_c, a.img_idx, R, L.cd, p.Crows, L.cd, cl.st));
    e = ep_none();
    e.bias = prm + L.bc0;
    UPNERF_TRY(mm(cl, p.Crows, L.cd, 1, prm + L.Wc0 + W, W + L.cd, 1, Bc, ldbias, 1, R, H, L.cd, &e, 0));
  }
  if (ph.rgb) {
    // per-ray bias of rgb_share_layer.0: W[:, F:] [PE(dir) | a_emb] + (W[:, :F] b_sf + b)
    const int pw = L.in_dir + L.ad;
    UPNERF_TRY(upnerf_posenc_fwd(a.rays + 3, 8, R, cfg.dir_L, k.band_dir, p.P, pw, L.in_dir, UPNERF_F32, cl.st));
    if (L.ad > 0) {
      if (cfg.encode_appearance) UPNERF_TRY(gather_rows(io.emb_a, a.img_idx, R, L.ad, p.P + L.in_dir, pw, cl.st));
      else UPNERF_REQUIRE(false, UPNERF_ERR_BAD_CONFIG, "appearance_dim > 0 with encode_appearance off");
    }
    const int front = cfg.encode_feat ? L.F : W;
    e = ep_none();
    e.bias = k.bq_const;
    UPNERF_TRY(mm(cl, p.P, pw, 1, prm + L.Wr0 + front, L.rgb_in, 1, Bq, ldbias, 1, R, H, pw, &e, 0));
  }
  if (stack && !reuse) {
    // rgb_share_layer.2 row-dots see only the Q half of the stacked output
    UPNERF_CHECK_CUDA(cudaMemsetAsync(k.hw3, 0, 3 * H2 * sizeof(float), cl.st));
    UPNERF_CHECK_CUDA(cudaMemcpy2DAsync(k.hw3 + H, H2 * sizeof(float), prm + L.Wr2, H * sizeof(float),
                                        H * sizeof(float), 3, cudaMemcpyDeviceToDevice, cl.st));
  }

  if (c.has_side()) UPNERF_CHECK_CUDA(cudaEventRecord(c.ev_pro[which], c.side));
  return UPNERF_OK;
}

// ------------------------------------------------------------------ one network, forward (after its prologue)
int pass_fwd(const Ctx& c, const upnerf_render_args& a, const NetLayout& L, const Phase& ph,
             const upnerf_pass_io& io, PassBufs& p, int which) {
  const upnerf_net_config& cfg = a.cfg;
  const float* prm = io.params;
  const int64_t M = p.M, R = p.R;
  const int S = p.S;
  Packed& k = p.pk;
  const Ctx cl = c.leaf();
  const int H2 = 2 * H;
  const bool stack = ph.cand && ph.rgb && c.dtype == UPNERF_BF16;
  float* Bc = p.Bc;
  float* Bq = stack ? p.Bc + H : p.Bq;
  upnerf_epilogue e = ep_none();

  // positional encoding of x = o + d z straight into the skip buffer [H4 | PE]
  void* PE = col(p.X4, W, c.es);
  UPNERF_TRY(upnerf_points_posenc_fwd(a.rays, p.z, R, S, cfg.xyz_L, k.band_xyz, PE, X4W, PEW, c.dtype, c.st));

  // trunk + xyz_encoding_final (+ share_sigma on layer 8)
  e = ep_none();
  if (c.dtype == UPNERF_BF16) {
    // one persistent tcgen05 kernel; activations stay in shared memory between layers
    upnerf_trunk_args ta;
    memset(&ta, 0, sizeof(ta));
    ta.pe = PE; ta.ld_pe = X4W;
    ta.wcat = k.Wcat; ta.ld_w = WCAT;
    for (int i = 0; i < 8; ++i) {
      ta.bias[i] = prm + L.bl[i];
      ta.out[i] = a.no_grad ? nullptr : p.Hs[i + 1];
      ta.ld_out[i] = (i + 1 == 4) ? X4W : W;
    }
    ta.bias[8] = prm + L.bf;
    ta.out[8] = p.HF; ta.ld_out[8] = W;
    ta.sigma_w = prm + L.Ws; ta.sigma_b = prm + L.bs;
    ta.s_sigma = p.ssig;
    ta.M = M;
    ta.relu_mask = p.relu_mask;
    UPNERF_TRY(upnerf_mlp_trunk_fwd_bf16(&ta, c.st));
  } else {
    e.act = 1;
    e.bias = prm + L.bl[0];
    UPNERF_TRY(linear(c, PE, X4W, k.W1, WCAT, p.Hs[1], W, M, W, PEW, e));
    for (int i = 1; i < 8; ++i) {
      e = ep_none();
      e.act = 1;
      e.bias = prm + L.bl[i];
      const void* in = p.Hs[i];
      const int64_t ldin = (i == 4) ? X4W : W;
      void* out = p.Hs[i + 1];
      const int64_t ldout = (i + 1 == 4) ? X4W : W;
      if (i == 7) {  // share_sigma rides on layer 8's epilogue
        e.n_heads = 1;
        e.head_w = prm + L.Ws;
        e.head_b = prm + L.bs;
        e.head_act = 1;
        e.head_out = p.ssig;
      }
      if (i == 4) UPNERF_TRY(linear(c, in, ldin, k.W5, WCAT, out, ldout, M, W, X4W, e));
      else UPNERF_TRY(linear(c, in, ldin, k.Wk[i], WCAT, out, ldout, M, W, W, e));
    }
    // xyz_encoding_final (no activation)
    e = ep_none();
    e.bias = prm + L.bf;
    UPNERF_TRY(linear(c, p.Hs[8], W, k.WF, WCAT, p.HF, W, M, W, W, e));
  }

  // Head layers on HF.  candidate_encoding.0 and (folded) rgb_share_layer.0 read the same input, so
  // with both live (phase 1) they run as ONE stacked 256-wide GEMM into the side-by-side buffer
  // [G1 | Q]; the per-ray biases and head operands come from the side stream (this pass's prologue).
  if (c.has_side()) UPNERF_CHECK_CUDA(cudaStreamWaitEvent(c.st, c.ev_pro[which], 0));
  if (stack) {
    e = ep_none();
    e.act = 1;
    e.ray_bias = Bc;
    e.rows_per_ray = S;
    e.n_heads = 3;
    e.head_col_begin = H;       // (the left half of hw3 is zero: skip it)
    e.head_w = k.hw3;
    e.head_b = prm + L.br2;
    e.head_act = 2;
    e.head_out = p.rgb;
    UPNERF_TRY(linear(c, p.HF, W, k.Wcq, W, p.G1, H2, M, H2, W, e));
  } else {
    if (ph.cand) {
      e = ep_none();
      e.act = 1;
      e.ray_bias = Bc;
      e.rows_per_ray = S;
      UPNERF_TRY(linear(c, p.HF, W, k.Wc1, W, p.G1, H2, M, H, W, e));
    }
    if (ph.rgb) {
      e = ep_none();
      e.act = 1;
      e.ray_bias = Bq;
      e.rows_per_ray = S;
      e.n_heads = 3;
      e.head_w = prm + L.Wr2;
      e.head_b = prm + L.br2;
      e.head_act = 2;
      e.head_out = p.rgb;
      UPNERF_TRY(linear(c, p.HF, W, k.Wq, W, p.Q, H2, M, H, W, e));
    }
  }
  if (ph.cand) {
    e = ep_none();
    e.act = 1;
    e.bias = prm + L.bc2;
    e.n_heads = 1;
    e.head_w = prm + L.Wcs;
    e.head_b = prm + L.bcs;
    e.head_act = 1;
    e.head_out = p.csig;
    UPNERF_TRY(linear(c, p.G1, H2, k.Wc2, H, p.G2, H, M, H, H, e));
  }

  // compositing
  upnerf_composite_args ca;
  memset(&ca, 0, sizeof(ca));
  ca.R = R; ca.S = S;
  ca.cand = ph.cand; ca.stat_rgb = ph.rgb; ca.feat_mode = ph.feat_mode; ca.dtype = c.dtype;
  ca.z = p.z; ca.s_sigma = p.ssig; ca.c_sigma = p.csig; ca.rgb = p.rgb;
  ca.hf = p.HF; ca.ld_hf = W; ca.g2 = p.G2; ca.ld_g2 = H;
  ca.c_weights = io.c_weights; ca.s_weights = io.s_weights; ca.c_depth = io.c_depth;
  ca.t_weight = io.t_weight; ca.s_depth = io.s_depth; ca.s_rgb = io.s_rgb;
  ca.hf_ray = p.HFr; ca.g2_ray = p.G2r; ca.ws_sum = p.Wsum; ca.wc_sum = p.Wcsum;
  UPNERF_TRY(upnerf_composite_fwd(&ca, c.st));

  if (ph.feat) {
    // feat = W_sf (sum w hF) + b_sf sum w  [+ W_cf (sum w' g2) + b_cf sum w']
    UPNERF_REQUIRE(io.feat, UPNERF_ERR_BAD_SHAPE, "feat output missing");
    // (side stream: the caller joins once per render, so the coarse projections overlap the fine pass)
    UPNERF_TRY(fork_side(c));
    e = ep_none();
    e.rank1_row = p.Wsum;
    e.rank1_col = prm + L.bsf;
    UPNERF_TRY(mm(cl, p.HFr, W, 1, prm + L.Wsf, W, 1, io.feat, L.F, 1, R, L.F, W, &e, 0));
    if (ph.cand) {
      e = ep_none();
      e.rank1_row = p.Wcsum;
      e.rank1_col = prm + L.bcf;
      UPNERF_TRY(mm(cl, p.G2r, H, 1, prm + L.Wcf, H, 1, io.feat, L.F, 1, R, L.F, H, &e, 1));
    }
  }
  return UPNERF_OK;
}

// ------------------------------------------------------------------ one network, backward
int pass_bwd(const Ctx& c, const upnerf_render_args& a, const NetLayout& L, const Phase& ph,
             const upnerf_pass_io& io, PassBufs& p, Scratch& s, int which) {
  const upnerf_net_config& cfg = a.cfg;
  const float* prm = io.params;
  // d_params == NULL: the network is frozen (test-time optimisation, models/nerf_system_optmize.py:
  // 253-266 trains only embedding_fine_a and se3_refine) -- every weight-gradient launch is skipped
  float* g = io.d_params;
  const bool wg = g != nullptr;
  float* const dWq = s.dWq2[which];
  float* const dbq = s.dbq2[which];
  const int64_t M = p.M, R = p.R;
  const int S = p.S;
  Packed& k = p.pk;
  upnerf_epilogue e;

  // 1. per-ray feature projections
  UPNERF_REQUIRE(!ph.feat || io.g_feat, UPNERF_ERR_BAD_SHAPE,
                 "render_bwd: g_feat is required in this phase (pass zeros when unused)");
  // Weight-gradient leaves (nothing later in the pass reads them) go to the side stream `cl`; the
  // main stream joins it once, before the split reduction at the end of the pass.
  const Ctx cl = c.leaf();
  UPNERF_TRY(fork_side(c));
  if (ph.feat) {
    const float* gf = io.g_feat;
    if (wg) {
      UPNERF_TRY(mm(cl, gf, 1, L.F, p.HFr, 1, W, g + L.Wsf, W, 1, L.F, W, R, nullptr, 0, split_for(R)));
      UPNERF_TRY(colsum_any(gf, L.F, p.Wsum, R, L.F, g + L.bsf, cl.st));
      if (ph.cand) {
        UPNERF_TRY(mm(cl, gf, 1, L.F, p.G2r, 1, H, g + L.Wcf, H, 1, L.F, H, R, nullptr, 0, split_for(R)));
        UPNERF_TRY(colsum_any(gf, L.F, p.Wcsum, R, L.F, g + L.bcf, cl.st));
      }
    }
    UPNERF_TRY(mm(c, gf, L.F, 1, prm + L.Wsf, 1, W, s.gHFr, W, 1, R, W, L.F, nullptr, 0));
    UPNERF_TRY(rowdot_head(gf, L.F, R, L.F, 1, prm + L.bsf, nullptr, 0, s.gWs, c.st));
    if (ph.cand) {
      UPNERF_TRY(mm(c, gf, L.F, 1, prm + L.Wcf, 1, H, s.gG2r, H, 1, R, H, L.F, nullptr, 0));
      UPNERF_TRY(rowdot_head(gf, L.F, R, L.F, 1, prm + L.bcf, nullptr, 0, s.gWc, c.st));
    }
  }
  const bool feat_grad = ph.feat;

  // 2. compositing backward
  upnerf_composite_args ca;
  memset(&ca, 0, sizeof(ca));
  ca.R = R; ca.S = S;
  ca.cand = ph.cand; ca.stat_rgb = ph.rgb; ca.dtype = c.dtype;
  ca.feat_mode = feat_grad ? ph.feat_mode : 0;
  ca.z = p.z; ca.s_sigma = p.ssig; ca.c_sigma = p.csig; ca.rgb = p.rgb;
  ca.hf = p.HF; ca.ld_hf = W; ca.g2 = p.G2; ca.ld_g2 = H;
  ca.g_c_weights = io.g_c_weights; ca.g_s_weights = io.g_s_weights; ca.g_c_depth = io.g_c_depth;
  ca.g_t_weight = io.g_t_weight; ca.g_s_depth = io.g_s_depth; ca.g_s_rgb = io.g_s_rgb;
  ca.g_hf_ray = s.gHFr; ca.g_g2_ray = s.gG2r; ca.g_ws_sum = s.gWs; ca.g_wc_sum = s.gWc;
  ca.w_csigma = prm + L.Wcs;
  ca.d_ssig_pre = s.dssig; ca.d_csig_pre = s.dcsig; ca.d_rgb = s.drgb;
  ca.d_hf = s.dHF; ca.ld_dhf = W; ca.d_g2pre = s.dG2p; ca.ld_dg2 = H;
  UPNERF_TRY(upnerf_composite_bwd(&ca, c.st));
  bool dhf_live = feat_grad;          // does dHF hold a gradient yet?
  bool dg2_from_feat = feat_grad && ph.cand;

  const int H2 = 2 * H;
  const bool stack = ph.cand && ph.rgb && c.dtype == UPNERF_BF16;
  const int pw = L.in_dir + L.ad;
  const int front = cfg.encode_feat ? L.F : W;

  // 3. rgb head, part 1: through the sigmoid / row-dots / ReLU -> dQ_pre (right half of [dG1p | dQp])
  if (ph.rgb) {
    UPNERF_TRY(rgb_head_bwd(p.Q, H2, p.rgb, s.drgb, prm + L.Wr2, R, S, s.dQp, H2, s.dBq,
                            wg ? g + L.Wr2 : s.gdump, wg ? g + L.br2 : s.gdump + 3 * H, c.dtype, c.st));
    // per-ray bias path: [PE(dir) | a_emb] columns of rgb_share_layer.0 and the appearance table
    UPNERF_TRY(fork_side(c));
    if (wg)
      UPNERF_TRY(mm(cl, s.dBq, 1, H, p.P, 1, pw, g + L.Wr0 + front, L.rgb_in, 1, H, pw, R, nullptr, 0, split_for(R)));
    if (L.ad > 0 && io.d_emb_a) {
      UPNERF_TRY(mm(cl, s.dBq, H, 1, prm + L.Wr0 + front + L.in_dir, 1, L.rgb_in, s.dP, L.ad, 1, R, L.ad, H, nullptr, 0));
      UPNERF_TRY(scatter_add_rows(s.dP, L.ad, a.img_idx, R, L.ad, io.d_emb_a, cl.st));
    }
    if (wg) {
      UPNERF_CHECK_CUDA(cudaMemsetAsync(dbq, 0, H * sizeof(float), cl.st));
      UPNERF_TRY(rowscale_colsum(s.dBq, H, nullptr, R, H, dbq, nullptr, UPNERF_F32, cl.st));
      UPNERF_TRY(rowscale_colsum(s.dBq, H, nullptr, R, H, g + L.br0, nullptr, UPNERF_F32, cl.st));
      if (cfg.encode_feat) UPNERF_CHECK_CUDA(cudaMemsetAsync(dWq, 0, H * W * sizeof(float), c.st));
    }
  }

  // 4. candidate head, part 1: candidate_sigma and candidate_encoding.2 -> dG1_pre (left half)
  if (ph.cand) {
    UPNERF_REQUIRE(dg2_from_feat, UPNERF_ERR_BAD_CONFIG, "candidate head without the feature path");
    // candidate_sigma: dw += sum dcsig g2, db += sum dcsig
    const Seg sgH{0, H, 0};
    if (wg) {
      UPNERF_TRY(fork_side(c));
      UPNERF_TRY(rowscale_colsum(p.G2, H, s.dcsig, M, H, g + L.Wcs, g + L.bcs, c.dtype, cl.st));
      UPNERF_TRY(wgrad(c, s.dG2p, H, p.G1, H2, g + L.Wc2, H, g + L.bc2, M, H, H, &sgH, 1));
    }
    e = ep_none();
    e.aux = p.G1; e.ldaux = H2; e.aux_mode = 2;
    UPNERF_TRY(linear(c, s.dG2p, H, k.Wc2T, H, s.dG1p, H2, M, H, H, e));
    UPNERF_TRY(ray_sum128(s.dG1p, H2, R, S, s.dBc, c.dtype, c.st));
    UPNERF_TRY(fork_side(c));
    if (wg) {
      UPNERF_TRY(mm(cl, s.dBc, 1, H, p.Crows, 1, L.cd, g + L.Wc0 + W, W + L.cd, 1, H, L.cd, R, nullptr, 0, split_for(R)));
      UPNERF_TRY(rowscale_colsum(s.dBc, H, nullptr, R, H, g + L.bc0, nullptr, UPNERF_F32, cl.st));
    }
    if (io.d_emb_c) {
      UPNERF_TRY(mm(cl, s.dBc, H, 1, prm + L.Wc0 + W, 1, W + L.cd, s.dCrows, L.cd, 1, R, L.cd, H, nullptr, 0));
      UPNERF_TRY(scatter_add_rows(s.dCrows, L.cd, a.img_idx, R, L.cd, io.d_emb_c, cl.st));
    }
  }

  // frozen weights and no pose gradient wanted: only the embedding tables needed a gradient and
  // they have it -- the trunk backward is dead work
  if (!wg && !a.d_rays) return join_side(c);

  // 4b. the two head layers on HF: weight gradients and dHF (+)= [dG1p | dQp] [Wc1; Wq]
  float* dWq_dst = cfg.encode_feat ? dWq : (wg ? g + L.Wr0 : nullptr);
  const int64_t ld_dWq = cfg.encode_feat ? W : L.rgb_in;
  const Seg sgW{0, W, 0};
  if (stack) {
    int src = 0, len = W, dst = 0;
    if (wg)
      UPNERF_TRY(wgrad_launch(s.dG1p, H2, p.HF, W, g + L.Wc0, W + L.cd, dWq_dst, ld_dWq, nullptr, M, H2, W, 1, &src,
                              &len, &dst, c.wb, c.st));
    e = ep_none();
    if (dhf_live) { e.aux = s.dHF; e.ldaux = W; e.aux_mode = 1; }
    UPNERF_TRY(linear(c, s.dG1p, H2, k.WcqT, H2, s.dHF, W, M, W, H2, e));
    dhf_live = true;
  } else {
    if (ph.rgb) {
      if (wg) UPNERF_TRY(wgrad(c, s.dQp, H2, p.HF, W, dWq_dst, ld_dWq, nullptr, M, H, W, &sgW, 1));
      e = ep_none();
      if (dhf_live) { e.aux = s.dHF; e.ldaux = W; e.aux_mode = 1; }
      UPNERF_TRY(linear(c, s.dQp, H2, k.WqT, H2, s.dHF, W, M, W, H, e));
      dhf_live = true;
    }
    if (ph.cand) {
      if (wg) UPNERF_TRY(wgrad(c, s.dG1p, H2, p.HF, W, g + L.Wc0, W + L.cd, nullptr, M, H, W, &sgW, 1));
      e = ep_none();
      if (dhf_live) { e.aux = s.dHF; e.ldaux = W; e.aux_mode = 1; }
      UPNERF_TRY(linear(c, s.dG1p, H2, k.Wc1T, H2, s.dHF, W, M, W, H, e));
      dhf_live = true;
    }
  }

  // 5. xyz_encoding_final + share_sigma  ->  dH8_pre
  if (wg) {
    UPNERF_TRY(fork_side(c));
    UPNERF_TRY(rowscale_colsum(p.Hs[8], W, s.dssig, M, W, g + L.Ws, g + L.bs, c.dtype, cl.st));
  }
  UPNERF_REQUIRE(dhf_live, UPNERF_ERR_BAD_CONFIG, "backward without any gradient into xyz_encoding_final");
  if (wg) {
    const Seg sgW{0, W, 0};
    UPNERF_TRY(wgrad(c, s.dHF, W, p.Hs[8], W, g + L.Wf, W, g + L.bf, M, W, W, &sgW, 1));
  }
  const void* PE = col(p.X4, W, c.es);
  const bool fused = c.dtype == UPNERF_BF16;
  if (fused) {
    // the whole data-gradient chain dHF -> dY8 .. dY1 in one persistent tcgen05 kernel
    upnerf_trunk_bwd_args ba;
    memset(&ba, 0, sizeof(ba));
    ba.d_hf = s.dHF; ba.ld_dhf = W;
    ba.d_ssig = s.dssig; ba.sigma_w = prm + L.Ws;
    ba.wcat_t = k.WcatT; ba.ld_w = WCATT;
    ba.relu_mask = p.relu_mask;
    for (int j = 0; j < 8; ++j) { ba.d_out[j] = s.dY[j]; ba.ld_dout[j] = W; }
    ba.M = M;
    UPNERF_TRY(upnerf_mlp_trunk_bwd_bf16(&ba, c.st));
  } else {
    e = ep_none();
    e.rank1_row = s.dssig; e.rank1_col = prm + L.Ws;
    e.aux = p.Hs[8]; e.ldaux = W; e.aux_mode = 2;
    UPNERF_TRY(linear(c, s.dHF, W, k.WFT, WCATT, s.dY[0], W, M, W, W, e));
  }

  // 6. trunk, layers 8..1: weight gradients (and, layer-wise mode, the data gradients)
  bool dpe_live = false;
  for (int i = 7; i >= 0; --i) {
    // dcur = dH_{i+1}_pre ; input of layer i+1 is Hs[i] (or PE for i == 0, X4 for i == 4)
    void* dcur = fused ? s.dY[7 - i] : s.dY[(7 - i) & 1];
    void* dnext = fused ? nullptr : s.dY[(8 - i) & 1];
    if (i == 4) {
      const Seg sg[2] = {{0, W, L.in_xyz}, {W, L.in_xyz, 0}};
      if (wg) UPNERF_TRY(wgrad(c, dcur, W, p.X4, X4W, g + L.Wl[4], W + L.in_xyz, g + L.bl[4], M, W, X4W, sg, 2));
      if (!fused) {
        e = ep_none();
        e.aux = p.X4; e.ldaux = X4W; e.aux_mode = 2;
        UPNERF_TRY(linear(c, dcur, W, k.W5T, WCATT, dnext, W, M, W, W, e));
      }
      if (a.d_rays && !fused) {
        e = ep_none();
        UPNERF_TRY(linear(c, dcur, W, k.W5peT, 2 * W, s.dPE, PEW, M, PEW, W, e));
        dpe_live = true;
      }
    } else if (i == 0) {
      const Seg sg{0, L.in_xyz, 0};
      if (wg) UPNERF_TRY(wgrad(c, dcur, W, PE, X4W, g + L.Wl[0], L.in_xyz, g + L.bl[0], M, W, PEW, &sg, 1));
      if (a.d_rays && fused) {
        // dPE = dY5 W5[:, pe] + dY1 W1 in one pass over [dY5 | dY1] (the two layers that read the encoding)
        e = ep_none();
        UPNERF_TRY(upnerf_gemm2_bf16(s.dY[3], W, W, dcur, W, W, k.W5peT, 2 * W, s.dPE, PEW, M, PEW, &e, c.st));
      } else if (a.d_rays) {
        e = ep_none();
        if (dpe_live) { e.aux = s.dPE; e.ldaux = PEW; e.aux_mode = 1; }
        UPNERF_TRY(linear(c, dcur, W, k.W1T, 2 * W, s.dPE, PEW, M, PEW, W, e));
      }
    } else {
      const Seg sg{0, W, 0};
      if (wg) UPNERF_TRY(wgrad(c, dcur, W, p.Hs[i], W, g + L.Wl[i], W, g + L.bl[i], M, W, W, &sg, 1));
      if (!fused) {
        e = ep_none();
        e.aux = p.Hs[i]; e.ldaux = W; e.aux_mode = 2;
        UPNERF_TRY(linear(c, dcur, W, k.WkT[i], WCATT, dnext, W, M, W, W, e));
      }
    }
  }

  // 7. positional encoding backward -> ray gradients
  if (a.d_rays)
    UPNERF_TRY(upnerf_points_posenc_bwd(s.dPE, PEW, a.rays, p.z, R, S, cfg.xyz_L, k.band_xyz, a.d_rays,
                                        c.dtype, c.st));

  // 8. every tcgen05 weight gradient of this pass has parked its split partials: one reduction
  //    launch sums them (fixed order) into the parameter gradients
  //    (the grouped weight-gradient launch does not depend on the side stream's leaves: it goes first, so they
  //    keep running beside it; the reduction writes parameter gradients and waits for them)
  if (c.wb) UPNERF_TRY(wgrad_flush(c.wb, c.st));
  UPNERF_TRY(join_side(c));
  if (c.wb) UPNERF_TRY(wgrad_reduce(c.wb, c.st));

  // 9. rgb head, part 2 (needs dWq from the reduction): chain rule through the folded matrix
  //    Wq = W_rgb0[:, :F] W_sf
  //    (side stream again; it trails into the next pass and is joined at the end of the render backward)
  if (wg && ph.rgb && cfg.encode_feat) {
    UPNERF_TRY(fork_side(c));
    UPNERF_TRY(mm(cl, dWq, W, 1, prm + L.Wsf, W, 1, g + L.Wr0, L.rgb_in, 1, H, L.F, W, nullptr, 1));
    UPNERF_TRY(mm(cl, prm + L.Wr0, 1, L.rgb_in, dWq, 1, W, g + L.Wsf, W, 1, L.F, W, H, nullptr, 1));
    // bq_const = W_rgb0[:, :F] b_sf + b_rgb0
    UPNERF_TRY(mm(cl, dbq, 0, 1, prm + L.Wr0, 1, L.rgb_in, g + L.bsf, 0, 1, 1, L.F, H, nullptr, 1));
    UPNERF_TRY(mm(cl, dbq, 1, 0, prm + L.bsf, 1, 0, g + L.Wr0, L.rgb_in, 1, H, L.F, 1, nullptr, 1));
  }
  return UPNERF_OK;
}

int check_common(const upnerf_render_args& a) {
  UPNERF_REQUIRE(upnerf_device_ok(), UPNERF_ERR_CUDA,
                 "upnerf_b200 needs a compute-capability 10.x GPU (sm_100a); there is no fallback");
  UPNERF_REQUIRE(a.rays && a.img_idx && a.coarse.params, UPNERF_ERR_BAD_SHAPE, "render: rays/img_idx/params missing");
  UPNERF_REQUIRE(a.n_importance == 0 || a.fine.params, UPNERF_ERR_BAD_SHAPE, "render: fine params missing");
  UPNERF_REQUIRE(a.workspace, UPNERF_ERR_WORKSPACE, "render: workspace missing");
  return UPNERF_OK;
}

}  // namespace
}  // namespace upnerf

extern "C" {

int64_t upnerf_nerf_param_count(const upnerf_net_config* cfg) {
  upnerf::NetLayout L;
  if (upnerf::make_layout(*cfg, &L) != 0) return -1;
  return L.total;
}

uint64_t upnerf_render_workspace_bytes(const upnerf_render_args* a) {
  upnerf::Plan pl;
  if (upnerf::make_plan(*a, nullptr, &pl) != 0) return 0;
  return pl.bytes;
}

int upnerf_render_fwd(const upnerf_render_args* a, void* stream) {
  using namespace upnerf;
  UPNERF_TRY(check_common(*a));
  Plan pl;
  UPNERF_TRY(make_plan(*a, a->workspace, &pl));
  UPNERF_REQUIRE(a->workspace_bytes >= pl.bytes, UPNERF_ERR_WORKSPACE,
                 "workspace too small: %llu < %llu bytes", (unsigned long long)a->workspace_bytes,
                 (unsigned long long)pl.bytes);
  const Ctx c = make_ctx(a->dtype, pl.es, as_stream(stream), nullptr);
  const Phase ph = make_phase(a->cfg, a->sched_mult);
  const int64_t R = a->n_rays;
  const int S = a->n_samples;

  UPNERF_TRY(upnerf_stratified_z(a->rays, a->perturb_rand, a->perturb, a->use_disp, R, S, pl.coarse.z, c.st));
  if (a->z_coarse)
    UPNERF_CHECK_CUDA(cudaMemcpyAsync(a->z_coarse, pl.coarse.z, R * S * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
  UPNERF_TRY(pass_prologue(c, *a, pl.L, ph, a->coarse, pl.coarse, 0));
  if (a->n_importance > 0) UPNERF_TRY(pass_prologue(c, *a, pl.L, ph, a->fine, pl.fine, 1));
  UPNERF_TRY(pass_fwd(c, *a, pl.L, ph, a->coarse, pl.coarse, 0));
  if (a->n_importance == 0) return join_side(c);

  // hierarchical resampling (models/rendering.py:262-307)
  const float* w0 = nullptr;
  const float* w1 = nullptr;
  int n0 = a->n_importance, n1 = 0;
  const bool cand_fine = a->cfg.encode_candidate && a->cfg.candidate_dim > 0;
  if (cand_fine && a->sched_mult == 0.f) {
    w0 = a->coarse.c_weights;
  } else if (cand_fine && a->sched_mult > 0.f && a->sched_mult < 1.f) {
    n1 = a->n_importance_static;
    n0 = a->n_importance - n1;
    w0 = a->coarse.c_weights;
    w1 = a->coarse.s_weights;
  } else {
    w0 = a->coarse.s_weights;
  }
  UPNERF_REQUIRE(w0 && (n1 == 0 || w1), UPNERF_ERR_BAD_SHAPE, "render_fwd: coarse weights output missing");
  UPNERF_REQUIRE(n0 >= 0 && n1 >= 0, UPNERF_ERR_BAD_SHAPE, "render_fwd: bad importance split");
  UPNERF_TRY(upnerf_resample_merge(pl.coarse.z, w0 + 1, w1 ? w1 + 1 : nullptr, S, a->u0, a->u1, n0, n1, R, S,
                                   1e-5f, pl.fine.z, c.st));
  if (a->z_fine)
    UPNERF_CHECK_CUDA(cudaMemcpyAsync(a->z_fine, pl.fine.z, R * pl.S_f * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
  UPNERF_TRY(pass_fwd(c, *a, pl.L, ph, a->fine, pl.fine, 1));
  return join_side(c);
}

// passes: bit 0 = the fine network's backward, bit 1 = the coarse network's.  Each call joins the side
// stream before it returns, so every gradient of the passes it ran is final on `stream` afterwards.
int upnerf_render_bwd_passes(const upnerf_render_args* a, int passes, void* stream) {
  using namespace upnerf;
  UPNERF_TRY(check_common(*a));
  Plan pl;
  UPNERF_TRY(make_plan(*a, a->workspace, &pl));
  UPNERF_REQUIRE(!a->no_grad, UPNERF_ERR_BAD_CONFIG, "render_bwd: the forward ran with no_grad (nothing was kept)");
  UPNERF_REQUIRE(a->workspace_bytes >= pl.bytes, UPNERF_ERR_WORKSPACE, "workspace too small");
  WgradBatch wb;
  memset(&wb, 0, sizeof(wb));
  wb.pool = pl.scratch.wg_pool;
  wb.pool_floats = pl.scratch.wg_pool_floats;
  wb.defer = 1;   // all weight gradients of a pass go out as ONE grouped launch in front of the split reduction
  const Ctx c = make_ctx(a->dtype, pl.es, as_stream(stream), wb.pool ? &wb : nullptr);
  const Phase ph = make_phase(a->cfg, a->sched_mult);
  // A pass none of whose outputs received a gradient contributes exact zeros everywhere (the fine
  // depths come from DETACHED coarse weights, models/rendering.py:268-276): it is skipped.  That is
  // the whole coarse pass of test-time optimisation, whose loss reads s_rgb_fine only
  // (models/nerf_system_optmize.py:128).
  auto live = [](const upnerf_pass_io& io) {
    return io.g_c_weights || io.g_s_weights || io.g_c_depth || io.g_s_depth || io.g_t_weight || io.g_feat ||
           io.g_s_rgb;
  };
  if ((passes & 1) && a->n_importance > 0 && live(a->fine))
    UPNERF_TRY(pass_bwd(c, *a, pl.L, ph, a->fine, pl.fine, pl.scratch, 0));
  if ((passes & 2) && live(a->coarse)) UPNERF_TRY(pass_bwd(c, *a, pl.L, ph, a->coarse, pl.coarse, pl.scratch, 1));
  return join_side(c);
}

int upnerf_render_bwd(const upnerf_render_args* a, void* stream) { return upnerf_render_bwd_passes(a, 3, stream); }

}  // extern "C"
