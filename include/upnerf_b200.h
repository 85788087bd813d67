/*
 * upnerf_b200 -- C ABI of the B200-native (sm_100a) UP-NeRF train/render hot path.
 *
 * The reference (mlvlab/UP-NeRF) is pure Python/PyTorch and has no FFI of its own; the
 * boundary it exposes for this path is a set of Python callables.
 * Each entry point below names the reference callable (file:line under the reference
 * tree) whose arithmetic it replaces.  The Python mirror of the reference interface
 * (upnerf_b200/models/rendering.py, models/nerf.py, utils/ray.py, utils/camera.py) binds
 * these symbols with ctypes; see INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are dense row-major unless a leading dimension (ld*) is given, in elements;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously;
 *   - return value: 0 on success, otherwise a upnerf_status code; the message of the
 *     last failure on the calling thread is returned by upnerf_last_error();
 *   - there is no CPU fallback: a missing GPU or unsupported shape is an error.
 */
#ifndef UPNERF_B200_H_
#define UPNERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum upnerf_status {
  UPNERF_OK = 0,
  UPNERF_ERR_BAD_SHAPE = 1,   /* a size/alignment the kernels do not support       */
  UPNERF_ERR_BAD_CONFIG = 2,  /* an architecture/phase combination not implemented */
  UPNERF_ERR_CUDA = 3,        /* a CUDA runtime/driver call failed                 */
  UPNERF_ERR_WORKSPACE = 4    /* caller-provided workspace too small               */
} upnerf_status;

typedef enum upnerf_dtype {
  UPNERF_F32 = 0,  /* validation mode: fp32 SIMT FMA GEMMs, fp32 activations       */
  UPNERF_BF16 = 1  /* production mode: tcgen05 bf16 MMA, fp32 accumulate in TMEM    */
} upnerf_dtype;

const char* upnerf_last_error(void);
int upnerf_version(void);
/* 1 if the current device is compute capability 10.x (tcgen05 available). */
int upnerf_device_ok(void);

/* Launch accounting for bench.py: total kernels launched by this library in the process,
 * and (while enabled) CUDA-event device time / launches / declared work per kernel family:
 * 0 gemm_tc, 1 wgrad_tc, 2 gemm_simt, 3 composite, 4 posenc, 5 sampling, 6 pose_rays,
 * 7 heads, 8 pack, 9 trunk_fwd (fused), 10 trunk_bwd (fused).  work = flop of the launch
 * (2*M*N*K per GEMM), bytes = its algorithmic memory traffic (every operand touched once);
 * both 0 for families that do not declare them. */
/* Measurement helper (tools/write_bw.py): fills `bytes` with an incompressible hash pattern using
 * streaming 16-byte stores -- the write-only HBM ceiling the store-heavy kernels are judged against. */
int upnerf_fill_pattern(void* dst, int64_t bytes, uint32_t seed, void* stream);
/* Measurement helper: TMA-store ceiling for 128 x 64 bf16 boxes written into a [rows, ld] bf16 matrix
 * (ld = 256: the fused trunk's activation-store pattern; ld = 64: contiguous 16 KB boxes), with 1..4
 * bulk stores in flight per CTA (depth 1 exposes the issue -> shared-memory-read-done latency). */
int upnerf_tma_store_probe(void* dst, int64_t rows, int64_t ld, int depth /* stores in flight, 1..4 */, void* stream);
long long upnerf_launch_count(void);
/* A CUDA graph captured from this library's launches was replayed: credit its `n` kernel nodes to the counter
 * (the library does not see replays; the host that captured the graph counted the launches under capture). */
void upnerf_launch_count_add(long long n);
void upnerf_profile_enable(int on);
int upnerf_profile_collect(double* ms, long long* launches, double* work, double* bytes, int ncat);

/* ------------------------------------------------------------------------------------
 * Dense layer primitive (the "one dense contraction" of the path).
 * Replaces nn.Linear (+ReLU/Softplus/Sigmoid) calls of NeRF.forward
 * (models/nerf.py:84-123) and their autograd backward.
 *
 *   C[m,n] = epi( sum_k A[m,k] * B[n,k] )          A:[M,K] lda, B:[N,K] ldb, C:[M,N] ldc
 *
 * Epilogue, applied in this order on the fp32 accumulator v:
 *   v += bias[n]; v += ray_bias[m / rows_per_ray, n]; v += rank1_row[m] * rank1_col[n];
 *   aux_mode 1: v += aux[m,n];   act 1: v = max(v,0);   aux_mode 2: v = aux[m,n] > 0 ? v : 0;
 *   heads: head_out[m,h] = head_act( sum_n v[m,n] * head_w[h,n] + head_b[h] ),  h < n_heads
 * ------------------------------------------------------------------------------------ */
typedef struct upnerf_epilogue {
  const float* bias;      /* [N] or NULL */
  const float* ray_bias;  /* [ceil(M/rows_per_ray), N] or NULL */
  int rows_per_ray;
  const float* rank1_row; /* [M] or NULL */
  const float* rank1_col; /* [N] */
  const void* aux;        /* [M, ldaux], element type of C, or NULL */
  int64_t ldaux;
  int aux_mode;           /* 0 none, 1 add, 2 relu-mask */
  int act;                /* 0 none, 1 relu */
  int n_heads;            /* 0..3 */
  const float* head_w;    /* [n_heads, N] */
  const float* head_b;    /* [n_heads] */
  int head_act;           /* 0 none, 1 softplus(beta=1,threshold=20), 2 sigmoid */
  float* head_out;        /* [M, n_heads] */
  int head_col_begin;     /* the heads read columns [head_col_begin, N) only (a multiple of 64; head_w entries of the
                             columns before it are ignored): the stacked candidate|rgb layer feeds its rgb row-dots
                             from the right half of its output */
} upnerf_epilogue;

/* bf16 in / bf16 out on tcgen05.  Requires K % 64 == 0, N % 64 == 0, N <= 256,
 * 16-byte aligned pointers and leading dimensions that are multiples of 8. */
int upnerf_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                     int64_t M, int N, int K, const upnerf_epilogue* ep, void* stream);
/* The same layer with its input split over TWO row-major operands along K: C = epi([A1 | A2] B^T), B [N, K1+K2].
 * One pass instead of two accumulating launches for sums of products over the same rows -- the positional-encoding
 * gradient dPE = dY5 W5[:, pe] + dY1 W1 (the two layers that read the encoding, models/nerf.py:84-93). */
int upnerf_gemm2_bf16(const void* A1, int64_t lda1, int K1, const void* A2, int64_t lda2, int K2, const void* B,
                      int64_t ldb, void* C, int64_t ldc, int64_t M, int N, const upnerf_epilogue* ep, void* stream);

/* Fused trunk forward on tcgen05 (bf16 in, fp32 accumulate, bf16 out): per 128-sample tile
 * PE -> xyz_encoding_1..8 (Linear 256 + ReLU, skip concat [PE | h] at layer 5) ->
 * xyz_encoding_final, plus share_sigma (row-dot + Softplus) -- reference models/nerf.py:84-93.
 * Activations stay in shared memory between layers; each layer output is stored once.
 *   pe    [M, 64]  bf16, row stride ld_pe (the 63-wide encoding, zero padded)
 *   wcat  [256, UPNERF_TRUNK_WCAT_COLS] bf16 K-major, row stride ld_w, columns
 *         [W1 (64) | W2 | W3 | W4 | W5 as [h (256) | PE (64)] | W6 | W7 | W8 | W_final]
 *   bias[l] [256] fp32; out[l] [M,256] bf16 row stride ld_out[l] (l = 0..7: H1..H8, 8: final);
 *   out[l] == NULL skips that store (inference keeps only the final output)
 *   s_sigma [M] fp32 = Softplus(H8 . sigma_w + sigma_b) */
#define UPNERF_TRUNK_LAYERS 9
#define UPNERF_TRUNK_WCAT_COLS 2176
typedef struct upnerf_trunk_args {
  const void* pe;
  int64_t ld_pe;
  const void* wcat;
  int64_t ld_w;
  const float* bias[UPNERF_TRUNK_LAYERS];
  const float* sigma_w;
  const float* sigma_b;
  void* out[UPNERF_TRUNK_LAYERS];
  int64_t ld_out[UPNERF_TRUNK_LAYERS];
  float* s_sigma;
  int64_t M;
  uint32_t* relu_mask; /* optional [upnerf_trunk_mask_words(M)]: ReLU bit masks of H1..H8 for the backward chain */
} upnerf_trunk_args;
int upnerf_mlp_trunk_fwd_bf16(const upnerf_trunk_args* a, void* stream);
int64_t upnerf_trunk_mask_words(int64_t M);

/* Fused backward data-gradient chain of the trunk (autograd backward of models/nerf.py:84-93
 * with respect to the activations), one persistent tcgen05 kernel:
 *   dY8 = (d_hf . W_final + d_ssig (x) sigma_w) * [H8 > 0];  dYl = (dY(l+1) . W(l+1)[h part]) * [Hl > 0]
 *   d_hf   [M,256] bf16 (gradient of xyz_encoding_final's output), d_ssig [M] fp32 (gradient of
 *          the pre-Softplus sigma, may be NULL)
 *   wcat_t [256, UPNERF_TRUNK_WCATT_COLS] bf16, row n = input feature, columns
 *          [W_final^T | W8^T | W7^T | W6^T | W5[h part]^T | W4^T | W3^T | W2^T]
 *   relu_mask: written by upnerf_mlp_trunk_fwd_bf16 on the same M
 *   d_out[j] [M,256] bf16 = dY(8-j), j = 0..7 (gradient w.r.t. the pre-activation of layer 8-j) */
#define UPNERF_TRUNK_BWD_LAYERS 8
#define UPNERF_TRUNK_WCATT_COLS 2048
typedef struct upnerf_trunk_bwd_args {
  const void* d_hf;
  int64_t ld_dhf;
  const float* d_ssig;
  const float* sigma_w;
  const void* wcat_t;
  int64_t ld_w;
  const uint32_t* relu_mask;
  void* d_out[UPNERF_TRUNK_BWD_LAYERS];
  int64_t ld_dout[UPNERF_TRUNK_BWD_LAYERS];
  int64_t M;
} upnerf_trunk_bwd_args;
int upnerf_mlp_trunk_bwd_bf16(const upnerf_trunk_bwd_args* a, void* stream);

/* Weight gradient on tcgen05:  dW[n, colmap(k)] += sum_m dY[m,n] * X[m,k]  (fp32 atomics),
 * db[n] += sum_m dY[m,n].  dY:[M,N] bf16, X:[M,K] bf16.  N % 128 == 0 (N <= 256),
 * K % 64 == 0 (K <= 320).  Column segments map packed K columns to parameter columns:
 * packed columns [seg_src[i], seg_src[i]+seg_len[i]) go to dW columns starting at
 * seg_dst[i]; columns not covered by a segment are dropped (padding). db may be NULL. */
int upnerf_wgrad_bf16(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW,
                      int64_t lddw, float* db, int64_t M, int N, int K, int n_seg,
                      const int* seg_src_host, const int* seg_len_host, const int* seg_dst_host,
                      void* stream);

/* Deterministic variant: every CTA parks its split partial in `workspace`
 * (>= upnerf_wgrad_det_workspace_bytes(N, K)) and a second launch sums the splits in a fixed order
 * -- bit-reproducible, no atomics.  upnerf_render_bwd uses this mode for every weight gradient,
 * with ONE reduction launch per network pass. */
uint64_t upnerf_wgrad_det_workspace_bytes(int N, int K);
int upnerf_wgrad_bf16_det(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW,
                          int64_t lddw, float* db, int64_t M, int N, int K, int n_seg,
                          const int* seg_src_host, const int* seg_len_host, const int* seg_dst_host,
                          void* workspace, uint64_t workspace_bytes, void* stream);

/* Same for two layers stacked along N (N = 256): output rows 0..127 accumulate into dW_lo
 * (row stride lddw_lo), rows 128..255 into dW_hi -- the candidate / rgb head layers share
 * their input (models/nerf.py:97,106), so one pass over it serves both weight gradients. */
int upnerf_wgrad2_bf16(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW_lo,
                       int64_t lddw_lo, float* dW_hi, int64_t lddw_hi, int64_t M, int K, int n_seg,
                       const int* seg_src_host, const int* seg_len_host, const int* seg_dst_host,
                       void* stream);

/* fp32 SIMT GEMM with arbitrary element strides (validation mode and the small per-ray
 * products):  C[m*scm + n*scn] = epi( sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk] ).
 * split_k > 1 splits the k range over grid.z and accumulates with atomicAdd into C
 * (epilogue ignored, C must be pre-initialised); accumulate != 0 adds to C instead of
 * overwriting.  aux is row-major with row stride ldaux. Heads are not supported here. */
int upnerf_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn,
                    int64_t sbk, float* C, int64_t scm, int64_t scn, int64_t M, int64_t N,
                    int64_t K, const upnerf_epilogue* ep, int accumulate, int split_k,
                    void* stream);

/* The same product on tcgen05 with tf32 operands (fp32 in memory, rounded to nearest tf32 -- 10-bit mantissa -- when
 * a tile is staged; fp32 accumulate, fp32 out): the production (bf16-mode) path of the small per-ray and
 * parameter-space products -- per-ray head biases (models/nerf.py:97-113 on the embeddings of
 * models/rendering.py:255-258), the feature projections after compositing (models/nerf.py:53,76) and their
 * gradients.  Same strides / split_k / accumulate semantics as upnerf_gemm_f32; the epilogue supports bias and the
 * rank-1 term only (anything else returns UPNERF_ERR_BAD_CONFIG). */
int upnerf_gemm_tf32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C,
                     int64_t scm, int64_t scn, int64_t M, int64_t N, int64_t K, const upnerf_epilogue* ep,
                     int accumulate, int split_k, void* stream);

/* ------------------------------------------------------------------------------------
 * (a) Pose refinement + ray casting.
 * Replaces, per ray: se3_refine(img_idx) -> Lie.se3_to_SE3 (utils/camera.py:87-98),
 * Pose.compose([refine, c2w]) (utils/camera.py:43-58), get_rays (utils/ray.py:30-67) and
 * the cat with ray_infos (models/nerf_system.py:158-166).
 *   se3_table [n_images,6] (NULL: no refinement, plain get_rays), img_idx [R] int64,
 *   c2w [R,3,4] or one [3,4] (c2w_is_single), directions [R,3], near_far [R,2] or NULL
 *   -> rays [R,8] = [o, d, near, far];  pose_out [R,3,4] optional (the composed pose).
 * Backward: d_rays [R,8] (columns 0..5 used) -> d_se3_table [n_images,6] (+=, atomics).
 * ------------------------------------------------------------------------------------ */
int upnerf_pose_rays_fwd(const float* se3_table, const int64_t* img_idx, const float* c2w,
                         int c2w_is_single, const float* directions, const float* near_far,
                         int64_t n_rays, float* rays, float* pose_out, void* stream);
int upnerf_pose_rays_bwd(const float* se3_table, const int64_t* img_idx, const float* c2w,
                         int c2w_is_single, const float* directions, int64_t n_rays,
                         const float* d_rays, float* d_se3_table, void* stream);

/* Stand-alone pieces at the reference's own granularity (each differentiable on its own):
 *   upnerf_se3_exp_*       Lie.se3_to_SE3   (utils/camera.py:87-98): wu [n,6] -> pose [n,3,4]
 *   upnerf_pose_compose_*  Pose.compose_pair (utils/camera.py:51-58): out = b o a; a pose given
 *                          as one (3,4) matrix broadcasts (x_is_single); d_a / d_b are [n,3,4]
 *                          (NULL to skip; a broadcast operand's gradient is not reduced here)
 *   upnerf_get_rays_bwd    gradient of get_rays (utils/ray.py:30-67) w.r.t. the pose(s):
 *                          d_c2w [n,3,4], or one [3,4] accumulated over rays when single. */
int upnerf_se3_exp_fwd(const float* wu, int64_t n, float* pose_out, void* stream);
int upnerf_se3_exp_bwd(const float* wu, const float* d_pose, int64_t n, float* d_wu, void* stream);
int upnerf_pose_compose_fwd(const float* pose_a, int a_is_single, const float* pose_b, int b_is_single,
                            int64_t n, float* out, void* stream);
int upnerf_pose_compose_bwd(const float* pose_a, int a_is_single, const float* pose_b, int b_is_single,
                            const float* d_out, int64_t n, float* d_a, float* d_b, void* stream);
int upnerf_get_rays_bwd(const float* c2w, int c2w_is_single, const float* directions, int64_t n_rays,
                        const float* d_rays, float* d_c2w, void* stream);

/* ------------------------------------------------------------------------------------
 * (d) Depth sampling.  models/rendering.py:231-249 (stratified), :7-50 (sample_pdf),
 * :262-307 (resample + sort-merge).  Uniform random numbers are INPUTS (drawn by the host
 * in the reference's order); u == NULL means det=True (linspace).
 * ------------------------------------------------------------------------------------ */
int upnerf_stratified_z(const float* rays, const float* perturb_rand, float perturb, int use_disp,
                        int64_t n_rays, int n_samples, float* z, void* stream);
/* bins [R, n_weights+1] (row stride ld_bins), weights [R, n_weights] (row stride ld_weights),
 * u [R,N] or NULL -> samples [R,N]; optional inds [R,N] int64 (searchsorted right=True
 * result) and cdf_out [R, n_weights+1]. */
int upnerf_sample_pdf(const float* bins, int64_t ld_bins, const float* weights, int64_t ld_weights,
                      const float* u, int64_t n_rays, int n_weights, int n_importance, float eps,
                      float* samples, int64_t* inds, float* cdf_out, void* stream);
/* torch.searchsorted(cdf, u, right=True) for row-wise sorted cdf [R,n_cdf], u [R,n_u]. */
int upnerf_searchsorted_right(const float* cdf, int n_cdf, const float* u, int n_u, int64_t n_rays,
                              int64_t* inds, void* stream);
/* z [R,S] sorted coarse depths; w0/w1: coarse weights ALREADY offset to column 1 (the
 * [:,1:-1] slice of models/rendering.py:271), row stride ld_w; draws n0 from w0 and n1 from
 * w1 (w1 NULL or n1 0: single draw) and writes the sorted union z_fine [R, S+n0+n1]. */
int upnerf_resample_merge(const float* z, const float* w0, const float* w1, int64_t ld_w,
                          const float* u0, const float* u1, int n0, int n1, int64_t n_rays,
                          int n_samples, float eps, float* z_fine, void* stream);

/* ------------------------------------------------------------------------------------
 * (b) Positional encoding with the coarse-to-fine mask (models/nerf.py:126-147).
 * band_w [L] are the per-band weights; upnerf_c2f_weights computes them on the device from
 * the NeRF.progress scalar (models/nerf.py:137-142) so the host never synchronises.
 * out rows: [x(3), per coordinate sin block (L) then cos block (L)], zero padded to `width`,
 * element type by `dtype`.  ld_x is the row stride of x in floats (3 for packed xyz, 8 to
 * encode the direction columns of a rays tensor in place).
 * ------------------------------------------------------------------------------------ */
int upnerf_c2f_weights(const float* progress_dev, float start, float end, int use_c2f, int L,
                       float* band_w, void* stream);
int upnerf_posenc_fwd(const float* x, int64_t ld_x, int64_t M, int L, const float* band_w, void* out,
                      int64_t ld_out, int width, int dtype, void* stream);
/* x = o + d*z formed in registers (models/rendering.py:251,308): rays [R,8], z [R,S]. */
int upnerf_points_posenc_fwd(const float* rays, const float* z, int64_t n_rays, int n_samples, int L,
                             const float* band_w, void* out, int64_t ld_out, int width, int dtype,
                             void* stream);
/* d_pe [R*S, ld_pe] -> d_rays[:,0:3] += sum_s dx, d_rays[:,3:6] += sum_s z dx. */
int upnerf_points_posenc_bwd(const void* d_pe, int64_t ld_pe, const float* rays, const float* z,
                             int64_t n_rays, int n_samples, int L, const float* band_w,
                             float* d_rays, int dtype, void* stream);

/* ------------------------------------------------------------------------------------
 * (c) Alpha compositing, forward and backward (models/rendering.py:124-219).
 * The 384-d feature heads are linear, so the kernel composites the hidden vectors that
 * feed them (hf: 256-d output of xyz_encoding_final, g2: 128-d output of
 * candidate_encoding) and the weight sums; see csrc/composite.cu.
 * ------------------------------------------------------------------------------------ */
typedef struct upnerf_composite_args {
  int64_t R;
  int S;
  int cand;      /* candidate pass on: sched_mult < 1 and the candidate head is encoded */
  int stat_rgb;  /* static rgb output on: sched_mult > 0 */
  int feat_mode; /* 0 none, 1 features with static weights (no candidate head), 2 candidate pass */
  int dtype;     /* element type of hf/g2/d_hf/d_g2pre */
  const float* z;        /* [R,S] */
  const float* s_sigma;  /* [R*S] post-softplus */
  const float* c_sigma;  /* [R*S] */
  const float* rgb;      /* [R*S,3] post-sigmoid */
  const void* hf;        /* [R*S, ld_hf] */
  int64_t ld_hf;
  const void* g2;        /* [R*S, ld_g2] */
  int64_t ld_g2;
  /* forward outputs */
  float* c_weights;  /* [R,S]  a T            (results["c_weights_*"]) */
  float* s_weights;  /* [R,S]  a^s T^s        (results["s_weights_*"]), may be NULL */
  float* c_depth;    /* [R] */
  float* t_weight;   /* [R] */
  float* s_depth;    /* [R] */
  float* s_rgb;      /* [R,3] */
  float* hf_ray;     /* [R,256] sum_i w_i hf_i */
  float* g2_ray;     /* [R,128] sum_i a^c_i T_i g2_i */
  float* ws_sum;     /* [R] sum_i w_i */
  float* wc_sum;     /* [R] */
  /* backward: upstream gradients (NULL = zero) */
  const float* g_c_weights;
  const float* g_s_weights;
  const float* g_c_depth;
  const float* g_t_weight;
  const float* g_s_depth;
  const float* g_s_rgb;
  const float* g_hf_ray;
  const float* g_g2_ray;
  const float* g_ws_sum;
  const float* g_wc_sum;
  const float* w_csigma; /* [128] candidate_sigma.0.weight, folded into d_g2pre */
  /* backward outputs */
  float* d_ssig_pre; /* [R*S] w.r.t. the pre-softplus static sigma */
  float* d_csig_pre; /* [R*S] */
  float* d_rgb;      /* [R*S,3] w.r.t. post-sigmoid rgb */
  void* d_hf;        /* [R*S, ld_dhf] */
  int64_t ld_dhf;
  void* d_g2pre;     /* [R*S, ld_dg2], already masked by g2 > 0 */
  int64_t ld_dg2;
} upnerf_composite_args;
int upnerf_composite_fwd(const upnerf_composite_args* a, void* stream);
int upnerf_composite_bwd(const upnerf_composite_args* a, void* stream);

/* ------------------------------------------------------------------------------------
 * The whole path: render_rays (models/rendering.py:53-314) with NeRF.forward
 * (models/nerf.py:80-124) for both networks, forward and backward.
 *
 * Parameters of one NeRF are passed as ONE flat fp32 buffer in the reference's
 * state_dict() order ("progress" first; upnerf_nerf_param_count/offsets describe it);
 * gradients are accumulated (+=) into a buffer of the same layout.
 * ------------------------------------------------------------------------------------ */
typedef struct upnerf_net_config {
  int D, W;               /* 8, 256 (only this trunk is implemented; skip at layer 5) */
  int xyz_L, dir_L;       /* xyz_L <= 10 */
  int encode_feat;        /* NeRF(encode_feat=...) */
  int feat_dim;
  int appearance_dim;     /* parameter layout */
  int candidate_dim;
  int encode_appearance;  /* run-time switches (tto clears encode_candidate) */
  int encode_candidate;
  int use_c2f;
  float c2f_start, c2f_end;
} upnerf_net_config;

typedef struct upnerf_pass_io {
  const float* params;  /* flat parameters */
  const float* emb_a;   /* [n_images, appearance_dim] or NULL */
  const float* emb_c;   /* [n_images, candidate_dim] or NULL */
  /* forward outputs (NULL when the phase does not produce them) */
  float* c_weights; float* s_weights; float* c_depth; float* s_depth; float* t_weight;
  float* feat; float* s_rgb;
  /* backward: upstream gradients (NULL = zero) */
  const float* g_c_weights; const float* g_s_weights; const float* g_c_depth;
  const float* g_s_depth; const float* g_t_weight; const float* g_feat; const float* g_s_rgb;
  /* backward outputs, accumulated */
  float* d_params; float* d_emb_a; float* d_emb_c;
} upnerf_pass_io;

typedef struct upnerf_render_args {
  upnerf_net_config cfg;
  int dtype;               /* upnerf_dtype */
  int64_t n_rays;
  int n_samples;           /* coarse samples per ray */
  int n_importance;        /* 0: coarse only */
  int n_importance_static; /* round(sched_mult*n_importance), computed by the host (Python round) */
  int n_images;
  float sched_mult;        /* 0: phase 0, 1: phase 2, otherwise phase 1 */
  int use_disp;
  float perturb;
  const float* rays;         /* [R,8] */
  const int64_t* img_idx;    /* [R] */
  const float* perturb_rand; /* [R,S] when perturb > 0 */
  const float* u0;           /* uniforms of the first sample_pdf call or NULL (det) */
  const float* u1;           /* second call (phase 1) */
  upnerf_pass_io coarse, fine;
  float* z_coarse;           /* optional output [R,S] */
  float* z_fine;             /* optional output [R,S+n_importance] */
  float* d_rays;             /* backward output [R,8], accumulated, or NULL */
  void* workspace;           /* activations saved by forward for backward + scratch */
  uint64_t workspace_bytes;
  int no_grad;               /* 1: forward only (inference): activations are not kept, the
                                workspace is ~4x smaller and upnerf_render_bwd is refused */
  int reuse_packed;          /* no_grad only.  1: the packed GEMM operands, folded head matrices and c2f band
                                weights a previous upnerf_render_fwd call left in THIS workspace (same
                                pointer, same sizes and configuration) are still valid -- parameters and
                                `progress` have not changed -- and are not rebuilt.  The chunk loop of a
                                full-image render (reference models/nerf_system.py:104-126) sets it for
                                every chunk after the first. */
} upnerf_render_args;

int64_t upnerf_nerf_param_count(const upnerf_net_config* cfg);
/* Bytes of workspace upnerf_render_fwd/bwd need for these sizes (0 on error). */
uint64_t upnerf_render_workspace_bytes(const upnerf_render_args* a);
int upnerf_render_fwd(const upnerf_render_args* a, void* stream);
int upnerf_render_bwd(const upnerf_render_args* a, void* stream);
/* The same backward one network pass at a time: passes bit 0 = fine, bit 1 = coarse (3 = upnerf_render_bwd).
 * Every gradient of the passes run is final on `stream` when the call returns, so a data-parallel caller can
 * start the all-reduce of the fine network's gradients while the coarse backward runs (train.py:72). */
int upnerf_render_bwd_passes(const upnerf_render_args* a, int passes, void* stream);

/* ------------------------------------------------------------------------------------
 * (f2) Per-ray tail of the train step in ONE launch.
 * Replaces the monocular-depth affine correction of NeRFSystem.training_step
 * (models/nerf_system.py:169-177), UPNeRFLoss.forward (losses.py:21-64) together with its
 * autograd backward, and the psnr of models/nerf_system.py:202-207.  Every gradient is written for
 * an upstream gradient of 1 (what manual_backward(loss) feeds).  Pointers of terms that are dead in
 * the phase may be NULL: sched_mult < 1 needs the depth / feature inputs, sched_mult > 0 the colour
 * ones; t_weight_* NULL = no candidate head; t_beta NULL = no TransientNet.
 *   losses[0..7] = l_depth_c, l_feat_c, l_rgb_c, l_depth_f, l_feat_f, l_rgb_f, l_beta, l_alpha
 *                  (0 when absent), losses[8] = their sum in the reference's order,
 *                  losses[9] = psnr of s_rgb_{fine|coarse} (0 when absent).
 *   d_depth_scale [n_images,2] is ACCUMULATED (fp32 atomics), all other gradients are overwritten.
 *   workspace: >= upnerf_tail_workspace_bytes(), zero-filled ONCE by the caller and then reused. */
#define UPNERF_TAIL_LOSS_SLOTS 16
typedef struct upnerf_tail_args {
  int64_t n_rays;
  int feat_dim;
  int has_fine;
  float sched_mult, depth_mult, alpha_reg, near_, far_;
  const int64_t* img_idx;     /* [R] */
  const float* inv_depths;    /* [R] */
  const float* depth_scale;   /* [n_images,2] = (log scale, shift) */
  const float* rgbs;          /* [R,3] */
  const float* feats;         /* [R,F] */
  const float *s_depth_c, *s_depth_f, *t_weight_c, *t_weight_f;   /* [R] */
  const float *feat_c, *feat_f;                                   /* [R,F] */
  const float *s_rgb_c, *s_rgb_f;                                 /* [R,3] */
  const float *t_beta, *t_alpha;                                  /* [R] */
  float* losses;              /* [UPNERF_TAIL_LOSS_SLOTS] */
  float *g_s_depth_c, *g_s_depth_f, *g_feat_c, *g_feat_f, *g_s_rgb_c, *g_s_rgb_f, *g_t_beta, *g_t_alpha;
  float* d_depth_scale;
  void* workspace;
  uint64_t workspace_bytes;
  const float* sched_mult_dev; /* optional DEVICE scalar: when non-NULL the kernel reads the schedule multiplier from
                                  it instead of `sched_mult` (CUDA-graph replay: the host rewrites it every step;
                                  `sched_mult` must still lie in the same phase -- 0, (0,1) or 1 -- as the value) */
} upnerf_tail_args;
uint64_t upnerf_tail_workspace_bytes(void);
int upnerf_tail_loss(const upnerf_tail_args* a, void* stream);

/* ------------------------------------------------------------------------------------
 * (f3) Fused Adam over one flat fp32 buffer.
 * Replaces torch.optim.Adam.step of the reference's two optimisers (models/nerf_system.py:41-73,
 * 188-195; utils/optim.py:20-44: Adam(eps 1e-8), no weight decay).  The buffer is a run of
 * consecutive segments (one per parameter tensor or run of tensors with the same history); a
 * segment with seg_live = 0 is left untouched -- the reference optimiser skips tensors whose
 * .grad is None in the current schedule phase -- a live one gets torch's single-tensor update
 * with its own step_size = lr / (1 - beta1^t) and bc2_sqrt = sqrt(1 - beta2^t). */
#define UPNERF_ADAM_MAX_SEGMENTS 64
typedef struct upnerf_adam_args {
  float* params;             /* [n] updated in place */
  const float* grads;        /* [n] */
  float* exp_avg;            /* [n] first moment, updated in place */
  float* exp_avg_sq;         /* [n] second moment, updated in place */
  int64_t n;
  int n_segments;
  int64_t seg_end[UPNERF_ADAM_MAX_SEGMENTS];      /* exclusive end offsets, ascending, last == n */
  float seg_step_size[UPNERF_ADAM_MAX_SEGMENTS];
  float seg_bc2_sqrt[UPNERF_ADAM_MAX_SEGMENTS];
  int seg_live[UPNERF_ADAM_MAX_SEGMENTS];
  double beta1, beta2, eps;   /* doubles: 1 - beta is rounded to fp32 from the double, as torch does */
  double decay_mul;           /* AdamW (tto, models/nerf_system_optmize.py:61): params *= 1 - lr*weight_decay
                                 before the update, as torch.optim.AdamW does; 0 or 1 = plain Adam */
  const float* dev_scalars;   /* optional DEVICE array [2*UPNERF_ADAM_MAX_SEGMENTS + 1] = seg_step_size | seg_bc2_sqrt |
                                 decay_mul: when non-NULL it overrides the three by-value fields, so a captured CUDA
                                 graph of the step can be replayed with the host rewriting the scalars every step */
} upnerf_adam_args;
int upnerf_adam_step(const upnerf_adam_args* a, void* stream);

/* ------------------------------------------------------------------------------------
 * (f1) GPU-resident training-ray batcher.
 * Replaces PhototourismDataset.__getitem__ (split "train", datasets/phototourism.py:420-454)
 * applied to every index of a batch plus torch's default_collate, i.e. what the reference's
 * DataLoader workers do per RAY in Python (2048+ calls per step) followed by the H2D copy.
 * The per-ray tables the reference builds once (datasets/phototourism.py:213-323: all_ray_infos,
 * all_directions, all_rgbs, all_pxl_coords, all_inv_depths, feat_maps) and the per-image start
 * poses (poses_dict, :181-211) stay resident in HBM; one launch gathers a batch:
 *   img_idx   = (int64) ray_infos[idx, 2]                       (:422)
 *   ray_infos = ray_infos[idx, :2], directions, rgbs, inv_depths = table[idx]   (:423-428, :451)
 *   c2w       = poses[img_idx]                                  (:427)
 *   feats     = the reference's 4-tap interpolation of feat_maps[img_idx] at
 *               pxl_coords[idx] * (h - 1)  (:430-450), INCLUDING its border behaviour: with
 *               y2 = min(h-1, y1+1) a sample exactly on the last row/column gets all-zero weights.
 *               Products and sums are separately rounded fp32 in the reference's order
 *               ((w11 p11 + w12 p12) + w21 p21) + w22 p22 -- results are bit-identical.
 * status (optional, device int32): bit 0 is set when an index or image id is out of range (the
 * reference raises IndexError; such rays are skipped here).  HBM-bound gather: 4*F*4 bytes read and
 * F*4 written per ray for the features + ~200 B of small fields (7.9 KB/ray at F = 384). */
typedef struct upnerf_ray_batch_args {
  int64_t n_rays;              /* R: batch size */
  int64_t n_total;             /* N: rows of the per-ray tables */
  int n_images, feat_h, feat_w, feat_dim;   /* the reference asserts feat_h == feat_w (:431) */
  const int64_t* idx;          /* [R] ray indices into the tables */
  const float* ray_infos;      /* [N,3] = (near, far, image index as float) */
  const float* directions;     /* [N,3] */
  const float* rgbs;           /* [N,3] */
  const float* pxl_coords;     /* [N,2] = (y, x) in [0,1]; NULL with feat_maps NULL */
  const float* inv_depths;     /* [N] or NULL */
  const float* feat_maps;      /* [n_images, feat_h, feat_w, feat_dim] or NULL */
  const float* poses;          /* [n_images,3,4] */
  float* out_ray_infos;        /* [R,2] */
  float* out_directions;       /* [R,3] */
  int64_t* out_img_idx;        /* [R] */
  float* out_c2w;              /* [R,3,4] */
  float* out_rgbs;             /* [R,3] */
  float* out_feats;            /* [R,feat_dim] or NULL */
  float* out_inv_depths;       /* [R] or NULL */
  int* status;                 /* optional */
} upnerf_ray_batch_args;
int upnerf_ray_batch_gather(const upnerf_ray_batch_args* a, void* stream);

/* ------------------------------------------------------------------------------------
 * TransientNet on the tensor cores.
 * Replaces TransientNet.forward (models/transient_net.py:27-38) and its autograd backward:
 *   h = feat_encoder(feats) (4 x Linear 256 + ReLU); t = ReLU(t_encoder([final_encoder(h) | embedding_t[img_idx]]));
 *   alpha = sigmoid(alpha_layer(h)); rgb = sigmoid(rgb_layer(t)); beta = softplus(beta_layer(t)) * alpha + beta_min.
 * bf16 operands on tcgen05 with fp32 accumulation for the six dense layers (upnerf_gemm_bf16 / the
 * weight-gradient kernel), fp32 row-dots for the N <= 3 heads.
 *   params[i]   fp32 tensors in state_dict order (:9-25): embedding_t.weight [n_images,128],
 *               feat_encoder.{0,2,4,6}.{weight,bias}, final_encoder.{weight,bias}, t_encoder.0.{weight,bias},
 *               alpha_layer.0.{weight,bias}, beta_layer.0.{weight,bias}, rgb_layer.0.{weight,bias}
 *   d_params[i] gradients, ACCUMULATED (fp32); a NULL entry skips that tensor
 *   alpha [R], beta [R], rgb [R,3] outputs; g_alpha / g_beta / g_rgb their upstream gradients (NULL = zero)
 *   workspace   >= upnerf_tnet_workspace_bytes(); upnerf_tnet_bwd reads what upnerf_tnet_fwd left in it */
#define UPNERF_TNET_PARAMS 19
typedef struct upnerf_tnet_args {
  int64_t n_rays;
  int n_images, feat_dim, hidden, transient_dim;
  float beta_min;
  const float* feats;          /* [R, feat_dim] */
  const int64_t* img_idx;      /* [R] */
  const float* params[UPNERF_TNET_PARAMS];
  float* d_params[UPNERF_TNET_PARAMS];
  float *alpha, *beta, *rgb;
  const float *g_alpha, *g_beta, *g_rgb;
  void* workspace;
  uint64_t workspace_bytes;
} upnerf_tnet_args;
uint64_t upnerf_tnet_workspace_bytes(const upnerf_tnet_args* a);
int upnerf_tnet_fwd(const upnerf_tnet_args* a, void* stream);
int upnerf_tnet_bwd(const upnerf_tnet_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UPNERF_B200_H_ */
