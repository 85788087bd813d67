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
} upnerf_epilogue;

/* bf16 in / bf16 out on tcgen05.  Requires K % 64 == 0, N % 64 == 0, N <= 256,
 * 16-byte aligned pointers and leading dimensions that are multiples of 8. */
int upnerf_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                     int64_t M, int N, int K, const upnerf_epilogue* ep, void* stream);

/* Weight gradient on tcgen05:  dW[n, colmap(k)] += sum_m dY[m,n] * X[m,k]  (fp32 atomics),
 * db[n] += sum_m dY[m,n].  dY:[M,N] bf16, X:[M,K] bf16.  N % 128 == 0 (N <= 256),
 * K % 64 == 0 (K <= 320).  Column segments map packed K columns to parameter columns:
 * packed columns [seg_src[i], seg_src[i]+seg_len[i]) go to dW columns starting at
 * seg_dst[i]; columns not covered by a segment are dropped (padding). db may be NULL. */
int upnerf_wgrad_bf16(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW,
                      int64_t lddw, float* db, int64_t M, int N, int K, int n_seg,
                      const int* seg_src_host, const int* seg_len_host, const int* seg_dst_host,
                      void* stream);

/* fp32 SIMT GEMM with arbitrary element strides (validation mode and the small per-ray
 * products):  C[m*scm + n*scn] = epi( sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk] ).
 * split_k > 1 splits the k range over grid.z and accumulates with atomicAdd into C
 * (epilogue ignored, C must be pre-initialised); accumulate != 0 adds to C instead of
 * overwriting.  aux uses the strides of C. Heads are not supported here. */
int upnerf_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn,
                    int64_t sbk, float* C, int64_t scm, int64_t scn, int64_t M, int64_t N,
                    int64_t K, const upnerf_epilogue* ep, int accumulate, int split_k,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UPNERF_B200_H_ */
