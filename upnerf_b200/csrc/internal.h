// Internal (non-ABI) declarations shared between the translation units of the library.
#pragma once
#include "common.h"

namespace upnerf {

// One strided copy fp32 -> T with optional transpose: dst[r, c] (or dst[c, r]) = src[r, c].
struct PackOp {
  const float* src;
  int64_t ld_src;
  void* dst;
  int64_t ld_dst;
  int rows, cols, transpose;
};
constexpr int kMaxPackOps = 40;
struct PackList {
  PackOp ops[kMaxPackOps];
  int n;
};

int run_pack(const PackList& list, int dtype, cudaStream_t st);

// Deterministic split reduction of the tcgen05 weight gradients (wgrad_tc.cu): the wgrad launches
// of one network pass park their per-split partial tiles in a pool and register a slot each;
// wgrad_reduce() then sums every slot into the parameter gradients with one launch.
constexpr int kMaxWgradSlots = 16;
struct WgradReduceSlot {
  const float* partial;      // [splits][chunks][K + 1][128]
  float* dW;
  int64_t lddw;
  float* dW_hi;              // destination of output rows 128..255 (stacked layers) or nullptr
  int64_t lddw_hi;
  float* db;
  int splits, chunks, K, n_seg;
  int seg_src[4], seg_len[4], seg_dst[4];
  int block_begin;           // first block of the reduce grid that works on this slot
};
struct WgradReduceList {
  WgradReduceSlot s[kMaxWgradSlots];
  int n;
};
// One weight gradient dW[n, map(k)] += dY^T X, db += colsum(dY) as the launcher takes it.
struct WgradProblem {
  const void *dY, *X;
  int64_t lddy, ldx;
  float *dW, *dW_hi, *db;
  int64_t lddw, lddw_hi;
  int64_t M;
  int N, K, n_seg;
  int seg_src[4], seg_len[4], seg_dst[4];
};
struct WgradBatch {
  WgradReduceList list;
  float* pool;
  uint64_t pool_floats, used;
  int blocks;
  // defer != 0: wgrad_launch() only queues the problem; wgrad_reduce() (or a full queue) launches all queued
  // problems as ONE grouped kernel whose grid is shared out in proportion to the bytes each one streams
  int defer;
  int n_pending;
  WgradProblem pending[kMaxWgradSlots];
};
// Upper bound of the pool one network pass of the D=8, W=256 architecture needs (floats).
uint64_t wgrad_pool_floats();
int wgrad_launch(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw,
                 float* dW_hi, int64_t lddw_hi, float* db, int64_t M, int N, int K, int n_seg,
                 const int* seg_src_host, const int* seg_len_host, const int* seg_dst_host,
                 WgradBatch* batch, void* stream);
int wgrad_flush(WgradBatch* batch, cudaStream_t st);
int wgrad_reduce(WgradBatch* batch, cudaStream_t st);
int gather_rows(const float* table, const int64_t* idx, int64_t R, int dim, float* out, int64_t ld_out,
                cudaStream_t st);
int scatter_add_rows(const float* src, int64_t ld_src, const int64_t* idx, int64_t R, int dim,
                     float* table, cudaStream_t st);
int ray_sum128(const void* X, int64_t ld, int64_t R, int S, float* out, int dtype, cudaStream_t st);
int rgb_head_bwd(const void* Q, int64_t ldq, const float* rgb, const float* d_rgb, const float* W2,
                 int64_t R, int S, void* dQ, int64_t lddq, float* d_raybias, float* dW2, float* db2,
                 int dtype, cudaStream_t st);
int rowscale_colsum(const void* X, int64_t ld, const float* s, int64_t M, int N, float* out,
                    float* out_s, int dtype, cudaStream_t st);
int rowdot_head(const float* X, int64_t ld, int64_t M, int N, int nh, const float* w, const float* b,
                int act, float* out, cudaStream_t st);

}  // namespace upnerf
