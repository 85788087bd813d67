// Host-side helpers shared by every translation unit of libupnerf_b200.so:
// error reporting behind the C ABI and TMA tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/upnerf_b200.h"

namespace upnerf {

// Records the message returned by upnerf_last_error() and returns `code`.
int set_error(int code, const char* fmt, ...);

#define UPNERF_CHECK_CUDA(expr)                                                              \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::upnerf::set_error(UPNERF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,            \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                \
  } while (0)

#define UPNERF_CHECK_LAUNCH(name)                                                            \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return ::upnerf::set_error(UPNERF_ERR_CUDA, "launch of %s failed: %s (%s:%d)", name,   \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                \
  } while (0)

#define UPNERF_REQUIRE(cond, code, ...)                                                      \
  do {                                                                                       \
    if (!(cond)) return ::upnerf::set_error(code, __VA_ARGS__);                              \
  } while (0)

#define UPNERF_TRY(expr)                                                                     \
  do {                                                                                       \
    int _s = (expr);                                                                         \
    if (_s != 0) return _s;                                                                  \
  } while (0)

// 2-D bf16 tensor map: `rows` x `cols` elements, row stride `ld` elements, box of
// box_rows x box_cols elements, 128-byte swizzle (box_cols * 2 bytes must be 128).
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld, uint32_t box_rows, uint32_t box_cols);

int sm_count();

// Launch accounting.  Every kernel launcher opens a LaunchScope: it bumps the library-wide
// launch counter and, while profiling is enabled, brackets the launch with CUDA events on
// the launching stream so that bench.py can attribute device time to kernel families.
enum LaunchCat {
  kCatGemmTc = 0, kCatWgradTc, kCatGemmSimt, kCatComposite, kCatPosenc, kCatSampling,
  kCatPoseRays, kCatHeads, kCatPack, kCatTrunkFwd, kCatTrunkBwd, kCatWgradReduce, kCatTnet, kCatGemmTf32, kNumCats
};
struct LaunchScope {
  int slot;
  cudaStream_t st;
  // work = flop of the launch, bytes = its ALGORITHMIC memory traffic (each operand touched once)
  LaunchScope(int cat, cudaStream_t stream, double work = 0.0, double bytes = 0.0);
  ~LaunchScope();
};

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace upnerf
