#include "common.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

namespace upnerf {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// cuTensorMapEncodeTiled is a driver entry point; resolve it through the runtime so the
// library has no link-time dependency on libcuda (this also builds on GPU-less hosts).
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  auto fn = encode_fn();
  UPNERF_REQUIRE(fn != nullptr, UPNERF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  UPNERF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, UPNERF_ERR_BAD_SHAPE,
                 "tensor base %p is not 16-byte aligned", base);
  UPNERF_REQUIRE((ld * 2) % 16 == 0, UPNERF_ERR_BAD_SHAPE,
                 "leading dimension %llu is not a multiple of 8 elements",
                 (unsigned long long)ld);
  UPNERF_REQUIRE(box_cols * 2 == 128 && box_rows <= 256, UPNERF_ERR_BAD_SHAPE,
                 "unsupported TMA box %u x %u", box_rows, box_cols);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UPNERF_REQUIRE(r == CUDA_SUCCESS, UPNERF_ERR_CUDA,
                 "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu)",
                 (int)r, (unsigned long long)rows, (unsigned long long)cols,
                 (unsigned long long)ld);
  return UPNERF_OK;
}

int sm_count() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  return n;
}

// ------------------------------------------------------------------ launch accounting
namespace {
constexpr int kMaxProf = 16384;
struct Prof {
  bool enabled = false;
  int n = 0;
  cudaEvent_t beg[kMaxProf], end[kMaxProf];
  bool made[kMaxProf] = {};
  int cat[kMaxProf];
  double work[kMaxProf];
  double bytes[kMaxProf];
};
Prof g_prof;
long long g_launches = 0;
}  // namespace

LaunchScope::LaunchScope(int cat, cudaStream_t stream, double work, double bytes) : slot(-1), st(stream) {
  ++g_launches;
  if (!g_prof.enabled || g_prof.n >= kMaxProf) return;
  slot = g_prof.n++;
  if (!g_prof.made[slot]) {
    cudaEventCreate(&g_prof.beg[slot]);
    cudaEventCreate(&g_prof.end[slot]);
    g_prof.made[slot] = true;
  }
  g_prof.cat[slot] = cat;
  g_prof.work[slot] = work;
  g_prof.bytes[slot] = bytes;
  cudaEventRecord(g_prof.beg[slot], st);
}
LaunchScope::~LaunchScope() {
  if (slot >= 0) cudaEventRecord(g_prof.end[slot], st);
}

}  // namespace upnerf

extern "C" {

long long upnerf_launch_count(void) { return upnerf::g_launches; }
void upnerf_launch_count_add(long long n) { upnerf::g_launches += n; }

void upnerf_profile_enable(int on) {
  upnerf::g_prof.enabled = on != 0;
  upnerf::g_prof.n = 0;
}

// Sums event-measured device time (ms), launch counts, declared flop and algorithmic bytes per
// kernel family since upnerf_profile_enable(1); synchronises on the recorded events.
int upnerf_profile_collect(double* ms, long long* launches, double* work, double* bytes, int ncat) {
  using namespace upnerf;
  for (int i = 0; i < ncat; ++i) {
    ms[i] = 0;
    launches[i] = 0;
    work[i] = 0;
    bytes[i] = 0;
  }
  for (int i = 0; i < g_prof.n; ++i) {
    if (cudaEventSynchronize(g_prof.end[i]) != cudaSuccess) return UPNERF_ERR_CUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.beg[i], g_prof.end[i]) != cudaSuccess) return UPNERF_ERR_CUDA;
    const int c = g_prof.cat[i];
    if (c < ncat) {
      ms[c] += t;
      launches[c] += 1;
      work[c] += g_prof.work[i];
      bytes[c] += g_prof.bytes[i];
    }
  }
  g_prof.n = 0;
  return UPNERF_OK;
}

const char* upnerf_last_error(void) { return upnerf::g_err; }

int upnerf_version(void) { return 100; }

int upnerf_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return 0;
  return major == 10 ? 1 : 0;
}

}  // extern "C"

// ------------------------------------------------------------------ measurement helper
namespace upnerf {
namespace {
__global__ void fill_pattern_kernel(uint4* dst, int64_t n16, uint32_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
    uint32_t h = static_cast<uint32_t>(i) * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    uint4 v = make_uint4(h, h * 3266489917u, h ^ 0x9E3779B9u, h + static_cast<uint32_t>(i >> 7));
    __stcs(dst + i, v);
  }
}
}  // namespace
}  // namespace upnerf

extern "C" int upnerf_fill_pattern(void* dst, int64_t bytes, uint32_t seed, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(dst && bytes > 0 && bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                 UPNERF_ERR_BAD_SHAPE, "fill_pattern: 16-byte aligned buffer and size required");
  fill_pattern_kernel<<<sm_count() * 16, 256, 0, as_stream(stream)>>>(static_cast<uint4*>(dst), bytes / 16, seed);
  UPNERF_CHECK_LAUNCH("fill_pattern_kernel");
  return UPNERF_OK;
}

// TMA-store ceiling probe (tools/write_bw.py): every CTA streams 128-row x 64-column bf16 boxes (16 KB,
// the activation box of the fused trunk kernels) from shared memory to a [rows, ld] bf16 matrix with
// up to four bulk stores in flight and no other work.  ld = 256 reproduces the trunk's pattern (128
// rows of 128 B at a 512 B stride); ld = 64 makes every box one contiguous 16 KB run.
#include "ptx_sm100.cuh"
namespace upnerf {
namespace {
__global__ void __launch_bounds__(128, 1)
tma_store_probe_kernel(const __grid_constant__ CUtensorMap map, int boxes_per_row, int64_t n_boxes, int depth) {
  extern __shared__ uint8_t probe_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(probe_raw) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < 4 * 16384 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u + blockIdx.x;
  ptx::fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = 0;
    const int reps = depth >> 8;     // high bits: passes over the footprint (L2-resident runs)
    depth &= 255;
    for (int rep = 0; rep <= reps; ++rep)
    for (int64_t b = blockIdx.x; b < n_boxes; b += gridDim.x, ++k) {
      const int col = static_cast<int>(b % boxes_per_row) * 64;
      const int row = static_cast<int>(b / boxes_per_row) * 128;
      ptx::tma_store_2d(&map, smem + (k & 3) * 16384, col, row);
      ptx::tma_store_commit();
      // depth = bulk stores allowed in flight (1: each store's shared-memory read is awaited alone)
      if (depth >= 4) ptx::tma_store_wait_read<3>();
      else if (depth == 3) ptx::tma_store_wait_read<2>();
      else if (depth == 2) ptx::tma_store_wait_read<1>();
      else ptx::tma_store_wait_read<0>();
    }
    ptx::tma_store_wait_all<0>();
  }
}
}  // namespace
}  // namespace upnerf

extern "C" int upnerf_tma_store_probe(void* dst, int64_t rows, int64_t ld, int depth, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(dst && rows > 0 && rows % 128 == 0 && ld >= 64 && ld % 64 == 0, UPNERF_ERR_BAD_SHAPE,
                 "tma_store_probe: rows %% 128 == 0 and ld %% 64 == 0 required");
  CUtensorMap map;
  UPNERF_TRY(make_tmap_bf16_2d(&map, dst, rows, ld, ld, 128, 64));
  static bool attr_set = false;
  if (!attr_set) {
    UPNERF_CHECK_CUDA(cudaFuncSetAttribute(tma_store_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           4 * 16384 + 1024));
    attr_set = true;
  }
  const int bpr = static_cast<int>(ld / 64);
  tma_store_probe_kernel<<<sm_count(), 128, 4 * 16384 + 1024, as_stream(stream)>>>(map, bpr, (rows / 128) * bpr, depth);
  UPNERF_CHECK_LAUNCH("tma_store_probe_kernel");
  return UPNERF_OK;
}
