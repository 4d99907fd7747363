// runtime.cu -- error reporting, device queries and TMA tensor-map encoding for libb200unet.so
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <mutex>

#include "b2u_internal.h"

namespace b2u {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
void reset_launch_count() { g_launches.store(0); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_nhwc(CUtensorMap* out, const void* base, int N, int H, int W, int C, int boxC, int boxW, int boxH,
                   CUtensorMapSwizzle swz, int cpitch) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(B2U_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if (cpitch <= 0) cpitch = C;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (cpitch * 2) % 16 != 0)
    return set_error(B2U_ERR_ARG, "tensor map: base %p / channel pitch %d not 16-byte aligned", base, cpitch);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cpitch * 2, (cuuint64_t)W * cpitch * 2, (cuuint64_t)H * W * cpitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B2U_ERR_DRIVER, "cuTensorMapEncodeTiled(nhwc N=%d H=%d W=%d C=%d pitch=%d box=%d,%d,%d) -> %d", N,
                     H, W, C, cpitch, boxC, boxW, boxH, (int)r);
  return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, int boxCols, int boxRows,
                 CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(B2U_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (cols * 2) % 16 != 0)
    return set_error(B2U_ERR_ARG, "tensor map 2d: base %p / row bytes %llu not 16-byte aligned", base,
                     (unsigned long long)(cols * 2));
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxCols, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B2U_ERR_DRIVER, "cuTensorMapEncodeTiled(2d cols=%llu rows=%llu box=%d,%d) -> %d",
                     (unsigned long long)cols, (unsigned long long)rows, boxCols, boxRows, (int)r);
  return 0;
}

}  // namespace b2u

namespace b2u { const char* last_error(); long long launch_count(); void reset_launch_count(); }
extern "C" {
const char* b2u_last_error(void) { return b2u::last_error(); }
int b2u_version(void) { return 100; }
int b2u_num_sms(void) { return b2u::num_sms(); }
long long b2u_launch_count(void) { return b2u::launch_count(); }
void b2u_reset_launch_count(void) { b2u::reset_launch_count(); }
}
