// hist.cu -- fast_hist: the n x n confusion matrix of get_miou / compute_mIoU (utils/utils_metrics.py:34-43).
//
//   k = (a >= 0) & (a < n);  hist = bincount(n * a[k] + b[k], minlength = n*n).reshape(n, n)
//
// Integer work, bit-exact.  HBM-bound: 2 bytes per pixel (uint8 ground truth + uint8 prediction), read as
// 16-byte vectors.  Each warp owns a private shared-memory histogram (no inter-warp contention); lanes that hit
// the same bin in one step are combined with __match_any_sync so the shared atomic count stays low for the
// few-class case; per-block histograms are flushed with 64-bit global atomics (integer adds: order-free, exact).
// Bins >= n*n (possible when b >= n, where numpy's reshape would raise) are counted in hist[n*n] so the host
// wrapper can raise the same error.
#include "b2u_internal.h"
#include "b2u_ptx.cuh"

namespace b2u {

constexpr int kHistWarps = 8;

template <typename T, int VEC>
__global__ void __launch_bounds__(kHistWarps * 32)
fast_hist_kernel(const T* __restrict__ a, const T* __restrict__ b, long long len, int n, unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int sh[];   // [warps][nbins + 1]
  const int nbins = n * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned int* mine = sh + warp * (nbins + 1);
  for (int i = threadIdx.x; i < kHistWarps * (nbins + 1); i += blockDim.x) sh[i] = 0;
  __syncthreads();

  auto count = [&](long long av, long long bv, bool live) {
    int bin = -1;
    if (live && av >= 0 && av < n) {
      const long long bb = static_cast<long long>(n) * av + bv;
      bin = (bb >= 0 && bb < nbins) ? static_cast<int>(bb) : nbins;   // nbins = overflow slot
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&mine[bin], __popc(peers));
  };

  const long long nvec = len / VEC;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < ((nvec + 31) / 32) * 32; i += stride) {
    const bool live = i < nvec;
    T av[VEC], bv[VEC];
    if (live) {
      if (sizeof(T) * VEC == 16) {
        *reinterpret_cast<uint4*>(av) = __ldg(reinterpret_cast<const uint4*>(a) + i);
        *reinterpret_cast<uint4*>(bv) = __ldg(reinterpret_cast<const uint4*>(b) + i);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) { av[e] = a[i * VEC + e]; bv[e] = b[i * VEC + e]; }
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) count(live ? static_cast<long long>(av[e]) : 0, live ? static_cast<long long>(bv[e]) : 0, live);
  }
  // tail (len % VEC elements), handled by block 0 warp 0
  if (blockIdx.x == 0 && warp == 0) {
    const long long t0 = nvec * VEC;
    for (long long base = t0; base < len; base += 32) {
      const long long i = base + lane;
      const bool live = i < len;
      count(live ? static_cast<long long>(a[i]) : 0, live ? static_cast<long long>(b[i]) : 0, live);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= nbins; i += blockDim.x) {
    unsigned long long s = 0;
    for (int w = 0; w < kHistWarps; ++w) s += sh[w * (nbins + 1) + i];
    if (s) atomicAdd(&hist[i], s);
  }
}

}  // namespace b2u

extern "C" {
using namespace b2u;

// hist: n*n + 1 uint64 on device, ACCUMULATED into (caller zeroes it); hist[n*n] counts out-of-range bins.
// dtype: 0 = uint8, 1 = int32, 2 = int64.  Blocks of 2^31 elements at most per call keep uint32 counters exact.
int b2u_fast_hist(const void* a, const void* b, long long len, int n, int dtype, unsigned long long* hist, void* stream) {
  if (n <= 0 || n > 64) return set_error(B2U_ERR_SHAPE, "fast_hist: 1 <= n <= 64 (got %d)", n);
  if (len < 0) return set_error(B2U_ERR_SHAPE, "fast_hist: negative length");
  if (len == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t sm = static_cast<size_t>(kHistWarps) * (n * n + 1) * sizeof(unsigned int);
  const long long chunk = 1ll << 31;
  for (long long off = 0; off < len; off += chunk) {
    const long long l = len - off < chunk ? len - off : chunk;
    int grid = 8 * num_sms();
    if (dtype == 0) {
      const bool aligned = ((reinterpret_cast<uintptr_t>(a) + off) % 16 == 0) && ((reinterpret_cast<uintptr_t>(b) + off) % 16 == 0);
      const unsigned char* pa = static_cast<const unsigned char*>(a) + off;
      const unsigned char* pb = static_cast<const unsigned char*>(b) + off;
      long long want = (l / 16 + kHistWarps * 32 - 1) / (kHistWarps * 32);
      if (want < 1) want = 1;
      if (grid > want) grid = static_cast<int>(want);
      if (aligned) fast_hist_kernel<unsigned char, 16><<<grid, kHistWarps * 32, sm, st>>>(pa, pb, l, n, hist);
      else         fast_hist_kernel<unsigned char, 1><<<grid, kHistWarps * 32, sm, st>>>(pa, pb, l, n, hist);
    } else if (dtype == 1) {
      fast_hist_kernel<int, 1><<<grid, kHistWarps * 32, sm, st>>>(static_cast<const int*>(a) + off, static_cast<const int*>(b) + off, l, n, hist);
    } else if (dtype == 2) {
      fast_hist_kernel<long long, 1><<<grid, kHistWarps * 32, sm, st>>>(static_cast<const long long*>(a) + off,
                                                                        static_cast<const long long*>(b) + off, l, n, hist);
    } else {
      return set_error(B2U_ERR_ARG, "fast_hist: dtype must be 0 (u8), 1 (i32) or 2 (i64)");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(B2U_ERR_CUDA, "fast_hist launch: %s", cudaGetErrorString(e));
    note_launch();
  }
  return 0;
}

}  // extern "C"
