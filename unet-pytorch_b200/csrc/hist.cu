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

// ---------------------------------------------------------------------------------------------------------------
// Fast paths for uint8 masks (the get_miou case).  Shared-memory atomics top out near 0.5 pixel/clock/SM, ~5 % of what
// HBM delivers (11.5 pixel/clock/SM at 2 B/pixel), so the hot variants count without atomics:
//
//  * n*n <= 32 bins ("vote"): per 32 pixels one __ballot_sync per bin; popcounts accumulate in registers.
//  * otherwise ("private"): every lane owns a byte counter per bin, laid out [bin][lane] in shared memory, so an
//    increment is a plain LDS.U8 / STS.U8 on an address no other lane ever touches; byte counters are folded into
//    per-warp 32-bit sums before they can overflow (every 240 pixels per lane).
// Both finish with per-block 64-bit global atomics (integer adds: exact, order-free).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kVoteMaxBins = 32;      // (historical name) largest n*n handled by the register-counter kernel

// n*n <= 32 bins: every lane keeps all its counters in registers, packed as 8-bit fields of NW 64-bit words
// (bin k -> word k/8, byte k%8), so counting a pixel is one shift and one 64-bit add on a word selected by predicate.
// Fields are widened into 32-bit registers every 240 pixels (before a byte can wrap).  Slot nbins counts out-of-range
// bins; ignored pixels add zero.
template <int NW>      // 64-bit words: (n*n + 1 + 7) / 8 <= NW
__global__ void __launch_bounds__(256)
fast_hist_packed_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, long long nvec, int n,
                        unsigned long long* __restrict__ hist) {
  const int nbins = n * n;
  unsigned long long pk[NW];
  unsigned int wide[NW * 8];
#pragma unroll
  for (int i = 0; i < NW; ++i) pk[i] = 0ull;
#pragma unroll
  for (int i = 0; i < NW * 8; ++i) wide[i] = 0u;
  auto widen = [&]() {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) wide[i * 8 + k] += static_cast<unsigned int>(pk[i] >> (8 * k)) & 0xffu;
      pk[i] = 0ull;
    }
  };
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  int since = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 va = __ldg(a + i), vb = __ldg(b + i);
    const unsigned wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int wsel = 0; wsel < 4; ++wsel) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned av = (wa[wsel] >> (8 * k)) & 0xffu, bv = (wb[wsel] >> (8 * k)) & 0xffu;
        unsigned bin = n * av + bv;
        if (bin > static_cast<unsigned>(nbins)) bin = nbins;                   // overflow slot
        const unsigned long long inc = av < static_cast<unsigned>(n) ? 1ull << (8 * (bin & 7u)) : 0ull;
        const unsigned word = bin >> 3;
#pragma unroll
        for (int q = 0; q < NW; ++q) pk[q] += (word == static_cast<unsigned>(q)) ? inc : 0ull;
      }
    }
    since += 16;
    if (since >= 240) { widen(); since = 0; }
  }
  widen();
  // lanes -> warp -> block -> global
  __shared__ unsigned int sblk[NW * 8];
  if (threadIdx.x < NW * 8) sblk[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int v = 0; v < NW * 8; ++v) {
    const unsigned t = __reduce_add_sync(0xffffffffu, wide[v]);
    if ((threadIdx.x & 31) == (v & 31) && t) atomicAdd(&sblk[v], t);
  }
  __syncthreads();
  if (threadIdx.x <= nbins && sblk[threadIdx.x]) atomicAdd(&hist[threadIdx.x], static_cast<unsigned long long>(sblk[threadIdx.x]));
}

__global__ void __launch_bounds__(256)
fast_hist_private_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, long long nvec, int n,
                         unsigned long long* __restrict__ hist, int warps) {
  extern __shared__ __align__(16) unsigned char smem_h[];
  const int nbins = n * n;
  const int slots = nbins + 2;                                  // + overflow slot + trash slot (ignored pixels)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // layout: [warps][slots][32] byte counters, then [warps][slots] u32 sums
  unsigned char* mine = smem_h + static_cast<size_t>(warp) * slots * 32;
  unsigned int* sums = reinterpret_cast<unsigned int*>(smem_h + static_cast<size_t>(warps) * slots * 32) + warp * slots;
  for (int i = lane; i < slots * 8; i += 32) reinterpret_cast<unsigned int*>(mine)[i] = 0;
  for (int i = lane; i < slots; i += 32) sums[i] = 0;
  __syncwarp();

  auto fold = [&]() {       // byte counters -> 32-bit sums; lane l owns slots l, l+32, ...
    __syncwarp();
    for (int s = lane; s < slots; s += 32) {
      uint4* row = reinterpret_cast<uint4*>(mine + s * 32);
      const uint4 r0 = row[0], r1 = row[1];
      unsigned t = 0;
      t = __dp4a(r0.x, 0x01010101u, t); t = __dp4a(r0.y, 0x01010101u, t); t = __dp4a(r0.z, 0x01010101u, t); t = __dp4a(r0.w, 0x01010101u, t);
      t = __dp4a(r1.x, 0x01010101u, t); t = __dp4a(r1.y, 0x01010101u, t); t = __dp4a(r1.z, 0x01010101u, t); t = __dp4a(r1.w, 0x01010101u, t);
      sums[s] += t;
      row[0] = make_uint4(0, 0, 0, 0); row[1] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
  };

  unsigned char* col = mine + lane;                              // this lane's counter column
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  int since_fold = 0;
  // i0 is the warp's first index: every lane of a warp runs the same number of iterations (fold() is warp-collective)
  for (long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + warp * 32; i0 < nvec; i0 += stride) {
    const long long i = i0 + lane;
    uint4 va = make_uint4(~0u, ~0u, ~0u, ~0u), vb = make_uint4(0, 0, 0, 0);      // a = 255: ignored
    if (i < nvec) { va = __ldg(a + i); vb = __ldg(b + i); }
    const unsigned wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int wsel = 0; wsel < 4; ++wsel) {
      // four pixels at a time: the four counter loads are independent (latency overlaps); pixels that share a bin
      // all store old + multiplicity, so aliasing stores agree.  Ignored pixels go to a trash slot: no branches.
      unsigned char* pc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int av = (wa[wsel] >> (8 * k)) & 0xff, bv = (wb[wsel] >> (8 * k)) & 0xff;
        int bin = n * av + bv;
        if (bin > nbins) bin = nbins;                 // out-of-range prediction -> overflow slot
        if (av >= n) bin = nbins + 1;                 // ignored ground truth -> trash slot
        pc[k] = col + bin * 32;
      }
      unsigned c[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) c[k] = *pc[k];
      const unsigned e01 = pc[0] == pc[1], e02 = pc[0] == pc[2], e03 = pc[0] == pc[3];
      const unsigned e12 = pc[1] == pc[2], e13 = pc[1] == pc[3], e23 = pc[2] == pc[3];
      *pc[0] = static_cast<unsigned char>(c[0] + 1 + e01 + e02 + e03);
      *pc[1] = static_cast<unsigned char>(c[1] + 1 + e01 + e12 + e13);
      *pc[2] = static_cast<unsigned char>(c[2] + 1 + e02 + e12 + e23);
      *pc[3] = static_cast<unsigned char>(c[3] + 1 + e03 + e13 + e23);
    }
    since_fold += 16;
    if (since_fold >= 240) { fold(); since_fold = 0; }       // byte counters hold at most 240 increments
  }
  fold();
  __syncthreads();
  // block total -> global
  for (int s = threadIdx.x; s <= nbins; s += blockDim.x) {      // the trash slot is not published
    unsigned long long t = 0;
    for (int w2 = 0; w2 < warps; ++w2)
      t += reinterpret_cast<unsigned int*>(smem_h + static_cast<size_t>(warps) * slots * 32)[w2 * slots + s];
    if (t) atomicAdd(&hist[s], t);
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
  if (sm > 48 * 1024) {
    static bool big_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !big_done[dev]) {
      cudaFuncSetAttribute(fast_hist_kernel<unsigned char, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      cudaFuncSetAttribute(fast_hist_kernel<unsigned char, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      cudaFuncSetAttribute(fast_hist_kernel<int, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      cudaFuncSetAttribute(fast_hist_kernel<long long, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      big_done[dev] = true;
    }
  }
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
      const long long nvec = l / 16;
      const int slots = n * n + 2;
      auto tail = [&]() {
        if (l % 16) fast_hist_kernel<unsigned char, 1><<<1, kHistWarps * 32, sm, st>>>(pa + nvec * 16, pb + nvec * 16, l % 16, n, hist);
      };
      if (aligned && nvec > 0 && n * n + 1 <= kVoteMaxBins) {
        int g2 = 8 * num_sms();
        const long long w2 = (nvec + 255) / 256;
        if (g2 > w2) g2 = static_cast<int>(w2);
        const uint4 *qa = reinterpret_cast<const uint4*>(pa), *qb = reinterpret_cast<const uint4*>(pb);
        if (n * n + 1 <= 8)       fast_hist_packed_kernel<1><<<g2, 256, 0, st>>>(qa, qb, nvec, n, hist);
        else if (n * n + 1 <= 24) fast_hist_packed_kernel<3><<<g2, 256, 0, st>>>(qa, qb, nvec, n, hist);
        else                      fast_hist_packed_kernel<4><<<g2, 256, 0, st>>>(qa, qb, nvec, n, hist);
        note_launch();
        tail();
      } else if (aligned && nvec > 0 && static_cast<size_t>(slots) * 36 * 2 <= 200 * 1024) {
        int warps = static_cast<int>((200 * 1024) / (static_cast<size_t>(slots) * 36));
        if (warps > 8) warps = 8;
        const size_t sm2 = static_cast<size_t>(warps) * slots * 36;
        static bool attr_done[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_done[dev]) {
          cudaFuncSetAttribute(fast_hist_private_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          attr_done[dev] = true;
        }
        int g2 = num_sms();
        const long long w2 = (nvec + warps * 32 - 1) / (warps * 32);
        if (g2 > w2) g2 = static_cast<int>(w2);
        fast_hist_private_kernel<<<g2, warps * 32, sm2, st>>>(reinterpret_cast<const uint4*>(pa), reinterpret_cast<const uint4*>(pb), nvec, n, hist, warps);
        note_launch();
        tail();
      } else if (aligned) {
        fast_hist_kernel<unsigned char, 16><<<grid, kHistWarps * 32, sm, st>>>(pa, pb, l, n, hist);
      } else {
        fast_hist_kernel<unsigned char, 1><<<grid, kHistWarps * 32, sm, st>>>(pa, pb, l, n, hist);
      }
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
