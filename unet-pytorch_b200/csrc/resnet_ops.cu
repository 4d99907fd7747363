// resnet_ops.cu -- the extra HBM-bound kernels the ResNet50 encoder needs (nets/resnet.py:100-176 of the reference).
//
//   im2col_stem      : 7x7 stride-2 pad-3 conv (nets/resnet.py:109) staged as im2col rows [N,H/2,W/2,192] bf16
//                      (147 real columns k = (r*7+s)*Cin + c) so the stem runs as a 1x1 tensor-core GEMM with K = 192
//   pack_weights_im2col : OIHW fp32 -> [Cout][Kpad] bf16 with the same column order
//   subsample2 / zero_insert2 : stride-2 convolutions are run as (stride-1 conv -> keep even pixels) for 3x3 and
//                      (keep even pixels -> 1x1 conv) for the downsample branch (nets/resnet.py:140); zero_insert2 is
//                      the adjoint used in their backward
//   maxpool3x3s2_ceil: nn.MaxPool2d(3, 2, padding=0, ceil_mode=True) (nets/resnet.py:113) fwd and bwd (gather form:
//                      every input pixel checks the <= 4 windows that contain it; first max in scan order wins)
//   add_bf16         : out = a + b (gradient accumulation where a tensor has two consumers)
#include "b2u_internal.h"
#include "b2u_ptx.cuh"

namespace b2u {

#define B2U_CHECK_LAUNCH(name)                                                                           \
  do {                                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                                \
    if (e__ != cudaSuccess) return b2u::set_error(B2U_ERR_CUDA, name " launch: %s", cudaGetErrorString(e__)); \
    b2u::note_launch();                                                                                  \
  } while (0)

__device__ __forceinline__ void r_unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 r_pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// ---------------------------------------------------------------------------------------------- stem im2col
constexpr int kStemPix = 32;     // output pixels per block
constexpr int kStemK = 192;      // padded K
__global__ void __launch_bounds__(256)
im2col_stem_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int N, int Cin, int H, int W, int Ho, int Wo) {
  __shared__ __align__(16) __nv_bfloat16 tile[kStemPix][kStemK + 8];
  const int segs = (Wo + kStemPix - 1) / kStemPix;
  const int seg = blockIdx.x % segs;
  const int row = blockIdx.x / segs;          // n * Ho + ho
  const int n = row / Ho, ho = row - n * Ho;
  const int wo0 = seg * kStemPix;
  const int K = 49 * Cin;
  for (int i = threadIdx.x; i < kStemK * kStemPix; i += blockDim.x) {
    const int k = i / kStemPix, pw = i - k * kStemPix;
    float v = 0.f;
    if (k < K) {
      const int tap = k / Cin, c = k - tap * Cin;
      const int hh = 2 * ho + tap / 7 - 3, ww = 2 * (wo0 + pw) + tap % 7 - 3;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(x + ((static_cast<size_t>(n) * Cin + c) * H + hh) * W + ww);
    }
    tile[pw][k] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(col) + (static_cast<size_t>(row) * Wo + wo0) * (kStemK / 8);
  for (int i = threadIdx.x; i < kStemPix * (kStemK / 8); i += blockDim.x) {
    const int pw = i / (kStemK / 8), q = i - pw * (kStemK / 8);
    if (wo0 + pw < Wo) out[i] = *reinterpret_cast<const uint4*>(&tile[pw][q * 8]);
  }
}

__global__ void pack_weights_im2col_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, int Cout, int Cin,
                                           int taps, int Kpad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * Kpad) return;
  const int co = idx / Kpad, k = idx - co * Kpad;
  float v = 0.f;
  if (k < taps * Cin) {
    const int tap = k / Cin, c = k - tap * Cin;
    v = w[(static_cast<size_t>(co) * Cin + c) * taps + tap];
  }
  wf[idx] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------- stride-2 helpers
__global__ void subsample2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int C8) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const unsigned t = blockIdx.y * blockDim.x + threadIdx.x;
  if (t >= static_cast<unsigned>(Wo) * C8) return;
  const int wo = t / C8, c = t - wo * C8;
  const int n = blockIdx.x / Ho, ho = blockIdx.x - n * Ho;
  y[static_cast<size_t>(blockIdx.x) * Wo * C8 + t] = __ldg(x + ((static_cast<size_t>(n) * H + 2 * ho) * W + 2 * wo) * C8 + c);
}
__global__ void zero_insert2_kernel(const uint4* __restrict__ y, uint4* __restrict__ x, int N, int H, int W, int C8) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const unsigned t = blockIdx.y * blockDim.x + threadIdx.x;
  if (t >= static_cast<unsigned>(W) * C8) return;
  const int w = t / C8, c = t - w * C8;
  const int n = blockIdx.x / H, h = blockIdx.x - n * H;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (!(h & 1) && !(w & 1)) v = __ldg(y + ((static_cast<size_t>(n) * Ho + h / 2) * Wo + w / 2) * C8 + c);
  x[static_cast<size_t>(blockIdx.x) * W * C8 + t] = v;
}

// ---------------------------------------------------------------------------------------------- 3x3 s2 ceil max-pool
__device__ __forceinline__ int pool_out(int in) { return (in - 3 + 1) / 2 + 1; }   // ceil((in-3)/2) + 1 for in >= 3
__global__ void maxpool3x3s2_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int Ho,
                                        int Wo, int C8) {
  const unsigned t = blockIdx.y * blockDim.x + threadIdx.x;
  if (t >= static_cast<unsigned>(Wo) * C8) return;
  const int wo = t / C8, c = t - wo * C8;
  const int n = blockIdx.x / Ho, ho = blockIdx.x - n * Ho;
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
  for (int r = 0; r < 3; ++r) {
    const int h = 2 * ho + r;
    if (h >= H) break;
    for (int s = 0; s < 3; ++s) {
      const int w = 2 * wo + s;
      if (w >= W) break;
      float f[8];
      r_unpack8(__ldg(x + ((static_cast<size_t>(n) * H + h) * W + w) * C8 + c), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
    }
  }
  y[static_cast<size_t>(blockIdx.x) * Wo * C8 + t] = r_pack8(m);
}
// dx[h,w] = sum over windows (ho,wo) containing (h,w) whose FIRST maximum (row-major scan) is at (h,w) of dy[ho,wo]
__global__ void maxpool3x3s2_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, uint4* __restrict__ dx,
                                        int N, int H, int W, int Ho, int Wo, int C8) {
  const unsigned t = blockIdx.y * blockDim.x + threadIdx.x;
  if (t >= static_cast<unsigned>(W) * C8) return;
  const int w = t / C8, c = t - w * C8;
  const int n = blockIdx.x / H, h = blockIdx.x - n * H;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float xv[8];
  const uint4* ximg = x + static_cast<size_t>(n) * H * W * C8 + c;
  r_unpack8(__ldg(ximg + (static_cast<size_t>(h) * W + w) * C8), xv);
  // windows containing h: ho with 2*ho <= h <= 2*ho + 2
  const int ho_lo = h >= 2 ? (h - 2 + 1) / 2 : 0, ho_hi = min(h / 2, Ho - 1);
  const int wo_lo = w >= 2 ? (w - 2 + 1) / 2 : 0, wo_hi = min(w / 2, Wo - 1);
  for (int ho = ho_lo; ho <= ho_hi; ++ho) {
    for (int wo = wo_lo; wo <= wo_hi; ++wo) {
      // is (h,w) the first maximum of window (ho,wo)?
      bool first[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) first[k] = true;
      for (int r = 0; r < 3; ++r) {
        const int hh = 2 * ho + r;
        if (hh >= H) break;
        for (int s = 0; s < 3; ++s) {
          const int ww = 2 * wo + s;
          if (ww >= W) break;
          if (hh == h && ww == w) continue;
          float f[8];
          r_unpack8(__ldg(ximg + (static_cast<size_t>(hh) * W + ww) * C8), f);
          const bool before = hh < h || (hh == h && ww < w);
#pragma unroll
          for (int k = 0; k < 8; ++k) first[k] = first[k] && (before ? f[k] < xv[k] : f[k] <= xv[k]);
        }
      }
      float g[8];
      r_unpack8(__ldg(dy + ((static_cast<size_t>(n) * Ho + ho) * Wo + wo) * C8 + c), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += first[k] ? g[k] : 0.f;
    }
  }
  dx[static_cast<size_t>(blockIdx.x) * W * C8 + t] = r_pack8(acc);
}

__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float fa[8], fb[8];
  r_unpack8(__ldg(a + i), fa);
  r_unpack8(__ldg(b + i), fb);
#pragma unroll
  for (int k = 0; k < 8; ++k) fa[k] += fb[k];
  out[i] = r_pack8(fa);
}

// y = relu(a + b): the residual join `x += residual; x = relu(x)` (nets/LightWeightUnet.py:52-53)
__global__ void add_relu_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float fa[8], fb[8];
  r_unpack8(__ldg(a + i), fa);
  r_unpack8(__ldg(b + i), fb);
#pragma unroll
  for (int k = 0; k < 8; ++k) fa[k] = fmaxf(fa[k] + fb[k], 0.f);
  out[i] = r_pack8(fa);
}
// dx = dy where y > 0, else 0 (gradient of the join for both of its inputs)
__global__ void relu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, uint4* __restrict__ dx, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float g[8], v[8];
  r_unpack8(__ldg(dy + i), g);
  r_unpack8(__ldg(y + i), v);
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.f ? g[k] : 0.f;
  dx[i] = r_pack8(g);
}

static inline dim3 rgrid(long long rows, int row_items, int block) {
  return dim3(static_cast<unsigned>(rows), static_cast<unsigned>((row_items + block - 1) / block), 1);
}

}  // namespace b2u

extern "C" {
using namespace b2u;

// 7x7 stride-2 pad-3 stem: x NCHW fp32 [N,Cin,H,W] (Cin <= 3) -> col [N,Ho,Wo,192] bf16, Ho = (H+6-7)/2+1
int b2u_im2col_stem(const float* x, void* col, int N, int Cin, int H, int W, void* stream) {
  if (N <= 0 || Cin <= 0 || 49 * Cin > kStemK || H < 1 || W < 1) return set_error(B2U_ERR_SHAPE, "im2col_stem: bad shape (Cin=%d)", Cin);
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long blocks = static_cast<long long>(N) * Ho * ((Wo + kStemPix - 1) / kStemPix);
  im2col_stem_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(col), N, Cin, H, W, Ho, Wo);
  B2U_CHECK_LAUNCH("im2col_stem");
  return 0;
}

// OIHW fp32 [Cout][Cin][taps] -> [Cout][Kpad] bf16, column k = tap*Cin + c, zero padded
int b2u_pack_weights_im2col(const float* w, void* wf, int Cout, int Cin, int taps, int Kpad, void* stream) {
  if (Cout <= 0 || Cin <= 0 || taps <= 0 || taps * Cin > Kpad) return set_error(B2U_ERR_SHAPE, "pack_weights_im2col: bad shape");
  pack_weights_im2col_kernel<<<(Cout * Kpad + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(wf), Cout, Cin, taps, Kpad);
  B2U_CHECK_LAUNCH("pack_weights_im2col");
  return 0;
}

// y[n,ho,wo,:] = x[n,2ho,2wo,:]   (H, W: dims of x; y is [N,ceil(H/2),ceil(W/2),C])
int b2u_subsample2(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C % 8 != 0) return set_error(B2U_ERR_SHAPE, "subsample2: bad shape");
  subsample2_kernel<<<rgrid(static_cast<long long>(N) * ((H + 1) / 2), ((W + 1) / 2) * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), N, H, W, C / 8);
  B2U_CHECK_LAUNCH("subsample2");
  return 0;
}
// adjoint of subsample2: x (H x W) = y at even (h,w), zero elsewhere
int b2u_zero_insert2(const void* y, void* x, int N, int H, int W, int C, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C % 8 != 0) return set_error(B2U_ERR_SHAPE, "zero_insert2: bad shape");
  zero_insert2_kernel<<<rgrid(static_cast<long long>(N) * H, W * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(y), static_cast<uint4*>(x), N, H, W, C / 8);
  B2U_CHECK_LAUNCH("zero_insert2");
  return 0;
}

// nn.MaxPool2d(3, 2, 0, ceil_mode=True): H, W >= 3; output dims ceil((H-3)/2)+1
int b2u_maxpool3x3s2_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  if (N <= 0 || H < 3 || W < 3 || C % 8 != 0) return set_error(B2U_ERR_SHAPE, "maxpool3x3s2: needs H,W >= 3 and C %% 8 == 0");
  const int Ho = (H - 3 + 1) / 2 + 1, Wo = (W - 3 + 1) / 2 + 1;
  maxpool3x3s2_fwd_kernel<<<rgrid(static_cast<long long>(N) * Ho, Wo * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), N, H, W, Ho, Wo, C / 8);
  B2U_CHECK_LAUNCH("maxpool3x3s2_fwd");
  return 0;
}
int b2u_maxpool3x3s2_bwd(const void* dy, const void* x, void* dx, int N, int H, int W, int C, void* stream) {
  if (N <= 0 || H < 3 || W < 3 || C % 8 != 0) return set_error(B2U_ERR_SHAPE, "maxpool3x3s2_bwd: needs H,W >= 3 and C %% 8 == 0");
  const int Ho = (H - 3 + 1) / 2 + 1, Wo = (W - 3 + 1) / 2 + 1;
  maxpool3x3s2_bwd_kernel<<<rgrid(static_cast<long long>(N) * H, W * (C / 8), 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dy), static_cast<const uint4*>(x), static_cast<uint4*>(dx), N, H, W, Ho, Wo, C / 8);
  B2U_CHECK_LAUNCH("maxpool3x3s2_bwd");
  return 0;
}

// out = a + b, n elements (multiple of 8), bf16; out may alias a or b
int b2u_add_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
  if (n <= 0 || n % 8 != 0) return set_error(B2U_ERR_SHAPE, "add_bf16: n must be a positive multiple of 8");
  const long long n8 = n / 8;
  add_bf16_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), n8);
  B2U_CHECK_LAUNCH("add_bf16");
  return 0;
}

// out = relu(a + b); out may alias a or b
int b2u_add_relu_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
  if (n <= 0 || n % 8 != 0) return set_error(B2U_ERR_SHAPE, "add_relu_bf16: n must be a positive multiple of 8");
  const long long n8 = n / 8;
  add_relu_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), n8);
  B2U_CHECK_LAUNCH("add_relu_bf16");
  return 0;
}

// dx = dy * (y > 0); dx may alias dy
int b2u_relu_bwd_bf16(const void* dy, const void* y, void* dx, long long n, void* stream) {
  if (n <= 0 || n % 8 != 0) return set_error(B2U_ERR_SHAPE, "relu_bwd_bf16: n must be a positive multiple of 8");
  const long long n8 = n / 8;
  relu_bwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dy), static_cast<const uint4*>(y), static_cast<uint4*>(dx), n8);
  B2U_CHECK_LAUNCH("relu_bwd_bf16");
  return 0;
}

}  // extern "C"
