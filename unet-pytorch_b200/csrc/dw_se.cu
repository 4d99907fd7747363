// dw_se.cu -- the memory-bound blocks of the Lightweight / UltraLightweight UNets of the reference:
//
//   dwconv3x3      depthwise 3x3 conv, pad 1 (nn.Conv2d(C, C, 3, padding=1, groups=C), nets/UltraLightweightUnet_large.py:9-10)
//                  fwd (+bias), dgrad (same kernel, taps flipped), wgrad + bias grad (per-channel reductions)
//   spatial_mean / spatial_dot   per-(image, channel) reductions over H*W: SE squeeze (AdaptiveAvgPool2d(1), :39)
//                  and its backward (sum of dy * x)
//   scale_nc       y = x * s[n][c] + a[n][c]: SE excitation (:52), its backward, Dropout2d (:78,97)
//   se_fc_fwd/bwd  the two tiny Linear layers + ReLU + Sigmoid of the SE block (:41-46) and their gradients
//
// All NHWC bf16 activations with 16-byte vectors (8 channels); reductions in fp32 with fixed order (deterministic).
#include "b2u_internal.h"
#include "b2u_ptx.cuh"

namespace b2u {

#define B2U_CHECK_LAUNCH(name)                                                                           \
  do {                                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                                \
    if (e__ != cudaSuccess) return b2u::set_error(B2U_ERR_CUDA, name " launch: %s", cudaGetErrorString(e__)); \
    b2u::note_launch();                                                                                  \
  } while (0)

#ifndef B2U_FP32_VALIDATION   // the fp32 validation build (validation_fp32.cu) keeps only the SE fully connected kernels of this file
__device__ __forceinline__ void d_unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 d_pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// ------------------------------------------------------------------------------------------ depthwise 3x3
// One thread = one image column x 8 channels, marching down a strip of R output rows: every input row is fetched once
// (three 16-byte pieces: columns x-1, x, x+1; the horizontal neighbours are other lanes' lines in L1) and feeds the three
// output rows it overlaps, held in a rolling register window; the chunk's 72 weights stay in registers.
// The fetches are cp.async copies into thread-private shared-memory slots, kDwStages rows deep, so the bytes in flight
// (what an HBM-bound kernel needs) cost no registers and no block-level synchronisation; out-of-image pieces are
// zero-filled by the copy itself (src-size 0).
// Traffic: (R+2)/R reads + 1 write per element.  w: fp32 [C][9]; FLIP uses w[8 - tap] (the data gradient).
constexpr int kDwRows = 8;
constexpr int kDwStages = 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool FLIP>
__global__ void __launch_bounds__(128, 3)
dwconv3x3_strip_kernel(const uint4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       uint4* __restrict__ y, int H, int W, int C8) {
  __shared__ uint4 slots[kDwStages][3][128];      // [stage][column piece][thread]
  const unsigned t = blockIdx.x * 128u + threadIdx.x;
  if (t >= static_cast<unsigned>(W) * C8) return;   // no block-level barrier below: early exit is safe
  const int wx = t / C8, c = t - wx * C8;
  const int h0 = blockIdx.y * kDwRows, n = blockIdx.z;
  const uint4* img = x + static_cast<size_t>(n) * H * W * C8 + c;
  const bool left = wx > 0, right = wx + 1 < W;
  auto fetch_row = [&](int i) {        // strip-relative input row i -> stage i % kDwStages
    const int r = h0 - 1 + i;
    const bool ok = r >= 0 && r < H;
    const uint4* row = img + (static_cast<size_t>(ok ? r : 0) * W + wx) * C8;
    const int st = i % kDwStages;
    cp_async16(smem_u32(&slots[st][0][threadIdx.x]), left ? row - C8 : row, ok && left);
    cp_async16(smem_u32(&slots[st][1][threadIdx.x]), row, ok);
    cp_async16(smem_u32(&slots[st][2][threadIdx.x]), right ? row + C8 : row, ok && right);
  };
#pragma unroll
  for (int i = 0; i < kDwStages - 1; ++i) { fetch_row(i); cp_async_commit(); }
  float wt[9][8];                       // [tap as applied][channel]
  {
    float wf[72];
    const float4* wp = reinterpret_cast<const float4*>(w + static_cast<size_t>(c) * 72);
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      const float4 v = __ldg(wp + i);
      wf[4 * i] = v.x; wf[4 * i + 1] = v.y; wf[4 * i + 2] = v.z; wf[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int k = 0; k < 8; ++k) wt[tap][k] = wf[k * 9 + (FLIP ? 8 - tap : tap)];
  }
  float b8[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) b8[k] = bias ? __ldg(bias + c * 8 + k) : 0.f;
  uint4* out = y + static_cast<size_t>(n) * H * W * C8 + t;
  float a0[8], a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a0[k] = b8[k]; a1[k] = b8[k]; }
#pragma unroll
  for (int i = 0; i < kDwRows + 2; ++i) {
    const int r = h0 - 1 + i;           // input row; feeds output rows r-1 (kernel row 2), r (1), r+1 (0)
    if (i + kDwStages - 1 < kDwRows + 2) fetch_row(i + kDwStages - 1);
    cp_async_commit();                  // one group per iteration (possibly empty) keeps the wait count uniform
    cp_async_wait<kDwStages - 1>();     // row i has landed
    const int st = i % kDwStages;
    float f[3][8];
    d_unpack8(slots[st][0][threadIdx.x], f[0]);
    d_unpack8(slots[st][1][threadIdx.x], f[1]);
    d_unpack8(slots[st][2][threadIdx.x], f[2]);
#pragma unroll
    for (int k = 0; k < 8; ++k) a2[k] = b8[k];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int k = 0; k < 8; ++k) {     // rows/columns outside the image arrive as zeros
        a0[k] = fmaf(f[s][k], wt[6 + s][k], a0[k]);
        a1[k] = fmaf(f[s][k], wt[3 + s][k], a1[k]);
        a2[k] = fmaf(f[s][k], wt[s][k], a2[k]);
      }
    const int o = r - 1;                // output row completed by this input row
    if (i >= 2 && o < H) out[static_cast<size_t>(o) * W * C8] = d_pack8(a0);
#pragma unroll
    for (int k = 0; k < 8; ++k) { a0[k] = a1[k]; a1[k] = a2[k]; }
  }
}

// Weight + bias gradient: same marching scheme and cp.async pipeline (four pieces per row: dy and three columns of x);
// thread = (column, 8-channel chunk) keeps the 9 tap sums and the bias sum of its chunk in registers over all its work
// items (image, strip of R rows, column tile), then the block folds its column lanes in shared memory and writes one
// partial [C][10] (9 taps + bias).  Traffic: (R+2)/R x + 1 dy reads.
constexpr int kDwgStages = 3;

__global__ void __launch_bounds__(128, 3)
dwconv3x3_wgrad_strip_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, float* __restrict__ partial, int N,
                             int H, int W, int C8) {
  __shared__ uint4 slots[kDwgStages][4][128];      // [stage][dy, x-1, x, x+1][thread]; reused by the final fold
  float* sred = reinterpret_cast<float*>(slots);   // [128][10]
  const int tid = threadIdx.x;
  const int cpt = C8 < 128 ? C8 : 128;     // chunks handled per pass
  const int xs = 128 / cpt;                // columns per block
  const int cc = tid % cpt, xi = tid / cpt;
  const int strips = (H + kDwRows - 1) / kDwRows, xtiles = (W + xs - 1) / xs;
  const long long items = static_cast<long long>(N) * strips * xtiles;
  for (int c0 = 0; c0 < C8; c0 += cpt) {
    const int c = c0 + cc;
    float acc[10][8];
#pragma unroll
    for (int a = 0; a < 10; ++a)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[a][k] = 0.f;
    if (c < C8 && xi < xs) {
      for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int xt = static_cast<int>(it % xtiles);
        const long long q = it / xtiles;
        const int h0 = static_cast<int>(q % strips) * kDwRows;
        const long long n = q / strips;
        const int wx = xt * xs + xi;
        if (wx >= W) continue;
        const bool left = wx > 0, right = wx + 1 < W;
        const uint4* xi_ = x + (n * H * W + wx) * C8 + c;
        const uint4* di_ = dy + (n * H * W + wx) * C8 + c;
        auto fetch_row = [&](int i) {
          const int r = h0 - 1 + i, o = r + 1;      // x row r pairs with dy rows r-1 (tap row 2), r (1), r+1 (0)
          const bool dok = i < kDwRows && o < H, xok = r >= 0 && r < H;
          const uint4* drow = di_ + static_cast<size_t>(dok ? o : 0) * W * C8;
          const uint4* row = xi_ + static_cast<size_t>(xok ? r : 0) * W * C8;
          const int st = i % kDwgStages;
          cp_async16(smem_u32(&slots[st][0][tid]), drow, dok);
          cp_async16(smem_u32(&slots[st][1][tid]), left ? row - C8 : row, xok && left);
          cp_async16(smem_u32(&slots[st][2][tid]), row, xok);
          cp_async16(smem_u32(&slots[st][3][tid]), right ? row + C8 : row, xok && right);
        };
#pragma unroll
        for (int i = 0; i < kDwgStages - 1; ++i) { fetch_row(i); cp_async_commit(); }
        float d0[8], d1[8], d2[8];          // dy rows r-1, r, r+1 (zero outside the strip / image)
#pragma unroll
        for (int k = 0; k < 8; ++k) { d0[k] = 0.f; d1[k] = 0.f; }
#pragma unroll
        for (int i = 0; i < kDwRows + 2; ++i) {
          if (i + kDwgStages - 1 < kDwRows + 2) fetch_row(i + kDwgStages - 1);
          cp_async_commit();
          cp_async_wait<kDwgStages - 1>();
          const int st = i % kDwgStages;
          d_unpack8(slots[st][0][tid], d2);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[9][k] += d2[k];
          float f[3][8];
          d_unpack8(slots[st][1][tid], f[0]); d_unpack8(slots[st][2][tid], f[1]); d_unpack8(slots[st][3][tid], f[2]);
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              acc[6 + s][k] = fmaf(d0[k], f[s][k], acc[6 + s][k]);
              acc[3 + s][k] = fmaf(d1[k], f[s][k], acc[3 + s][k]);
              acc[s][k] = fmaf(d2[k], f[s][k], acc[s][k]);
            }
#pragma unroll
          for (int k = 0; k < 8; ++k) { d0[k] = d1[k]; d1[k] = d2[k]; }
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();                       // the slots become the fold buffer
    // fold the column lanes, one channel element at a time.  Narrow tensors have up to 64 column lanes per chunk and only
    // cpt * 10 outputs, so each output is first summed by `lpo` lanes in parallel and then by one thread.
    const int nout = cpt * 10;
    const int lpo = nout >= 128 ? 1 : 128 / nout;           // lanes per output (block-uniform)
    float* sfold = sred + 128 * 10;                         // [128], inside the 24 KB slot area
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int a = 0; a < 10; ++a) sred[tid * 10 + a] = acc[a][k];
      __syncthreads();
      if (lpo == 1) {
        for (int o = tid; o < nout; o += 128) {
          const int oc = o / 10, oa = o - oc * 10;
          float t = 0.f;
          for (int r = 0; r < xs; ++r) t += sred[(r * cpt + oc) * 10 + oa];
          if (c0 + oc < C8) partial[(static_cast<size_t>(blockIdx.x) * C8 * 8 + (c0 + oc) * 8 + k) * 10 + oa] = t;
        }
      } else {
        float t = 0.f;
        if (tid < nout * lpo) {
          const int o = tid / lpo, part = tid - o * lpo;
          const int oc = o / 10, oa = o - oc * 10;
          for (int r = part; r < xs; r += lpo) t += sred[(r * cpt + oc) * 10 + oa];
        }
        sfold[tid] = t;
        __syncthreads();
        if (tid < nout) {
          const int oc = tid / 10, oa = tid - oc * 10;
          float u = 0.f;
          for (int q = 0; q < lpo; ++q) u += sfold[tid * lpo + q];
          if (c0 + oc < C8) partial[(static_cast<size_t>(blockIdx.x) * C8 * 8 + (c0 + oc) * 8 + k) * 10 + oa] = u;
        }
      }
      __syncthreads();
    }
  }
}

// dw[c][a] / db[c] = sum_r partial[r][c][10]  (fixed order: deterministic)
__global__ void __launch_bounds__(256)
dw_reduce_split_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db, int rows, int L) {
  __shared__ float sred[8][33];
  const int col = threadIdx.x & 31, lane_r = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + col;
  float acc = 0.f;
  if (i < L)
    for (int r = lane_r; r < rows; r += 8) acc += __ldg(partial + static_cast<size_t>(r) * L + i);
  sred[lane_r][col] = acc;
  __syncthreads();
  if (lane_r == 0 && i < L) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sred[k][col];
    const int c = i / 10, a = i - c * 10;
    if (a < 9) { if (dw) dw[c * 9 + a] = t; }
    else if (db) db[c] = t;
  }
}

// ------------------------------------------------------------------------------------------ per-(n, c) reductions
// MODE 0: sum x; MODE 1: sum a*b.  grid (slabs, N); partial[n][slab][C]
constexpr int kSpatialSlabs = 32;
template <int MODE>
__global__ void __launch_bounds__(256)
spatial_colsum_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, float* __restrict__ partial, long long HW, int C8) {
  extern __shared__ float sred[];          // [256][8]
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const int cpt = C8 < 256 ? C8 : 256;
  const int lanes = 256 / cpt;
  const int cc = tid % cpt, rr = tid / cpt;
  const uint4* pa = a + static_cast<size_t>(n) * HW * C8;
  const uint4* pb = MODE == 1 ? b + static_cast<size_t>(n) * HW * C8 : nullptr;
  for (int c0 = 0; c0 < C8; c0 += cpt) {
    const int c = c0 + cc;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < C8 && rr < lanes) {
      for (long long p = static_cast<long long>(blockIdx.x) * lanes + rr; p < HW; p += static_cast<long long>(gridDim.x) * lanes) {
        float f[8];
        d_unpack8(__ldg(pa + p * C8 + c), f);
        if (MODE == 1) {
          float g[8];
          d_unpack8(__ldg(pb + p * C8 + c), g);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(f[k], g[k], acc[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sred[tid * 8 + k] = acc[k];
    __syncthreads();
    if (rr == 0 && c < C8) {
      for (int r = 1; r < lanes; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += sred[(r * cpt + cc) * 8 + k];
      float* out = partial + (static_cast<size_t>(n) * gridDim.x + blockIdx.x) * C8 * 8 + c * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) out[k] = acc[k];
    }
    __syncthreads();
  }
}
__global__ void spatial_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int slabs, int C, float scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (c >= C) return;
  float t = 0.f;
  for (int s = 0; s < slabs; ++s) t += partial[(static_cast<size_t>(n) * slabs + s) * C + c];
  out[static_cast<size_t>(n) * C + c] = t * scale;
}

// y[n,p,c] = x[n,p,c] * s[n][c] + a[n][c]
__global__ void scale_nc_kernel(const uint4* __restrict__ x, const float* __restrict__ s, const float* __restrict__ a,
                                uint4* __restrict__ y, int HW, int C8) {
  const unsigned t = blockIdx.y * blockDim.x + threadIdx.x;      // position inside the image: pixel * C8 + chunk
  const int n = blockIdx.x;
  if (t >= static_cast<unsigned>(HW) * C8) return;
  const int c = t % C8;
  float f[8];
  const size_t i = static_cast<size_t>(n) * HW * C8 + t;
  d_unpack8(__ldg(x + i), f);
  const float* sp = s + static_cast<size_t>(n) * C8 * 8 + c * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], __ldg(sp + k), a ? __ldg(a + static_cast<size_t>(n) * C8 * 8 + c * 8 + k) : 0.f);
  y[i] = d_pack8(f);
}

#endif  // !B2U_FP32_VALIDATION

// ------------------------------------------------------------------------------------------ SE fully connected part
// one block per image: hidden = relu(W1 p + b1), scale = sigmoid(W2 hidden + b2)
__global__ void se_fc_fwd_kernel(const float* __restrict__ pooled, const float* __restrict__ w1, const float* __restrict__ b1,
                                 const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ hidden,
                                 float* __restrict__ scale, int C, int Cp, int R) {
  extern __shared__ float sm[];            // [C] pooled, [R] hidden
  const int n = blockIdx.x;
  float* sp = sm; float* sh = sm + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sp[c] = pooled[static_cast<size_t>(n) * Cp + c];
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float acc = b1[r];
    for (int c = 0; c < C; ++c) acc = fmaf(w1[static_cast<size_t>(r) * C + c], sp[c], acc);
    acc = fmaxf(acc, 0.f);
    sh[r] = acc;
    hidden[static_cast<size_t>(n) * R + r] = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    float v = 0.f;                         // padded channels carry zeros: their scale is irrelevant, keep it 0
    if (c < C) {
      float acc = b2[c];
      for (int r = 0; r < R; ++r) acc = fmaf(w2[static_cast<size_t>(c) * R + r], sh[r], acc);
      v = 1.f / (1.f + expf(-acc));
    }
    scale[static_cast<size_t>(n) * Cp + c] = v;
  }
}
// backward, per image: dpre2 = ds * s (1-s); dh = W2^T dpre2; dpre1 = dh (h > 0); dp = W1^T dpre1 / HW-scaling by caller
__global__ void se_fc_bwd_image_kernel(const float* __restrict__ dscale, const float* __restrict__ hidden,
                                       const float* __restrict__ scale, const float* __restrict__ w1, const float* __restrict__ w2,
                                       float* __restrict__ dpre2, float* __restrict__ dpre1, float* __restrict__ dpooled, int C,
                                       int Cp, int R, float dp_scale) {
  extern __shared__ float sm[];            // [C] dpre2, [R] dpre1
  const int n = blockIdx.x;
  float* s2 = sm; float* s1 = sm + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s = scale[static_cast<size_t>(n) * Cp + c];
    const float v = dscale[static_cast<size_t>(n) * Cp + c] * s * (1.f - s);
    s2[c] = v;
    dpre2[static_cast<size_t>(n) * C + c] = v;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(w2[static_cast<size_t>(c) * R + r], s2[c], acc);
    acc = hidden[static_cast<size_t>(n) * R + r] > 0.f ? acc : 0.f;
    s1[r] = acc;
    dpre1[static_cast<size_t>(n) * R + r] = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    float acc = 0.f;
    if (c < C)
      for (int r = 0; r < R; ++r) acc = fmaf(w1[static_cast<size_t>(r) * C + c], s1[r], acc);
    dpooled[static_cast<size_t>(n) * Cp + c] = acc * dp_scale;
  }
}
// weight gradients over the batch: dW2[c][r] = sum_n dpre2[n][c] h[n][r]; dW1[r][c] = sum_n dpre1[n][r] p[n][c]
__global__ void se_fc_bwd_weights_kernel(const float* __restrict__ dpre2, const float* __restrict__ dpre1,
                                         const float* __restrict__ hidden, const float* __restrict__ pooled,
                                         float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                         float* __restrict__ db2, int N, int C, int Cp, int R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * R) {
    const int c = i / R, r = i % R;
    float a2 = 0.f, a1 = 0.f;
    for (int n = 0; n < N; ++n) {
      a2 = fmaf(dpre2[static_cast<size_t>(n) * C + c], hidden[static_cast<size_t>(n) * R + r], a2);
      a1 = fmaf(dpre1[static_cast<size_t>(n) * R + r], pooled[static_cast<size_t>(n) * Cp + c], a1);
    }
    if (dw2) dw2[static_cast<size_t>(c) * R + r] = a2;
    if (dw1) dw1[static_cast<size_t>(r) * C + c] = a1;
  }
  if (i < C && db2) { float t = 0.f; for (int n = 0; n < N; ++n) t += dpre2[static_cast<size_t>(n) * C + i]; db2[i] = t; }
  if (i < R && db1) { float t = 0.f; for (int n = 0; n < N; ++n) t += dpre1[static_cast<size_t>(n) * R + i]; db1[i] = t; }
}

#ifndef B2U_FP32_VALIDATION
static inline dim3 rgrid(long long rows, int row_items, int block) {
  return dim3(static_cast<unsigned>(rows), static_cast<unsigned>((row_items + block - 1) / block), 1);
}
#endif

}  // namespace b2u

extern "C" {
using namespace b2u;

#ifndef B2U_FP32_VALIDATION
int b2u_dwconv3x3_fwd(const void* x, const float* w, const float* bias, void* y, int N, int H, int W, int C, int flip,
                      void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C % 8 != 0) return set_error(B2U_ERR_SHAPE, "dwconv3x3: bad shape");
  if (N > 65535 || (H + kDwRows - 1) / kDwRows > 65535) return set_error(B2U_ERR_SHAPE, "dwconv3x3: N or H too large");
  const dim3 grid((static_cast<unsigned>(W) * (C / 8) + 127) / 128, (H + kDwRows - 1) / kDwRows, N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (reinterpret_cast<uintptr_t>(w) % 16 != 0) return set_error(B2U_ERR_ARG, "dwconv3x3: weights must be 16-byte aligned");
  if (flip) dwconv3x3_strip_kernel<true><<<grid, 128, 0, st>>>(static_cast<const uint4*>(x), w, bias, static_cast<uint4*>(y), H, W, C / 8);
  else      dwconv3x3_strip_kernel<false><<<grid, 128, 0, st>>>(static_cast<const uint4*>(x), w, bias, static_cast<uint4*>(y), H, W, C / 8);
  B2U_CHECK_LAUNCH("dwconv3x3");
  return 0;
}

static const int kDwBlocks = 3 * 148;
size_t b2u_dwconv3x3_wgrad_workspace(int C) { return (static_cast<size_t>(kDwBlocks) + 1) * C * 10 * sizeof(float); }

// dw: [C][9] fp32, db: [C] (either may be NULL)
int b2u_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, float* db, void* ws, size_t ws_bytes, int N, int H, int W,
                        int C, void* stream);

int b2u_spatial_reduce_workspace_floats(int N, int C) { return N * kSpatialSlabs * C; }

// out[n][c] = scale * sum over the H*W pixels of image n of a (b == NULL) or a*b
int b2u_spatial_reduce(const void* a, const void* b, float* out, void* ws, size_t ws_bytes, int N, long long HW, int C,
                       float scale, void* stream) {
  if (N <= 0 || HW <= 0 || C % 8 != 0) return set_error(B2U_ERR_SHAPE, "spatial_reduce: bad shape");
  if (!ws || ws_bytes < static_cast<size_t>(N) * kSpatialSlabs * C * sizeof(float)) return set_error(B2U_ERR_ARG, "spatial_reduce: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(kSpatialSlabs, N);
  if (b) spatial_colsum_kernel<1><<<grid, 256, 256 * 8 * sizeof(float), st>>>(static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<float*>(ws), HW, C / 8);
  else   spatial_colsum_kernel<0><<<grid, 256, 256 * 8 * sizeof(float), st>>>(static_cast<const uint4*>(a), nullptr, static_cast<float*>(ws), HW, C / 8);
  B2U_CHECK_LAUNCH("spatial_colsum");
  spatial_finalize_kernel<<<dim3((C + 127) / 128, N), 128, 0, st>>>(static_cast<const float*>(ws), out, kSpatialSlabs, C, scale);
  B2U_CHECK_LAUNCH("spatial_finalize");
  return 0;
}

// y = x * s[n][c] + a[n][c]  (a nullable); y may alias x
int b2u_scale_nc(const void* x, const float* s, const float* a, void* y, int N, long long HW, int C, void* stream) {
  if (N <= 0 || HW <= 0 || C % 8 != 0 || HW * (C / 8) >= (1ll << 31)) return set_error(B2U_ERR_SHAPE, "scale_nc: bad shape");
  scale_nc_kernel<<<rgrid(N, static_cast<int>(HW * (C / 8)), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), s, a, static_cast<uint4*>(y), static_cast<int>(HW), C / 8);
  B2U_CHECK_LAUNCH("scale_nc");
  return 0;
}

#endif  // !B2U_FP32_VALIDATION

// SE fully connected part.  pooled/scale/dscale/dpooled are [N][Cp] (Cp >= C: channel padding of the activations),
// hidden [N][R]; w1 [R][C], w2 [C][R] (nn.Linear layouts).
int b2u_se_fc_fwd(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2, float* hidden,
                  float* scale, int N, int C, int Cp, int R, void* stream) {
  if (N <= 0 || C <= 0 || R <= 0 || Cp < C || (C + R) * sizeof(float) > 48 * 1024) return set_error(B2U_ERR_SHAPE, "se_fc_fwd: bad shape");
  se_fc_fwd_kernel<<<N, 256, (C + R) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(pooled, w1, b1, w2, b2, hidden, scale, C, Cp, R);
  B2U_CHECK_LAUNCH("se_fc_fwd");
  return 0;
}
// scratch: N*(C+R) floats.  dpooled = W1^T dpre1 * dp_scale (dp_scale = 1/HW turns it into the per-pixel addend)
int b2u_se_fc_bwd(const float* dscale, const float* pooled, const float* hidden, const float* scale, const float* w1,
                  const float* w2, float* dpooled, float* dw1, float* db1, float* dw2, float* db2, float* scratch, int N,
                  int C, int Cp, int R, float dp_scale, void* stream) {
  if (N <= 0 || C <= 0 || R <= 0 || Cp < C || (C + R) * sizeof(float) > 48 * 1024) return set_error(B2U_ERR_SHAPE, "se_fc_bwd: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* dpre2 = scratch; float* dpre1 = scratch + static_cast<size_t>(N) * C;
  se_fc_bwd_image_kernel<<<N, 256, (C + R) * sizeof(float), st>>>(dscale, hidden, scale, w1, w2, dpre2, dpre1, dpooled, C, Cp, R, dp_scale);
  B2U_CHECK_LAUNCH("se_fc_bwd_image");
  if (dw1 || dw2 || db1 || db2) {
    se_fc_bwd_weights_kernel<<<(C * R + 255) / 256, 256, 0, st>>>(dpre2, dpre1, hidden, pooled, dw1, db1, dw2, db2, N, C, Cp, R);
    B2U_CHECK_LAUNCH("se_fc_bwd_weights");
  }
  return 0;
}

#ifndef B2U_FP32_VALIDATION
int b2u_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, float* db, void* ws, size_t ws_bytes, int N, int H, int W,
                        int C, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C % 8 != 0) return set_error(B2U_ERR_SHAPE, "dwconv3x3_wgrad: bad shape");
  if (!ws || ws_bytes < b2u_dwconv3x3_wgrad_workspace(C)) return set_error(B2U_ERR_ARG, "dwconv3x3_wgrad: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(ws);
  const int C8 = C / 8, cpt = C8 < 128 ? C8 : 128, xs = 128 / cpt;
  const long long items = static_cast<long long>(N) * ((H + kDwRows - 1) / kDwRows) * ((W + xs - 1) / xs);
  const int blocks = static_cast<int>(items < kDwBlocks ? items : kDwBlocks);
  dwconv3x3_wgrad_strip_kernel<<<blocks, 128, 0, st>>>(static_cast<const uint4*>(x), static_cast<const uint4*>(dy),
                                                                              partial, N, H, W, C8);
  B2U_CHECK_LAUNCH("dwconv3x3_wgrad");
  dw_reduce_split_kernel<<<(C * 10 + 31) / 32, 256, 0, st>>>(partial, dw, db, blocks, C * 10);
  B2U_CHECK_LAUNCH("dw_reduce_split");
  return 0;
}
#endif  // !B2U_FP32_VALIDATION

}  // extern "C"
