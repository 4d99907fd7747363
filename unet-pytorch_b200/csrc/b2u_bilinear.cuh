// b2u_bilinear.cuh -- the arithmetic of nn.UpsamplingBilinear2d(scale_factor=2) (bilinear, align_corners=True;
// nets/unet.py:13 of the reference), shared by the standalone upsample kernels (elementwise.cu) and by the decoder conv
// whose producer warps interpolate the low-resolution tensor straight into the A-operand stage (conv_igemm.cu), so both
// paths round identically: horizontal lerp in fp32, then the vertical one, one rounding to bf16 at the end.
#pragma once
#include "b2u_ptx.cuh"

namespace b2u {

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

__device__ __forceinline__ void src_index(int o, float scale, int in_size, int& i0, int& i1, float& lam) {
  const float src = scale * static_cast<float>(o);   // ATen area_pixel_compute_source_index(align_corners=True)
  i0 = static_cast<int>(src);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  lam = src - static_cast<float>(i0);
}
// For scale 2 with align_corners=True, output index o reads low-res indices i0(o), i1(o) that always lie in
// {j, j + 1} with j = floor((o - 1) / 2) (checked exhaustively in float arithmetic for every size up to 1024), so
// output rows 2i+1 and 2i+2 interpolate between the same two source rows i and i+1.  weight_of(o, i) is the weight of
// low-res index i in output o under ATen's formula.
__device__ __forceinline__ float weight_of(int o, int i, float scale, int in_size, int out_size) {
  if (o < 0 || o >= out_size) return 0.f;
  int i0, i1; float lam;
  src_index(o, scale, in_size, i0, i1, lam);
  return (i0 == i ? 1.f - lam : 0.f) + (i1 == i ? lam : 0.f);
}
// weights of low-res rows j and j + 1 in output row o (one index computation for both)
__device__ __forceinline__ void pair_weights(int o, int j, float scale, int in_size, int out_size, float& wa, float& wb) {
  wa = wb = 0.f;
  if (o < 0 || o >= out_size) return;
  int i0, i1; float lam;
  src_index(o, scale, in_size, i0, i1, lam);
  wa = (i0 == j ? 1.f - lam : 0.f) + (i1 == j ? lam : 0.f);
  wb = (i0 == j + 1 ? 1.f - lam : 0.f) + (i1 == j + 1 ? lam : 0.f);
}
// horizontal step: v = (1 - lw) a + lw b over 8 channels
__device__ __forceinline__ void hlerp8(const uint4& a, const uint4& b, float w0l, float lw, float* v) {
  float fa[8], fb[8];
  unpack8(a, fa); unpack8(b, fb);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = fmaf(lw, fb[k], w0l * fa[k]);
}
// vertical step: out = wa va + wb vb over 8 channels, rounded to bf16
__device__ __forceinline__ uint4 vlerp8(const float* va, const float* vb, float wa, float wb) {
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = fmaf(wb, vb[k], wa * va[k]);
  return pack8(v);
}

}  // namespace b2u
