// b2u_internal.h -- declarations shared by the translation units of libb200unet.so
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define B2U_OK 0
#define B2U_ERR_SHAPE 1
#define B2U_ERR_CUDA 2
#define B2U_ERR_DRIVER 3
#define B2U_ERR_NCCL 4
#define B2U_ERR_ARG 5

namespace b2u {

// records a thread-local message, returns `code`
int set_error(int code, const char* fmt, ...);
int num_sms();
// counts kernel launches made by this library (bench.py reports them as gpu_launches)
void note_launch(int n = 1);

// NHWC bf16 tensor viewed as a rank-4 TMA tensor (C, W, H, N); `cpitch` = elements between pixels.
int make_tmap_nhwc(CUtensorMap* out, const void* base, int N, int H, int W, int C, int boxC, int boxW, int boxH,
                   CUtensorMapSwizzle swz, int cpitch = 0);
// row-major bf16 matrix [rows][cols] viewed as rank-2 (cols, rows)
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, int boxCols, int boxRows,
                 CUtensorMapSwizzle swz);

struct ConvLaunch {
  const void* x0 = nullptr; int C0 = 0;     // source 0 (NHWC bf16)
  const void* x1 = nullptr; int C1 = 0;     // optional source 1 (virtual concat), same N,H,W
  const void* up_low = nullptr;             // instead of x1: source 1 = this [N,H/2,W/2,C1] tensor, bilinearly up-sampled 2x inside the kernel
  void* up_out = nullptr;                   // with up_low, nullable: by-product copy of the up-sampled tensor [N,H,W,C1]
  const void* wpacked = nullptr;            // bf16 [Cout][taps*(C0+C1)]
  const float* bias = nullptr;              // fp32 [Cout] or null
  const float* scale = nullptr;             // fp32 [Cout] or null: y = acc * scale + bias (eval-mode BatchNorm folded into the conv)
  void* y0 = nullptr; void* y1 = nullptr;   // outputs (NHWC bf16); y1 receives channels >= split_c
  int split_c = 0;
  const __nv_bfloat16* mask = nullptr; int mask_c = 0;
  const unsigned long long* mask_bits = nullptr;   // flags bit3: the mask as [N,H,W,Cout/64] 64-bit words, one bit per channel
  unsigned long long* bits_out = nullptr;          // forward + ReLU: receives (y > 0) in that layout
  int N = 0, H = 0, W = 0, Cout = 0, taps = 9;
  int flags = 0;                            // bit0 relu, bit1 mask, bit2 classifier head (fp32 NCHW logits), bit3 bit mask
  float* stat_partial = nullptr;            // optional [m tiles][2][Cout] fp32: per-tile column sums (z, z^2) of the stored output
  float* head_out = nullptr; int head_cls = 0;   // bit2: logits [N][head_cls][H][W]; Cout must be 64 = [hi(32) | lo(32)] weights
  int bn_override = 0;
  int tile_flags = 0;                       // bit0: force one M tile per CTA step (debug / tests)
};
int launch_conv(const ConvLaunch& a, cudaStream_t st);
int conv_m_tiles(int N, int H, int W, int Cout, int taps, int bn_override);   // rows of ConvLaunch::stat_partial
int stacked_tiles(int bn, int taps, int H, int tile_flags, bool masked, bool decoder);   // M tiles per CTA step

struct WgradLaunch {
  const void* x0 = nullptr; int C0 = 0;
  const void* x1 = nullptr; int C1 = 0;
  const void* dz = nullptr; int Cout = 0;   // NHWC bf16 gradient wrt pre-activation
  float* partial = nullptr;                 // workspace [splits][taps][Cout][Cin] fp32
  int N = 0, H = 0, W = 0, taps = 9;
  int splits = 0;                           // 0: choose
};

}  // namespace b2u
