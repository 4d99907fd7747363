"""Times the Cout = 64 tensor-core convolutions at 512x512 under the tile variants the C ABI exposes (bn_override bit 16 =
one 8x16-pixel M tile per CTA step, bit 17 = at most two stacked).  Usage: python scripts/tile_variants_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_pytorch_b200 as b2u  # noqa: E402

ops = b2u.ops
dev = torch.device("cuda:0")
N, H, W = 16, 512, 512


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for c0, c1 in ((64, 0), (64, 128)):
    x0 = torch.randn((N, H, W, c0), device=dev).bfloat16()
    x1 = torch.randn((N, H, W, c1), device=dev).bfloat16() if c1 else None
    w = torch.randn((64, c0 + c1, 3, 3), device=dev) * 0.05
    wf, wd = ops.pack_weights(w)
    bias = torch.zeros(64, device=dev)
    out = torch.empty((N, H, W, 64), device=dev, dtype=torch.bfloat16)
    mask = torch.randn((N, H, W, c0), device=dev).bfloat16()
    dz = torch.randn((N, H, W, 64), device=dev).bfloat16()
    dx = torch.empty((N, H, W, c0), device=dev, dtype=torch.bfloat16)
    flops = 2.0 * N * H * W * 64 * (c0 + c1) * 9
    ref = ops.conv_fprop(x0, wf, bias, 64, x1=x1).float()
    for flag, name in ((2 << 16, "two stacked M tiles"), (1 << 16, "one M tile"), (0, "default (four stacked if unmasked)")):
        got = ops.conv_fprop(x0, wf, bias, 64, x1=x1, bn=flag).float()
        assert torch.equal(got, ref), name
        t = timed(lambda: ops.conv_fprop(x0, wf, bias, 64, x1=x1, out=out, bn=flag))
        print(f"fprop {c0}+{c1}->64  {name:22s} {t:.3f} ms  {flops / t / 1e9:.0f} TFLOP/s")
    if not c1:
        wd0 = wd
        refm = ops.conv_dgrad(dz, wd0, c0, mask=mask).float()
        for flag, name in ((2 << 16, "two stacked M tiles"), (1 << 16, "one M tile"), (0, "default (four stacked if unmasked)")):
            assert torch.equal(ops.conv_dgrad(dz, wd0, c0, mask=mask, bn=flag).float(), refm), name
            t = timed(lambda: ops.conv_dgrad(dz, wd0, c0, mask=mask, out0=dx, bn=flag))
            print(f"dgrad 64->{c0} masked  {name:22s} {t:.3f} ms  {flops / t / 1e9:.0f} TFLOP/s")
            t = timed(lambda: ops.conv_dgrad(dz, wd0, c0, out0=dx, bn=flag))
            print(f"dgrad 64->{c0} plain   {name:22s} {t:.3f} ms  {flops / t / 1e9:.0f} TFLOP/s")
