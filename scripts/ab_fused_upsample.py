#!/usr/bin/env python
"""A/B of the decoder conv with the bilinear up-sampling folded into its operand load (b2u_decoder_conv_fprop) against the
two-kernel path (b2u_upsample2x_fwd + b2u_conv_fprop over the virtual concat), at the four unetUp shapes of the
headline Unet-VGG16 step (batch 16, 512x512) -- nets/unet.py:16-18 of the reference.  CUDA events, medians of 20 runs;
the tensors of one shape total 0.3-1.7 GB, far beyond the 126 MB L2.

  python scripts/ab_fused_upsample.py [--json out.json]
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_pytorch_b200 as b2u  # noqa: E402

SHAPES = [  # name, H (= W) of the conv, C_skip, C_low, C_out
    ("up_concat4.conv1", 64, 512, 512, 512),
    ("up_concat3.conv1", 128, 256, 512, 256),
    ("up_concat2.conv1", 256, 128, 256, 128),
    ("up_concat1.conv1", 512, 64, 128, 64),
]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    ops, dev = b2u.ops, torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    N, rows = args.batch, []
    for name, hw, c0, c1, co in SHAPES:
        skip = torch.randn(N, hw, hw, c0, generator=g).to(torch.bfloat16).to(dev)
        low = torch.randn(N, hw // 2, hw // 2, c1, generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn(co, c0 + c1, 3, 3, generator=g) / ((c0 + c1) * 9) ** 0.5).to(dev)
        bias = torch.zeros(co, device=dev)
        wf, _ = ops.pack_weights(w, want_dgrad=False)
        up = torch.empty(N, hw, hw, c1, dtype=torch.bfloat16, device=dev)
        out = torch.empty(N, hw, hw, co, dtype=torch.bfloat16, device=dev)
        t_up = timed(lambda: ops.upsample2x(low, out=up))
        t_conv = timed(lambda: ops.conv_fprop(skip, wf, bias, co, relu=True, x1=up, out=out))
        ref = out.clone()
        t_fused_train = timed(lambda: ops.decoder_conv_fprop(skip, low, wf, bias, co, relu=True, out=out, up_out=up))
        same = bool(torch.equal(out.view(torch.int16), ref.view(torch.int16)))
        t_fused_infer = timed(lambda: ops.decoder_conv_fprop(skip, low, wf, bias, co, relu=True, out=out))
        flop = 2.0 * N * hw * hw * co * (c0 + c1) * 9
        rows.append({"layer": name, "shape": f"{N}x{hw}x{hw} {c0}+{c1}->{co}", "upsample_ms": round(t_up, 4), "conv_ms": round(t_conv, 4),
                     "separate_ms": round(t_up + t_conv, 4), "fused_with_byproduct_ms": round(t_fused_train, 4),
                     "fused_inference_ms": round(t_fused_infer, 4), "bit_identical": same,
                     "conv_tflops": round(flop / t_conv * 1e-9, 1), "fused_tflops": round(flop / t_fused_train * 1e-9, 1)})
        print(rows[-1], flush=True)
    tot = {k: round(sum(r[k] for r in rows), 4) for k in ("separate_ms", "fused_with_byproduct_ms", "fused_inference_ms")}
    print("total", tot)
    if args.json:
        json.dump({"rows": rows, "total": tot}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
