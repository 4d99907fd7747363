"""Achieved HBM bandwidth of the memory-bound kernels, one at a time, on tensors larger than L2 (CUDA events, 10 reps
after 3 warm-ups).  Bytes are ALGORITHMIC (each operand read once, each result written once), so GB/s below the copy
figure means re-reads or latency-bound code.   python scripts/membound_bench.py [--out profiles/rN_membound.json]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_pytorch_b200 import ops

BF = torch.bfloat16


def ev(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); ap.add_argument("--only", default=None); ap.add_argument("--shapes", default=None, help="e.g. 16x512x512x32,8x256x256x16")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    res = {}
    shapes = ((16, 512, 512, 64), (16, 128, 128, 256), (16, 512, 512, 32)) if not args.shapes else [tuple(int(v) for v in t.split('x')) for t in args.shapes.split(',')]
    for (N, H, W, C) in shapes:
        shp = (N, H, W, C)
        T = N * H * W * C * 2          # bytes of one bf16 tensor
        x = torch.randn(shp, device=dev).to(BF); dy = torch.randn(shp, device=dev).to(BF)
        y = torch.empty_like(x); z = torch.randn(shp, device=dev).to(BF)
        gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
        rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
        wdw = torch.randn(C, 9, device=dev); bdw = torch.zeros(C, device=dev)
        s_nc = torch.rand(N, C, device=dev)
        _, mean, invstd = ops.bn_fwd_train(z, gamma, beta, rm, rv, out=y)
        half = torch.empty((N, H // 2, W // 2, C), dtype=BF, device=dev)
        dhalf = torch.randn((N, H // 2, W // 2, C), device=dev).to(BF)
        cases = {
            "torch_copy": (lambda: y.copy_(x), 2 * T),
            "bn_fwd_train": (lambda: ops.bn_fwd_train(z, gamma, beta, rm, rv, out=y), 3 * T),
            "bn_bwd": (lambda: ops.bn_bwd(dy, None, z, gamma, mean, invstd, beta=beta, out=y), 5 * T),
            "bn_bwd_with_y": (lambda: ops.bn_bwd(dy, x, z, gamma, mean, invstd, out=y), 7 * T),
            "dwconv3x3_fwd": (lambda: ops.dwconv3x3(x, wdw, bdw, out=y), 2 * T),
            "dwconv3x3_wgrad": (lambda: ops.dwconv3x3_wgrad(x, dy), 2 * T),
            "scale_nc": (lambda: ops.scale_nc(x, s_nc, out=y), 2 * T),
            "spatial_reduce": (lambda: ops.spatial_reduce(x), T),
            "spatial_reduce_dot": (lambda: ops.spatial_reduce(x, dy), 2 * T),
            "add_bf16": (lambda: ops.add_bf16(x, dy, out=y), 3 * T),
            "bias_grad": (lambda: ops.bias_grad(dy), T),
            "maxpool2x2_fwd": (lambda: ops.maxpool2x2(x, out=half), T + T // 4),
            "maxpool2x2_bwd": (lambda: ops.maxpool2x2_bwd(dhalf, x, relu_mask=False, out=y), 2 * T + T // 4),
            "upsample2x_fwd": (lambda: ops.upsample2x(half, out=y), T + T // 4),
            "upsample2x_bwd": (lambda: ops.upsample2x_bwd(dy, out=half), T + T // 4),
        }
        for name, (fn, nbytes) in cases.items():
            if args.only and args.only not in name:
                continue
            ms = ev(fn)
            res[f"{name}|{N}x{H}x{W}x{C}"] = {"ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1), "alg_bytes": nbytes}
        del x, dy, y, z, half, dhalf
        torch.cuda.empty_cache()
    base = max(v["GBps"] for k, v in res.items() if k.startswith("torch_copy")) if not args.only else None
    for k, v in res.items():
        if base:
            v["frac_of_copy"] = round(v["GBps"] / base, 3)
        print(f"{k:44s} {v['ms']:8.3f} ms {v['GBps']:8.1f} GB/s" + (f"  {v['frac_of_copy']:.2f}" if base else ""))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
