#!/usr/bin/env python
"""Warm 512x512 parity report of every model family against the float64 reference (tests/warm_parity.py).

  python scripts/parity_512.py [--families a,b,...] [--batch 8] [--warm 200] [--out gpurun_out/parity_512.json]

Prints one block per family: the product bf16 path, the reference in fp32, and the reference under torch bf16 autocast,
each against the reference evaluated in float64 on the same warm weights and the same structured batch."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEFAULT = [("unet_vgg", 21, False), ("unet_vgg", 2, True), ("unet_resnet50", 21, False), ("traditional", 2, True),
           ("traditional", 21, False), ("lightweight", 2, True), ("ultralight_large", 2, True), ("ultralight", 21, False),
           ("ultralight_large_optimized", 4, False)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", default=None)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=512)
    ap.add_argument("--warm", type=int, default=200)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_512.json"))
    ap.add_argument("--no-autocast", action="store_true")
    args = ap.parse_args()
    import unet_pytorch_b200 as b2u
    import warm_parity as WP
    cases = DEFAULT
    if args.families:
        want = args.families.split(",")
        cases = [c for c in DEFAULT if c[0] in want]
    rows = []
    for fam, C, medical in cases:
        print(f"== {fam} nc={C} {'medical' if medical else 'voc-like'} {args.hw}x{args.hw} batch {args.batch}", flush=True)
        r = WP.measure(b2u, fam, C, hw=args.hw, batch=args.batch, warm_steps=args.warm, medical=medical, log=lambda s: print(s, flush=True),
                       with_autocast=not args.no_autocast)
        for k in ("ours_bf16", "autocast_bf16", "ref_fp32"):
            if k in r:
                v = r[k]
                print(f"  {k:14s} logits {v['logits']:.2e}  grad global {v['grad_global']:.2e}  median {v['grad_median']:.2e}  "
                      f"worst {v['grad_worst']:.2e} ({v['grad_worst_name']})  argmax {100 * v['argmax_all']:.3f} %", flush=True)
        print(f"  loss fp64 {r['loss_fp64']:.5f} ours {r['loss_ours']:.5f}  ({r['seconds']:.0f} s)", flush=True)
        rows.append(r)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
