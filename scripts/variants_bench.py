"""Side measurements for BASELINE configs[2] and configs[3] (not the headline bench line): training img/s of the
Unet-ResNet50, TraditionalUnet and UltraLightweightUnet variants at 512x512 on one B200, with the step's time split by
C-ABI entry point (CUDA events around every call) so the memory-bound kernels can be read against HBM bandwidth.
    python scripts/variants_bench.py [--models a,b] [--batch 16] [--out profiles/rN_variants.json]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_pytorch_b200 as b2u
from unet_pytorch_b200 import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="lightweight,ultralight_large,ultralight,ultralight_large_optimized,traditional,unet_resnet50")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--by-shape", action="store_true")
    ap.add_argument("--cpu-baseline", action="store_true", help="also time the CPU port (bench.py's cpu_baseline leg) per model")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    res = {"device": torch.cuda.get_device_name(0), "batch": args.batch, "image": "512x512", "classes": args.classes}
    for model in args.models.split(","):
        C = 21 if model == "unet_resnet50" else args.classes
        tr = b2u.UnetTrainer(num_classes=C, device=dev, model=model, lr=1e-4)
        imgs, pngs = b2u.synthetic.make_inputs(args.batch, C, 512, 512, seed=3)
        imgs, pngs = imgs.to(dev), pngs.to(dev)
        for _ in range(3):
            tr.train_step(imgs, pngs)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            tr.train_step(imgs, pngs)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        # inference forward (model.eval(): BatchNorm folded into the conv epilogues), logits only
        for _ in range(2):
            tr.engine.forward(imgs, tr.tensors, save=False, training=False)
        torch.cuda.synchronize()
        a.record()
        for _ in range(args.steps):
            tr.engine.forward(imgs, tr.tensors, save=False, training=False)
        b.record(); torch.cuda.synchronize()
        ms_inf = a.elapsed_time(b) / args.steps
        with _lib.CallProfile(by_shape=args.by_shape) as prof:
            tr.train_step(imgs, pngs)
        calls = prof.read()
        tot = sum(v["ms"] for v in calls.values())
        top = sorted(calls.items(), key=lambda kv: -kv[1]["ms"])[:40 if args.by_shape else 25]
        res[model] = {"classes": C, "ms_per_step": ms, "img_per_s": args.batch * 1e3 / ms,
                      "params": int(sum(p.numel() for p in tr.params.values())),
                      "eval_forward_ms": ms_inf, "eval_forward_img_per_s": args.batch * 1e3 / ms_inf,
                      "profiled_step_ms_sum": tot,
                      "by_entry_point": {k: {"launches": v["launches"], "ms": round(v["ms"], 4)} for k, v in top}}
        if args.cpu_baseline:
            import bench
            res[model]["cpu_baseline"] = bench.variant_cpu_img_per_s(model, C)
            res[model]["speedup_vs_cpu_port"] = res[model]["img_per_s"] / res[model]["cpu_baseline"]["value"]
        loss = float(tr.last[0]) if tr.last is not None else None
        res[model]["loss_finite"] = loss is None or loss == loss
        del tr
        torch.cuda.empty_cache()
    txt = json.dumps(res, indent=1)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt)


if __name__ == "__main__":
    main()
