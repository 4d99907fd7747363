"""BASELINE configs[0] and configs[4] side measurements (not the headline bench line):
   * inference FPS of Unet-VGG16 (21 classes, 512x512) at batch 1 and 64 in the shape of predict.py's fps mode
     (unet.py:240-257: forward + per-pixel class decision + result on the host), and
   * get_miou's fast_hist accumulated over 1000 synthetic 512x512 mask pairs (utils_metrics.py:74-95), n = 21, 4, 2,
     with the numpy reference timed beside it on a bounded sample.
   Writes one JSON document to stdout / --out."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import unet_pytorch_b200 as b2u
from unet_pytorch_b200 import ops


def ev_time(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    res = {"device": torch.cuda.get_device_name(0)}
    # ---------------- inference (fps mode)
    C = 21
    model = b2u.Unet(num_classes=C)
    model.load_state_dict(b2u.synthetic.make_params(C))
    model = model.to(dev).eval()
    for B in (1, 64):
        imgs, _ = b2u.synthetic.make_inputs(B, C, 512, 512, seed=B)
        himgs = imgs.pin_memory()
        out_host = torch.empty((B, 512, 512), dtype=torch.uint8).pin_memory()

        def frame():
            with torch.no_grad():
                x = himgs.to(dev, non_blocking=True)            # H2D of the pre-processed frame(s)
                pr = model(x)                                   # forward
                mask = ops.argmax_u8(pr)                        # argmax(softmax(z)) == argmax(z), on the device
                out_host.copy_(mask, non_blocking=True)         # 1 byte/pixel back to the host
            torch.cuda.current_stream().synchronize()
        ms = ev_time(frame, 30 if B == 1 else 5)
        res[f"infer_b{B}"] = {"ms_per_call": ms, "fps": B * 1e3 / ms,
                              "loop": "H2D fp32 frame(s) + forward + device argmax + D2H uint8 mask, one sync per call"}
    # ---------------- fast_hist over 1k masks
    def make_masks(n_masks, n, seed):          # 3 % ignore (255), 20 % of predictions re-drawn
        rng = np.random.default_rng(seed)
        gt = rng.integers(0, n, size=(n_masks, 512, 512), dtype=np.uint8)
        pred = gt.copy()
        redraw = rng.random((n_masks, 512, 512)) < 0.2
        pred[redraw] = rng.integers(0, n, size=int(redraw.sum()), dtype=np.uint8)
        gt[rng.random((n_masks, 512, 512)) < 0.03] = 255
        return gt, pred

    def np_fast_hist(a, b, n):                 # the reference's own numpy formulation (utils_metrics.py:34-43)
        k = (a >= 0) & (a < n)
        return np.bincount(n * a[k].astype(int) + b[k], minlength=n ** 2).reshape(n, n)

    for n in (21, 4, 2):
        gt, pred = make_masks(100, n, seed=n)                            # 100 distinct masks, cycled 10x = 1000 pairs
        dg, dp = torch.from_numpy(gt).to(dev), torch.from_numpy(pred).to(dev)
        hist = torch.zeros(n * n + 1, dtype=torch.int64, device=dev)

        def run_1k():
            hist.zero_()
            for rep in range(10):
                for i in range(100):
                    ops.fast_hist_accumulate(dg[i].reshape(-1), dp[i].reshape(-1), n, hist)
        ms = ev_time(run_1k, 3, warm=1)
        def run_1k_batched():
            hist.zero_()
            for rep in range(10):
                ops.fast_hist_accumulate(dg.reshape(-1), dp.reshape(-1), n, hist)
        msb = ev_time(run_1k_batched, 5, warm=1)
        got = hist[:-1].cpu().numpy().reshape(n, n)
        t0 = time.perf_counter()
        ref = np.zeros((n, n))
        for i in range(100):
            ref += np_fast_hist(gt[i].flatten(), pred[i].flatten(), n)
        cpu_s = time.perf_counter() - t0
        assert np.array_equal(got, (ref * 10).astype(np.int64)), "fast_hist mismatch"
        px = 1000 * 512 * 512
        res[f"fast_hist_n{n}"] = {
            "per_mask_launch": {"ms_per_1k_masks": ms, "masks_per_s": 1e6 / ms, "GBps": 2 * px / ms / 1e6},
            "batched_100_masks_per_launch": {"ms_per_1k_masks": msb, "masks_per_s": 1e6 / msb, "GBps": 2 * px / msb / 1e6,
                                             "frac_of_hbm_peak": 2 * px / msb / 1e6 / peaks["hbm_gbs"]},
            "numpy_1_thread": {"masks_per_s": 100 / cpu_s, "sample": "100 masks"},
            "bit_exact": True, "mIoU": float(np.nanmean(b2u.per_class_iu(got.astype(np.float64))))}
    txt = json.dumps(res, indent=1)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt)


if __name__ == "__main__":
    main()
