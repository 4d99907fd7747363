"""fast_hist over 1000 distinct 512x512 uint8 mask pairs (BASELINE configs[4]): one call on the contiguous stack, the pointer
table (one launch), the per-mask launch loop of the reference's compute_mIoU, and the fused argmax+hist on logits."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import unet_pytorch_b200 as b2u
from unet_pytorch_b200 import ops
import bench

dev = torch.device("cuda:0")
peaks = bench._peaks()
res = {}
for n in (21, 4, 2):
    g = torch.Generator(device=dev).manual_seed(n)
    M = 1000
    gt = torch.randint(0, n, (M, 512, 512), dtype=torch.uint8, device=dev, generator=g)
    pred = torch.where(torch.rand((M, 512, 512), device=dev, generator=g) < 0.2,
                       torch.randint(0, n, (M, 512, 512), dtype=torch.uint8, device=dev, generator=g), gt)
    gt = torch.where(torch.rand((M, 512, 512), device=dev, generator=g) < 0.03, torch.full_like(gt, 255), gt)
    hist = torch.zeros(n * n + 1, dtype=torch.int64, device=dev)
    ms_one = bench._ev_ms(lambda: ops.fast_hist_accumulate(gt.reshape(-1), pred.reshape(-1), n, hist), 10, 3)
    pairs = [(gt[i].reshape(-1), pred[i].reshape(-1)) for i in range(M)]
    table = ops.HistTable(pairs, dev)
    ms_tab = bench._ev_ms(lambda: ops.fast_hist_batch(table, n, hist), 10, 3)
    ms_loop = bench._ev_ms(lambda: [ops.fast_hist_accumulate(a, b, n, hist) for a, b in pairs], 2, 1)
    hist.zero_(); ops.fast_hist_batch(table, n, hist)
    got = hist.cpu().numpy()
    keep = gt < n
    want = torch.bincount(gt[keep].long() * n + pred[keep].long(), minlength=n * n).cpu().numpy()
    gb = 2 * gt.numel() / 1e9
    res[f"n{n}"] = {"bit_exact": bool(got[-1] == 0 and np.array_equal(got[:-1], want)),
                    "one_call": {"ms": ms_one, "GBps": gb / ms_one * 1e3, "frac_hbm": gb / ms_one * 1e3 / peaks["hbm"]},
                    "pointer_table_one_launch": {"ms": ms_tab, "GBps": gb / ms_tab * 1e3, "frac_hbm": gb / ms_tab * 1e3 / peaks["hbm"]},
                    "per_mask_launch_loop": {"ms": ms_loop, "GBps": gb / ms_loop * 1e3, "masks_per_s": M / ms_loop * 1e3}}
    del gt, pred, pairs, table, keep
    torch.cuda.empty_cache()
C = 21
logits = torch.randn(16, C, 512, 512, device=dev)
gt = torch.randint(0, C, (16, 512, 512), dtype=torch.uint8, device=dev)
hist = torch.zeros(C * C + 1, dtype=torch.int64, device=dev)
pred = torch.empty((16, 512, 512), dtype=torch.uint8, device=dev)
ms = bench._ev_ms(lambda: ops.argmax_hist(logits, gt, C, hist=hist, pred=pred), 10, 3)
ms2 = bench._ev_ms(lambda: (ops.argmax_u8(logits, out=pred), ops.fast_hist_accumulate(gt.reshape(-1), pred.reshape(-1), C, hist)), 10, 3)
gb = (logits.numel() * 4 + 2 * gt.numel()) / 1e9
res["argmax_hist_b16_c21"] = {"fused_ms": ms, "fused_GBps": gb / ms * 1e3, "frac_hbm": gb / ms * 1e3 / peaks["hbm"], "two_kernels_ms": ms2}
print(json.dumps(res, indent=1))
