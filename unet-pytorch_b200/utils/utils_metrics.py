"""Drop-ins for the metric functions of the reference's utils/utils_metrics.py: f_score (:12-31), fast_hist
(:34-43) and the per-class ratios (:45-55).  fast_hist and f_score run on the GPU (csrc/hist.cu,
csrc/head_loss.cu); the n-element ratios stay float64 numpy on the host exactly like the reference."""
import numpy as np
import torch

from .. import ops


def f_score(inputs, target, beta=1, smooth=1e-5, threhold=0.5):
    n, c, h, w = inputs.size()
    nt, ht, wt, ct = target.size()
    if not inputs.is_cuda:
        raise RuntimeError("f_score: CUDA tensors required (no CPU fallback)")
    x = inputs.detach()
    x = (x if x.dtype == torch.float32 else x.float()).contiguous()
    if h != ht and w != wt:       # utils/utils_metrics.py:15-16
        x = ops.resize_bilinear(x, (ht, wt))
    oh = target.to(device=x.device, dtype=torch.float32).contiguous()
    fin = ops.loss_fwd(x, target=None, onehot=oh, cls_w=None, beta=float(beta), smooth=float(smooth), thr=float(threhold))
    return fin[3]


_NP2T = {np.dtype("uint8"): torch.uint8, np.dtype("int32"): torch.int32, np.dtype("int64"): torch.int64}


def _as_device_flat(v, device):
    if isinstance(v, torch.Tensor):
        t = v.reshape(-1)
    else:
        arr = np.ascontiguousarray(np.asarray(v).reshape(-1))
        if arr.dtype not in _NP2T:
            if not np.issubdtype(arr.dtype, np.integer) and arr.dtype != np.bool_:
                raise TypeError("fast_hist: integer label arrays expected")
            arr = arr.astype(np.int64)
        t = torch.from_numpy(arr)
    if t.dtype not in (torch.uint8, torch.int32, torch.int64):
        t = t.long()
    return t.to(device, non_blocking=True).contiguous()


def fast_hist_device(a, b, n, hist=None, device=None):
    """Accumulates the n x n confusion matrix of (a = ground truth, b = prediction) into a device int64 vector of
    n*n + 1 entries (last = count of out-of-range bins).  Use this form to fold many images without a sync."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    ta, tb = _as_device_flat(a, device), _as_device_flat(b, device)
    if ta.dtype != tb.dtype:
        ta, tb = ta.long(), tb.long()
    if hist is None:
        hist = torch.zeros(n * n + 1, dtype=torch.int64, device=device)
    return ops.fast_hist_accumulate(ta, tb, n, hist)


def fast_hist(a, b, n):
    """np.bincount(n * a[k] + b[k], minlength=n**2).reshape(n, n) with k = (a >= 0) & (a < n), bit-exact.
    Returns an int64 ndarray like the reference; raises ValueError where numpy's reshape would."""
    hist = fast_hist_device(a, b, n).cpu().numpy()
    if hist[-1] != 0:
        raise ValueError("cannot reshape array of size > n**2 into shape ({0},{0}): predictions contain labels "
                         "that push n*a+b past n**2".format(n))
    return hist[:-1].reshape(n, n)


def per_class_iu(hist):
    return np.diag(hist) / np.maximum((hist.sum(1) + hist.sum(0) - np.diag(hist)), 1)


def per_class_PA_Recall(hist):
    return np.diag(hist) / np.maximum(hist.sum(1), 1)


def per_class_Precision(hist):
    return np.diag(hist) / np.maximum(hist.sum(0), 1)


def per_Accuracy(hist):
    return np.sum(np.diag(hist)) / np.maximum(np.sum(hist), 1)
