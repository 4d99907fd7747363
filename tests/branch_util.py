"""Test helpers: compare a CUDA run with the float64 oracle ON THE SAME PIECEWISE-LINEAR BRANCH.

The loss of the BatchNorm + ReLU + max-pool networks is only piecewise smooth: millions of pre-activations, a few within
rounding of zero, and near-ties inside pooling windows.  Two precisions disagree about a handful of such decisions, and one
flipped decision at a small layer moves every gradient below it by ~5e-3.  `oracle.branch` replays
the decisions of the CUDA run (its ReLU masks and max-pool winners, read back from the engine's saved activations) inside the
oracle, so what is compared is the arithmetic, not the coin flips -- and the number of differing decisions is reported."""
import torch
import torch.nn.functional as F

from oracle import unet_oracle as O


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def global_rel(grads, ref, keys=None):
    keys = list(ref) if keys is None else keys
    num = sum((grads[k].double().cpu() - ref[k].double()).pow(2).sum().item() for k in keys)
    den = sum(ref[k].double().pow(2).sum().item() for k in keys)
    return (num / den) ** 0.5


def _nchw(t):
    return t.float().permute(0, 3, 1, 2).cpu()


def engine_branch(eng):
    """{site: ReLU mask | max-pool arg-max indices} of the forward `eng` just ran, keyed like oracle.branch.  Channel padding
    of the engine's tensors is still present (trim with shapes from a recorded oracle run)."""
    pin = {}
    if hasattr(eng, "program"):                      # GraphEngine: ResNet50 / Lightweight / UltraLightweight
        T = eng.saved[0]
        for ins in eng.program:
            if ins["op"] == "bn" and ins.get("relu", True):
                pin[ins["bn"]] = _nchw(T[ins["out"]].data) > 0
            elif ins["op"] == "conv" and ins.get("relu"):
                pin[ins["w"]] = _nchw(T[ins["out"]].data) > 0
            elif ins["op"] == "addrelu":
                pin[ins["out"]] = _nchw(T[ins["out"]].data) > 0
            elif ins["op"] == "pool3":
                pin[ins["out"]] = F.max_pool2d(_nchw(T[ins["x"]].data), 3, 2, ceil_mode=True, return_indices=True)[1]
            elif ins["op"] == "pool2":
                pin[ins["out"]] = F.max_pool2d(_nchw(T[ins["x"]].data), 2, 2, return_indices=True)[1]
    else:                                            # UNetEngine with BatchNorm: TraditionalUnet
        acts = eng.saved[0]
        for c in eng.convs:
            pin[c.bn] = _nchw(acts[c.name]) > 0
        for bi in range(1, len(eng.enc)):
            pin[f"pool{bi}"] = F.max_pool2d(_nchw(acts[eng.enc[bi - 1][-1].name]), 2, 2, return_indices=True)[1]
    return pin


def compare_on_branch(tag, step, sd, imgs, weights, grads, eng, skip=lambda k: False):
    """step(sd, imgs, weights) -> (loss, logits, grads, stats) of the oracle.  Runs it in float64 on its own branch (site shapes,
    its own decisions), then in float64 and in fp32 on the branch `eng` took.  Returns a dict of distances:
    ours / worst (CUDA gradients vs float64), ref32 / worst32 (torch fp32 vs float64, same branch), relu_flips / pool_flips
    (decisions that differ from float64's own branch), l64 / z64 (float64 loss and logits on the branch)."""
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    own = {}
    with O.branch(record=own):
        step(sd64, imgs.double(), weights.double())
    pin = engine_branch(eng)
    assert sorted(pin) == sorted(own), sorted(set(pin) ^ set(own))
    pin = {k: v[:, :own[k].shape[1]].contiguous() for k, v in pin.items()}      # drop the engine's zero-padded channels
    with O.branch(pin=pin):
        l64, z64, g64, _ = step(sd64, imgs.double(), weights.double())
        _, _, g32p, _ = step(sd, imgs, weights)
    live = [k for k in g64 if not skip(k)]
    out = dict(
        ours=global_rel(grads, g64, live), ref32=global_rel(g32p, g64, live),
        worst=max(rel(grads[k], g64[k]) for k in live), worst32=max(rel(g32p[k], g64[k]) for k in live),
        relu_flips=sum(int((own[k] != pin[k]).sum()) for k in pin if own[k].dtype == torch.bool),
        pool_flips=sum(int((own[k] != pin[k]).sum()) for k in pin if own[k].dtype != torch.bool),
        relu_sites=sum(v.numel() for k, v in pin.items() if own[k].dtype == torch.bool), l64=l64, z64=z64, g64=g64)
    print(f"{tag}: CUDA vs float64 {out['ours']:.2e} (worst tensor {out['worst']:.2e}); torch fp32 on the same branch vs float64 "
          f"{out['ref32']:.2e} (worst {out['worst32']:.2e}); {out['relu_flips']} of {out['relu_sites']} ReLU signs and "
          f"{out['pool_flips']} max-pool winners differ from float64's own branch")
    return out
