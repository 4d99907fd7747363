"""CPU: pins the oracle (oracle/unet_oracle.py, oracle/fast_hist.c) against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py -> tests/golden/)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sample(flat):
    if flat.numel() <= 4096:
        return flat
    return flat[torch.linspace(0, flat.numel() - 1, 4096).long()]


@pytest.mark.parametrize("tag", ["nc2_medical", "nc21_cedice", "nc4_focaldice"])
def test_model_matches_reference_golden(tag, golden_dir):
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    g = np.load(os.path.join(golden_dir, f"unet_vgg_{tag}.npz"))
    C, n, h, w, seed, medical, dice, focal = [int(v) for v in g["meta"]]
    params = O.make_params(C, seed=11)
    imgs, pngs = O.make_inputs(n, C, h, w, seed=seed, medical=bool(medical))
    loss, logits, grads = O.train_step(params, imgs, pngs, torch.from_numpy(g["cls_w"]), C, dice=bool(dice), focal=bool(focal))
    ref = torch.from_numpy(g["logits"])
    assert ((logits - ref).norm() / ref.norm()).item() <= 1e-5
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    fs = O.f_score(logits, O.one_hot(pngs, C)).item()
    assert abs(fs - float(g["f_score"])) <= 1e-5
    for name, gr in grads.items():
        gn = float(g["gnorm:" + name])
        assert abs(gr.double().norm().item() - gn) <= 2e-4 * gn + 1e-12, name
        s = _sample(gr.reshape(-1))
        ref_s = torch.from_numpy(g["g:" + name])
        assert (s - ref_s).norm().item() <= 2e-4 * ref_s.norm().item() + 1e-9, name


@pytest.mark.parametrize("C", [21, 4, 2])
def test_losses_match_reference_golden(C, golden_dir):
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    logits = torch.from_numpy(g[f"C{C}:logits"]).requires_grad_(True)
    png = torch.from_numpy(g[f"C{C}:png"])
    w = torch.from_numpy(g[f"C{C}:w"])
    oh = O.one_hot(png, C)
    ce = O.ce_loss(logits, png, w, C)
    fo = O.focal_loss(logits, png, w, C)
    di = O.dice_loss(logits, oh)
    fs = O.f_score(logits, oh)
    vals = g[f"C{C}:vals"]
    for got, want in zip((ce, fo, di, fs), vals):
        assert abs(got.item() - want) <= 1e-5 * max(abs(want), 1e-3)
    for loss, key in ((ce, "g_ce"), (fo, "g_focal"), (di, "g_dice")):
        gr, = torch.autograd.grad(loss, logits, retain_graph=True)
        ref = torch.from_numpy(g[f"C{C}:{key}"])
        assert ((gr - ref).norm() / ref.norm()).item() <= 1e-5


def _c_oracle():
    so = os.path.join(ROOT, "oracle", "_ref", "libfasthist.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.oracle_fast_hist_u8.restype = ctypes.c_longlong
    lib.oracle_fast_hist_u8.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("n", [2, 4, 21])
def test_fast_hist_matches_reference_golden(n, golden_dir):
    g = np.load(os.path.join(golden_dir, "fast_hist.npz"))
    gt, pred = O.make_masks(3, n, h=64, w=96, seed=n)
    hist = np.zeros((n, n), np.int64)
    chist = np.zeros(n * n, np.int64)
    lib = _c_oracle()
    for i in range(3):
        hist += O.fast_hist(gt[i].flatten(), pred[i].flatten(), n)
        a = np.ascontiguousarray(gt[i].flatten()); b = np.ascontiguousarray(pred[i].flatten())
        assert lib.oracle_fast_hist_u8(a.ctypes.data, b.ctypes.data, a.size, n, chist.ctypes.data) == 0
    assert np.array_equal(hist, g[f"n{n}:hist"])
    assert np.array_equal(chist.reshape(n, n), g[f"n{n}:hist"])
    assert np.array_equal(O.per_class_iu(hist), g[f"n{n}:iou"])            # float64 of exact ints: bit-exact
    assert np.array_equal(O.per_class_PA_Recall(hist), g[f"n{n}:recall"])
    assert np.array_equal(O.per_class_Precision(hist), g[f"n{n}:precision"])
    assert np.nanmean(O.per_class_iu(hist)) == float(g[f"n{n}:miou"])


def test_fast_hist_edge_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "fast_hist.npz"))
    assert np.array_equal(O.fast_hist(np.full(1000, 255, np.uint8), np.zeros(1000, np.uint8), 21), g["allignore:hist"])
    assert np.array_equal(O.fast_hist(np.full(1000, 3, np.uint8), np.full(1000, 3, np.uint8), 21), g["single:hist"])
    assert np.array_equal(O.fast_hist(np.zeros(0, np.uint8), np.zeros(0, np.uint8), 4), g["empty:hist"])
    with pytest.raises(ValueError):       # b >= n pushes a bin past n*n: numpy's reshape raises, like the reference
        O.fast_hist(np.array([1], np.uint8), np.array([200], np.uint8), 2)


@pytest.mark.parametrize("tag", ["nc4_focaldice", "nc21_cedice"])
def test_traditional_unet_matches_reference_golden(tag, golden_dir):
    """TraditionalUnet (conv + BatchNorm + ReLU) restatement against the reference's own module, train and eval mode."""
    g = np.load(os.path.join(golden_dir, f"traditional_{tag}.npz"))
    C, n, h, w, seed, dice, focal = [int(v) for v in g["meta"]]
    sd = O.make_trad_params(C, seed=11)
    imgs, pngs = O.make_inputs(n, C, h, w, seed=seed)
    loss, logits, grads, stats = O.trad_train_step(sd, imgs, pngs, torch.from_numpy(g["cls_w"]), C, dice=bool(dice), focal=bool(focal))
    ref = torch.from_numpy(g["logits"])
    assert ((logits - ref).norm() / ref.norm()).item() <= 1e-5
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for name, gr in grads.items():
        gn = float(g["gnorm:" + name])
        if name.endswith(".double_conv.0.bias") or name.endswith(".double_conv.3.bias"):
            # a conv bias in front of a train-mode BatchNorm has an exactly-zero gradient in exact arithmetic: both
            # sides only hold rounding noise
            assert gr.norm().item() <= 1e-5 and gn <= 1e-5, name
            continue
        assert abs(gr.double().norm().item() - gn) <= 2e-4 * gn + 1e-12, name
        s = _sample(gr.reshape(-1))
        ref_s = torch.from_numpy(g["g:" + name])
        assert (s - ref_s).norm().item() <= 3e-4 * ref_s.norm().item() + 2e-6, name
    for name, b in stats.items():                       # running_mean / running_var / num_batches_tracked after one step
        want = torch.from_numpy(np.asarray(g["buf:" + name]))
        assert torch.allclose(b.double(), want.double(), rtol=1e-5, atol=1e-6), name
    sd_after = dict(sd)
    sd_after.update(stats)
    with torch.no_grad():
        ev, _ = O.trad_forward(sd_after, imgs, training=False)
    ref_ev = torch.from_numpy(g["logits_eval"])
    assert ((ev - ref_ev).norm() / ref_ev.norm()).item() <= 1e-5


def test_resnet50_unet_matches_reference_golden(golden_dir):
    """Unet(backbone='resnet50') restatement (bottlenecks, BatchNorm, ceil-mode pool, up_conv) against the reference."""
    g = np.load(os.path.join(golden_dir, "unet_resnet50_nc21_cedice.npz"))
    C, n, h, w, seed, dice = [int(v) for v in g["meta"]]
    sd = O.make_resnet_unet_params(C, seed=11)
    assert sum(v.numel() for k, v in O.resnet_unet_param_shapes(21).items() for v in [torch.empty(v)]) == 43_934_101
    imgs, pngs = O.make_inputs(n, C, h, w, seed=seed)
    loss, logits, grads, stats = O.resnet_unet_train_step(sd, imgs, pngs, torch.from_numpy(g["cls_w"]), C, dice=bool(dice))
    ref = torch.from_numpy(g["logits"])
    assert ((logits - ref).norm() / ref.norm()).item() <= 1e-5
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for name, gr in grads.items():
        gn = float(g["gnorm:" + name])
        assert abs(gr.double().norm().item() - gn) <= 5e-4 * gn + 1e-9, name
    for key in g.files:
        if key.startswith("buf:"):
            assert torch.allclose(stats[key[4:]], torch.from_numpy(g[key]), rtol=1e-5, atol=1e-6), key
    sd_after = dict(sd); sd_after.update(stats)
    with torch.no_grad():
        ev, _ = O.resnet_unet_forward(sd_after, imgs, training=False)
    ref_ev = torch.from_numpy(g["logits_eval"])
    assert ((ev - ref_ev).norm() / ref_ev.norm()).item() <= 1e-5


ULU_FIXTURES = [("ultralight", "nc21_cedice"), ("ultralight_large", "nc4_focaldice"), ("ultralight_large_optimized", "nc21_cedice")]


@pytest.mark.parametrize("variant,tag", ULU_FIXTURES)
def test_ultralight_unet_matches_reference_golden(golden_dir, variant, tag):
    """UltraLightweightUnet / _large / _large_optimized restatement (depthwise-separable blocks, SE, Dropout2d replayed from
    the mask the reference drew) against the reference's own forward/backward."""
    g = np.load(os.path.join(golden_dir, f"{variant}_{tag}.npz"))
    C, n, h, w, seed, dice, focal = [int(v) for v in g["meta"]]
    sd = O.make_ulu_params(C, variant, seed=11)
    imgs, pngs = O.make_inputs(n, C, h, w, seed=seed)
    mask = torch.from_numpy(g["drop_mask"]) if "drop_mask" in g.files else None
    loss, logits, grads, stats = O.ulu_train_step(sd, imgs, pngs, torch.from_numpy(g["cls_w"]), C, variant, dice=bool(dice),
                                                  focal=bool(focal), drop_mask=mask)
    ref = torch.from_numpy(g["logits"])
    assert ((logits - ref).norm() / ref.norm()).item() <= 1e-5
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for name, gr in grads.items():
        gn = float(g["gnorm:" + name])
        if ".conv.0.bias" in name or "pointwise.bias" in name or "depthwise.bias" in name:
            # a constant per channel ahead of BatchNorm (directly, or through the linear pointwise conv): true gradient is zero
            assert gr.abs().max().item() <= 1e-4 and gn <= 1e-3, name      # fp32 cancellation residue
            continue
        assert abs(gr.double().norm().item() - gn) <= 5e-4 * gn + 1e-9, name
        flat = gr.reshape(-1)
        samp = flat if flat.numel() <= 1024 else flat[torch.linspace(0, flat.numel() - 1, 1024).long()]
        r = torch.from_numpy(g["g:" + name])
        assert ((samp - r).norm() / r.norm().clamp_min(1e-12)).item() <= 2e-3, name
    for key in g.files:
        if key.startswith("buf:"):
            assert torch.allclose(stats[key[4:]], torch.from_numpy(g[key]), rtol=1e-5, atol=1e-6), key
    sd_after = dict(sd); sd_after.update(stats)
    with torch.no_grad():
        ev, _ = O.ulu_forward(sd_after, imgs, variant, training=False)
    ref_ev = torch.from_numpy(g["logits_eval"])
    assert ((ev - ref_ev).norm() / ref_ev.norm()).item() <= 1e-5


def _load_checkpoint_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "ultralight_large_optimized_checkpoint_eval.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    C, n, h, w, seed = [int(v) for v in g["meta"]]
    imgs, _ = O.make_inputs(n, C, h, w, seed=seed)
    return sd, imgs, torch.from_numpy(g["logits"]), C


def test_ultralight_trained_checkpoint_eval(golden_dir):
    """The reference's shipped checkpoint (Submit_result/model.pth) through the restatement in eval mode."""
    sd, imgs, ref, C = _load_checkpoint_fixture(golden_dir)
    assert list(sd.keys()) == list(O.make_ulu_params(C, "ultralight_large_optimized").keys())
    with torch.no_grad():
        out, _ = O.ulu_forward(sd, imgs, "ultralight_large_optimized", training=False)
    assert ((out - ref).norm() / ref.norm()).item() <= 1e-5


def _lw_masks(g):
    return {k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("drop:")}


@pytest.mark.parametrize("tag", ["nc4_focaldice", "nc21_cedice"])
def test_lightweight_unet_matches_reference_golden(golden_dir, tag):
    """LightweightUnet restatement (ConvBlock, SE ResidualBlock, ten Dropout2d sites replayed from the reference's own draws,
    half-resolution logits resized inside the losses) against the reference's forward/backward."""
    g = np.load(os.path.join(golden_dir, f"lightweight_{tag}.npz"))
    C, n, h, w, seed, dice, focal = [int(v) for v in g["meta"]]
    sd = O.make_lw_params(C, seed=11)
    imgs, pngs = O.make_inputs(n, C, h, w, seed=seed)
    loss, logits, grads, stats = O.lw_train_step(sd, imgs, pngs, torch.from_numpy(g["cls_w"]), C, dice=bool(dice), focal=bool(focal),
                                                 drop_masks=_lw_masks(g))
    ref = torch.from_numpy(g["logits"])
    assert tuple(logits.shape) == (n, C, h // 2, w // 2)
    assert ((logits - ref).norm() / ref.norm()).item() <= 1e-5
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    full = O.resize_logits(logits, h, w)
    assert abs(O.f_score(full, O.one_hot(pngs, C)).item() - float(g["f_score"])) <= 1e-5
    for name, gr in grads.items():
        gn = float(g["gnorm:" + name])
        base = name.rsplit(".", 1)[0]
        if name.endswith(".bias") and (base.endswith(".conv.0") or base.endswith(".conv1") or base.endswith(".conv2")):
            assert gr.abs().max().item() <= 1e-4 and gn <= 1e-3, name      # conv bias ahead of BatchNorm: zero gradient
            continue
        assert abs(gr.double().norm().item() - gn) <= 5e-4 * gn + 1e-9, name
        flat = gr.reshape(-1)
        samp = flat if flat.numel() <= 1024 else flat[torch.linspace(0, flat.numel() - 1, 1024).long()]
        r = torch.from_numpy(g["g:" + name])
        assert ((samp - r).norm() / r.norm().clamp_min(1e-12)).item() <= 2e-3, name
    for key in g.files:
        if key.startswith("buf:"):
            assert torch.allclose(stats[key[4:]], torch.from_numpy(g[key]), rtol=1e-5, atol=1e-6), key
    sd_after = dict(sd); sd_after.update(stats)
    with torch.no_grad():
        ev, _ = O.lw_forward(sd_after, imgs, training=False)
    ref_ev = torch.from_numpy(g["logits_eval"])
    assert ((ev - ref_ev).norm() / ref_ev.norm()).item() <= 1e-5


def test_predictor_tail_matches_reference(golden_dir):
    """unet.py::Unet.get_miou_png (letterbox, forward, softmax, crop, cv2 INTER_LINEAR resize, argmax) restated without cv2."""
    g = np.load(os.path.join(golden_dir, "predictor_vgg_nc21.npz"))
    C, ih, iw = [int(v) for v in g["meta"]]
    mask, pr = O.predictor_mask(O.make_predictor_params(C, seed=11), g["image"], (ih, iw))
    assert mask.shape == g["mask"].shape
    agree = (mask == g["mask"])
    confident = g["margin"].astype(np.float32) > 1e-3
    assert agree[confident].all() and agree.mean() >= 0.9999
    top2 = np.sort(pr, axis=-1)[..., -2:]
    assert np.abs((top2[..., 1] - top2[..., 0]) - g["margin"].astype(np.float32)).max() <= 2e-3      # fp16-stored margins


def test_branch_record_and_replay():
    """oracle.branch (test aid of the fp32 validation tests): replaying a run's own recorded ReLU masks / max-pool winners
    reproduces that run exactly, and on one pinned branch torch's fp32 agrees with float64 far better than two free runs do
    whenever they disagree about a decision."""
    C = 4
    sd = O.make_trad_params(C, seed=11)
    imgs, pngs = O.make_inputs(1, C, 32, 32, seed=3)
    w = torch.ones(C)
    rec = {}
    with O.branch(record=rec):
        l0, z0, g0, _ = O.trad_train_step(sd, imgs, pngs, w, C, dice=True)
    assert len(rec) == 14 + 3 and all(v.dtype in (torch.bool, torch.int64) for v in rec.values())
    with O.branch(pin=rec):
        l1, z1, g1, _ = O.trad_train_step(sd, imgs, pngs, w, C, dice=True)
    assert torch.equal(z0, z1) and all(torch.equal(g0[k], g1[k]) for k in g0)
    # outside the context nothing is pinned or recorded
    assert O._BRANCH == {"pin": None, "record": None}
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    with O.branch(pin=rec):
        _, z64, g64, _ = O.trad_train_step(sd64, imgs.double(), pngs, w.double(), C, dice=True)
    num = sum((g0[k].double() - g64[k]).pow(2).sum().item() for k in g0 if g64[k].abs().max() > 1e-6)
    den = sum(g64[k].pow(2).sum().item() for k in g0 if g64[k].abs().max() > 1e-6)
    assert ((z0.double() - z64).norm() / z64.norm()).item() <= 1e-5
    assert (num / den) ** 0.5 <= 2e-3          # torch's fp32 BatchNorm backward: ~1e-4 of float64


@pytest.mark.parametrize("family", ["resnet50", "ultralight_large", "lightweight"])
def test_branch_sites_of_the_graph_families(family):
    """Every ReLU / max-pool site of the other BatchNorm families is keyed for oracle.branch, and replaying a run's own
    decisions is the identity (the GPU tests replay the CUDA engine's decisions under the same keys)."""
    C = 4
    if family == "resnet50":
        sd = O.make_resnet_unet_params(C, seed=11)
        step = lambda: O.resnet_unet_train_step(sd, imgs, pngs, torch.ones(C), C, dice=True)
        n_relu, n_pool = 1 + 16 * 3 + 8 + 2, 1          # stem, 16 bottlenecks x 3, 4 decoder stages x 2, up_conv x 2; stem pool
    elif family == "ultralight_large":
        sd = O.make_ulu_params(C, family, seed=11)
        step = lambda: O.ulu_train_step(sd, imgs, pngs, torch.ones(C), C, family, dice=True)
        n_relu, n_pool = 9 * 2, 4                        # 9 LightConvBlocks x 2 BatchNorm+ReLU; 4 pools
    else:
        sd = O.make_lw_params(C, seed=11)
        step = lambda: O.lw_train_step(sd, imgs, pngs, torch.ones(C), C, dice=True)
        n_relu, n_pool = 10 * 3, 5                       # 10 x (ConvBlock ReLU + ResidualBlock bn1 ReLU + join ReLU); 5 pools
    imgs, pngs = O.make_inputs(1, C, 64, 64, seed=2)
    rec = {}
    with O.branch(record=rec):
        l0, z0, g0, _ = step()
    assert sum(v.dtype == torch.bool for v in rec.values()) == n_relu
    assert sum(v.dtype == torch.int64 for v in rec.values()) == n_pool
    with O.branch(pin=rec):
        l1, z1, g1, _ = step()
    assert torch.allclose(z0, z1, rtol=0, atol=0) and all(torch.equal(g0[k], g1[k]) for k in g0)
