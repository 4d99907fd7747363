"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.
Run in the build container only (the reference does not travel to the GPU box):  python oracle/make_golden.py
The synthetic weights/inputs come from oracle.unet_oracle.make_params / make_inputs so tests can regenerate them.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
# utils/utils_metrics.py imports matplotlib at module level (line 5); plotting is never reached on this path
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))

from nets.unet import Unet as RefUnet                                   # noqa: E402
from nets.TraditionalUnet import TraditionalUnet as RefTraditional      # noqa: E402
from nets.UltraLightweightUnet import UltraLightweightUnet as RefULU                           # noqa: E402
from nets.UltraLightweightUnet_large import UltraLightweightUnet_large as RefULULarge         # noqa: E402
from nets.UltraLightweightUnet_large_optimized import UltraLightweightUnet_large_optimized as RefULUOpt   # noqa: E402
from nets.LightWeightUnet import LightweightUnet as RefLightweight                            # noqa: E402
from nets.unet_training import CE_Loss, Dice_loss, Focal_Loss           # noqa: E402
from utils.utils_metrics import f_score, fast_hist, per_class_iu, per_class_PA_Recall, per_class_Precision  # noqa: E402

from oracle import unet_oracle as O                                     # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(8)


def model_case(tag, num_classes, n, h, w, seed, medical, cls_w, dice, focal):
    params = O.make_params(num_classes, seed=11)
    model = RefUnet(num_classes=num_classes, pretrained=False, backbone="vgg")
    model.load_state_dict(params)
    model.train()
    imgs, pngs = O.make_inputs(n, num_classes, h, w, seed=seed, medical=medical)
    labels = torch.eye(num_classes + 1)[pngs]
    weights = torch.tensor(cls_w, dtype=torch.float32)
    out = model(imgs)
    loss = Focal_Loss(out, pngs, weights, num_classes=num_classes) if focal else CE_Loss(out, pngs, weights, num_classes=num_classes)
    parts = {"ce_or_focal": loss.item()}
    if dice:
        d = Dice_loss(out, labels)
        parts["dice"] = d.item()
        loss = loss + d
    with torch.no_grad():
        fs = f_score(out, labels).item()
    loss.backward()
    rec = {"logits": out.detach().numpy().astype(np.float32), "loss": np.float64(loss.item()), "f_score": np.float64(fs),
           "cls_w": np.asarray(cls_w, np.float32), "meta": np.asarray([num_classes, n, h, w, seed, int(medical), int(dice), int(focal)])}
    for k, v in parts.items():
        rec["loss_" + k] = np.float64(v)
    for name, p in model.named_parameters():
        g = p.grad.detach()
        rec["gnorm:" + name] = np.float64(g.double().norm().item())
        flat = g.reshape(-1)
        # evenly spaced sample of up to 4096 entries (full tensor for biases and the head)
        if flat.numel() <= 4096:
            rec["g:" + name] = flat.numpy().astype(np.float32)
        else:
            idx = torch.linspace(0, flat.numel() - 1, 4096).long()
            rec["g:" + name] = flat[idx].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, f"unet_vgg_{tag}.npz"), **rec)
    print(tag, "loss", loss.item(), "f_score", fs, "logits", out.shape)


def traditional_case(tag, num_classes, n, h, w, seed, cls_w, dice, focal):
    sd = O.make_trad_params(num_classes, seed=11)
    model = RefTraditional(in_channels=3, num_classes=num_classes)
    model.load_state_dict(sd)
    model.train()
    imgs, pngs = O.make_inputs(n, num_classes, h, w, seed=seed)
    labels = torch.eye(num_classes + 1)[pngs]
    weights = torch.tensor(cls_w, dtype=torch.float32)
    out = model(imgs)
    loss = Focal_Loss(out, pngs, weights, num_classes=num_classes) if focal else CE_Loss(out, pngs, weights, num_classes=num_classes)
    if dice:
        loss = loss + Dice_loss(out, labels)
    loss.backward()
    rec = {"logits": out.detach().numpy().astype(np.float32), "loss": np.float64(loss.item()),
           "cls_w": np.asarray(cls_w, np.float32), "meta": np.asarray([num_classes, n, h, w, seed, int(dice), int(focal)])}
    for name, p in model.named_parameters():
        g = p.grad.detach().reshape(-1)
        rec["gnorm:" + name] = np.float64(g.double().norm().item())
        rec["g:" + name] = (g if g.numel() <= 4096 else g[torch.linspace(0, g.numel() - 1, 4096).long()]).numpy().astype(np.float32)
    for name, b in model.named_buffers():
        rec["buf:" + name] = b.detach().numpy()
    model.eval()
    with torch.no_grad():
        rec["logits_eval"] = model(imgs).numpy().astype(np.float32)       # eval mode: running statistics after one step
    np.savez_compressed(os.path.join(OUT, f"traditional_{tag}.npz"), **rec)
    print("traditional", tag, "loss", loss.item())


def resnet_case(tag, num_classes, n, h, w, seed, cls_w, dice):
    sd = O.make_resnet_unet_params(num_classes, seed=11)
    model = RefUnet(num_classes=num_classes, pretrained=False, backbone="resnet50")
    model.load_state_dict(sd)
    model.train()
    imgs, pngs = O.make_inputs(n, num_classes, h, w, seed=seed)
    labels = torch.eye(num_classes + 1)[pngs]
    weights = torch.tensor(cls_w, dtype=torch.float32)
    out = model(imgs)
    loss = CE_Loss(out, pngs, weights, num_classes=num_classes)
    if dice:
        loss = loss + Dice_loss(out, labels)
    loss.backward()
    rec = {"logits": out.detach().numpy().astype(np.float32), "loss": np.float64(loss.item()),
           "cls_w": np.asarray(cls_w, np.float32), "meta": np.asarray([num_classes, n, h, w, seed, int(dice)])}
    for name, p in model.named_parameters():
        g = p.grad.detach().reshape(-1)
        rec["gnorm:" + name] = np.float64(g.double().norm().item())
        rec["g:" + name] = (g if g.numel() <= 1024 else g[torch.linspace(0, g.numel() - 1, 1024).long()]).numpy().astype(np.float32)
    for name, b in model.named_buffers():
        if name.endswith("running_mean") and (".bn3" in name or name == "resnet.bn1.running_mean"):
            rec["buf:" + name] = b.detach().numpy()
    model.eval()
    with torch.no_grad():
        rec["logits_eval"] = model(imgs).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, f"unet_resnet50_{tag}.npz"), **rec)
    print("resnet50", tag, "loss", loss.item())


ULU_REF = {"ultralight": RefULU, "ultralight_large": RefULULarge, "ultralight_large_optimized": RefULUOpt}


def ulu_case(variant, tag, num_classes, n, h, w, seed, cls_w, dice, focal):
    """UltraLightweightUnet family in train mode.  The bridge Dropout2d draws from torch's RNG, so a forward hook records
    the multiplier the reference applied (output / input per (sample, channel)); the CUDA path replays that mask."""
    sd = O.make_ulu_params(num_classes, variant, seed=11)
    model = ULU_REF[variant](num_classes=num_classes)
    model.load_state_dict(sd)
    model.train()
    imgs, pngs = O.make_inputs(n, num_classes, h, w, seed=seed)
    labels = torch.eye(num_classes + 1)[pngs]
    weights = torch.tensor(cls_w, dtype=torch.float32)
    rec = {}
    captured = {}

    def hook(mod, inp, out):
        x = inp[0].detach()
        amax = x.abs().amax(dim=(2, 3))
        ratio = (out.detach().abs().amax(dim=(2, 3)) / amax.clamp_min(1e-30))
        keep = 1.0 - mod.p
        captured["mask"] = torch.where(amax > 0, (ratio > 0.5).float() / keep, torch.full_like(ratio, 1.0 / keep))
        captured["dead"] = int((amax == 0).sum())
    torch.manual_seed(seed)
    hnd = model.dropout.register_forward_hook(hook)
    out = model(imgs)
    hnd.remove()
    if "mask" in captured:
        rec["drop_mask"] = captured["mask"].numpy().astype(np.float32)
        # a bridge channel that is identically zero (dead ReLU) hides its dropout decision, which then cannot matter:
        # its value and its gradient are zero under either outcome
    loss = Focal_Loss(out, pngs, weights, num_classes=num_classes) if focal else CE_Loss(out, pngs, weights, num_classes=num_classes)
    if dice:
        loss = loss + Dice_loss(out, labels)
    with torch.no_grad():
        fs = f_score(out, labels).item()
    loss.backward()
    rec.update({"logits": out.detach().numpy().astype(np.float32), "loss": np.float64(loss.item()), "f_score": np.float64(fs),
                "cls_w": np.asarray(cls_w, np.float32), "meta": np.asarray([num_classes, n, h, w, seed, int(dice), int(focal)])})
    for name, p in model.named_parameters():
        g = p.grad.detach().reshape(-1)
        rec["gnorm:" + name] = np.float64(g.double().norm().item())
        rec["g:" + name] = (g if g.numel() <= 1024 else g[torch.linspace(0, g.numel() - 1, 1024).long()]).numpy().astype(np.float32)
    for name, b in model.named_buffers():
        if name.endswith("conv.4.running_mean") or name.endswith("conv.4.running_var"):
            rec["buf:" + name] = b.detach().numpy()
    model.eval()
    with torch.no_grad():
        rec["logits_eval"] = model(imgs).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, f"{variant}_{tag}.npz"), **rec)
    print(variant, tag, "loss", loss.item(), "dropmask" if "drop_mask" in rec else "")


def lightweight_case(tag, num_classes, n, h, w, seed, cls_w, dice, focal):
    """LightweightUnet in train mode: logits at H/2 x W/2 (the reference's losses resize them); the ten Dropout2d draws
    are recorded by forward hooks in call order and replayed by the CUDA path."""
    sd = O.make_lw_params(num_classes, seed=11)
    model = RefLightweight(num_classes=num_classes)
    model.load_state_dict(sd)
    model.train()
    imgs, pngs = O.make_inputs(n, num_classes, h, w, seed=seed)
    labels = torch.eye(num_classes + 1)[pngs]
    weights = torch.tensor(cls_w, dtype=torch.float32)
    masks = {}
    counters = {"backbone": 0}

    def make_hook(site_fn):
        def hook(mod, inp, out):
            x = inp[0].detach()
            amax = x.abs().amax(dim=(2, 3))
            ratio = out.detach().abs().amax(dim=(2, 3)) / amax.clamp_min(1e-30)
            keep = 1.0 - mod.p
            masks[site_fn()] = torch.where(amax > 0, (ratio > 0.5).float() / keep, torch.full_like(ratio, 1.0 / keep))
        return hook

    def backbone_site():
        counters["backbone"] += 1
        return f"feat{counters['backbone']}"
    hooks = [model.backbone.dropout.register_forward_hook(make_hook(backbone_site)),
             model.final_conv[1].register_forward_hook(make_hook(lambda: "final_conv.1"))]
    for k in (4, 3, 2, 1):
        hooks.append(getattr(model, f"up_concat{k}").dropout.register_forward_hook(make_hook(lambda k=k: f"up_concat{k}.drop")))
    torch.manual_seed(seed)
    out = model(imgs)
    for hnd in hooks:
        hnd.remove()
    assert len(masks) == 10
    loss = Focal_Loss(out, pngs, weights, num_classes=num_classes) if focal else CE_Loss(out, pngs, weights, num_classes=num_classes)
    if dice:
        loss = loss + Dice_loss(out, labels)
    with torch.no_grad():
        fs = f_score(out, labels).item()
    loss.backward()
    rec = {"logits": out.detach().numpy().astype(np.float32), "loss": np.float64(loss.item()), "f_score": np.float64(fs),
           "cls_w": np.asarray(cls_w, np.float32), "meta": np.asarray([num_classes, n, h, w, seed, int(dice), int(focal)])}
    for site, m in masks.items():
        rec["drop:" + site] = m.numpy().astype(np.float32)
    for name, p in model.named_parameters():
        g = p.grad.detach().reshape(-1)
        rec["gnorm:" + name] = np.float64(g.double().norm().item())
        rec["g:" + name] = (g if g.numel() <= 1024 else g[torch.linspace(0, g.numel() - 1, 1024).long()]).numpy().astype(np.float32)
    for name, b in model.named_buffers():
        if name.endswith("bn2.running_mean") or name.endswith("bn2.running_var"):
            rec["buf:" + name] = b.detach().numpy()
    model.eval()
    with torch.no_grad():
        rec["logits_eval"] = model(imgs).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, f"lightweight_{tag}.npz"), **rec)
    print("lightweight", tag, "loss", loss.item(), "f_score", fs)


def predictor_case():
    """The reference's predictor class (unet.py::Unet, CPU) on a seeded non-square RGB image: get_miou_png's class map at
    the original size (letterbox -> net -> softmax -> crop -> cv2.resize INTER_LINEAR -> argmax) and the top-2 probability
    margin per pixel (random-init logits are near-tied in places; the GPU test scores confident pixels)."""
    import contextlib, io, tempfile
    import cv2
    from PIL import Image
    import unet as ref_unet_module
    C = 21
    sd = O.make_predictor_params(C, seed=11)
    g = torch.Generator().manual_seed(41)
    low = torch.rand(1, 3, 10, 14, generator=g)
    img = (torch.nn.functional.interpolate(low, size=(300, 420), mode="bilinear", align_corners=False)[0] * 255).round().byte()
    img = img.permute(1, 2, 0).contiguous().numpy()                    # H x W x 3 uint8
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "w.pth")
        torch.save(sd, path)
        with contextlib.redirect_stdout(io.StringIO()):
            pred = ref_unet_module.Unet(model_path=path, num_classes=C, backbone="vgg", input_shape=[256, 256], mix_type=1, cuda=False)
            mask = np.array(pred.get_miou_png(Image.fromarray(img)))
            seg = np.array(pred.detect_image(Image.fromarray(img)))
        # the probabilities the reference arg-maxes, for the margin map (same steps as unet.py:324-336)
        boxed, nw, nh = ref_unet_module.resize_image(Image.fromarray(img), (256, 256))
        data = np.expand_dims(np.transpose(np.array(boxed, np.float32) / 255.0, (2, 0, 1)), 0)
        with torch.no_grad():
            pr = torch.softmax(pred.net(torch.from_numpy(data))[0].permute(1, 2, 0), dim=-1).numpy()
        pr = pr[(256 - nh) // 2:(256 - nh) // 2 + nh, (256 - nw) // 2:(256 - nw) // 2 + nw]
        pr = cv2.resize(pr, (420, 300), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(pr.argmax(-1), mask)
        top2 = np.sort(pr, axis=-1)[..., -2:]
    np.savez_compressed(os.path.join(OUT, "predictor_vgg_nc21.npz"), image=img, mask=mask.astype(np.uint8), seg=seg.astype(np.uint8),
                        margin=(top2[..., 1] - top2[..., 0]).astype(np.float16), meta=np.asarray([C, 256, 256]))
    print("predictor ok", mask.shape, np.bincount(mask.reshape(-1), minlength=C)[:8])


def checkpoint_case():
    """The reference's own trained checkpoint (Submit_result/model.pth = UltraLightweightUnet_large_optimized, 4 classes,
    all keys match) in eval mode on a seeded input: trained BatchNorm statistics make this a well-conditioned fixture
    (the bf16-storage model sits at 1.1e-2 of fp32).  The weights travel with the fixture (3.7 MB fp32)."""
    sd = torch.load("/root/reference/Submit_result/model.pth", map_location="cpu")
    model = RefULUOpt(num_classes=4)
    model.load_state_dict(sd)
    model.eval()
    imgs, _ = O.make_inputs(2, 4, 128, 128, seed=5)
    with torch.no_grad():
        out = model(imgs)
    rec = {"logits": out.numpy().astype(np.float32), "meta": np.asarray([4, 2, 128, 128, 5])}
    for k, v in sd.items():
        rec["sd:" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "ultralight_large_optimized_checkpoint_eval.npz"), **rec)
    print("checkpoint eval fixture ok", out.abs().max().item())


def loss_case():
    g = torch.Generator().manual_seed(7)
    rec = {}
    for C, cw in ((21, None), (4, [1, 15, 1.5, 2]), (2, [1, 1])):
        n, h, w = 2, 24, 40
        logits = torch.randn(n, C, h, w, generator=g) * 2
        png = torch.randint(0, C + 1, (n, h, w), generator=g)
        weights = torch.ones(C) if cw is None else torch.tensor(cw, dtype=torch.float32)
        labels = torch.eye(C + 1)[png]
        lg = logits.clone().requires_grad_(True)
        ce = CE_Loss(lg, png, weights, num_classes=C)
        fo = Focal_Loss(lg, png, weights, num_classes=C)
        di = Dice_loss(lg, labels)
        fs = f_score(lg, labels)
        g_ce, = torch.autograd.grad(ce, lg, retain_graph=True)
        g_fo, = torch.autograd.grad(fo, lg, retain_graph=True)
        g_di, = torch.autograd.grad(di, lg)
        rec[f"C{C}:logits"] = logits.numpy(); rec[f"C{C}:png"] = png.numpy(); rec[f"C{C}:w"] = weights.numpy()
        rec[f"C{C}:vals"] = np.asarray([ce.item(), fo.item(), di.item(), fs.item()], np.float64)
        rec[f"C{C}:g_ce"] = g_ce.numpy(); rec[f"C{C}:g_focal"] = g_fo.numpy(); rec[f"C{C}:g_dice"] = g_di.numpy()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **rec)
    print("losses ok")


def hist_case():
    rec = {}
    for n in (2, 4, 21):
        gt, pred = O.make_masks(3, n, h=64, w=96, seed=n)
        hist = np.zeros((n, n))
        for i in range(gt.shape[0]):
            hist += fast_hist(gt[i].flatten(), pred[i].flatten(), n)     # utils_metrics.py:95
        rec[f"n{n}:hist"] = hist.astype(np.int64)
        rec[f"n{n}:iou"] = per_class_iu(hist); rec[f"n{n}:recall"] = per_class_PA_Recall(hist)
        rec[f"n{n}:precision"] = per_class_Precision(hist)
        rec[f"n{n}:miou"] = np.float64(np.nanmean(per_class_iu(hist)))
    # adversarial: all ignored, single class, empty
    a = np.full(1000, 255, np.uint8); b = np.zeros(1000, np.uint8)
    rec["allignore:hist"] = fast_hist(a, b, 21).astype(np.int64)
    a = np.full(1000, 3, np.uint8); b = np.full(1000, 3, np.uint8)
    rec["single:hist"] = fast_hist(a, b, 21).astype(np.int64)
    rec["empty:hist"] = fast_hist(np.zeros(0, np.uint8), np.zeros(0, np.uint8), 4).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "fast_hist.npz"), **rec)
    print("hist ok")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    # config 1 shape: num_classes=2, batch 2, Medical_Datasets-shaped, CE only (train_medical.py: dice_loss=False)
    model_case("nc2_medical", 2, 2, 64, 64, 0, True, [1, 1], dice=False, focal=False)
    # config 2 shape: 21 classes, CE + Dice, ~2 % ignore pixels
    model_case("nc21_cedice", 21, 2, 64, 96, 1, False, [1] * 21, dice=True, focal=False)
    # TraditionalUnet_Train-style loss settings: focal + dice, weights [1,15,1.5,2]
    model_case("nc4_focaldice", 4, 1, 32, 32, 2, False, [1, 15, 1.5, 2], dice=True, focal=True)
    # TraditionalUnet_Train.py settings: focal loss, class weights [1,15,...] (lines 236, 245); and plain CE + Dice
    traditional_case("nc4_focaldice", 4, 2, 64, 64, 3, [1, 15, 1.5, 2], dice=True, focal=True)
    traditional_case("nc21_cedice", 21, 2, 32, 64, 4, [1] * 21, dice=True, focal=False)
    # BASELINE configs[2]: Unet-ResNet50, 21 classes, CE + Dice (batch >= 2 because of BatchNorm, train.py:139-140)
    resnet_case("nc21_cedice", 21, 2, 64, 64, 7, [1] * 21, dice=True)
    # UltraLightweightUnet family (UltraLightweightUnet*_Train.py): focal+dice with class weights, and CE+dice
    ulu_case("ultralight", "nc21_cedice", 21, 2, 64, 64, 8, [1] * 21, dice=True, focal=False)
    ulu_case("ultralight_large", "nc4_focaldice", 4, 2, 64, 64, 9, [1, 15, 1.5, 2], dice=True, focal=True)
    ulu_case("ultralight_large_optimized", "nc21_cedice", 21, 2, 32, 64, 10, [1] * 21, dice=True, focal=False)
    # LightweightUnet_Train.py-style settings (focal + dice with class weights) and plain CE + Dice
    lightweight_case("nc4_focaldice", 4, 2, 64, 64, 12, [1, 15, 1.5, 2], dice=True, focal=True)
    lightweight_case("nc21_cedice", 21, 2, 64, 96, 13, [1] * 21, dice=True, focal=False)
    checkpoint_case()
    predictor_case()
    loss_case()
    hist_case()
