"""Synthetic weights and batches for benchmarks and smoke runs (the reference ships no VGG/ResNet UNet checkpoint
and no dataset).  Mirrors SURVEY.md 8(d): images in [0,1] quantised to k/255, label maps that are
low-frequency class fields correlated with the image, ~2 % ignore pixels (= num_classes, as utils/dataloader.py:43
produces from the VOC border)."""
import torch
import torch.nn.functional as F

from .engine import vgg_unet_param_shapes


def make_params(num_classes, seed=11, in_channels=3, gain=0.5):
    """He-scaled deterministic weights: tensor k of the state_dict comes from a generator seeded seed*1000+k."""
    params = {}
    for k, (name, shape) in enumerate(vgg_unet_param_shapes(num_classes, in_channels).items()):
        g = torch.Generator().manual_seed(seed * 1000 + k)
        if len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            params[name] = torch.randn(shape, generator=g) * (gain * (2.0 / fan_in) ** 0.5)
        else:
            params[name] = torch.randn(shape, generator=g) * 0.05
    return params


def make_inputs(n, num_classes, h, w, seed=0, medical=False):
    g = torch.Generator().manual_seed(seed)
    if medical:
        base = F.interpolate(torch.rand(n, 1, max(h // 64, 2), max(w // 64, 2), generator=g), size=(h, w), mode="bilinear",
                             align_corners=True)
        img = (base + 0.1 * torch.rand(n, 1, h, w, generator=g)).clamp(0, 1)
        img = torch.round(img * 255) / 255
        png = (base[:, 0] > 0.5).long()
        return img.repeat(1, 3, 1, 1).contiguous(), png
    fields = F.interpolate(torch.rand(n, num_classes, max(h // 32, 2), max(w // 32, 2), generator=g), size=(h, w),
                           mode="bilinear", align_corners=True)
    png = fields.argmax(1)
    img = torch.stack([fields[:, k % num_classes] for k in range(3)], 1) * 0.7 + 0.3 * torch.rand(n, 3, h, w, generator=g)
    img = torch.round(img.clamp(0, 1) * 255) / 255
    ign = torch.rand(n, h, w, generator=g) < 0.02
    png = torch.where(ign, torch.full_like(png, num_classes), png)
    return img.contiguous(), png.contiguous()


def random_state_dict(param_shapes, buffer_shapes, seed=0):
    """Random-init state_dict for any engine (the benches' stand-in for nets/unet_training.py::weights_init, lines 60-80,
    and the modules' own _initialize_weights): He-normal dense convs, depthwise taps around 1/3, Linear 1/sqrt(fan_in),
    BatchNorm weight 1 / bias 0, biases 0, fresh running statistics."""
    g = torch.Generator().manual_seed(seed)
    bn = {n[:-len(".running_mean")] for n in buffer_shapes if n.endswith(".running_mean")}
    sd = {}
    for name, shape in param_shapes.items():
        base = name.rsplit(".", 1)[0]
        if base in bn:
            sd[name] = torch.ones(shape) if name.endswith(".weight") else torch.zeros(shape)
        elif len(shape) == 4 and shape[1] == 1 and shape[2] == 3:
            sd[name] = (1.0 / 3.0) * (1.0 + 0.5 * torch.randn(shape, generator=g))
        elif len(shape) == 4:
            sd[name] = torch.randn(shape, generator=g) * (2.0 / (shape[1] * shape[2] * shape[3])) ** 0.5
        elif len(shape) == 2:
            sd[name] = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
        else:
            sd[name] = torch.zeros(shape)
    for name, shape in buffer_shapes.items():
        if name.endswith("running_var"):
            sd[name] = torch.ones(shape)
        elif name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.long)
        else:
            sd[name] = torch.zeros(shape)
    return sd
