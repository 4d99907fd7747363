"""The "library kernel to beat" of SURVEY.md 8(d): the same Unet-VGG16 training step (21 classes, 512x512, batch 16: forward,
CE + Dice, backward, Adam) written with stock torch.nn modules and run by PyTorch eager on this GPU with cuDNN
(bf16 autocast, channels_last, cudnn.benchmark) -- NOT part of the product, only a yardstick next to bench.py's number.
    python scripts/torch_eager_baseline.py [--steps 20] [--batch 16] [--fp32]"""
import argparse, json
import torch
import torch.nn as nn
import torch.nn.functional as F


class VGGUnet(nn.Module):
    def __init__(self, C):
        super().__init__()
        cfg = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]
        layers, cin = [], 3
        for v in cfg:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*layers)
        self.taps = (3, 8, 15, 22, 29)       # outputs of the last ReLU of each stage
        self.up = nn.UpsamplingBilinear2d(scale_factor=2)
        dec = [(1024, 512), (768, 256), (384, 128), (192, 64)]
        self.dec = nn.ModuleList(nn.Sequential(nn.Conv2d(i, o, 3, padding=1), nn.ReLU(inplace=True),
                                               nn.Conv2d(o, o, 3, padding=1), nn.ReLU(inplace=True)) for i, o in dec)
        self.final = nn.Conv2d(64, C, 1)

    def forward(self, x):
        feats = []
        for i, l in enumerate(self.features):
            x = l(x)
            if i in self.taps:
                feats.append(x)
        y = feats[4]
        for k, d in enumerate(self.dec):
            y = d(torch.cat([feats[3 - k], self.up(y)], 1))
        return self.final(y)


def loss_fn(logits, png, C):
    logits = logits.float()
    ce = F.cross_entropy(logits, png, ignore_index=C)
    p = torch.softmax(logits, 1)
    t = F.one_hot(png, C + 1)[..., :C].permute(0, 3, 1, 2).float()
    tp = (t * p).sum((0, 2, 3)); fp = p.sum((0, 2, 3)) - tp; fn = t.sum((0, 2, 3)) - tp
    dice = 1 - ((2 * tp + 1e-5) / (2 * tp + fn + fp + 1e-5)).mean()
    return ce + dice


def measure(steps=20, batch=16, fp32=False):
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.backends.cudnn.benchmark = True
    C = 21
    model = VGGUnet(C).to(dev).to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    x = torch.rand(batch, 3, 512, 512, device=dev).contiguous(memory_format=torch.channels_last)
    png = torch.randint(0, C + 1, (batch, 512, 512), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not fp32):
            out = model(x)
        loss = loss_fn(out, png, C)
        loss.backward()
        opt.step()
        return loss
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return {"what": "torch eager + cuDNN, Unet-VGG16 21 classes 512x512 train step", "dtype": "fp32" if fp32 else "bf16 autocast",
            "batch": batch, "ms_per_step": ms, "img_per_s": batch * 1e3 / ms, "torch": torch.__version__,
            "cudnn": torch.backends.cudnn.version(), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20); ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--fp32", action="store_true")
    args = ap.parse_args()
    print(json.dumps(measure(args.steps, args.batch, args.fp32)))


if __name__ == "__main__":
    main()
