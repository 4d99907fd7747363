"""A/B on one box: VGG training step with db fused into the wgrad kernel (bias warps) vs the separate bias_grad pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_pytorch_b200 as b2u

dev = torch.device("cuda:0")
C = 21
tr = b2u.UnetTrainer(num_classes=C, device=dev, model="unet_vgg", lr=1e-4)
imgs, pngs = b2u.synthetic.make_inputs(16, C, 512, 512, seed=3)
imgs, pngs = imgs.to(dev), pngs.to(dev)


def run(flag, steps=15):
    tr.engine.fuse_bias_grad = flag
    for _ in range(3):
        tr.train_step(imgs, pngs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        tr.train_step(imgs, pngs)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


for rep in range(3):
    print("fused", round(run(True), 3), "separate", round(run(False), 3))
