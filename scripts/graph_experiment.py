"""Experiment: how much of the Unet-VGG16 step is launch gap?  Captures train_step in a CUDA graph and compares replay with
eager issue (the captured Adam step count is frozen, so this is a timing experiment only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_pytorch_b200 as b2u

dev = torch.device("cuda:0")
C = 21
tr = b2u.UnetTrainer(num_classes=C, device=dev, model="unet_vgg", lr=1e-4)
imgs, pngs = b2u.synthetic.make_inputs(16, C, 512, 512, seed=3)
imgs, pngs = imgs.to(dev), pngs.to(dev)
for _ in range(3):
    tr.train_step(imgs, pngs)
torch.cuda.synchronize()


def timeit(fn, steps=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


print("eager", round(timeit(lambda: tr.train_step(imgs, pngs)), 3))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    tr.train_step(imgs, pngs)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        out = tr.train_step(imgs, pngs)
    torch.cuda.synchronize()
    print("graph replay", round(timeit(g.replay), 3), "loss", out.tolist())
    print("eager again", round(timeit(lambda: tr.train_step(imgs, pngs)), 3))
except Exception as e:
    print("capture failed:", repr(e)[:500])
