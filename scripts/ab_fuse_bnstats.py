"""A/B on one box: BatchNorm statistics from the conv epilogue vs BatchNorm's own statistics pass, per model family."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_pytorch_b200 as b2u

dev = torch.device("cuda:0")
for model, C in (("traditional", 2), ("unet_resnet50", 21), ("ultralight_large", 2), ("lightweight", 2)):
    tr = b2u.UnetTrainer(num_classes=C, device=dev, model=model, lr=1e-4)
    imgs, pngs = b2u.synthetic.make_inputs(16, C, 512, 512, seed=3)
    imgs, pngs = imgs.to(dev), pngs.to(dev)

    def run(min_k, min_cout, steps=10):
        tr.engine.fuse_bn_stats = True
        tr.engine.bn_stats_min_k, tr.engine.bn_stats_min_cout = min_k, min_cout
        for _ in range(3):
            tr.train_step(imgs, pngs)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            tr.train_step(imgs, pngs)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / steps
    BIG = 1 << 30
    for rep in range(2):
        print(model, "all", round(run(0, 0), 3), "k1024|c256", round(run(1024, 256), 3), "k2304|c256", round(run(2304, 256), 3),
              "k1024|c128", round(run(1024, 128), 3), "k1024", round(run(1024, BIG), 3), "none", round(run(BIG, BIG), 3), flush=True)
    del tr
    torch.cuda.empty_cache()
