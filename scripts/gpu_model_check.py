"""Per-tensor gradient diagnosis of the full hot path on a B200 (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_pytorch_b200 as b2u
from oracle import unet_oracle as O

dev = torch.device("cuda:0")
torch.set_num_threads(16)

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()

def grel(a, b):
    num = sum((a[k].float().cpu() - b[k]).double().pow(2).sum().item() for k in b)
    den = sum(b[k].double().pow(2).sum().item() for k in b)
    return (num / den) ** 0.5

for seed, lr in ((0, 0.0), (5, 0.0), (5, 1e-4)):
    C, n, h, w = 21, 2, 64, 64
    params = O.make_params(C, seed=11)
    imgs, pngs = O.make_inputs(n, C, h, w, seed=seed)
    l32, z32, g32 = O.train_step(params, imgs, pngs, torch.ones(C), C, dice=True)
    lbf, zbf, gbf = O.train_step_bf16_storage(params, imgs, pngs, torch.ones(C), C, dice=True)
    tr = b2u.UnetTrainer(num_classes=C, device=dev, state_dict=params, lr=lr)
    out = tr.train_step(imgs.to(dev), pngs.to(dev)).cpu()
    g1 = {k: v.clone() for k, v in tr.grads.items()}
    print(f"== seed {seed} lr {lr}: loss {out[0].item():.6f} ref {l32.item():.6f}; grads vs fp32 {grel(g1, g32):.3e}, vs bf16 model {grel(g1, gbf):.3e}; model vs fp32 {grel(gbf, g32):.3e}")
    if lr == 0.0:
        tr.train_step(imgs.to(dev), pngs.to(dev))
        print("   run-to-run diff:", grel({k: v for k, v in tr.grads.items()}, {k: v.cpu() for k, v in g1.items()}))
    for k in g32:
        e32, ebf = rel(g1[k], g32[k]), rel(g1[k], gbf[k])
        flag = " <<<" if ebf > 3e-2 else ""
        print(f"   {k:28s} vs fp32 {e32:.3e}  vs bf16-model {ebf:.3e}  |g| {g32[k].norm().item():.3e}{flag}")
