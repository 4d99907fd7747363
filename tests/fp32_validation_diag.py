"""Per-tensor gradient distances of the fp32 validation build against the float64 oracle on the build's own ReLU branch
(development aid for test_fp32_validation_gpu.py; test tooling: it imports the oracle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_pytorch_b200 as b2u
from oracle import unet_oracle as O

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()

dev = torch.device("cuda:0")
b2u.ops.set_validation_fp32(True)
C, n, h, w, seed = 4, 2, 64, 64, 3
focal, dice = True, True
sd = O.make_trad_params(C, seed=11)
imgs, pngs = O.make_inputs(n, C, h, w, seed=seed)
weights = torch.tensor([1, 15, 1.5, 2], dtype=torch.float32)
l32, z32, g32, s32 = O.trad_train_step(sd, imgs, pngs, weights, C, dice=dice, focal=focal)
model = b2u.TraditionalUnet(in_channels=3, num_classes=C)
model.load_state_dict(sd)
model = model.train().to(dev)
out = model(imgs.to(dev))
loss = (b2u.Focal_Loss if focal else b2u.CE_Loss)(out, pngs.to(dev), weights.to(dev), num_classes=C)
if dice:
    loss = loss + b2u.Dice_loss(out, O.one_hot(pngs, C).to(dev))
loss.backward()
eng = model._engine_for(dev)
acts = eng.saved[0]
masks = {c.bn: (acts[c.name][..., :c.cout] > 0).permute(0, 3, 1, 2).cpu() for c in eng.convs}
pools = {}
for bi in range(1, len(eng.enc)):
    c = eng.enc[bi - 1][-1]
    pools[f"pool{bi}"] = torch.nn.functional.max_pool2d(acts[c.name][..., :c.cout].permute(0, 3, 1, 2).cpu(), 2, return_indices=True)[1]
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
pin = dict(masks); pin.update(pools)
with O.branch(pin=pin):
    l64, z64, g64, _ = O.trad_train_step(sd64, imgs.double(), pngs, weights.double(), C, dice=dice, focal=focal)
own = {}
with O.branch(record=own):
    _, _, g64o, _ = O.trad_train_step(sd64, imgs.double(), pngs, weights.double(), C, dice=dice, focal=focal)
print("logits vs f64", rel(out, z64), "flips", {k: int((own[k] != masks[k]).sum()) for k in masks if int((own[k] != masks[k]).sum())},
      "pool flips", {k: int((own[k] != pools[k]).sum()) for k in pools})
# torch on the GPU in fp32 (cuDNN, TF32 off) as a third opinion
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
sdg = {k: v.to(dev) for k, v in sd.items()}
lg, zg, gg, _ = O.trad_train_step(sdg, imgs.to(dev), pngs.to(dev), weights.to(dev), C, dice=dice, focal=focal)
for k, p in model.named_parameters():
    print(f"   {k:44s} build {rel(p.grad, g64[k]):.2e}   torch-cpu-f32 {rel(g32[k], g64o[k]):.2e}   torch-gpu-f32 {rel(gg[k], g64o[k]):.2e}")
