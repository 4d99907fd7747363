"""CPU: host-side mirror of the reference interface -- module tree / state_dict contract, freeze semantics,
LR schedule, flat-buffer layout and bucket planning."""
import math

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O


def test_state_dict_contract(b2u):
    m = b2u.Unet(num_classes=21, pretrained=False, backbone="vgg")
    shapes = O.param_shapes(21)
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())            # 44 tensors, reference order
    assert all(tuple(sd[k].shape) == shapes[k] and sd[k].dtype == torch.float32 for k in sd)
    assert sum(v.numel() for v in sd.values()) == 24_892_437
    m.load_state_dict(O.make_params(21))
    assert m.backbone == "vgg" and m.up_conv is None and hasattr(m, "vgg") and hasattr(m, "final")


def test_backbone_errors(b2u):
    with pytest.raises(ValueError):
        b2u.Unet(num_classes=2, backbone="mobilenet")          # nets/unet.py:34


def test_resnet50_and_traditional_state_dict_contract(b2u):
    m = b2u.Unet(num_classes=21, backbone="resnet50")
    shapes = O.resnet_unet_param_shapes(21)
    assert [n for n, _ in m.named_parameters()] == list(shapes.keys())
    assert all(tuple(p.shape) == shapes[n] for n, p in m.named_parameters())
    assert sum(p.numel() for p in m.parameters()) == 43_934_101
    assert m.up_conv is not None and hasattr(m, "resnet")
    m.load_state_dict(O.make_resnet_unet_params(21))
    m.freeze_backbone()
    assert all(not p.requires_grad for p in m.resnet.parameters()) and m.final.weight.requires_grad
    t = b2u.TraditionalUnet(in_channels=3, num_classes=21)
    assert [n for n, _ in t.named_parameters()] == list(O.trad_param_shapes(21).keys())
    assert sum(p.numel() for p in t.parameters()) == 1_950_357        # Submit_result figure quoted in BASELINE.md
    t.load_state_dict(O.make_trad_params(21))
    t.freeze_encoder()
    assert not t.inc.double_conv[0].weight.requires_grad and t.outc.weight.requires_grad


def test_freeze_unfreeze(b2u):
    m = b2u.Unet(num_classes=2)
    m.freeze_backbone()
    assert all(not p.requires_grad for p in m.vgg.parameters())
    assert all(p.requires_grad for n, p in m.named_parameters() if not n.startswith("vgg."))
    m.unfreeze_backbone()
    assert all(p.requires_grad for p in m.parameters())


def test_weights_init_matches_class_name_rule(b2u):
    from unet_pytorch_b200.nets.unet_training import weights_init
    m = b2u.Unet(num_classes=2)
    torch.manual_seed(0)
    weights_init(m)
    w = m.up_concat1.conv1.weight
    assert abs(w.std().item() - 0.02) < 2e-3 and abs(w.mean().item()) < 1e-3


def test_lr_scheduler_closed_forms(b2u):
    from unet_pytorch_b200.nets.unet_training import get_lr_scheduler
    f = get_lr_scheduler("cos", 1e-4, 1e-6, 100)
    assert f(0) == pytest.approx(1e-5)                        # warm-up start = max(0.1 * lr, 1e-6)
    assert f(3) == pytest.approx(1e-4)
    assert f(99) == pytest.approx(1e-6)
    mid = 1e-6 + 0.5 * (1e-4 - 1e-6) * (1 + math.cos(math.pi * (50 - 3) / (100 - 3 - 5)))
    assert f(50) == pytest.approx(mid)
    s = get_lr_scheduler("step", 1e-4, 1e-6, 100)
    assert s(0) == pytest.approx(1e-4) and s(99) == pytest.approx(1e-6)


def test_flat_layout_and_buckets(b2u):
    from unet_pytorch_b200.trainer import FlatBuckets, _backward_order
    shapes = b2u.vgg_unet_param_shapes(21)
    order = _backward_order(list(shapes))
    assert order[0] == "final.weight" and order[-1] == "vgg.features.0.bias"
    assert order.index("up_concat1.conv2.weight") < order.index("up_concat1.conv1.weight") < order.index("up_concat2.conv2.weight")
    lay = FlatBuckets(shapes, order, torch.device("cpu"), bucket_bytes=16 << 20)
    assert lay.total % 4 == 0 and lay.total >= 24_892_437
    covered = 0
    for s, e, names in lay.buckets:
        assert s == covered and e > s
        covered = e
    assert covered == lay.total
    flat = lay.new_buffer()
    views = lay.views(flat)
    views["final.bias"].fill_(3.0)
    o, n, _ = lay.offsets["final.bias"]
    assert flat[o:o + n].eq(3.0).all() and flat.sum().item() == 3.0 * n
    for k, v in views.items():
        assert v.data_ptr() % 16 == 0


def test_per_class_metrics_match_oracle(b2u):
    rng = np.random.default_rng(0)
    hist = rng.integers(0, 1000, size=(21, 21)).astype(np.float64)
    assert np.array_equal(b2u.per_class_iu(hist), O.per_class_iu(hist))
    assert np.array_equal(b2u.per_class_PA_Recall(hist), O.per_class_PA_Recall(hist))
    assert np.array_equal(b2u.per_class_Precision(hist), O.per_class_Precision(hist))
