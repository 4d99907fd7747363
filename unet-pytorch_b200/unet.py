"""Drop-in for the reference's predictor class `unet.py::Unet` (lines 19-357): same keyword configuration
(`model_path`, `num_classes`, `backbone`, `input_shape`, `mix_type`, `cuda`), same `detect_image`, `get_FPS`,
`get_miou_png` results -- with the per-frame tail moved onto the GPU.

Reference per frame: logits -> softmax -> .cpu().numpy() (H x W x C fp32, 22 MB at C = 21, 512^2) -> crop of the letterbox
bars -> cv2.resize of the probabilities to the original size -> argmax (unet.py:131-148).  Here: logits stay on the device,
one kernel (`b2u_softmax_resize_argmax_u8`) does softmax + crop + INTER_LINEAR resize + argmax, and a uint8 mask (1 byte
per pixel) is the only thing copied back.  Host-side image handling (RGB conversion, BICUBIC letterbox with grey bars,
colour blending) follows utils/utils.py:12-34 and unet.py:150-203 with PIL/numpy, as in the reference."""
import colorsys
import time

import numpy as np
import torch
from PIL import Image

from . import ops
from .nets.unet import Unet as unet

def _voc_palette(n):
    """The PASCAL VOC class palette (bit-interleaved colour code of the class index) -- the table of unet.py:59-63."""
    out = []
    for i in range(n):
        r = g = b = 0
        c = i
        for j in range(8):
            r |= ((c >> 0) & 1) << (7 - j)
            g |= ((c >> 1) & 1) << (7 - j)
            b |= ((c >> 2) & 1) << (7 - j)
            c >>= 3
        out.append((r, g, b))
    return out


def cvtColor(image):                                   # utils/utils.py:12-17
    if len(np.shape(image)) == 3 and np.shape(image)[2] == 3:
        return image
    return image.convert("RGB")


def resize_image(image, size):                         # utils/utils.py:22-34: undistorted resize onto a grey canvas
    iw, ih = image.size
    w, h = size
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    canvas = Image.new("RGB", size, (128, 128, 128))
    canvas.paste(image.resize((nw, nh), Image.BICUBIC), ((w - nw) // 2, (h - nh) // 2))
    return canvas, nw, nh


class Unet(object):
    _defaults = {
        "model_path": "model_data/unet_vgg_voc.pth",
        "num_classes": 21,
        "backbone": "vgg",
        "input_shape": [512, 512],
        "mix_type": 0,
        "cuda": True,
    }

    def __init__(self, **kwargs):
        self.__dict__.update(self._defaults)
        self.state_dict = None                         # optional: weights passed in memory instead of model_path
        for name, value in kwargs.items():
            setattr(self, name, value)
        if self.num_classes <= 21:
            self.colors = _voc_palette(22)
        else:
            hsv = [(x / self.num_classes, 1.0, 1.0) for x in range(self.num_classes)]
            self.colors = [tuple(int(v * 255) for v in colorsys.hsv_to_rgb(*t)) for t in hsv]
        self.generate()

    def generate(self, onnx=False):                    # unet.py:86-98
        if not self.cuda or not torch.cuda.is_available():
            raise RuntimeError("the B200 predictor needs CUDA (cuda=True and a visible GPU); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.net = unet(num_classes=self.num_classes, backbone=self.backbone)
        sd = self.state_dict if self.state_dict is not None else torch.load(self.model_path, map_location="cpu")
        self.net.load_state_dict(sd)
        self.net = self.net.eval().to(self.device)

    # ------------------------------------------------------------------ shared pieces
    def _letterbox(self, image):
        image = cvtColor(image)
        oh, ow = np.array(image).shape[:2]
        boxed, nw, nh = resize_image(image, (self.input_shape[1], self.input_shape[0]))
        data = np.expand_dims(np.transpose(np.array(boxed, np.float32) / 255.0, (2, 0, 1)), 0)     # preprocess_input
        return image, torch.from_numpy(data), nw, nh, oh, ow

    def _crop(self, nw, nh):
        return (int((self.input_shape[0] - nh) // 2), int((self.input_shape[1] - nw) // 2), nh, nw)

    def predict_mask(self, image):
        """uint8 class map at the original image size (the `pr` of unet.py:148 / 340)."""
        image, data, nw, nh, oh, ow = self._letterbox(image)
        with torch.no_grad():
            logits = self.net(data.to(self.device, non_blocking=True))
            mask = ops.softmax_resize_argmax_u8(logits, self._crop(nw, nh), (oh, ow))
        return image, mask[0].cpu().numpy()

    def predict_mask_device(self, image):
        """The same class map, left on the device (uint8 [H, W] CUDA tensor)."""
        _, data, nw, nh, oh, ow = self._letterbox(image)
        with torch.no_grad():
            logits = self.net(data.to(self.device, non_blocking=True))
            return ops.softmax_resize_argmax_u8(logits, self._crop(nw, nh), (oh, ow))[0]

    def get_miou(self, images, gts):
        """get_miou.py:45-65 + utils_metrics.compute_mIoU (:58-126) without the PNG round trip: every prediction stays on
        the GPU and is folded into one device-resident confusion matrix (`fast_hist`), read back once at the end.
        images: iterable of PIL images; gts: iterable of label maps (PIL 'L'/'P' images or uint8 arrays, 255 = ignore).
        Returns (hist int64 [n, n], IoUs, PA_Recall, Precision) like compute_mIoU."""
        from .utils.utils_metrics import per_class_iu, per_class_PA_Recall, per_class_Precision
        n = self.num_classes
        hist = torch.zeros(n * n + 1, dtype=torch.int64, device=self.device)
        pairs = []

        def flush():
            if pairs:
                ops.fast_hist_batch(pairs, n, hist)       # one launch for the whole group of masks
                pairs.clear()
        for image, gt in zip(images, gts):
            pred = self.predict_mask_device(image)
            g = torch.from_numpy(np.ascontiguousarray(np.array(gt, dtype=np.uint8))).to(self.device, non_blocking=True)
            if g.numel() != pred.numel():          # utils_metrics.py:88-93: mismatching pairs are skipped
                continue
            pairs.append((g.reshape(-1), pred.reshape(-1).contiguous()))
            if len(pairs) == 256:
                flush()
        flush()
        h = hist.cpu().numpy()
        if h[-1] != 0:
            raise ValueError("predictions outside [0, num_classes)")
        h = h[:-1].reshape(n, n)
        return h, per_class_iu(h), per_class_PA_Recall(h), per_class_Precision(h)

    # ------------------------------------------------------------------ reference API
    def detect_image(self, image, count=False, name_classes=None):
        """unet.py:100-203: the class map (GPU) rendered with the reference's three mix types; `count` returns the per-class
        pixel counts through `self.last_counts` (the reference prints them as a table)."""
        old_img, pr = self.predict_mask(image)
        if count:
            self.last_counts = np.bincount(pr.reshape(-1), minlength=self.num_classes)
        if self.mix_type not in (0, 1, 2):
            return old_img
        if self.mix_type == 2:                      # keep the pixels of every non-background class
            return Image.fromarray(np.where((pr != 0)[..., None], np.asarray(old_img), 0).astype(np.uint8))
        seg = Image.fromarray(np.asarray(self.colors, np.uint8)[pr])
        return Image.blend(old_img, seg, 0.7) if self.mix_type == 0 else seg

    def get_FPS(self, image, test_interval):           # unet.py:205-258: forward + per-pixel class + crop, result on the host
        _, data, nw, nh, _, _ = self._letterbox(image)
        cy, cx, ch, cw = self._crop(nw, nh)
        pinned = data.pin_memory()
        host = torch.empty((1, ch, cw), dtype=torch.uint8).pin_memory()

        def frame():
            with torch.no_grad():
                logits = self.net(pinned.to(self.device, non_blocking=True))
                host.copy_(ops.argmax_u8(logits)[:, cy:cy + ch, cx:cx + cw], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        frame()
        t1 = time.time()
        for _ in range(test_interval):
            frame()
        return (time.time() - t1) / test_interval

    def get_miou_png(self, image):                     # unet.py:298-344
        _, pr = self.predict_mask(image)
        return Image.fromarray(np.uint8(pr))
