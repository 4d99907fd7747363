"""Execution engine of the VGG16-backbone UNet hot path (nets/unet.py:62-78 + nets/vgg.py:21-31 of the reference).

The engine owns the bf16 NHWC activation buffers and the packed bf16 weights and launches the CUDA kernels of
libb200unet.so in the order of the reference's forward, then the hand-derived backward:

  forward   im2col(first conv) -> 13 x [conv3x3+bias+ReLU] with 4 max-pools -> 4 decoder stages
            (upsample2x -> conv over the *virtual* concat [skip, up] -> conv) -> 1x1 head -> logits (NCHW fp32)
  backward  head bwd -> per decoder stage: wgrad/dgrad conv2, wgrad conv1, split dgrad conv1 (d skip | d up),
            upsample adjoint (+ReLU mask) -> per encoder block: wgrad, dgrad(+ReLU mask), pool bwd (+skip grad, +mask)

Gradients are written straight into caller-provided fp32 OIHW tensors (normally views of one flat buffer), and
`on_grads_ready(names)` is called after the launches that complete each layer's gradients so a data-parallel
trainer can start that bucket's all-reduce while the remaining dgrad/wgrad kernels still run.
"""
import torch

from . import ops

# (features index, Cin, Cout) and 'M' for nn.MaxPool2d(2,2): cfgs['D'] of nets/vgg.py:62-64 without the last pool
VGG16_CFG = [(0, 3, 64), (2, 64, 64), "M", (5, 64, 128), (7, 128, 128), "M", (10, 128, 256), (12, 256, 256),
             (14, 256, 256), "M", (17, 256, 512), (19, 512, 512), (21, 512, 512), "M", (24, 512, 512), (26, 512, 512),
             (28, 512, 512)]
# (module name, C_skip, C_up, C_out): in_filters [192, 384, 768, 1024], out_filters [64, 128, 256, 512] (nets/unet.py:28-45)
DECODER_CFG = [("up_concat4", 512, 512, 512), ("up_concat3", 256, 512, 256), ("up_concat2", 128, 256, 128),
               ("up_concat1", 64, 128, 64)]


def vgg_unet_param_shapes(num_classes, in_channels=3):
    """name -> shape, in the reference's state_dict order (44 tensors)."""
    shapes = {}
    for item in VGG16_CFG:
        if item == "M":
            continue
        i, cin, cout = item
        cin = in_channels if i == 0 else cin
        shapes[f"vgg.features.{i}.weight"] = (cout, cin, 3, 3)
        shapes[f"vgg.features.{i}.bias"] = (cout,)
    for name, cs, cu, co in DECODER_CFG:
        shapes[f"{name}.conv1.weight"] = (co, cs + cu, 3, 3)
        shapes[f"{name}.conv1.bias"] = (co,)
        shapes[f"{name}.conv2.weight"] = (co, co, 3, 3)
        shapes[f"{name}.conv2.bias"] = (co,)
    shapes["final.weight"] = (num_classes, 64, 1, 1)
    shapes["final.bias"] = (num_classes,)
    return shapes


class _Conv:
    __slots__ = ("name", "cin", "cout", "first", "c0", "c1", "wf", "wd", "version")

    def __init__(self, name, cin, cout, first=False, c0=None, c1=0):
        self.name, self.cin, self.cout, self.first = name, cin, cout, first
        self.c0 = cin if c0 is None else c0
        self.c1 = c1
        self.wf = self.wd = None
        self.version = None


class VGGUnetEngine:
    def __init__(self, num_classes, in_channels=3, device=None):
        if in_channels * 9 > 64:
            raise ValueError("in_channels must be <= 7 (first-layer im2col is 64 columns wide)")
        if not 1 <= num_classes <= 32:
            raise ValueError("num_classes must be in [1, 32]")
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.device = device
        self.enc = []      # list of lists (blocks) of _Conv
        block = []
        for item in VGG16_CFG:
            if item == "M":
                self.enc.append(block)
                block = []
            else:
                i, cin, cout = item
                block.append(_Conv(f"vgg.features.{i}", in_channels if i == 0 else cin, cout, first=(i == 0)))
        self.enc.append(block)
        self.dec = []
        for name, cs, cu, co in DECODER_CFG:
            self.dec.append((_Conv(f"{name}.conv1", cs + cu, co, c0=cs, c1=cu), _Conv(f"{name}.conv2", co, co)))
        self.convs = [c for b in self.enc for c in b] + [c for pair in self.dec for c in pair]
        # The wgrad kernel can produce db in the same pass (an N=16 MMA against a ones tile); measured on B200 it costs
        # more than the separate HBM-bound column-sum kernel (work units with the extra MMAs become the stragglers of
        # the static schedule: +40 % wgrad time vs +1.2 ms for bias_grad), so it is off by default.
        self.fuse_bias_grad = False
        self._pack_key = self._pack_versions = self._pack_table = None
        self._pack_total = 0
        self._bufs = {}
        self._shape = None
        self.saved = None
        self._ws = {}

    # ------------------------------------------------------------------ buffers
    def _buf(self, key, shape, dtype=torch.bfloat16):
        t = self._bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    def _workspace(self, key, nbytes):
        t = self._ws.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
            self._ws[key] = t
        return t

    def release(self):
        self._bufs.clear()
        self._ws.clear()
        self.saved = None

    # ------------------------------------------------------------------ weights
    def pack(self, params, need_dgrad=True):
        """(Re)packs fp32 OIHW weights to bf16 operands when any of them changed: one launch for the whole model,
        driven by a device-resident table of (weight, operand) pointers that is rebuilt only if a pointer moved."""
        key = tuple(params[c.name + ".weight"].data_ptr() for c in self.convs) + (need_dgrad,)
        versions = tuple(params[c.name + ".weight"]._version for c in self.convs)
        if self._pack_key == key and self._pack_versions == versions and self._pack_versions is not None:
            return
        dev = params[self.convs[0].name + ".weight"].device
        if self._pack_key != key:
            import struct
            blob, start = b"", 0
            for c in self.convs:
                w = params[c.name + ".weight"]
                if c.first:
                    if c.wf is None:
                        c.wf = torch.empty((c.cout, 64), dtype=torch.bfloat16, device=dev)
                    count = (c.cout + 31) // 32                      # work blocks of this layer
                else:
                    if c.wf is None:
                        c.wf = torch.empty((c.cout, 9 * c.cin), dtype=torch.bfloat16, device=dev)
                    if need_dgrad and c.wd is None:
                        c.wd = torch.empty((c.cin, 9 * c.cout), dtype=torch.bfloat16, device=dev)
                    count = (c.cout // 32) * (c.cin // 32)
                wd_ptr = c.wd.data_ptr() if (need_dgrad and not c.first) else 0
                blob += struct.pack("<QQQqiiii", w.data_ptr(), c.wf.data_ptr(), wd_ptr, start, c.cout, c.cin,
                                    1 if c.first else 9, 1 if c.first else 0)
                start += count
            self._pack_table = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
            self._pack_total = start
            self._pack_key = key
        ops.check(ops.lib().b2u_pack_weights_multi(self._pack_table.data_ptr(), len(self.convs), self._pack_total,
                                                   ops.stream_ptr()))
        self._pack_versions = versions

    def invalidate_packed_weights(self):
        """Call after writing parameters behind torch's back (raw-pointer optimizer kernels)."""
        self._pack_versions = None

    # ------------------------------------------------------------------ forward
    def forward(self, x, params, save=True):
        """x: NCHW fp32 CUDA [N, Cin, H, W] with H, W multiples of 16 -> logits NCHW fp32."""
        if not x.is_cuda:
            raise ValueError("VGGUnetEngine.forward: input must be a CUDA tensor (no CPU fallback)")
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        N, C, H, W = x.shape
        if C != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {C}")
        if H % 16 or W % 16:
            raise ValueError("input height and width must be multiples of 16")
        self.device = x.device
        self.pack(params, need_dgrad=save)
        A = {}
        col = self._buf("col", (N, H, W, 64))
        ops.lib().b2u_im2col_first(x.data_ptr(), col.data_ptr(), N, C, H, W, ops.stream_ptr())
        A["col"] = col
        cur = col
        h, w = H, W
        feats = []
        for bi, block in enumerate(self.enc):
            if bi > 0:
                pooled = self._buf(f"pool{bi}", (N, h // 2, w // 2, cur.shape[3]))
                ops.maxpool2x2(cur, out=pooled)
                A[f"pool{bi}"] = pooled
                cur = pooled
                h, w = h // 2, w // 2
            for c in block:
                out = self._buf(c.name, (N, h, w, c.cout))
                ops.conv_fprop(cur, c.wf, params[c.name + ".bias"], c.cout, taps=1 if c.first else 9, relu=True, out=out)
                A[c.name] = out
                cur = out
            feats.append(cur)
        low = feats[4]
        for si, (c1, c2) in enumerate(self.dec):
            skip = feats[3 - si]
            n_, hl, wl, cl = low.shape
            up = self._buf(f"up{si}", (N, 2 * hl, 2 * wl, cl))
            ops.upsample2x(low, out=up)
            A[f"up{si}"] = up
            o1 = self._buf(c1.name, (N, 2 * hl, 2 * wl, c1.cout))
            ops.conv_fprop(skip, c1.wf, params[c1.name + ".bias"], c1.cout, taps=9, relu=True, x1=up, out=o1)
            A[c1.name] = o1
            o2 = self._buf(c2.name, (N, 2 * hl, 2 * wl, c2.cout))
            ops.conv_fprop(o1, c2.wf, params[c2.name + ".bias"], c2.cout, taps=9, relu=True, out=o2)
            A[c2.name] = o2
            low = o2
        wfin = params["final.weight"]
        logits = ops.head_fwd(low, wfin.reshape(self.num_classes, 64), params["final.bias"])
        if save:
            self.saved = (A, feats, (N, H, W))
        return logits

    # ------------------------------------------------------------------ backward
    def backward(self, dlogits, params, grads, trainable=None, on_grads_ready=None):
        """dlogits: NCHW fp32, or bf16 [N,H,W,64] from ops.loss_bwd(nhwc64=True).  grads: name -> fp32 tensor to
        overwrite (missing / not in `trainable`: skipped).
        Returns nothing; the input image gets no gradient (the reference never asks for one)."""
        if self.saved is None:
            raise RuntimeError("backward() without a saved forward()")
        A, feats, (N, H, W) = self.saved
        if trainable is None:
            trainable = set(grads.keys())
        order = [c.name for c in self.convs] + ["final"]
        want = {n: ((n + ".weight") in trainable or (n + ".bias") in trainable) for n in order}
        # data gradients are needed down to the first (in execution order) trainable conv
        first_trainable = next((i for i, n in enumerate(order) if want[n]), len(order))
        need_dx = {n: i > first_trainable for i, n in enumerate(order)}   # does layer n have to produce dx?

        def ready(*names):
            if on_grads_ready is not None:
                on_grads_ready([n for n in names if n in grads])

        def wgrad(c, x0, dz, x1=None):
            if not want[c.name]:
                return
            wn, bn = c.name + ".weight", c.name + ".bias"
            want_w = wn in trainable and wn in grads
            want_b = bn in trainable and bn in grads
            fuse_b = want_w and want_b and self.fuse_bias_grad
            if want_w:
                need = ops.lib().b2u_conv_wgrad_workspace(dz.shape[0], dz.shape[1], dz.shape[2],
                                                          64 if c.first else c.cin, c.cout, 1 if c.first else 9)
                ops.conv_wgrad(x0, dz, taps=1 if c.first else 9, x1=x1, first_cin=c.cin if c.first else 0,
                               dw=grads[wn], db=grads[bn] if fuse_b else None, ws=self._workspace("wgrad", need))
            if want_b and not fuse_b:
                ops.bias_grad(dz, db=grads[bn], ws=self._workspace("bias", ops.lib().b2u_bias_grad_workspace(c.cout)))
            ready(wn, bn)

        last = A[self.dec[-1][1].name]
        wfin = params["final.weight"].reshape(self.num_classes, 64)
        fw, fb = "final.weight" in trainable and "final.weight" in grads, "final.bias" in trainable and "final.bias" in grads
        dz = self._buf("g:" + self.dec[-1][1].name, last.shape) if need_dx["final"] else None
        if dlogits.dtype == torch.bfloat16:
            # [N,H,W,64] = [hi | lo] split dlogits from loss_bwd(nhwc64=True): the head's backward runs on the tensor
            # cores as a 1x1 dgrad (+ReLU mask) and a 1x1 wgrad (+bias) over 2 x 32 padded classes
            dl = dlogits.contiguous()
            if need_dx["final"]:
                wd_head = ops.pack_head_dgrad(wfin, wd=self._buf("head:wd", (64, 64)))
                ops.conv_dgrad(dl, wd_head, 64, taps=1, mask=last, out0=dz)
            if fw or fb:
                dw64 = self._buf("head:dw", (64, 64, 1, 1), torch.float32)
                db64 = self._buf("head:db", (64,), torch.float32)
                need = ops.lib().b2u_conv_wgrad_workspace(dl.shape[0], dl.shape[1], dl.shape[2], 64, 64, 1)
                ops.conv_wgrad(last, dl, taps=1, dw=dw64, db=db64, ws=self._workspace("wgrad", need))
                C = self.num_classes     # rows [0,32) came from the hi half of dlogits, rows [32,64) from the lo half
                if fw:
                    torch.add(dw64[:C], dw64[32:32 + C], out=grads["final.weight"])
                if fb:
                    torch.add(db64[:C], db64[32:32 + C], out=grads["final.bias"])
        else:
            dl = dlogits.contiguous()
            if dl.dtype != torch.float32:
                dl = dl.float()
            ops.head_bwd(dl, last, wfin, need_dx=need_dx["final"], need_dw=fw or fb, relu_mask=True, dx=dz,
                         dw=grads["final.weight"] if fw else None, db=grads["final.bias"] if fb else None,
                         ws=self._workspace("head", ops.lib().b2u_head_bwd_workspace()))
        ready("final.weight", "final.bias")
        if dz is None:
            return

        dskips = [None] * 4     # gradient wrt feat1..feat4 coming from the decoder
        for si in range(3, -1, -1):
            c1, c2 = self.dec[si]
            skip = feats[3 - si]
            up = A[f"up{si}"]
            o1 = A[c1.name]
            low = feats[4] if si == 0 else A[self.dec[si - 1][1].name]
            # conv2
            wgrad(c2, o1, dz)
            if not need_dx[c2.name]:
                return
            dz1 = self._buf("g:" + c1.name, o1.shape)
            ops.conv_dgrad(dz, c2.wd, c2.cin, mask=o1, out0=dz1)
            # conv1 over [skip, up]
            wgrad(c1, skip, dz1, x1=up)
            if not need_dx[c1.name]:
                return
            enc_trainable = first_trainable < len([c for b in self.enc for c in b])
            dup = self._buf(f"g:up{si}", up.shape)
            if enc_trainable:
                dsk = self._buf(f"g:skip{3 - si}", skip.shape)
                ops.conv_dgrad(dz1, c1.wd, c1.c0, C1=c1.c1, out0=dsk, out1=dup)
                dskips[3 - si] = dsk
            else:
                # frozen encoder: only the up-sampled half of the concat needs a gradient
                wd_up = c1.wd[c1.c0:]
                ops.conv_dgrad(dz1, wd_up, c1.c1, out0=dup)
            if si == 0 and not enc_trainable:
                return
            dz = self._buf("g:low" + str(si), low.shape)
            ops.upsample2x_bwd(dup, ylow=low, out=dz)

        # encoder, deepest block first; dz = gradient wrt the pre-activation of the block's last conv
        for bi in range(4, -1, -1):
            block = self.enc[bi]
            for ci in range(len(block) - 1, -1, -1):
                c = block[ci]
                if ci > 0:
                    xin = A[block[ci - 1].name]
                elif bi > 0:
                    xin = A[f"pool{bi}"]
                else:
                    xin = A["col"]
                wgrad(c, xin, dz)
                if not need_dx[c.name]:
                    return
                if ci > 0:
                    nxt = self._buf("g:" + block[ci - 1].name, xin.shape)
                    ops.conv_dgrad(dz, c.wd, c.cin, mask=xin, out0=nxt)
                    dz = nxt
                else:
                    dpool = self._buf(f"g:pool{bi}", xin.shape)
                    ops.conv_dgrad(dz, c.wd, c.cin, out0=dpool)
                    y = feats[bi - 1]
                    nxt = self._buf(f"g:feat{bi - 1}", y.shape)
                    ops.maxpool2x2_bwd(dpool, y, dskip=dskips[bi - 1], relu_mask=True, out=nxt)
                    dz = nxt
