"""Execution engine of the UNet hot path: one configurable encoder/decoder schedule that runs both the
VGG16-backbone Unet (nets/unet.py:62-78 + nets/vgg.py:21-31 of the reference: conv+bias+ReLU) and the
conv+BatchNorm+ReLU family (nets/TraditionalUnet.py:5-93).

The engine owns the bf16 NHWC activation buffers and the packed bf16 weights and launches the CUDA kernels of
libb200unet.so in the order of the reference's forward, then the hand-derived backward:

  forward   im2col(first conv) -> encoder blocks of [conv3x3 (+bias) (+BN) +ReLU] separated by 2x2 max-pools ->
            decoder stages (upsample2x -> conv over the *virtual* concat [skip, up] -> conv) -> 1x1 head -> logits
  backward  head bwd -> per decoder stage: (BN bwd,) wgrad/dgrad conv2, wgrad conv1, split dgrad conv1 (d skip | d up),
            upsample adjoint -> per encoder block: wgrad, dgrad, pool bwd (+skip grad); ReLU masks are fused into the
            dgrad / pool / upsample kernels (plain nets) or applied by the BN backward kernel (BN nets)

Channel counts that are not multiples of 64 (the 32-channel layers of TraditionalUnet) are zero-padded to 64 in the
activation buffers and operands; padded channels stay exactly zero in both directions.

Gradients are written straight into caller-provided fp32 OIHW tensors (normally views of one flat buffer), and
`on_grads_ready(names)` is called after the launches that complete each layer's gradients so a data-parallel
trainer can start that bucket's all-reduce while the remaining dgrad/wgrad kernels still run.
"""
import contextlib
import functools
import os
import struct

import torch

from . import ops

# (features index, Cin, Cout) and 'M' for nn.MaxPool2d(2,2): cfgs['D'] of nets/vgg.py:62-64 without the last pool
VGG16_CFG = [(0, 3, 64), (2, 64, 64), "M", (5, 64, 128), (7, 128, 128), "M", (10, 128, 256), (12, 256, 256),
             (14, 256, 256), "M", (17, 256, 512), (19, 512, 512), (21, 512, 512), "M", (24, 512, 512), (26, 512, 512),
             (28, 512, 512)]
# (module name, C_skip, C_up, C_out): in_filters [192, 384, 768, 1024], out_filters [64, 128, 256, 512] (nets/unet.py:28-45)
DECODER_CFG = [("up_concat4", 512, 512, 512), ("up_concat3", 256, 512, 256), ("up_concat2", 128, 256, 128),
               ("up_concat1", 64, 128, 64)]


def pad64(c):
    return (c + 63) // 64 * 64


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


class ConvSpec:
    """One conv3x3 (+BN) + ReLU layer.  name/bn: parameter-name prefixes; c0/c1: real channels of the two sources
    of a virtual concat (c1 = 0: single source)."""
    __slots__ = ("name", "bn", "cin", "cout", "first", "c0", "c1", "wf", "wd", "cout_p", "c0_p", "c1_p")

    def __init__(self, name, cin, cout, first=False, c0=None, c1=0, bn=None):
        self.name, self.bn, self.cin, self.cout, self.first = name, bn, cin, cout, first
        self.c0 = cin if c0 is None else c0
        self.c1 = c1
        self.cout_p = pad64(cout)
        self.c0_p = 64 if first else pad64(self.c0)
        self.c1_p = pad64(c1) if c1 else 0
        self.wf = self.wd = None

    @property
    def padded(self):
        return self.cout_p != self.cout or (not self.first and (self.c0_p != self.c0 or self.c1_p != self.c1))


class UNetConfig:
    """enc: list of blocks (lists of ConvSpec) separated by 2x2 max-pools; dec: list of (conv1, conv2) stages, stage i
    consumes the output of stage i-1 (or the last encoder block) upsampled 2x plus encoder block len(enc)-2-i as skip;
    head: (name, real input channels)."""

    def __init__(self, enc, dec, head, num_classes, in_channels=3):
        self.enc, self.dec, self.head, self.num_classes, self.in_channels = enc, dec, head, num_classes, in_channels
        self.bn = any(c.bn for b in enc for c in b)


def vgg_unet_config(num_classes, in_channels=3):
    enc, block = [], []
    for item in VGG16_CFG:
        if item == "M":
            enc.append(block)
            block = []
        else:
            i, cin, cout = item
            block.append(ConvSpec(f"vgg.features.{i}", in_channels if i == 0 else cin, cout, first=(i == 0)))
    enc.append(block)
    dec = [(ConvSpec(f"{name}.conv1", cs + cu, co, c0=cs, c1=cu), ConvSpec(f"{name}.conv2", co, co))
           for name, cs, cu, co in DECODER_CFG]
    return UNetConfig(enc, dec, ("final", 64), num_classes, in_channels)


def traditional_unet_config(num_classes, in_channels=3):
    """nets/TraditionalUnet.py:45-93: DoubleConv = (conv3x3+bias, BN, ReLU) x 2; widths 32-64-128-256; Up concatenates
    [skip, up] (:41) like unetUp."""
    def double(prefix, cin, cout, first=False, c0=None, c1=0):
        return [ConvSpec(f"{prefix}.double_conv.0", cin, cout, first=first, c0=c0, c1=c1, bn=f"{prefix}.double_conv.1"),
                ConvSpec(f"{prefix}.double_conv.3", cout, cout, bn=f"{prefix}.double_conv.4")]
    enc = [double("inc", in_channels, 32, first=True), double("down1.maxpool_conv.1", 32, 64),
           double("down2.maxpool_conv.1", 64, 128), double("down3.maxpool_conv.1", 128, 256)]
    dec = [tuple(double("up1.conv", 384, 128, c0=128, c1=256)), tuple(double("up2.conv", 192, 64, c0=64, c1=128)),
           tuple(double("up3.conv", 96, 32, c0=32, c1=64))]
    return UNetConfig(enc, dec, ("outc", 32), num_classes, in_channels)


class UNetEngine:
    def __init__(self, cfg, device=None):
        if cfg.in_channels * 9 > 64:
            raise ValueError("in_channels must be <= 7 (first-layer im2col is 64 columns wide)")
        if not 1 <= cfg.num_classes <= 32:
            raise ValueError("num_classes must be in [1, 32]")
        self.cfg = cfg
        self.num_classes = cfg.num_classes
        self.in_channels = cfg.in_channels
        self.device = device
        self.enc, self.dec, self.bn = cfg.enc, cfg.dec, cfg.bn
        self.head_name, self.head_cin = cfg.head
        self.convs = [c for b in self.enc for c in b] + [c for pair in self.dec for c in pair]
        self.eps, self.momentum = 1e-5, 0.1                 # nn.BatchNorm2d defaults (the reference never changes them)
        # The wgrad kernel can produce db in the same pass (an N=16 MMA against a ones tile); measured on B200 it costs
        # more than the separate HBM-bound column-sum kernel (work units with the extra MMAs become the stragglers of
        # the static schedule: +40 % wgrad time vs +1.2 ms for bias_grad), so it is off by default.
        # BatchNorm statistics from the conv epilogue (one pass over z less); only where the reduction is long enough for the
        # extra epilogue work to hide behind the main loop (scripts/ab_fuse_bnstats.py)
        self.fuse_bn_stats = True
        self.center_pre_bn = os.environ.get("B2U_CENTER_PRE_BN", "1") == "1"      # pre-BatchNorm tensors stored centred on the running mean
        self.sync_bn_group = None        # torch.distributed group: BatchNorm statistics over all ranks (SyncBatchNorm)
        self.bn_stats_min_k = 1024
        self.bn_stats_min_cout = 256
        # db from the wgrad kernel's bias warps (3x3 layers) instead of a separate pass over dz: an A/B on one box
        # (scripts/ab_fuse_bias.py: 23.2-23.9 vs 23.4-23.5 ms/step) shows no gain, so the separate pass stays the default
        self.fuse_bias_grad = False
        # bilinear 2x up-sampling + concat folded into the decoder conv's operand load (b2u_decoder_conv_fprop).  Measured at
        # the four unetUp shapes of the headline step (scripts/ab_fused_upsample.py, profiles/r2_fused_upsample.json): with
        # N tiles of 128 / 256 output channels the fused conv beats upsample + conv (0.42 / 0.65 / 0.70 ms against 0.42 / 0.67 /
        # 0.77); with the N = 64 tile (64+128 -> 64 at 512 x 512) the five interpolation warps cannot keep up with a main loop
        # that short (1.32 against 1.07 ms), so that stage keeps the separate pass.  B2U_FUSE_UPSAMPLE: 0 = never,
        # 1 = where it pays (default), 2 = every decoder stage (tests compare all three bit for bit).
        self.fuse_upsample = int(os.environ.get("B2U_FUSE_UPSAMPLE", "1"))
        # Bias gradients of the wide layers (Cout >= 128; the Cout = 64 ones get theirs from the swapped-role wgrad kernel) from
        # the data-gradient launch that PRODUCES their dz: its epilogue leaves per-tile column sums of the masked gradient it
        # stores (b2u_conv_dgrad_stats, the BatchNorm-statistics machinery) and b2u_bias_from_stats folds them, instead of
        # b2u_bias_grad re-reading dz from HBM (1.2 GB per headline step).  Those dgrads have K >= 1152, long enough to hide the
        # extra epilogue work.  Layers whose dz comes from the pool / upsample adjoints keep the separate pass.
        self.bias_from_dgrad = os.environ.get("B2U_BIAS_FROM_DGRAD", "1") == "1"
        self.bias_in_wgrad_rest = os.environ.get("B2U_BIAS_WGRAD_REST", "0") == "1"
        # ReLU backward from bit masks (plain conv + ReLU nets, training forward): every conv epilogue also writes (y > 0) as one
        # bit per channel ([N,H,W,C/64] 64-bit words, 1/16 of the bytes of y) and the masked data gradients / the head backward
        # read those 8 bytes per pixel and block instead of 128 bytes of y -- 2.8 GB less HBM traffic per headline step, and the
        # HBM-bound 64-channel data gradients at 512 x 512 then tile like unmasked launches (four stacked M tiles, no cp.async
        # mask stream).  Same-call A/B: 22.13 / 22.15 ms per step without, 21.77 / 21.73 with, results bit-identical
        # (test_relu_bit_masks*).  B2U_RELU_BITS=0 restores the bf16 masks (the pool / upsample adjoints always read y).
        self.relu_bits = os.environ.get("B2U_RELU_BITS", "1") == "1"
        # Weight/bias gradients on a second stream (plain conv+ReLU nets): wgrad_L depends only on dz_L and the saved
        # activation, not on the dgrad chain, so its launches are queued on a side stream behind an event and the
        # HBM-bound glue of the main chain (pool / upsample adjoints, column-sum folds) shares the SMs with tensor-core-bound
        # kernels of the other stream instead of running alone.  Same-box A/B (bench.py --kernels-only, 12 steps, twice each,
        # round 2 final build): 21.79 / 21.67 ms per step with it, 22.05 / 21.87 without (-1 %; round 1 measured -0.9 %).
        # Per-launch CUDA-event times are not a kernel's own duration while two streams overlap, so the roofline pass of
        # bench.py (ops.KernelTimer) runs single-stream: backward() checks ops.timing_active().  B2U_WGRAD_STREAM=0 turns it off.
        self.wgrad_stream = os.environ.get("B2U_WGRAD_STREAM", "1") == "1"
        self._side = None
        self._pack_key = self._pack_versions = self._pack_table = None
        self._pack_total = 0
        self._bufs = {}
        self.saved = None
        self._ws = {}

    # ------------------------------------------------------------------ buffers
    def _buf(self, key, shape, dtype=None, zero=False):
        dtype = ops.act_dtype() if dtype is None else dtype      # bf16; fp32 under the fp32 validation build
        t = self._bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
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

    def _padded_vec(self, key, vec, n, fill=0.0):
        """fp32 per-channel vector padded to n entries (persistent buffer, real part refreshed from `vec`)."""
        if vec.numel() == n:
            return vec
        t = self._bufs.get(key)
        if t is None or t.numel() != n:
            t = torch.full((n,), fill, dtype=torch.float32, device=self.device)
            self._bufs[key] = t
        t[:vec.numel()].copy_(vec)
        return t

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
            blob, start = b"", 0
            for c in self.convs:
                w = params[c.name + ".weight"]
                ctot_p = c.c0_p + c.c1_p
                if c.first:
                    if c.wf is None:
                        c.wf = torch.zeros((c.cout_p, 64), dtype=ops.act_dtype(), device=dev)
                    count = (c.cout + 31) // 32                      # work blocks of this layer
                else:
                    if c.wf is None:
                        c.wf = torch.zeros((c.cout_p, 9 * ctot_p), dtype=ops.act_dtype(), device=dev)
                    if need_dgrad and c.wd is None:
                        c.wd = torch.zeros((ctot_p, 9 * c.cout_p), dtype=ops.act_dtype(), device=dev)
                    count = ((c.cout + 31) // 32) * ((c.cin + 31) // 32)
                wd_ptr = c.wd.data_ptr() if (need_dgrad and not c.first) else 0
                blob += struct.pack("<QQQqiiiiiiii", w.data_ptr(), c.wf.data_ptr(), wd_ptr, start, c.cout, c.cin,
                                    1 if c.first else 9, 1 if c.first else 0, c.c0, c.c0_p, ctot_p, c.cout_p)
                start += count
            self._pack_table = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
            self._pack_total = start
            self._pack_key = key
        ops.check(ops.lib().b2u_pack_weights_multi(self._pack_table.data_ptr(), len(self.convs), self._pack_total,
                                                   ops.stream_ptr()))
        self._pack_versions = versions

    def param_shapes(self):
        """name -> shape of every trainable tensor, in the reference module's state_dict order."""
        shapes = {}
        for c in self.convs:
            shapes[c.name + ".weight"] = (c.cout, c.cin, 3, 3)
            shapes[c.name + ".bias"] = (c.cout,)
            if c.bn:
                shapes[c.bn + ".weight"] = (c.cout,)
                shapes[c.bn + ".bias"] = (c.cout,)
        shapes[self.head_name + ".weight"] = (self.num_classes, self.head_cin, 1, 1)
        shapes[self.head_name + ".bias"] = (self.num_classes,)
        return shapes

    def buffer_shapes(self):
        """BatchNorm buffers (running statistics), state_dict names."""
        shapes = {}
        for c in self.convs:
            if c.bn:
                shapes[c.bn + ".running_mean"] = (c.cout,)
                shapes[c.bn + ".running_var"] = (c.cout,)
                shapes[c.bn + ".num_batches_tracked"] = ()
        return shapes

    def backward_param_order(self):
        """Parameter names in the order backward() completes their gradients (head first, first conv last)."""
        layers = []
        for si in range(len(self.dec) - 1, -1, -1):
            layers += [self.dec[si][1], self.dec[si][0]]
        for bi in range(len(self.enc) - 1, -1, -1):
            layers += list(reversed(self.enc[bi]))
        out = [self.head_name + ".weight", self.head_name + ".bias"]
        for c in layers:
            if c.bn:
                out += [c.bn + ".weight", c.bn + ".bias"]
            out += [c.name + ".weight", c.name + ".bias"]
        return out

    def invalidate_packed_weights(self):
        """Call after writing parameters behind torch's back (raw-pointer optimizer kernels)."""
        self._pack_versions = None

    # ------------------------------------------------------------------ one conv (+BN) + ReLU layer, forward
    def _layer_fwd(self, c, x0, params, A, training, x1=None, save=True, low=None, up_out=None):
        """low (instead of x1): the second source is upsample2x(low), interpolated inside the conv kernel
        (ops.decoder_conv_fprop); up_out then optionally receives the up-sampled tensor as a by-product."""
        n, h, w, _ = x0.shape
        bias = self._padded_vec("b:" + c.name, params[c.name + ".bias"], c.cout_p)
        taps = 1 if c.first else 9

        def conv(bias_, relu, out, stats=None, scale=None):
            if low is not None:
                return ops.decoder_conv_fprop(x0, low, c.wf, bias_, c.cout_p, relu=relu, out=out, up_out=up_out, scale=scale, stats=stats)
            if scale is not None:
                return ops.conv_fprop_scaled(x0, c.wf, scale, bias_, c.cout_p, taps=taps, relu=relu, x1=x1, out=out)
            return ops.conv_fprop(x0, c.wf, bias_, c.cout_p, taps=taps, relu=relu, x1=x1, out=out, stats=stats)

        if not c.bn:
            out = self._buf(c.name, (n, h, w, c.cout_p))
            if save and self.relu_bits and ops.act_dtype() == torch.bfloat16:
                bits = self._buf("bits:" + c.name, (n, h, w, c.cout_p // 64), torch.int64)
                ops.conv_fprop_relu_bits(x0, c.wf, bias, c.cout_p, bits, taps=taps, x1=x1, low=low, up_out=up_out, out=out)
                A["bits:" + c.name] = bits
            else:
                conv(bias, True, out)
            A[c.name] = out
            return out
        if not training and not save:
            # pure inference: BatchNorm is an affine map with fixed statistics, folded into the conv epilogue
            # (y = relu(acc * s + b')): no z tensor, no BatchNorm pass.  An eval-mode forward that will be back-propagated
            # (model.eval() + loss.backward()) takes the unfused path below so the BatchNorm backward has z
            gamma = self._padded_vec("g:" + c.bn, params[c.bn + ".weight"], c.cout_p, fill=1.0)
            beta = self._padded_vec("bt:" + c.bn, params[c.bn + ".bias"], c.cout_p)
            rmp = self._padded_vec("rm:" + c.bn, params[c.bn + ".running_mean"], c.cout_p)
            rvp = self._padded_vec("rv:" + c.bn, params[c.bn + ".running_var"], c.cout_p, fill=1.0)
            sc, bs = ops.bn_fold(gamma, beta, rmp, rvp, bias, self.eps, scale=self._buf("fs:" + c.bn, (c.cout_p,), torch.float32),
                                 bias=self._buf("fb:" + c.bn, (c.cout_p,), torch.float32))
            out = self._buf(c.name, (n, h, w, c.cout_p))
            conv(bs, True, out, scale=sc)
            A[c.name] = out
            return out
        z = self._buf("z:" + c.name, (n, h, w, c.cout_p))
        rm, rv = params[c.bn + ".running_mean"], params[c.bn + ".running_var"]
        # Training: z is stored CENTRED on the running mean -- the shift rides in the conv bias (fp32, before the bf16
        # rounding), BatchNorm is shift-invariant, and only the running-mean update adds it back.  A channel whose mean is large
        # against its spread otherwise loses |mean|/std * 2^-9 of relative accuracy in bf16 (the dominant error site of the
        # depthwise nets on warm weights, profiles/r2_precision_sites.txt); at initialisation the running mean is 0: no change.
        centered = training and self.center_pre_bn and ops.act_dtype() == torch.bfloat16
        if centered:
            rmp0 = self._padded_vec("rm:" + c.bn, rm, c.cout_p)
            bias = torch.sub(bias, rmp0, out=self._buf("bc:" + c.name, (c.cout_p,), torch.float32))
        # training: the conv epilogue also emits the BatchNorm statistics of z (per-tile sums), so BatchNorm skips its
        # statistics pass over z
        stats, rows = None, 0
        kdim = taps * (x0.shape[3] + (x1.shape[3] if x1 is not None else 0) + (low.shape[3] if low is not None else 0))
        if training and self.fuse_bn_stats and self.sync_bn_group is None and (kdim >= self.bn_stats_min_k or c.cout_p >= self.bn_stats_min_cout):
            rows = ops.conv_stat_rows(n, h, w, c.cout_p, taps, bn=(1 << 18) if low is not None else 0)      # bit 18: the decoder conv's tiles
            stats = self._workspace("bnstat", rows * 2 * c.cout_p * 4)[:rows * 2 * c.cout_p * 4].view(torch.float32)
        conv(bias, False, z, stats=stats)
        gamma = self._padded_vec("g:" + c.bn, params[c.bn + ".weight"], c.cout_p, fill=1.0)
        beta = self._padded_vec("bt:" + c.bn, params[c.bn + ".bias"], c.cout_p)
        out = self._buf(c.name, (n, h, w, c.cout_p))
        ws = self._workspace("bn", ops.lib().b2u_bn_workspace(c.cout_p))
        if training and self.sync_bn_group is not None:
            rmp = self._padded_vec("rm:" + c.bn, rm, c.cout_p)
            rvp = self._padded_vec("rv:" + c.bn, rv, c.cout_p, fill=1.0)
            _, mean, invstd = ops.bn_fwd_train_sync(z, gamma, beta, rmp, rvp, self.sync_bn_group, self.eps, self.momentum, True,
                                                    out=out, ws=ws, centered=centered)
            if c.cout_p != c.cout or rmp is not rm:
                rm.copy_(rmp[:c.cout]); rv.copy_(rvp[:c.cout])
            nbt = params.get(c.bn + ".num_batches_tracked")
            if nbt is not None:
                nbt.add_(1)
            A["bn:" + c.name] = (z, mean, invstd, gamma, beta)
        elif training:
            if c.cout_p == c.cout:
                _, mean, invstd = ops.bn_fwd_train(z, gamma, beta, rm, rv, self.eps, self.momentum, True, out=out, ws=ws,
                                                   stats=stats, stat_rows=rows, centered=centered)
            else:
                rmp = self._padded_vec("rm:" + c.bn, rm, c.cout_p)
                rvp = self._padded_vec("rv:" + c.bn, rv, c.cout_p, fill=1.0)
                _, mean, invstd = ops.bn_fwd_train(z, gamma, beta, rmp, rvp, self.eps, self.momentum, True, out=out, ws=ws,
                                                   stats=stats, stat_rows=rows, centered=centered)
                rm.copy_(rmp[:c.cout])
                rv.copy_(rvp[:c.cout])
            nbt = params.get(c.bn + ".num_batches_tracked")
            if nbt is not None:
                nbt.add_(1)
            A["bn:" + c.name] = (z, mean, invstd, gamma, beta)
        else:
            rmp = self._padded_vec("rm:" + c.bn, rm, c.cout_p)
            rvp = self._padded_vec("rv:" + c.bn, rv, c.cout_p, fill=1.0)
            ops.bn_fwd_eval(z, gamma, beta, rmp, rvp, self.eps, True, out=out, ws=ws)
            A["bn:" + c.name] = (z, (rmp, rvp), None, gamma, beta)      # eval-mode backward: statistics are constants
        A[c.name] = out
        return out

    # ------------------------------------------------------------------ forward
    def forward(self, x, params, save=True, training=None, trainable=None):
        """x: NCHW fp32 CUDA [N, Cin, H, W] -> logits NCHW fp32.  training (BN nets): batch statistics + running-stat
        update (default: same as `save`).  trainable is accepted for interface parity with GraphEngine (backward decides
        what to skip from the gradient names it is given)."""
        if not x.is_cuda:
            raise ValueError("UNetEngine.forward: input must be a CUDA tensor (no CPU fallback)")
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        N, C, H, W = x.shape
        if C != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {C}")
        div = 2 ** (len(self.enc) - 1)
        if H % div or W % div:
            raise ValueError(f"input height and width must be multiples of {div}")
        if training is None:
            training = save
        self.device = x.device
        self.pack(params, need_dgrad=save)
        A = {}
        col = self._buf("col", (N, H, W, 64))
        ops.check(ops.lib().b2u_im2col_first(x.data_ptr(), col.data_ptr(), N, C, H, W, ops.stream_ptr()))
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
                cur = self._layer_fwd(c, cur, params, A, training, save=save)
            feats.append(cur)
        low = feats[-1]
        for si, (c1, c2) in enumerate(self.dec):
            skip = feats[len(self.enc) - 2 - si]
            _, hl, wl, cl = low.shape
            if self.fuse_upsample == 2 or (self.fuse_upsample == 1 and c1.cout_p >= 128):
                # the decoder conv interpolates `low` inside its producer warps (nets/unet.py:16-18 in one kernel); the
                # up-sampled tensor is written only when a backward pass will need it as the weight gradient's operand
                # (the fp32 validation build always takes the buffer: it keeps the two steps apart)
                keep = save or ops.act_dtype() == torch.float32
                up = self._buf(f"up{si}", (N, 2 * hl, 2 * wl, cl)) if keep else None
                o1 = self._layer_fwd(c1, skip, params, A, training, save=save, low=low, up_out=up)
            else:
                up = self._buf(f"up{si}", (N, 2 * hl, 2 * wl, cl))
                ops.upsample2x(low, out=up)
                o1 = self._layer_fwd(c1, skip, params, A, training, x1=up, save=save)
            A[f"up{si}"] = up
            low = self._layer_fwd(c2, o1, params, A, training, save=save)
        wh = self._head_weight(params)
        # 1x1 classifier on the tensor cores: [hi | lo] bf16 split of the fp32 weights, fp32 NCHW logits from the epilogue
        if ops.act_dtype() == torch.float32:       # fp32 validation build: plain fp32 1x1 conv
            logits = ops.head_fwd(low, wh, params[self.head_name + ".bias"])
        else:
            wf_head = ops.pack_head_fprop(wh, wf=self._buf("head:wf", (64, 64)))
            logits = ops.head_fwd_tc(low, wf_head, params[self.head_name + ".bias"], self.num_classes)
        if save:
            self.saved = (A, feats, (N, H, W))
        return logits

    def _head_weight(self, params):
        w = params[self.head_name + ".weight"].reshape(self.num_classes, self.head_cin)
        if self.head_cin == 64:
            return w
        t = self._buf("head:w", (self.num_classes, 64), torch.float32, zero=True)
        t[:, :self.head_cin].copy_(w)
        return t

    # ------------------------------------------------------------------ backward
    def backward(self, dlogits, params, grads, trainable=None, on_grads_ready=None):
        """dlogits: NCHW fp32, or bf16 [N,H,W,64] from ops.loss_bwd(nhwc64=True).  grads: name -> fp32 tensor to
        overwrite (missing / not in `trainable`: skipped).
        Returns nothing; the input image gets no gradient (the reference never asks for one)."""
        if self.saved is None:
            raise RuntimeError("backward() without a saved forward()")
        side = None
        if self.wgrad_stream and not self.bn and dlogits.is_cuda and not ops.timing_active():
            if self._side is None or self._side.device != dlogits.device:
                self._side = torch.cuda.Stream(device=dlogits.device)
            side = self._side
            # size the shared workspaces before the side stream starts using them (a reallocation mid-backward would hand
            # the old block back to the allocator while a kernel of the other stream may still read it)
            A, _, (N, H, W) = self.saved
            need = 0
            for c in self.convs:
                t = A[c.name]
                need = max(need, ops.lib().b2u_conv_wgrad_workspace(t.shape[0], t.shape[1], t.shape[2],
                                                                     64 if c.first else c.c0_p + c.c1_p, c.cout_p, 1 if c.first else 9))
            need = max(need, ops.lib().b2u_conv_wgrad_workspace(N, H, W, 64, 64, 1))
            self._workspace("wgrad", need)
            self._workspace("bias", ops.lib().b2u_bias_grad_workspace(max(c.cout_p for c in self.convs)))
            side.wait_stream(torch.cuda.current_stream())
        try:
            self._backward(dlogits, params, grads, trainable, on_grads_ready, side)
        finally:
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)      # optimizer / next forward see every gradient

    def _backward(self, dlogits, params, grads, trainable, on_grads_ready, side):
        A, feats, (N, H, W) = self.saved
        if trainable is None:
            trainable = set(grads.keys())
        bn = self.bn
        order = [c.name for c in self.convs] + [self.head_name]

        def wanted(c):
            names = [c.name + ".weight", c.name + ".bias"] + ([c.bn + ".weight", c.bn + ".bias"] if c.bn else [])
            return any(n in trainable for n in names)
        want = {c.name: wanted(c) for c in self.convs}
        want[self.head_name] = (self.head_name + ".weight") in trainable or (self.head_name + ".bias") in trainable
        # data gradients are needed down to the first (in execution order) trainable conv
        first_trainable = next((i for i, n in enumerate(order) if want[n]), len(order))
        need_dx = {n: i > first_trainable for i, n in enumerate(order)}   # does layer n have to produce dx?
        n_enc = sum(len(b) for b in self.enc)
        enc_trainable = first_trainable < n_enc

        def ready(*names):
            if on_grads_ready is not None:
                on_grads_ready([n for n in names if n in grads])

        def grad_stream():
            """Context for the launches that produce parameter gradients: the side stream, ordered after everything queued on
            the main stream so far (the dz they read), or a no-op."""
            if side is None:
                return contextlib.nullcontext()
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            side.wait_event(ev)
            return torch.cuda.stream(side)

        def has(n):
            return n in trainable and n in grads

        db_stats = {}      # conv name -> (per-tile column sums of its dz, rows), left behind by the dgrad that produced dz

        def dgrad_to(below, dz, wd, C0, mask, out0):
            """Masked data gradient that becomes `below`'s dz; also records dz's column sums when `below` wants a bias gradient
            that would otherwise need a pass over dz.  The ReLU mask is `below`'s bit mask when the forward left one."""
            bits = A.get("bits:" + below.name) if mask is not None else None
            stats = None
            if (self.bias_from_dgrad and not bn and ops.act_dtype() == torch.bfloat16 and want[below.name] and has(below.name + ".bias")
                    and not below.padded and not below.first and below.cout_p >= 128 and not self.fuse_bias_grad):
                n_, h_, w_, _ = out0.shape
                rows = ops.conv_dgrad_stat_rows(n_, h_, w_, C0, 9, masked=mask is not None and bits is None)
                nbytes = rows * 2 * C0 * 4
                stats = self._workspace("dbstat:" + below.name, nbytes)[:nbytes].view(torch.float32)
                db_stats[below.name] = (stats, rows)
            if bits is not None:
                ops.conv_dgrad_bits(dz, wd, C0, bits, out0=out0, stats=stats)
            else:
                ops.conv_dgrad(dz, wd, C0, mask=mask, out0=out0, stats=stats)

        def layer_bwd(c, x0, g, x1=None):
            """g: gradient wrt the layer's output (BN nets: wrt y, unmasked; plain nets: wrt the pre-activation, already
            masked).  Runs BN backward if any, wgrad, bias grad; returns dz (the conv's output gradient)."""
            dz = g
            if c.bn:
                z, mean, invstd, gamma, beta = A["bn:" + c.name]
                wn, bnn = c.bn + ".weight", c.bn + ".bias"
                dgam = self._buf("dg:" + c.bn, (c.cout_p,), torch.float32)
                dbet = self._buf("db:" + c.bn, (c.cout_p,), torch.float32)
                if invstd is None:      # eval-mode forward: running statistics, mask from the saved output
                    ops.bn_bwd_eval(g, A[c.name], z, gamma, beta, mean[0], mean[1], self.eps, relu=True, out=g, dgamma=dgam,
                                    dbeta=dbet, ws=self._workspace("bn", ops.lib().b2u_bn_workspace(c.cout_p)))
                else:
                    bwd = ops.bn_bwd if self.sync_bn_group is None else functools.partial(ops.bn_bwd_sync, group=self.sync_bn_group)
                    bwd(g, None, z, gamma, mean, invstd, relu=True, out=g, dgamma=dgam, dbeta=dbet, beta=beta,   # mask from z
                        ws=self._workspace("bn", ops.lib().b2u_bn_workspace(c.cout_p)))
                if has(wn):
                    grads[wn].copy_(dgam[:c.cout])
                if has(bnn):
                    grads[bnn].copy_(dbet[:c.cout])
                ready(wn, bnn)
            if not want[c.name]:
                return dz
            with grad_stream():
                param_grads(c, x0, dz, x1)
            return dz

        def param_grads(c, x0, dz, x1):
            wn, bn_ = c.name + ".weight", c.name + ".bias"
            want_w, want_b = has(wn), has(bn_)
            # Cout = 64: db is free in the swapped-role wgrad kernel.  bias_in_wgrad_rest: the wide layers whose dz comes from the
            # pool / upsample adjoints (no producing dgrad to take column sums from) let the wgrad kernel's bias warps sum the dz
            # tiles while they sit in shared memory, instead of a separate pass over dz
            fuse_b = (want_w and want_b and not c.padded and not c.bn and not c.first
                      and (self.fuse_bias_grad or c.cout_p == 64 or (self.bias_in_wgrad_rest and c.name not in db_stats)))
            taps = 1 if c.first else 9
            if want_w:
                ctot_p = 64 if c.first else c.c0_p + c.c1_p
                need = ops.lib().b2u_conv_wgrad_workspace(dz.shape[0], dz.shape[1], dz.shape[2], ctot_p, c.cout_p, taps)
                ws = self._workspace("wgrad", need)
                if not c.padded:
                    ops.conv_wgrad(x0, dz, taps=taps, x1=x1, first_cin=c.cin if c.first else 0, dw=grads[wn],
                                   db=grads[bn_] if fuse_b else None, ws=ws)
                else:
                    if c.first:
                        tmp = self._buf("dw:" + c.name, (c.cout_p, c.cin, 3, 3), torch.float32)
                        ops.conv_wgrad(x0, dz, taps=1, first_cin=c.cin, dw=tmp, ws=ws)
                        grads[wn].copy_(tmp[:c.cout])
                    else:
                        tmp = self._buf("dw:" + c.name, (c.cout_p, ctot_p, 3, 3), torch.float32)
                        ops.conv_wgrad(x0, dz, taps=9, x1=x1, dw=tmp, ws=ws)
                        if c.c1:
                            grads[wn][:, :c.c0].copy_(tmp[:c.cout, :c.c0])
                            grads[wn][:, c.c0:].copy_(tmp[:c.cout, c.c0_p:c.c0_p + c.c1])
                        else:
                            grads[wn].copy_(tmp[:c.cout, :c.c0])
            if want_b and c.bn:
                # a bias in front of BatchNorm: dz = a (g - mean(g) - xhat mean(g xhat)) sums to zero over the batch
                # exactly, so the gradient is 0 (the reference's autograd returns fp32 cancellation residue ~1e-8)
                grads[bn_].zero_()
            elif want_b and not fuse_b and c.name in db_stats:
                st_, rows_ = db_stats.pop(c.name)
                ops.bias_from_stats(st_, rows_, c.cout_p, db=grads[bn_])
            elif want_b and not fuse_b:
                if c.cout_p == c.cout:
                    ops.bias_grad(dz, db=grads[bn_], ws=self._workspace("bias", ops.lib().b2u_bias_grad_workspace(c.cout_p)))
                else:
                    tmpb = self._buf("dbias:" + c.name, (c.cout_p,), torch.float32)
                    ops.bias_grad(dz, db=tmpb, ws=self._workspace("bias", ops.lib().b2u_bias_grad_workspace(c.cout_p)))
                    grads[bn_].copy_(tmpb[:c.cout])
            ready(wn, bn_)

        # ---- head
        hn = self.head_name
        last = A[self.dec[-1][1].name]
        wh = self._head_weight(params)
        C = self.num_classes
        fw, fb = has(hn + ".weight"), has(hn + ".bias")
        g = self._buf("g:" + self.dec[-1][1].name, last.shape) if need_dx[hn] else None
        if dlogits.dtype == torch.bfloat16 and dlogits.dim() == 4 and dlogits.shape[-1] == 64 and ops.act_dtype() == torch.bfloat16:
            # [N,H,W,64] = [hi | lo] split dlogits from loss_bwd(nhwc64=True): the head's backward runs on the tensor
            # cores as a 1x1 dgrad (+ReLU mask) and a 1x1 wgrad (+bias) over 2 x 32 padded classes
            dl = dlogits.contiguous()
            if need_dx[hn]:
                wd_head = ops.pack_head_dgrad(wh, wd=self._buf("head:wd", (64, 64)))
                hbits = None if bn else A.get("bits:" + self.dec[-1][1].name)
                if hbits is not None:
                    ops.conv_dgrad_bits(dl, wd_head, 64, hbits, taps=1, out0=g)
                else:
                    ops.conv_dgrad(dl, wd_head, 64, taps=1, mask=None if bn else last, out0=g)
            if fw or fb:
                dw64 = self._buf("head:dw", (64, 64, 1, 1), torch.float32)
                db64 = self._buf("head:db", (64,), torch.float32)
                need = ops.lib().b2u_conv_wgrad_workspace(dl.shape[0], dl.shape[1], dl.shape[2], 64, 64, 1)
                with grad_stream():
                    ops.conv_wgrad(last, dl, taps=1, dw=dw64, db=db64, ws=self._workspace("wgrad", need))
                    if fw:      # rows [0,32) came from the hi half of dlogits, rows [32,64) from the lo half
                        torch.add(dw64[:C, :self.head_cin], dw64[32:32 + C, :self.head_cin], out=grads[hn + ".weight"])
                    if fb:
                        torch.add(db64[:C], db64[32:32 + C], out=grads[hn + ".bias"])
        else:
            dl = dlogits.contiguous()
            if dl.dtype != torch.float32:
                dl = dl.float()
            if self.head_cin == 64:
                dwt, dbt = (grads[hn + ".weight"] if fw else None), (grads[hn + ".bias"] if fb else None)
            else:
                dwt = self._buf("head:dwf", (C, 64, 1, 1), torch.float32) if fw else None
                dbt = grads[hn + ".bias"] if fb else None
            ops.head_bwd(dl, last, wh, need_dx=need_dx[hn], need_dw=fw or fb, relu_mask=not bn, dx=g, dw=dwt, db=dbt,
                         ws=self._workspace("head", ops.lib().b2u_head_bwd_workspace()))
            if fw and self.head_cin != 64:
                grads[hn + ".weight"].copy_(dwt[:, :self.head_cin])
        with grad_stream():          # (orders the side stream after the head's launches, whichever stream ran them)
            ready(hn + ".weight", hn + ".bias")
        if g is None:
            return

        # ---- decoder, last stage first
        n_dec = len(self.dec)
        dskips = [None] * (len(self.enc) - 1)     # gradient wrt the encoder features coming from the decoder
        for si in range(n_dec - 1, -1, -1):
            c1, c2 = self.dec[si]
            fi = len(self.enc) - 2 - si
            skip = feats[fi]
            up = A[f"up{si}"]
            o1 = A[c1.name]
            low = feats[-1] if si == 0 else A[self.dec[si - 1][1].name]
            dz = layer_bwd(c2, o1, g)
            if not need_dx[c2.name]:
                return
            g1 = self._buf("g:" + c1.name, o1.shape)
            dgrad_to(c1, dz, c2.wd, c2.c0_p, None if bn else o1, g1)
            dz1 = layer_bwd(c1, skip, g1, x1=up)
            if not need_dx[c1.name]:
                return
            dup = self._buf(f"g:up{si}", up.shape)
            if enc_trainable:
                dsk = self._buf(f"g:skip{fi}", skip.shape)
                ops.conv_dgrad(dz1, c1.wd, c1.c0_p, C1=c1.c1_p, out0=dsk, out1=dup)
                dskips[fi] = dsk
            else:
                # frozen encoder: only the up-sampled half of the concat needs a gradient
                ops.conv_dgrad(dz1, c1.wd[c1.c0_p:], c1.c1_p, out0=dup)
            if si == 0 and not enc_trainable:
                return
            g = self._buf("g:low" + str(si), low.shape)
            ops.upsample2x_bwd(dup, ylow=None if bn else low, out=g)

        # ---- encoder, deepest block first; g = gradient wrt the block's last conv output
        for bi in range(len(self.enc) - 1, -1, -1):
            block = self.enc[bi]
            for ci in range(len(block) - 1, -1, -1):
                c = block[ci]
                if ci > 0:
                    xin = A[block[ci - 1].name]
                elif bi > 0:
                    xin = A[f"pool{bi}"]
                else:
                    xin = A["col"]
                dz = layer_bwd(c, xin, g)
                if not need_dx[c.name]:
                    return
                if ci > 0:
                    nxt = self._buf("g:" + block[ci - 1].name, xin.shape)
                    dgrad_to(block[ci - 1], dz, c.wd, c.c0_p, None if bn else xin, nxt)
                    g = nxt
                else:
                    dpool = self._buf(f"g:pool{bi}", xin.shape)
                    ops.conv_dgrad(dz, c.wd, c.c0_p, out0=dpool)
                    y = feats[bi - 1]
                    nxt = self._buf(f"g:feat{bi - 1}", y.shape)
                    ops.maxpool2x2_bwd(dpool, y, dskip=dskips[bi - 1], relu_mask=not bn, out=nxt)
                    g = nxt


class VGGUnetEngine(UNetEngine):
    def __init__(self, num_classes, in_channels=3, device=None):
        super().__init__(vgg_unet_config(num_classes, in_channels), device=device)


class TraditionalUnetEngine(UNetEngine):
    def __init__(self, num_classes, in_channels=3, device=None):
        super().__init__(traditional_unet_config(num_classes, in_channels), device=device)
