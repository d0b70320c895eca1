"""SAUNet graph + wrappers: the nn.Module surface of the reference's models/models.py
(``SAUNet`` :264-401, ``DecoderBlock`` :203-237, ``conv3x3_bn_relu`` :118-123, ``ModelBuilder`` :143-166,
``SegmentationModule`` :80-109), re-built on the B200 CUDA path.

The modules own ``nn.Parameter``s / buffers under the reference's exact state_dict keys (1708 entries incl. the
aliased ``encoder.features.*`` <-> ``conv1.0.* / conv2.* ...`` pairs), so optimizers, ``group_weight`` and
checkpoints from the reference's train.py keep working.  ``forward`` runs the whole graph as ONE autograd node
whose forward/backward are hand-written sequences of libsaunet_b200.so kernels on NHWC buffers
(saunet_b200.engine / saunet_b200.blocks).  No torch arithmetic, no CPU fallback.
"""
import math
import os
import warnings

import torch
import torch.nn as nn

from saunet_b200 import _C, engine
from saunet_b200.blocks import (basic_block_body, cat_copy, conv_bn_relu_body, conv_op, decoder_block_body,
                                dense_block_body, dual_att_body, gsconv_body, maxpool2, transition_body, _check_bn,
                                _stat)
from saunet_b200.engine import (ACT_NONE, ACT_SIGMOID, _round4, affine_act, bilinear, bn_backward, bn_finalize,
                                bn_stats_slot, channel_stats, conv2d, conv2d_bwd)
from . import GSConv as gsc
from .attention_blocks import DualAttBlock
from .densenet import densenet121
from .norm import Norm2d
from .resnet import BasicBlock as ResBlock


class SegmentationModuleBase(nn.Module):
    """Metric helpers of models/models.py:21-78 (host-side torch ops; outside the kernel path)."""

    def pixel_acc(self, pred, label, num_class):
        preds = torch.argmax(pred, dim=1)
        valid = label >= 1
        acc_sum = torch.sum(valid & (preds == label))
        pixel_sum = torch.sum(valid)
        acc = acc_sum.float() / (pixel_sum.float() + 1e-10)
        jaccard = []
        for i in range(1, num_class):
            v = label == i
            p = preds == i
            anb = torch.sum(v & p).float()
            j = anb / (torch.sum(v).float() + torch.sum(p).float() - anb + 1e-10)
            # reference: `j if j <= 1 else 0` (a host branch = a device sync per class); same value, no sync
            jaccard.append(torch.where(j <= 1, j, torch.zeros_like(j)))
        return acc, jaccard

    def jaccard(self, pred, label):
        anb = torch.sum(pred.long() & label)
        return anb / (pred.view(-1).sum().float() + label.view(-1).sum().float() - anb)


class SegmentationModule(SegmentationModuleBase):
    """models/models.py:80-109: ``forward(feed_dict, epoch, *, segSize=None, return_att=False)``."""

    def __init__(self, crit, unet, num_class):
        super().__init__()
        self.crit = crit
        self.unet = unet
        self.num_class = num_class

    def forward(self, feed_dict, epoch, *, segSize=None, return_att=False):
        if segSize is None:          # training
            p = self.unet(feed_dict["image"])
            loss = self.crit(p, feed_dict["mask"], epoch=epoch)
            # metrics (models/models.py:92): counted by the loss kernel's own pass over the logits when `crit` is this
            # repo's DualLoss; any other criterion takes the reference's torch expression
            fm = getattr(self.crit, "fused_metrics", None)
            acc = fm(p[0]) if fm is not None and self.num_class == p[0].shape[1] else None
            if acc is None:
                label = feed_dict["mask"][0].long().to(p[0].device)
                acc = self.pixel_acc(torch.round(nn.functional.softmax(p[0].detach(), dim=1)).long(), label,
                                     self.num_class)
            return loss, acc
        if segSize is True:          # test
            maps = None
            if return_att:
                x_out, edge_out, maps = self.unet(feed_dict["image"], return_att=True)
            else:
                x_out, edge_out = self.unet(feed_dict["image"], return_att=False)
            return nn.functional.softmax(x_out, dim=1), maps
        p = self.unet(feed_dict["image"], return_att=return_att)      # inference
        loss = self.crit((p[0], p[1]), (feed_dict["mask"][0].long().unsqueeze(0), feed_dict["mask"][1].unsqueeze(0)))
        return nn.functional.softmax(p[0], dim=1), loss


def conv3x3_bn_relu(in_planes, out_planes, stride=1):
    return nn.Sequential(nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1),
                         nn.BatchNorm2d(out_planes),
                         nn.ReLU(inplace=True))


class ModelBuilder():
    def build_unet(self, num_class=1, arch="albunet", weights=""):
        arch = arch.lower()
        if arch == "saunet":
            unet = SAUNet(num_classes=num_class)
        else:
            raise Exception("Architecture undefined!")
        if len(weights) > 0:
            unet.load_state_dict(torch.load(weights, map_location=lambda storage, loc: storage), strict=False)
            print("Loaded pretrained UNet weights.")
        print("Loaded weights for unet")
        return unet


class DecoderBlock(nn.Module):
    def __init__(self, in_channels, middle_channels, out_channels, is_deconv=True):
        super().__init__()
        self.in_channels = in_channels
        if not is_deconv:
            raise NotImplementedError("saunet_b200 DecoderBlock: only is_deconv=True (what SAUNet builds) is implemented")
        self.block = nn.Sequential(
            conv3x3_bn_relu(in_channels, middle_channels),
            nn.ConvTranspose2d(middle_channels, out_channels, kernel_size=4, stride=2, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))
        for m in self.modules():     # models/models.py:223-234 (Conv2d only; the ConvTranspose2d keeps torch's default)
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _body(self, tp, x, out=None):
        return decoder_block_body(tp, self, x, out)

    def forward(self, x):
        return engine.run(self, lambda tp, xb: [self._body(tp, xb)], [x])[0]


class SAUNet(nn.Module):
    def __init__(self, num_classes=4, num_filters=32, pretrained=True, is_deconv=True):
        super().__init__()
        self.num_classes = num_classes
        self.pool = nn.MaxPool2d(2, 2)
        self.encoder = densenet121(pretrained=False)
        self.relu = nn.ReLU(inplace=True)
        self.sigmoid = nn.Sigmoid()
        # shape stream
        self.c3 = nn.Conv2d(256, 1, kernel_size=1)
        self.c4 = nn.Conv2d(512, 1, kernel_size=1)
        self.c5 = nn.Conv2d(1024, 1, kernel_size=1)
        self.d0 = nn.Conv2d(128, 64, kernel_size=1)
        self.res1 = ResBlock(64, 64)
        self.d1 = nn.Conv2d(64, 32, kernel_size=1)
        self.res2 = ResBlock(32, 32)
        self.d2 = nn.Conv2d(32, 16, kernel_size=1)
        self.res3 = ResBlock(16, 16)
        self.d3 = nn.Conv2d(16, 8, kernel_size=1)
        self.fuse = nn.Conv2d(8, 1, kernel_size=1, padding=0, bias=False)
        self.cw = nn.Conv2d(2, 1, kernel_size=1, padding=0, bias=False)
        self.gate1 = gsc.GatedSpatialConv2d(32, 32)
        self.gate2 = gsc.GatedSpatialConv2d(16, 16)
        self.gate3 = gsc.GatedSpatialConv2d(8, 8)
        self.expand = nn.Sequential(nn.Conv2d(1, num_filters, kernel_size=1), Norm2d(num_filters), nn.ReLU(inplace=True))
        # encoder aliases (models/models.py:304-313): same modules under a second name
        f = self.encoder.features
        self.conv1 = nn.Sequential(f.conv0, f.norm0)
        self.conv2 = f.denseblock1
        self.conv2t = f.transition1
        self.conv3 = f.denseblock2
        self.conv3t = f.transition2
        self.conv4 = f.denseblock3
        self.conv4t = f.transition3
        self.conv5 = nn.Sequential(f.denseblock4, f.norm5)
        # decoder
        self.center = conv3x3_bn_relu(1024, num_filters * 8 * 2)
        self.dec5 = DualAttBlock(inchannels=[512, 1024], outchannels=512)
        self.dec4 = DualAttBlock(inchannels=[512, 512], outchannels=256)
        self.dec3 = DualAttBlock(inchannels=[256, 256], outchannels=128)
        self.dec2 = DualAttBlock(inchannels=[128, 128], outchannels=64)
        self.dec1 = DecoderBlock(64, 48, num_filters, is_deconv)
        self.dec0 = conv3x3_bn_relu(num_filters * 2, num_filters)
        self.final = nn.Conv2d(num_filters, self.num_classes, kernel_size=1)
        if pretrained:
            path = os.environ.get("SAUNET_DENSENET121_WEIGHTS", "")
            if path and os.path.exists(path):
                self.encoder.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
            else:
                warnings.warn("SAUNet(pretrained=True): no ImageNet DenseNet-121 checkpoint available offline "
                              "(set SAUNET_DENSENET121_WEIGHTS=<state_dict.pth>); encoder stays randomly initialised")

    # ------------------------------------------------------------------
    def _body(self, tp, xb, x_nchw, return_att):
        """models/models.py:326-394 on NHWC buffers.  Concat buffers are preallocated and producers write their
        channel slice in place (no torch.cat); transitions write straight into the next dense block's buffer."""
        f = self.encoder.features
        B, H, W = xb.B, xb.H, xb.W
        # ---- encoder: conv0 + norm0 (no relu0 / pool0, :304-305) ----
        # The 3-channel image is carried as a zero-padded 4-channel NHWC buffer (and conv0's weight as a zero-padded
        # [64,4,7,7] copy) so the stem's forward and weight gradient run on the tensor-core kernels (Cin % 4 == 0).
        C0 = f.conv0.weight.shape[0]
        n0m = _check_bn(f.norm0)
        st0 = _stat(tp, n0m, C0)
        w0 = f.conv0.weight
        Cin0, KH0, KW0 = w0.shape[1], w0.shape[2], w0.shape[3]
        pad_c = (-Cin0) % 4
        if pad_c:
            # ONE persistent padded copy, refreshed IN PLACE whenever conv0.weight moved (version / optimizer-step
            # generation) and on every forward recorded into a CUDA graph (so the refresh is part of the graph and the
            # graph never reads a freed buffer)
            ent = getattr(self, "_conv0_padded", None)
            if ent is None or ent[1] != w0.data_ptr() or ent[2].device != w0.device:
                ent = [None, w0.data_ptr(), torch.zeros(C0, Cin0 + pad_c, KH0, KW0, dtype=w0.dtype, device=w0.device)]
                self._conv0_padded = ent
            gen = engine.weight_generation(w0)
            if ent[0] != gen or engine.FORCE_PACK:
                with torch.no_grad():
                    ent[2][:, :Cin0].copy_(w0)
                ent[0] = gen
            w0p = ent[2]
            xin = tp.new(B, H, W, Cin0 + pad_c)
            xin.s.t.zero_()
            _C.call("saunet_nchw_to_nhwc", x_nchw.data_ptr(), xin.ptr, xin.ld, B, Cin0, H * W, tp.stream)
        else:
            w0p, xin = w0, xb
        t0, r0 = conv2d(tp, xin, w0p, None, stride=f.conv0.stride[0], pad=f.conv0.padding[0], stat=st0)
        bn0 = bn_finalize(tp, n0m, C0, st0, t0.npix)
        stages = [(f.denseblock1, f.transition1), (f.denseblock2, f.transition2), (f.denseblock3, f.transition3),
                  (f.denseblock4, None)]
        Hc, Wc = t0.H, t0.W
        X = tp.new(B, Hc, Wc, f.denseblock1.out_features)
        first = X.slice(0, C0)
        affine_act(tp, t0, bn0.state, first, ACT_NONE)

        def bwd_stem(first=first):
            g = tp.grad(first)
            if g is None:
                return
            dt0 = tp.new(t0.B, t0.H, t0.W, C0)
            bn_backward(tp, bn0, g, t0, None, ACT_NONE, dt0, 0)
            if not w0.requires_grad:
                return
            if not pad_c:
                conv2d_bwd(tp, r0, dt0, None)
                return
            # weight gradient in the padded layout, then the real channels are added into conv0.weight's gradient
            taps, Cp = KH0 * KW0, Cin0 + pad_c
            dwp = torch.zeros(taps * Cp * C0, dtype=torch.float32, device=tp.device)
            engine.wgrad(tp, dt0, xin, dwp.data_ptr(), KH0, KW0, dt0.H, dt0.W, sy=r0.stride, sx=r0.stride, offy=-r0.pad,
                         offx=-r0.pad)
            g4 = torch.empty(C0 * Cp * taps, dtype=torch.float32, device=tp.device)
            _C.call("saunet_unpack_wgrad", dwp.data_ptr(), g4.data_ptr(), C0, Cp, KH0, KW0, 0, tp.stream)
            _C.call("saunet_copy_slice", g4.data_ptr(), Cp * taps, tp.pgrad(w0), Cin0 * taps, Cin0 * taps, C0, 1, tp.stream)
        tp.on_backward(bwd_stem)

        feats = []
        cin = C0
        for db, tr in stages:
            sums = None
            if db.training:
                sums = bn_stats_slot(tp, X.C)
                channel_stats(tp, first, sums[0], sums[1])
            dense_block_body(tp, db, X, cin, sums)
            if tr is not None:
                cout = tr.conv.weight.shape[0]
                nxt_blk = stages[len(feats) + 1][0]
                Hc, Wc = Hc // 2, Wc // 2
                Xn = tp.new(B, Hc, Wc, nxt_blk.out_features)
                first = Xn.slice(0, cout)
                transition_body(tp, tr, X, sums, first)
                feats.append(first)
                X, cin = Xn, cout
        conv2, conv3, conv4 = feats
        # norm5 (no ReLU after it, :312-313) written straight into dec5's concat buffer
        n5 = _check_bn(f.norm5)
        C5 = X.C
        cen_out = self.center[0].weight.shape[0]
        M5 = tp.new(B, Hc, Wc, C5 + cen_out)
        conv5 = M5.slice(0, C5)
        bn5 = bn_finalize(tp, n5, C5, sums if (n5.training or not n5.track_running_stats) else None, X.npix)
        affine_act(tp, X, bn5.state, conv5, ACT_NONE)
        X4 = X

        def bwd_norm5():
            g = tp.grad(conv5)
            if g is None:
                return
            gX, acc = tp.gw(X4)
            bn_backward(tp, bn5, g, X4, None, ACT_NONE, gX, acc)
        tp.on_backward(bwd_norm5)

        # The shape stream (everything at 256x256: HBM-shaped, latency-bound launches) and the decoder (tensor-bound, small
        # grids in its deep stages) are independent until dec0 reads cat[dec1, edge]: the forward issues the shape stream
        # on a side stream so the two share the GPU, and so does the backward (Tape.side_section: gradients the shape
        # stream accumulates into the shared encoder features go to private buffers folded in at the join).  D0 is
        # allocated first, on the main stream, because both write a slice of it.
        nf = self.expand[0].weight.shape[0]
        dec1_C = self.dec1.block[1].weight.shape[1]
        D0 = tp.new(B, H, W, dec1_C + nf)
        with tp.side_section(3):
            # ---- shape stream (:337-356) ----
            d0o = conv_op(tp, conv2, self.d0.weight, self.d0.bias)
            ss = bilinear(tp, d0o, tp.new(B, H, W, d0o.C))
            ss = basic_block_body(tp, self.res1, ss)
            gates = []
            for dconv, cconv, feat, gate, res in ((self.d1, self.c3, conv3, self.gate1, self.res2),
                                                  (self.d2, self.c4, conv4, self.gate2, self.res3),
                                                  (self.d3, self.c5, conv5, self.gate3, None)):
                C = dconv.weight.shape[0]
                xg = tp.new(B, H, W, C + 1, ld=_round4(C + 1))
                conv_op(tp, ss, dconv.weight, dconv.bias, y=xg.slice(0, C))
                clo = conv_op(tp, feat, cconv.weight, cconv.bias)
                bilinear(tp, clo, xg.slice(C, 1))
                ss, alphas = gsconv_body(tp, gate, xg.slice(0, C), xg.slice(C, 1))
                gates.append(alphas)
                if res is not None:
                    ss = basic_block_body(tp, res, ss)
            # fuse -> (identity resize, :355) -> sigmoid
            edge_out = conv_op(tp, ss, self.fuse.weight, self.fuse.bias, act=ACT_SIGMOID)
            # ---- Canny fusion (:358-369): on-device, non-differentiable ----
            ec = tp.new(B, H, W, 2)
            cat_copy(tp, edge_out, ec.slice(0, 1))
            canny = tp.new(B, H, W, 1)
            nbytes = _C.load().saunet_canny_workspace_bytes(B, H, W)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=tp.device)
            _C.call("saunet_canny_fwd", x_nchw.data_ptr(), B, x_nchw.shape[1], H, W, 10, 100, canny.ptr, ws.data_ptr(),
                    nbytes, tp.stream)
            engine.copy_slice(tp, canny, ec.slice(1, 1))
            acts = conv_op(tp, ec, self.cw.weight, self.cw.bias, act=ACT_SIGMOID)
            conv_bn_relu_body(tp, self.expand, acts, out=D0.slice(dec1_C, nf))
        # ---- decoder (:372-384) ----
        def decoder_chain():
            center = conv_bn_relu_body(tp, self.center, maxpool2(tp, conv5))
            dec, atts = center, []
            for blk, skip_src, M in ((self.dec5, None, M5), (self.dec4, conv4, None), (self.dec3, conv3, None),
                                     (self.dec2, conv2, None)):
                if M is None:
                    M = tp.new(B, 2 * skip_src.H, 2 * skip_src.W, skip_src.C + dec.C)
                    bilinear(tp, skip_src, M.slice(0, skip_src.C))
                skip = M.slice(0, M.C - dec.C)
                dec, att = dual_att_body(tp, blk, dec, skip, mcat=M)
                atts.append(att)
            decoder_block_body(tp, self.dec1, dec, out=D0.slice(0, dec1_C))
            return atts
        atts = decoder_chain()
        tp.join_sides()                  # (the shape stream above ran on a side stream, see below)
        dec0 = conv_bn_relu_body(tp, self.dec0, D0)
        x_out = conv_op(tp, dec0, self.final.weight, self.final.bias)
        outs = [x_out, edge_out]
        if return_att:
            # (the reference resizes the attention maps on every call and drops them; only done on request here)
            for att in reversed(atts):          # att2, att3, att4, att5
                outs.append(bilinear(tp, att, tp.new(B, H, W, 1)))
            outs.extend(gates)
        return outs

    def forward(self, x, return_att=False):
        if x.requires_grad:
            raise RuntimeError("saunet_b200 SAUNet: gradients w.r.t. the input image are not on the training path")
        xc = x.contiguous()
        outs = engine.run(self, lambda tp, xb: self._body(tp, xb, xc, return_att), [x])
        if return_att:
            return outs[0], outs[1], list(outs[2:])
        return outs[0], outs[1]
