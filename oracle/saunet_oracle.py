"""CPU oracle: functional restatement of the reference SAUNet hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line it restates; all paths are relative to /root/reference.
The arithmetic is floating point, so the restatement is written with plain
torch CPU functional ops in fp32 (or fp64 when ``dtype=torch.float64``); the
integer Canny step is restated in C (oracle/canny_oracle.c).

The oracle is driven by a *state_dict with the reference's key names* -- it
holds no modules and no parameters of its own.  The encoder arithmetic lives in
a third-party dependency of the reference (torchvision ``densenet121``,
requirements.txt:2 unpinned; 0.26.0 in the build image): its published
algorithm (torchvision/models/densenet.py:31-133,160-210) is restated in
``dense_layer`` / ``dense_block`` / ``transition`` below.

Pinning: tests/golden/make_golden.py runs the REAL reference (imported from
/root/reference in the build container) on the same seeded weights/inputs and
commits its outputs; tests/test_oracle_vs_golden.py checks this file against
them.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


class BNRecorder:
    """Collects the running-stat updates a train-mode forward would make
    (Appendix A of SURVEY.md; torch.nn.functional.batch_norm semantics)."""

    def __init__(self):
        self.updates = {}


def batch_norm(sd, prefix, x, training, momentum=0.1, rec=None):
    """nn.BatchNorm2d forward (train: batch stats, biased var for normalisation,
    unbiased for the running buffer; eval: running stats).  Used by every
    BatchNorm on the path; momentum 0.001 for the six BasicBlock norms
    (lib/nn/modules/batchnorm.py:39,58-61)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if not training:
        return F.batch_norm(x, rm.to(x.dtype), rv.to(x.dtype), w, b, False, momentum, BN_EPS)
    rm2, rv2 = rm.detach().clone().to(x.dtype), rv.detach().clone().to(x.dtype)
    y = F.batch_norm(x, rm2, rv2, w, b, True, momentum, BN_EPS)
    if rec is not None:
        rec.updates[prefix + ".running_mean"] = rm2
        rec.updates[prefix + ".running_var"] = rv2
    return y


def conv(sd, prefix, x, stride=1, padding=0):
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride, padding)


def conv_t(sd, prefix, x):
    """nn.ConvTranspose2d(k=4, s=2, p=1) (attention_blocks.py:179-183, models.py:211)."""
    return F.conv_transpose2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=2, padding=1)


def interp(x, size=None, scale=None):
    """F.interpolate(mode='bilinear', align_corners=True) (models/models.py:337-389)."""
    return F.interpolate(x, size=size, scale_factor=scale, mode="bilinear", align_corners=True)


# ---------------------------------------------------------------- encoder
def dense_layer(sd, p, feats, training, rec):
    """torchvision densenet.py:31-93 (_DenseLayer; drop_rate 0)."""
    x = torch.cat(feats, 1)
    x = conv(sd, p + ".conv1", F.relu(batch_norm(sd, p + ".norm1", x, training, rec=rec)))
    x = conv(sd, p + ".conv2", F.relu(batch_norm(sd, p + ".norm2", x, training, rec=rec)), padding=1)
    return x


def dense_block(sd, p, x, n_layers, training, rec):
    """torchvision densenet.py:96-124 (_DenseBlock)."""
    feats = [x]
    for i in range(n_layers):
        feats.append(dense_layer(sd, "%s.denselayer%d" % (p, i + 1), feats, training, rec))
    return torch.cat(feats, 1)


def transition(sd, p, x, training, rec):
    """torchvision densenet.py:127-133 (_Transition): BN-ReLU-1x1-AvgPool2."""
    x = conv(sd, p + ".conv", F.relu(batch_norm(sd, p + ".norm", x, training, rec=rec)))
    return F.avg_pool2d(x, 2, 2)


# ---------------------------------------------------------------- shape stream
def basic_block(sd, p, x, training, rec):
    """models/resnet.py:30-59 (BasicBlock, stride 1, no downsample);
    SynchronizedBatchNorm2d outside DataParallel == F.batch_norm with
    momentum 0.001 (lib/nn/modules/batchnorm.py:39,58-61)."""
    out = conv(sd, p + ".conv1", x, padding=1)
    out = F.relu(batch_norm(sd, p + ".bn1", out, training, 0.001, rec))
    out = conv(sd, p + ".conv2", out, padding=1)
    out = batch_norm(sd, p + ".bn2", out, training, 0.001, rec)
    return F.relu(out + x)


def gated_spatial_conv(sd, p, x, g, training, rec):
    """models/GSConv.py:38-57 (GatedSpatialConv2d)."""
    a = torch.cat([x, g], 1)
    a = batch_norm(sd, p + "._gate_conv.0", a, training, rec=rec)
    a = F.relu(conv(sd, p + "._gate_conv.1", a))
    a = conv(sd, p + "._gate_conv.3", a)
    alphas = torch.sigmoid(batch_norm(sd, p + "._gate_conv.4", a, training, rec=rec))
    out = F.conv2d(x * (alphas + 1), sd[p + ".weight"], sd.get(p + ".bias"))
    return out, alphas


# ---------------------------------------------------------------- decoder
def conv3x3_bn_relu(sd, p, x, training, rec):
    """models/models.py:118-123."""
    return F.relu(batch_norm(sd, p + ".1", conv(sd, p + ".0", x, padding=1), training, rec=rec))


def dual_att_block(sd, p, lo, skip, training, rec):
    """models/attention_blocks.py:208-238 (DualAttBlock) with _MRF (:175-206),
    SpatialAttentionBlock (:145-173) and SEModule (:28-57)."""
    up = F.relu(batch_norm(sd, p + ".mrf.up.1", conv_t(sd, p + ".mrf.up.0", lo), training, rec=rec))
    fused = torch.cat([skip, up], 1)
    fused = conv3x3_bn_relu(sd, p + ".c3x3rb", fused, training, rec)
    c = conv(sd, p + ".spatialAttn.down", fused)
    c = conv(sd, p + ".spatialAttn.phi", F.relu(batch_norm(sd, p + ".spatialAttn.bn", c, training, rec=rec)))
    spatial = torch.sigmoid(c)
    s = F.adaptive_avg_pool2d(fused, 1)
    s = F.relu(conv(sd, p + ".channelAttn.fc1", s))
    s = torch.sigmoid(conv(sd, p + ".channelAttn.fc2", s))
    channel = fused * s
    return (spatial.expand_as(channel) + 1) * channel, spatial


def decoder_block(sd, p, x, training, rec):
    """models/models.py:203-237 (DecoderBlock, is_deconv=True)."""
    x = conv3x3_bn_relu(sd, p + ".block.0", x, training, rec)
    x = conv_t(sd, p + ".block.1", x)
    return F.relu(batch_norm(sd, p + ".block.2", x, training, rec=rec))


def image_to_u8(x):
    """models/models.py:359: np.mean(x.cpu().numpy(), axis=1).astype(np.uint8).
    float32 sequential sum over the channel axis / 3, then the x86 float->uint8
    cast (truncate toward zero to int32, keep the low 8 bits)."""
    import numpy as np
    a = x.detach().to(torch.float32).cpu().numpy()
    m = a[:, 0].copy()
    for c in range(1, a.shape[1]):
        m = m + a[:, c]
    m = (m / np.float32(a.shape[1])).astype(np.float32)
    return (np.trunc(m).astype(np.int64) & 0xFF).astype(np.uint8)


def canny_map(x):
    """models/models.py:358-364: per-sample cv2.Canny(im, 10, 100) -> float
    [B,1,H,W] in {0,255}.  Uses the C restatement (oracle/canny_oracle.c)."""
    from . import canny as _canny
    import numpy as np
    im = image_to_u8(x)
    out = np.stack([_canny.canny_u8(im[i], 10, 100) for i in range(im.shape[0])])
    return torch.from_numpy(out[:, None].astype(np.float32))


def saunet_forward(sd, x, training=True, canny=None, return_att=False, rec=None):
    """models/models.py:326-394 (SAUNet.forward)."""
    E = "encoder.features."
    size = x.shape[2:]
    # encoder (models.py:330-334); conv1 = conv0 + norm0, no relu0/pool0 (:304-305)
    conv1 = batch_norm(sd, E + "norm0", conv(sd, E + "conv0", x, stride=2, padding=3), training, rec=rec)
    conv2 = transition(sd, E + "transition1", dense_block(sd, E + "denseblock1", conv1, 6, training, rec), training, rec)
    conv3 = transition(sd, E + "transition2", dense_block(sd, E + "denseblock2", conv2, 12, training, rec), training, rec)
    conv4 = transition(sd, E + "transition3", dense_block(sd, E + "denseblock3", conv3, 24, training, rec), training, rec)
    conv5 = batch_norm(sd, E + "norm5", dense_block(sd, E + "denseblock4", conv4, 16, training, rec), training, rec=rec)
    # shape stream (models.py:337-356)
    ss = interp(conv(sd, "d0", conv2), size=size)
    ss = basic_block(sd, "res1", ss, training, rec)
    c3 = interp(conv(sd, "c3", conv3), size=size)
    ss = conv(sd, "d1", ss)
    ss, g1 = gated_spatial_conv(sd, "gate1", ss, c3, training, rec)
    ss = basic_block(sd, "res2", ss, training, rec)
    ss = conv(sd, "d2", ss)
    c4 = interp(conv(sd, "c4", conv4), size=size)
    ss, g2 = gated_spatial_conv(sd, "gate2", ss, c4, training, rec)
    ss = basic_block(sd, "res3", ss, training, rec)
    ss = conv(sd, "d3", ss)
    c5 = interp(conv(sd, "c5", conv5), size=size)
    ss, g3 = gated_spatial_conv(sd, "gate3", ss, c5, training, rec)
    ss = conv(sd, "fuse", ss)
    ss = interp(ss, size=size)
    edge_out = torch.sigmoid(ss)
    # canny fusion (models.py:358-369); non-differentiable
    if canny is None:
        canny = canny_map(x)
    canny = canny.to(x.dtype)
    acts = torch.sigmoid(conv(sd, "cw", torch.cat([edge_out, canny], 1)))
    edge = F.relu(batch_norm(sd, "expand.1", conv(sd, "expand.0", acts), training, rec=rec))
    # decoder (models.py:372-384)
    conv2u = interp(conv2, scale=2)
    conv3u = interp(conv3, scale=2)
    conv4u = interp(conv4, scale=2)
    center = conv3x3_bn_relu(sd, "center", F.max_pool2d(conv5, 2, 2), training, rec)
    dec5, att5 = dual_att_block(sd, "dec5", center, conv5, training, rec)
    dec4, att4 = dual_att_block(sd, "dec4", dec5, conv4u, training, rec)
    dec3, att3 = dual_att_block(sd, "dec3", dec4, conv3u, training, rec)
    dec2, att2 = dual_att_block(sd, "dec2", dec3, conv2u, training, rec)
    dec1 = decoder_block(sd, "dec1", dec2, training, rec)
    dec0 = conv3x3_bn_relu(sd, "dec0", torch.cat([dec1, edge], 1), training, rec)
    x_out = conv(sd, "final", dec0)
    if return_att:
        maps = [interp(att2, scale=2), interp(att3, scale=4), interp(att4, scale=8),
                interp(att5, scale=16), g1, g2, g3]
        return x_out, edge_out, maps
    return x_out, edge_out


# ---------------------------------------------------------------- loss
CE_WEIGHTS = (1.0, 4.0, 5.0, 1.0)


def dice_loss(true, logits, eps=1e-7):
    """loss.py:51-88, multi-class branch (:79-88)."""
    c = logits.shape[1]
    onehot = F.one_hot(true.long(), c).permute(0, 3, 1, 2).to(logits.dtype)
    probas = F.softmax(logits, dim=1)
    inter = torch.sum(probas * onehot, (0, 2, 3))
    card = torch.sum(probas + onehot, (0, 2, 3))
    return 1 - (2.0 * inter / (card + eps)).mean()


def dual_loss(seg, edge_in, seg_t, edge_t, parts=False):
    """loss.py:149-159 (DualLoss.forward): dice + weighted CE + edge BCE."""
    w = torch.tensor(CE_WEIGHTS, dtype=seg.dtype, device=seg.device)[: seg.shape[1]]
    ce = F.cross_entropy(seg, seg_t.long(), weight=w)
    dice = dice_loss(seg_t, seg)
    bce = F.binary_cross_entropy(edge_in, edge_t.to(edge_in.dtype))
    total = dice + ce + bce
    return (total, dice, ce, bce) if parts else total


# ---------------------------------------------------------------- drivers
def prepare_params(state_dict, dtype=torch.float32, requires_grad=False):
    """Detached (optionally leaf, grad-requiring) copies of the float tensors."""
    sd = {}
    for k, v in state_dict.items():
        if v.is_floating_point():
            t = v.detach().clone().to(dtype)
            if requires_grad and not k.rsplit(".", 1)[-1].startswith(("running_", "_tmp", "_running")):
                t.requires_grad_(True)
            sd[k] = t
        else:
            sd[k] = v.detach().clone()
    return sd


def train_step(state_dict, image, seg_t, edge_t, canny=None, dtype=torch.float32, training=True, return_att=False):
    """fwd + DualLoss + bwd (train.py:95-104).  Returns dict with logits, edge,
    loss parts, grads (by state_dict key, canonical ``encoder.features.*`` names
    only) and BN running-stat updates."""
    sd = prepare_params(state_dict, dtype, requires_grad=True)
    # the reference aliases encoder tensors under two names; drive the oracle by
    # the canonical name only so each parameter has one leaf.
    rec = BNRecorder()
    x = image.to(dtype)
    maps = None
    if return_att:
        seg, edge, maps = saunet_forward(sd, x, training=training, canny=canny, return_att=True, rec=rec)
    else:
        seg, edge = saunet_forward(sd, x, training=training, canny=canny, rec=rec)
    total, dice, ce, bce = dual_loss(seg, edge, seg_t, edge_t, parts=True)
    total.backward()
    grads = {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.requires_grad and v.grad is not None}
    out = {"logits": seg.detach(), "edge": edge.detach(), "loss": total.detach(), "dice": dice.detach(),
           "ce": ce.detach(), "bce": bce.detach(), "grads": grads, "bn_updates": rec.updates}
    if maps is not None:
        out["maps"] = [t.detach() for t in maps]
    return out


def pixel_metrics(logits, label, num_class):
    """SegmentationModule's training-branch metrics, models/models.py:51-74 applied as in :92 to
    round(softmax(logits)): (pixel accuracy over label >= 1, [Jaccard of class 1 .. num_class-1])."""
    pred = torch.round(F.softmax(logits, dim=1)).long()
    _, preds = torch.max(pred, dim=1)
    valid = (label >= 1).long()
    acc = torch.sum(valid * (preds == label).long()).float() / (torch.sum(valid).float() + 1e-10)
    jac = []
    for i in range(1, num_class):
        v = (label == i).long()
        p = (preds == i).long()
        anb = torch.sum(v * p)
        j = anb.float() / (torch.sum(v).float() + torch.sum(p).float() - anb.float() + 1e-10)
        jac.append(j if j <= 1 else torch.zeros(()))
    return acc, jac
