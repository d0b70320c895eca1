"""3-D volume inference (BASELINE configs[4]) and loader-side data prep (row N4).

The reference (test_and_pack.py:98-137, train.py:25-64) walks a volume ONE slice at a time with a device
synchronisation, an argmax and a `.cpu()` per slice.  Here the z axis is the batch axis: each rank takes the slices
``r::world`` of the stack (slices are independent in eval mode), runs them as one batch through the CUDA path, takes the
per-pixel argmax on the device (saunet_argmax_u8, 1 byte per pixel) and the ranks exchange the uint8 label maps with one
``all_gather``.  No data-path collective other than that gather exists.

``undo_crop`` / ``resample_to_orig`` restate the packing geometry of test_and_pack.py:31-76 in numpy (the reference uses
PIL / skimage, absent here); they are host-side and outside any timed region.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _C
from .parallel import shard_batch


def argmax_u8(logits):
    """[B,C,H,W] fp32 CUDA logits (channels_last, as SAUNet returns them) -> uint8 [B,H,W] label map on the device."""
    if not logits.is_cuda:
        raise RuntimeError("saunet_b200.inference: CUDA tensors required; there is no CPU fallback")
    B, C, H, W = logits.shape
    v = logits.permute(0, 2, 3, 1)
    if not v.is_contiguous():
        v = v.contiguous()
    out = torch.empty(B, H, W, dtype=torch.uint8, device=logits.device)
    _C.call("saunet_argmax_u8", v.data_ptr(), C, C, B * H * W, out.data_ptr(),
            torch.cuda.current_stream(logits.device).cuda_stream, nbytes=(4.0 * C + 1) * B * H * W)
    return out


def edge_ground_truth(seg, radius=2, num_classes=3):
    """Edge target of the loader (data/ac17_dataloader.py:231-258) from an int64 label map [B,H,W] on the device ->
    fp32 [B,1,H,W]; replaces two scipy distance transforms per class per slice on the host."""
    if not seg.is_cuda:
        raise RuntimeError("saunet_b200.inference: CUDA tensors required; there is no CPU fallback")
    seg = seg.to(torch.int64).contiguous()
    B, H, W = seg.shape
    out = torch.empty(B, 1, H, W, dtype=torch.float32, device=seg.device)
    _C.call("saunet_edge_gt", seg.data_ptr(), B, H, W, radius, num_classes, out.data_ptr(),
            torch.cuda.current_stream(seg.device).cuda_stream, nbytes=12.0 * B * H * W)
    return out


def shard_slices(n_slices, rank, world):
    """Slices of the z stack rank ``rank`` runs (r, r+world, ...) and the interleave order that undoes the gather."""
    return shard_batch(n_slices, rank, world)


def merge_gathered(parts, n_slices, world):
    """parts[r] = label maps of the slices r::world (padded to the longest shard) -> [n_slices, H, W] in z order."""
    out = torch.empty((n_slices,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype, device=parts[0].device)
    for r in range(world):
        idx = shard_batch(n_slices, r, world)
        out[idx] = parts[r][:len(idx)]
    return out


@torch.no_grad()
def predict_volume(unet, volume, group=None, max_batch=32):
    """volume: [Z, C, H, W] (or [C, H, W, Z] as the reference loader yields, test_and_pack.py:105-109) fp32 on this
    rank's device -> uint8 label volume [Z, H, W], identical on every rank.  ``unet`` must be in eval mode."""
    if unet.training:
        raise RuntimeError("predict_volume: put the model in eval mode (BatchNorm running statistics) first")
    if volume.dim() == 4 and volume.shape[0] in (1, 3) and volume.shape[-1] not in (1, 3) and volume.shape[1] == volume.shape[2]:
        volume = volume.permute(3, 0, 1, 2)             # [C,H,W,Z] -> [Z,C,H,W]
    Z = volume.shape[0]
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    mine = shard_slices(Z, rank, world)
    x = volume[mine].contiguous()
    labels = []
    for i in range(0, x.shape[0], max_batch):
        seg, _edge = unet(x[i:i + max_batch])
        labels.append(argmax_u8(seg))
    lab = torch.cat(labels) if labels else torch.empty((0,) + tuple(volume.shape[2:]), dtype=torch.uint8, device=volume.device)
    if world == 1:
        return lab
    longest = (Z + world - 1) // world
    pad = torch.zeros((longest,) + tuple(lab.shape[1:]), dtype=torch.uint8, device=lab.device)
    pad[:lab.shape[0]] = lab
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return merge_gathered(parts, Z, world)


def dice_per_class(pred, ref, num_classes=4):
    """Dice of two label volumes per class 1..num_classes-1 (numpy / torch, host or device)."""
    pred, ref = torch.as_tensor(pred), torch.as_tensor(ref)
    out = []
    for c in range(1, num_classes):
        a, b = pred == c, ref == c
        den = int(a.sum()) + int(b.sum())
        out.append(2.0 * int((a & b).sum()) / den if den else 1.0)
    return out


# ---- packing geometry (host, numpy): test_and_pack.py:28-76 ----------------------------------------------------
def _round_num(x):
    return int(x) + 1 if (x - int(x)) >= 0.5 else int(x)


def undo_crop(orig_hw, pred):
    """Inverse of the loader's PaddingCenterCrop for one slice (test_and_pack.py:31-61): ``pred`` [th, tw] is pasted
    back into a zero canvas of the original slice size ``orig_hw`` = (h, w) (centre crop undone by zero padding, zero
    padding undone by centre cropping)."""
    h, w = orig_hw
    th, tw = pred.shape
    if w >= tw and h >= th:
        x1, y1 = _round_num((w - tw) / 2.0), _round_num((h - th) / 2.0)
        out = np.zeros((h, w), dtype=pred.dtype)
        # PIL ImageOps.expand(border=(left, top, right, bottom)) with right = x1 - rem_x, bottom = y1 - rem_y
        out[y1:y1 + th, x1:x1 + tw] = pred
        return out[:y1 + th + (y1 - (h - th) % 2), :x1 + tw + (x1 - (w - tw) % 2)]
    pad_h, pad_w = max(th - h, 0), max(tw - w, 0)
    b = [pad_w // 2, pad_h // 2, pad_w // 2 + w, pad_h // 2 + h]
    if pad_w == 0:
        b[2] = tw
    if pad_h == 0:
        b[3] = th
    crop = pred[b[1]:b[3], b[0]:b[2]]
    x1, y1 = max(_round_num((w - tw) / 2.0), 0), max(_round_num((h - th) / 2.0), 0)
    rem_w = (w - tw) % 2 if (w - tw) >= 0 else 0
    rem_h = (h - th) % 2 if (h - th) >= 0 else 0
    ch, cw = crop.shape
    out = np.zeros((y1 + ch + y1 - rem_h, x1 + cw + x1 - rem_w), dtype=pred.dtype)
    out[y1:y1 + ch, x1:x1 + cw] = crop
    return out


def resample_nearest(vol, shape):
    """Order-0 resize of a label volume to ``shape`` (test_and_pack.py:69-73: skimage.transform.resize(order=0,
    mode='constant', preserve_range=True)): output voxel centres mapped back to input voxel centres, nearest voxel."""
    idx = []
    for n_in, n_out in zip(vol.shape, shape):
        c = (np.arange(n_out) + 0.5) * (n_in / float(n_out)) - 0.5
        idx.append(np.clip(np.rint(c).astype(np.int64), 0, n_in - 1))
    return vol[np.ix_(*idx)]
