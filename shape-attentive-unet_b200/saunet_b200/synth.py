"""Synthetic ACDC-shaped inputs and deterministic random-init weights.

There is no network, so neither the ACDC dataset nor the ImageNet DenseNet-121
checkpoint is available.  This module produces

* slices shaped like the reference loader's output
  (data/ac17_dataloader.py:146-148 per-slice z-score, :208-221 grayscale
  replicated to 3 channels, :236-258 radius-2 edge band via two EDTs), and
* a state_dict whose VALUES depend only on (key name, shape, seed) -- not on
  module construction order or the global torch RNG -- so the reference model
  (in the build container) and this repo's model (on the GPU box) can be
  loaded with bit-identical weights without shipping 130 MB of tensors.

Host-side only (numpy / torch CPU); nothing here is on the timed path.
"""
import zlib

import numpy as np
import torch


def synthetic_batch(batch, size=256, seed=304, num_classes=4):
    """-> dict(image [B,3,S,S] f32, seg [B,S,S] i64, edge [B,1,S,S] f32)."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(batch, 1, size, size, generator=g)
    # add low-frequency structure so Canny / the shape stream see real edges
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    seg = torch.zeros(batch, size, size, dtype=torch.int64)
    for b in range(batch):
        r = torch.rand(6, generator=g)
        cy = size * (0.35 + 0.3 * float(r[0]))
        cx = size * (0.35 + 0.3 * float(r[1]))
        r3 = size * (0.10 + 0.08 * float(r[2]))          # class 3 (outer disc)
        r2 = r3 * (0.45 + 0.25 * float(r[3]))            # class 2 nested inside 3
        r1 = size * (0.06 + 0.05 * float(r[4]))          # class 1 adjacent disc
        d = torch.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
        d1 = torch.sqrt((yy - cy) ** 2 + (xx - (cx + r3 + r1 * 0.8)) ** 2)
        s = seg[b]
        s[d1 < r1] = 1
        s[d < r3] = 3
        s[d < r2] = 2
        if num_classes < 4:
            s.clamp_(max=num_classes - 1)
        x1[b, 0] += 1.5 * (s > 0).float() + 0.8 * (s == 2).float()
    # per-slice z-score (data/ac17_dataloader.py:146-148)
    m = x1.mean(dim=(1, 2, 3), keepdim=True)
    sd = x1.std(dim=(1, 2, 3), keepdim=True)
    x1 = (x1 - m) / sd
    image = x1.repeat(1, 3, 1, 1).contiguous()
    edge = torch.from_numpy(np.stack([mask_to_edges(seg[b].numpy()) for b in range(batch)])).float()
    return {"image": image, "seg": seg, "edge": edge}


def mask_to_edges(mask, radius=2, num_classes=3):
    """Restates data/ac17_dataloader.py:231-258: union over classes 1..3 of
    pixels whose distance to the class boundary is in (0, radius]."""
    from scipy.ndimage import distance_transform_edt
    onehot = np.stack([(mask == i) for i in range(1, num_classes + 1)]).astype(np.float64)
    pad = np.pad(onehot, ((0, 0), (1, 1), (1, 1)), mode="constant", constant_values=0)
    edgemap = np.zeros(mask.shape, dtype=np.float64)
    for i in range(num_classes):
        dist = distance_transform_edt(pad[i]) + distance_transform_edt(1.0 - pad[i])
        dist = dist[1:-1, 1:-1]
        dist[dist > radius] = 0
        edgemap += dist
    return (edgemap > 0).astype(np.uint8)[None]


def _gen_for(key, seed):
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synthetic_state_dict(template_state_dict, seed=0):
    """Deterministic values for every entry of ``template_state_dict``.

    Entries that alias the same storage (the reference exposes every encoder
    tensor twice: ``encoder.features.*`` and ``conv1.0.* / conv2.* ...``,
    models/models.py:304-313) get the value generated for the FIRST key.
    """
    out, seen = {}, {}
    for key, t in template_state_dict.items():
        ptr = (t.data_ptr(), tuple(t.shape)) if t.numel() > 0 else (key, ())
        if ptr in seen:
            out[key] = out[seen[ptr]]
            continue
        seen[ptr] = key
        g = _gen_for(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        shape = tuple(t.shape)
        if leaf == "num_batches_tracked":
            v = torch.zeros(shape, dtype=t.dtype)
        elif leaf == "_running_iter":
            v = torch.ones(shape, dtype=t.dtype)
        elif leaf in ("running_mean", "_tmp_running_mean"):
            v = 0.1 * torch.randn(shape, generator=g)
        elif leaf in ("running_var", "_tmp_running_var"):
            v = 0.5 + torch.rand(shape, generator=g)
        elif t.dim() >= 2:      # conv / conv-transpose / linear weights
            fan_in = int(np.prod(shape[1:]))
            if t.dim() == 4 and ".up.0." in key or key.endswith("block.1.weight"):
                # ConvTranspose2d weight is [C_in, C_out, 4, 4]; each output pixel sees 4 taps
                fan_in = shape[0] * 4
            v = torch.randn(shape, generator=g) * float(np.sqrt(2.0 / max(fan_in, 1)))
        elif leaf == "weight":  # norm scale
            v = 0.5 + torch.rand(shape, generator=g)
        else:                   # biases (conv and norm)
            v = 0.1 * torch.randn(shape, generator=g)
        out[key] = v.to(t.dtype)
    return out
