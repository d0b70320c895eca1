"""Shared helpers for the parity tests (host side only)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def key_contract():
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as f:
        return json.load(f)


def template_state_dict():
    """Zero tensors with the reference's state_dict keys/shapes (incl. aliasing:
    an alias pair is represented by the SAME tensor object)."""
    contract = key_contract()
    alias = alias_map()
    sd = {}
    for e in contract["keys"]:
        k = e["key"]
        canon = alias.get(k, k)
        if canon in sd:
            sd[k] = sd[canon]
        else:
            sd[k] = torch.zeros(e["shape"], dtype=getattr(torch, e["dtype"]))
    return sd


_ALIAS_PREFIX = [("conv1.0.", "encoder.features.conv0."), ("conv1.1.", "encoder.features.norm0."),
                 ("conv2.", "encoder.features.denseblock1."), ("conv2t.", "encoder.features.transition1."),
                 ("conv3.", "encoder.features.denseblock2."), ("conv3t.", "encoder.features.transition2."),
                 ("conv4.", "encoder.features.denseblock3."), ("conv4t.", "encoder.features.transition3."),
                 ("conv5.0.", "encoder.features.denseblock4."), ("conv5.1.", "encoder.features.norm5.")]


def alias_map():
    """alias key -> canonical encoder.features.* key (models/models.py:304-313)."""
    out = {}
    for e in key_contract()["keys"]:
        k = e["key"]
        for a, c in _ALIAS_PREFIX:
            if k.startswith(a):
                out[k] = c + k[len(a):]
    return out


def rel_err(a, b):
    """max |a-b| / max |b|  (SURVEY.md section 8c normalised criterion)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def median_rel(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float(((a - b).abs() / b.abs().clamp_min(1e-12)).median())
