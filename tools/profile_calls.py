#!/usr/bin/env python
"""Per-C-ABI-call timing of one training step (CUDA events around every call): which geometry is slow."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "shape-attentive-unet_b200"))
import torch
import bench
from saunet_b200 import _C

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
seg_mod, unet, arena = bench.build_ours(dev, B)
hb = bench.host_batch(B, 0)
feed = {k: v.to(dev) for k, v in hb.items()}
def step():
    arena.zero()
    loss, acc = seg_mod({"image": feed["image"], "mask": (feed["seg"], feed["edge"])}, 0)
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
_C.PROFILE = []
step(); torch.cuda.synchronize()
prof, _C.PROFILE = _C.PROFILE, None
agg = collections.OrderedDict()
for name, a, b, fl, nb, tag, _k in prof:
    e = agg.setdefault((name, tag), [0.0, 0, 0.0, 0.0]); e[0] += a.elapsed_time(b); e[1] += 1; e[2] += fl; e[3] += nb
tot = sum(e[0] for e in agg.values())
print("total %.2f ms over %d calls" % (tot, len(prof)))
byname = collections.Counter()
for (name, tag), e in agg.items():
    byname[name] += e[0]
print("  ".join("%s %.1f" % (k.replace("saunet_", ""), v) for k, v in byname.most_common(12)))
FILTER = os.environ.get("PROFILE_FILTER", "")
for (name, tag), e in [kv for kv in sorted(agg.items(), key=lambda kv: -kv[1][0]) if FILTER in kv[0][0]][:int(os.environ.get("PROFILE_TOP", "45"))]:
    print("%7.3f ms %4d x  %6.1f TF/s %7.1f GB/s  %-24s %s" % (e[0], e[1], e[2] / e[0] / 1e9 if e[0] else 0, e[3] / e[0] / 1e6 if e[0] else 0, name.replace("saunet_", ""), tag))
