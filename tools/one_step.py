#!/usr/bin/env python
"""Warm up, then run ONE bench step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "shape-attentive-unet_b200"))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
seg_mod, unet, arena = bench.build_ours(dev, B)
feed = {k: v.to(dev) for k, v in bench.host_batch(B, 0).items()}
def step():
    arena.zero()
    loss, acc = seg_mod({"image": feed["image"], "mask": (feed["seg"], feed["edge"])}, 0)
    loss.backward()
for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
