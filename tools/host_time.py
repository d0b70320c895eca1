#!/usr/bin/env python
"""Host enqueue time vs device time per step (is the step host-bound?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "shape-attentive-unet_b200"))
import torch, bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
seg_mod, unet, arena = bench.build_ours(dev, B)
feed = {k: v.to(dev) for k, v in bench.host_batch(B, 0).items()}
def step():
    arena.zero()
    loss, acc = seg_mod({"image": feed["image"], "mask": (feed["seg"], feed["edge"])}, 0)
    loss.backward()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("host enqueue %.1f ms, total %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
