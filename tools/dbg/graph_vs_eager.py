import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "shape-attentive-unet_b200"))
import torch
import bench
from saunet_b200.graphs import GraphedStep
dev = torch.device("cuda", 0)
seg_mod, unet, arena = bench.build_ours(dev, 16)
feed = {k: v.to(dev) for k, v in bench.host_batch(16, 0).items()}
def step():
    arena.zero()
    loss, acc = seg_mod({"image": feed["image"], "mask": (feed["seg"], feed["edge"])}, 0)
    loss.backward()
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager %.2f ms" % timed(step))
g = GraphedStep(seg_mod, arena, feed)
print("graph replay only %.2f ms" % timed(lambda: g.graph.replay()))
print("graph + copies %.2f ms" % timed(lambda: g(feed)))
