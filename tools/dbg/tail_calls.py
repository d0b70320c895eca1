import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "shape-attentive-unet_b200"))
import torch
from models.attention_blocks import DualAttBlock
from models.GSConv import GatedSpatialConv2d
from saunet_b200 import _C
dev = torch.device("cuda", 0)
C = int(sys.argv[1]); B, S = 32, 128
def nhwc(*s): return torch.randn(*s, device=dev).contiguous(memory_format=torch.channels_last)
blk = DualAttBlock(inchannels=[C, C], outchannels=C).to(dev).train()
lo, skip = nhwc(B, C, S // 2, S // 2).requires_grad_(True), nhwc(B, C, S, S).requires_grad_(True)
gs = GatedSpatialConv2d(C, C).to(dev).train()
x, g = nhwc(B, C, S, S).requires_grad_(True), nhwc(B, 1, S, S).requires_grad_(True)
def run_blk():
    o, sp = blk([lo, skip]); (o.sum() + sp.sum()).backward()
def run_gs():
    o, a = gs(x, g); (o.sum() + a.sum()).backward()
for name, fn in (("dualatt", run_blk), ("gsconv", run_gs)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    _C.PROFILE = []
    fn(); torch.cuda.synchronize()
    prof, _C.PROFILE = _C.PROFILE, None
    print("==", name, "C", C)
    for n, a, b, fl, nb, tag, k in prof:
        if name == "gsconv" or tag.startswith("tail:"):
            print("  %.3f ms  %-28s %-28s %s" % (a.elapsed_time(b), n.replace("saunet_", ""), k, tag))
