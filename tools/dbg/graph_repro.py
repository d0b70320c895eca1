import os, sys, warnings, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "shape-attentive-unet_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import template_state_dict
from saunet_b200 import synth
from models import SAUNet, SegmentationModule
from loss import DualLoss
from saunet_b200.graphs import GraphedStep
from saunet_b200.parallel import GradArena
DEV = "cuda:0"
def model():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = SAUNet(num_classes=4, pretrained=False)
    m.load_state_dict(synth.synthetic_state_dict(template_state_dict(), seed=0))
    return m.to(DEV).train()
class DataSGD(torch.optim.Optimizer):
    def __init__(self, params, lr): super().__init__(params, dict(lr=lr))
    def step(self, closure=None):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None: p.data.copy_(p.data - g["lr"] * p.grad.data)
def run(variant):
    unet = model()
    seg_mod = SegmentationModule(DualLoss(), unet, 4).to(DEV).train()
    arena = GradArena(unet)
    d = {k: v.to(DEV) for k, v in synth.synthetic_batch(2, 64, seed=304).items()}
    feed = {"image": d["image"], "mask": (d["seg"], d["edge"])}
    if "eager" in variant:
        for _ in range(2):
            if "zerograd" in variant: seg_mod.zero_grad()
            else: arena.zero()
            loss, _ = seg_mod(feed, 0); loss.backward()
    if "opt" in variant:
        DataSGD(unet.parameters(), 0.05).step()
    if "nograd" in variant:
        with torch.no_grad(): unet(d["image"])
    try:
        g = GraphedStep(seg_mod, arena, d)
        print(variant, "OK", float(g(d)))
    except Exception as e:
        print(variant, "FAIL", repr(e)[:200])
        torch.cuda.synchronize()
if os.environ.get("DBG_CAPTURE"):
    from saunet_b200 import _C, engine
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    orig = _C.call
    state = {"bad": False, "n": 0}
    def status():
        st = ctypes.c_int(0)
        err = rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
        return err, st.value
    def call(name, *a, **k):
        before = status()
        orig(name, *a, **k)
        after = status()
        state["n"] += 1
        if not state["bad"] and (after[0] != 0 or after[1] == 2 or before[0] != 0 or before[1] == 2):
            state["bad"] = True
            print("FIRST BAD at call", state["n"], name, "before", before, "after", after, "tag", k.get("tag"))
            traceback.print_stack(limit=12)
    _C.call = call
    engine._C.call = call
for v in sys.argv[1:]:
    run(v)
