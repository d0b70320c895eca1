#!/usr/bin/env python
"""BASELINE configs[3] / [4] companions:
  (a) DualAttBlock + GatedSpatialConv2d micro-benchmark: C in {64,128,256,512} at 128x128 output maps, batch 32 --
      whole block forward and forward+backward through the public nn.Module API, device-timed; reported against the
      algorithmic byte counts of section 8(d): DualAttBlock attention TAIL fwd = (2*C*HW + HW)*B*4 B ("1-pass") and
      GSConv fwd = (2C+2)*HW*B*4 B, backward 2x the forward figure; plus the whole-block conv FLOPs.
  (b) volume inference: a 16-slice 256x256 stack as one eval-mode batch (argmax on device), slices/s.
Run on a B200:  python tools/microbench_blocks.py [--quick]
"""
import json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "shape-attentive-unet_b200"))
import torch
from models.attention_blocks import DualAttBlock
from models.GSConv import GatedSpatialConv2d
from models import SAUNet
from saunet_b200 import synth

dev = torch.device("cuda", 0)
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
quick = "--quick" in sys.argv


def timed(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def nhwc(*shape):
    return torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)


out = {"hbm_peak_gbs": HBM, "dualatt": [], "gsconv": []}
B, S = (8 if quick else 32), 128
for C in ((64, 128) if quick else (64, 128, 256, 512)):
    blk = DualAttBlock(inchannels=[C, C], outchannels=C).to(dev).train()
    lo, skip = nhwc(B, C, S // 2, S // 2).requires_grad_(True), nhwc(B, C, S, S).requires_grad_(True)
    def fwd():
        with torch.no_grad(): blk([lo, skip])
    def fwdbwd():
        o, s = blk([lo, skip]); (o.sum() + s.sum()).backward()
    f, fb = timed(fwd), timed(fwdbwd)
    HW = S * S
    tail = (2 * C * HW + HW) * B * 4
    flops = 2.0 * B * HW * (9 * 2 * C * C + 4 * C * C + C * C // 4)        # c3x3rb + convT (4 taps/output) + spatial down
    out["dualatt"].append({"C": C, "B": B, "fwd_ms": round(f, 3), "fwdbwd_ms": round(fb, 3),
                           "tail_bytes_1pass_fwd": tail, "tail_gbs_if_all_time_were_tail": round(tail / f / 1e6, 1),
                           "block_fwd_tflops": round(flops / f / 1e9, 1), "block_fwdbwd_tflops": round(3 * flops / fb / 1e9, 1)})
    gs = GatedSpatialConv2d(C, C).to(dev).train()
    x, g = nhwc(B, C, S, S).requires_grad_(True), nhwc(B, 1, S, S).requires_grad_(True)
    def gfwd():
        with torch.no_grad(): gs(x, g)
    def gfb():
        o, a = gs(x, g); (o.sum() + a.sum()).backward()
    f, fb = timed(gfwd), timed(gfb)
    byt = (2 * C + 2) * HW * B * 4
    out["gsconv"].append({"C": C, "B": B, "fwd_ms": round(f, 3), "fwdbwd_ms": round(fb, 3), "fwd_alg_bytes": byt,
                          "fwd_gbs": round(byt / f / 1e6, 1), "fwd_frac_hbm": round(byt / f / 1e6 / HBM, 3),
                          "fwdbwd_gbs": round(3 * byt / fb / 1e6, 1), "fwdbwd_frac_hbm": round(3 * byt / fb / 1e6 / HBM, 3)})
    del blk, gs, lo, skip, x, g
    torch.cuda.empty_cache()

# the GatedSpatialConv2d instances SAUNet actually builds (models/models.py:295-297): C = 32 / 16 / 8 at 256x256
out["gsconv_model_sizes"] = []
for C in (32, 16, 8):
    Bm, Sm = 16, 256
    gs = GatedSpatialConv2d(C, C).to(dev).train()
    x, g = nhwc(Bm, C, Sm, Sm).requires_grad_(True), nhwc(Bm, 1, Sm, Sm).requires_grad_(True)
    def gfwd():
        with torch.no_grad(): gs(x, g)
    def gfb():
        o, a = gs(x, g); (o.sum() + a.sum()).backward()
    f, fb = timed(gfwd), timed(gfb)
    byt = (2 * C + 2) * Sm * Sm * Bm * 4
    out["gsconv_model_sizes"].append({"C": C, "B": Bm, "HW": Sm, "fwd_ms": round(f, 3), "fwdbwd_ms": round(fb, 3), "fwd_alg_bytes": byt,
                                      "fwd_gbs": round(byt / f / 1e6, 1), "fwd_frac_hbm": round(byt / f / 1e6 / HBM, 3),
                                      "fwdbwd_gbs": round(3 * byt / fb / 1e6, 1), "fwdbwd_frac_hbm": round(3 * byt / fb / 1e6 / HBM, 3)})
    del gs, x, g
    torch.cuda.empty_cache()

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = SAUNet(num_classes=4, pretrained=False)
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), seed=0))
m = m.to(dev).eval()
vol = synth.synthetic_batch(16, 256, seed=7)["image"].to(dev)
def infer():
    with torch.no_grad():
        seg, edge = m(vol)
        return seg.argmax(1)
ms = timed(infer, 10)
out["volume_inference"] = {"slices": 16, "ms_per_volume": round(ms, 3), "slices_per_s": round(16 / ms * 1e3, 1),
                           "note": "eval-mode BN (running statistics), one 16-slice 256x256 stack per call, argmax on device"}
print(json.dumps(out, indent=1))
