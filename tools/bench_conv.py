#!/usr/bin/env python
"""Micro-benchmark of one conv geometry through the C-ABI (for ncu captures):
   python tools/bench_conv.py fwd|wgrad B H W Cin Cout k [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shape-attentive-unet_b200"))
import torch
from saunet_b200 import engine
from saunet_b200.engine import Tape, conv, wgrad, packed, packed_tc

kind = sys.argv[1]
B, H, W, Cin, Cout, k = map(int, sys.argv[2:8])
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 5
engine.set_precision(os.environ.get("SAUNET_PRECISION", "3xtf32"))
dev = torch.device("cuda", 0)
tp = Tape(dev, False)
x = tp.new(B, H, W, Cin); x.s.t.normal_()
y = tp.new(B, H, W, Cout); y.s.t.normal_()
w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
st = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
dw = torch.zeros(k * k * Cin * Cout, device=dev)
pro_state = torch.cat([0.5 + torch.rand(Cin), 0.3 * torch.randn(Cin)]).to(dev) if os.environ.get('PRO') else None
prof = None
if os.environ.get("PROF") and kind == "fwd":
    # (the library reads SAUNET_TC_PROF once per process, at its first launch: set it before anything runs)
    prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    os.environ["SAUNET_TC_PROF"] = str(prof.data_ptr())
def run():
    if kind == "fwd":
        conv(tp, x, packed(tp, w, 0), Cout, k, k, y, H, W, offy=-(k // 2), offx=-(k // 2),
             stat=None if os.environ.get('NOSTAT') else (st.data_ptr(), st.data_ptr() + 8 * Cout), wtc=packed_tc(tp, w, 0, k * k, Cin, Cout, M=B * H * W, cm=engine._wants_cm(x, k, k, 1, k // 2)),
             pro=pro_state.data_ptr() if pro_state is not None else 0, pro_relu=1)
    else:
        wgrad(tp, y, x, dw.data_ptr(), k, k, H, W, offy=-(k // 2), offx=-(k // 2))
for _ in range(2): run()
torch.cuda.synchronize()
# time a CUDA graph of `iters` launches: the Python/ctypes launch path (~50-100 us per call) must not be what is measured
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    tp.stream = side.cuda_stream
    run(); side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(iters): run()
    graph.replay(); side.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    graph.replay()
    e1.record(side); side.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = 2.0 * B * H * W * k * k * Cin * Cout
if os.environ.get("PROF") and kind == "fwd":
    prof.zero_()
    tp.stream = torch.cuda.current_stream().cuda_stream
    run(); torch.cuda.synchronize()
    pr = prof.view(148, 16).double().cpu()
    pr = pr[pr[:, 0] > 0]
    names_tma = ["total", "tma:wait_patch_empty", "xform:wait_raw_full", "xform:work", "mma:wait_tmem_empty", "mma:wait_patch_full",
                 "mma:wait_b_full", "mma:issue+commit", "load:wait_b_empty", "epi:wait_tmem_full", "epi:work"]
    names = ["total", "prod:wait_empty", "prod:transform(+load latency)", "prod:fence+arrive", "mma:wait_tmem_empty", "mma:wait_full_a",
             "mma:wait_full_b", "mma:issue+commit", "epi:wait_tmem_full", "epi:work", "epi:tmem_ld", "epi:alu+st", "epi:fence+bar"]
    if os.environ.get("PROF") == "tma":
        names = names_tma
    tot = float(pr[:, 0].mean())
    print("  per-role cycles (mean over %d CTAs, kernel = %.0f cycles): " % (pr.shape[0], tot) +
          ", ".join("%s %.0f%%" % (n, 100 * float(pr[:, i].mean()) / tot) for i, n in enumerate(names) if i))
print("%s B%d %dx%d %d->%d k%d: %.3f ms  %.1f TFLOP/s" % (kind, B, H, W, Cin, Cout, k, ms, fl / ms / 1e9))
