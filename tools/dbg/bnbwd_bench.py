import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "shape-attentive-unet_b200"))
import torch
from saunet_b200 import engine
from saunet_b200.engine import Tape, conv, packed, packed_tc
B, H, W, Cmid, Cin, ld = 16, 128, 128, 128, int(sys.argv[1]), 256
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
tp = Tape(dev, False)
dy = tp.new(B, H, W, Cmid); dy.s.t.normal_()
X = tp.new(B, H, W, ld); X.s.t.normal_()
G = tp.new(B, H, W, ld); G.s.t.normal_()
w = torch.randn(Cmid, Cin, 1, 1, device=dev) / Cmid ** 0.5
state = torch.cat([0.5 + torch.rand(Cin), 0.3 * torch.randn(Cin), torch.randn(Cin), torch.rand(Cin) + 0.5]).to(dev)
sums = torch.zeros(2 * Cin, dtype=torch.float64, device=dev)
x = X.slice(0, Cin); g = G.slice(0, Cin)
def run():
    conv(tp, dy, packed(tp, w, 1), Cin, 1, 1, g, H, W, acc=1, stat=(sums.data_ptr(), sums.data_ptr() + 8 * Cin),
         wtc=packed_tc(tp, w, 1, 1, Cmid, Cin, wide=True), epi=(x, state.data_ptr(), flags))
def run_plain():
    conv(tp, dy, packed(tp, w, 1), Cin, 1, 1, g, H, W, acc=1, wtc=packed_tc(tp, w, 1, 1, Cmid, Cin, wide=True))
for name, fn in (("fused flags=%d" % flags, run), ("plain acc", run_plain)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        tp.stream = side.cuda_stream
        fn(); side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(10): fn()
        graph.replay(); side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side); graph.replay(); e1.record(side); side.synchronize()
    print("%s Cin=%d: %.3f ms" % (name, Cin, e0.elapsed_time(e1) / 10))
    tp.stream = torch.cuda.current_stream().cuda_stream
