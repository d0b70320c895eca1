"""CPU-side checks of the drop-in boundary: the state_dict / module-tree contract of the reference, the C-ABI
library's exported symbols, and that the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import warnings

import pytest
import torch

from helpers import ROOT, key_contract


def _model():
    from models import SAUNet
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return SAUNet(num_classes=4, pretrained=False)


def test_state_dict_contract_matches_reference():
    d = key_contract()
    m = _model()
    sd = m.state_dict()
    ref = [(k["key"], tuple(k["shape"]), k["dtype"]) for k in d["keys"]]
    mine = [(k, tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()]
    assert mine == ref
    assert [n for n, _ in m.named_children()] == d["children"]
    assert [k for k, _ in m.named_parameters()] == d["param_names"]
    assert sum(p.numel() for p in m.parameters()) == d["n_params"] == 32896505
    # aliases share storage (models/models.py:304-313)
    assert sd["conv1.0.weight"].data_ptr() == sd["encoder.features.conv0.weight"].data_ptr()
    assert sd["conv5.1.running_var"].data_ptr() == sd["encoder.features.norm5.running_var"].data_ptr()


def test_module_types_seen_by_train_py():
    """train.py:166-185 (group_weight) sorts parameters by isinstance checks on these torch classes."""
    from torch.nn.modules.batchnorm import _BatchNorm
    from torch.nn.modules.conv import _ConvNd
    from models.GSConv import GatedSpatialConv2d
    m = _model()
    assert isinstance(m.gate1, _ConvNd) and isinstance(m.gate1, GatedSpatialConv2d)
    bns = [x for x in m.modules() if isinstance(x, _BatchNorm)]
    assert len(bns) == 150
    assert sorted({b.momentum for b in bns}) == [0.001, 0.1]
    assert sum(1 for b in bns if b.momentum == 0.001) == 6


def test_library_exports_every_declared_symbol():
    from saunet_b200 import _C
    hdr = open(os.path.join(ROOT, "include", "saunet_b200.h")).read()
    declared = set(re.findall(r"\b(saunet_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("saunet_conv_desc")
    declared.discard("saunet_wgrad_desc")
    assert declared == set(_C.ALL_SYMBOLS)
    lib = ctypes.CDLL(_C.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s
    assert _C.load().saunet_version() >= 100


def test_product_path_has_no_cpu_fallback():
    from loss import DualLoss
    m = _model()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        DualLoss()((torch.zeros(1, 4, 8, 8), torch.zeros(1, 1, 8, 8)), (torch.zeros(1, 8, 8), torch.zeros(1, 1, 8, 8)))
    with pytest.raises(Exception, match="Architecture undefined"):
        from models import ModelBuilder
        ModelBuilder().build_unet(num_class=4, arch="albunet")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "shape-attentive-unet_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_grad_arena_packed_table_layout():
    """saunet_b200.parallel.GradArena (host logic, CPU): every 4-D (conv) parameter gets a packed-gradient image after
    the parameter-layout gradients, 16-byte aligned, and the device table handed to saunet_unpack_wgrad_multi lists
    them with exclusive prefix sums of their element counts (include/saunet_b200.h: saunet_unpack_entry)."""
    import ctypes as C
    from saunet_b200 import _C
    from saunet_b200.parallel import GradArena
    m = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.ConvTranspose2d(8, 6, 4), torch.nn.Conv2d(6, 5, 1))
    arena = GradArena(m)
    convs = [p for p in m.parameters() if p.dim() == 4]
    assert arena._unpack_n == len(convs) == 3
    assert arena._unpack_total == sum(p.numel() for p in convs)
    assert arena.flat.numel() == arena.n_grad and arena.flat_all.numel() == arena.n_grad + arena.packed.numel()
    assert arena.packed.numel() >= arena._unpack_total
    raw = bytes(arena._unpack_table.numpy().tobytes())
    assert len(raw) == C.sizeof(_C.UnpackEntry) * 3 == 40 * 3
    ents = (_C.UnpackEntry * 3).from_buffer_copy(raw)
    first = 0
    for e, p in zip(ents, convs):
        assert (e.first, e.A, e.Bc, e.T) == (first, p.shape[0], p.shape[1], p.shape[2] * p.shape[3])
        assert e.grad_off == arena.offsets[id(p)] and e.grad_off % 4 == 0 and e.packed_off % 4 == 0
        assert arena.packed_ptr(p) == arena.packed.data_ptr() + 4 * e.packed_off
        assert p.grad.data_ptr() == arena.flat.data_ptr() + 4 * e.grad_off
        first += p.numel()
    assert arena.packed_ptr(m[1].weight) is None              # BatchNorm weights have no packed image
    arena.flat_all.fill_(1.0)
    arena.zero()
    assert float(arena.flat.abs().sum()) == 0.0 and float(arena.packed.min()) == 1.0     # zero() leaves the (self-clearing) images alone


def test_grad_arena_keeps_one_bucket_schedule_per_tape_length():
    """The overlap schedule (which backward op finalises which gradient bucket) is learnt per tape length.  A differently
    shaped pass that only ONE rank runs (bench.py's serialised profiling pass on rank 0) must not evict the regular step's
    schedule: that rank would fall back to a learning pass without overlapped collectives while the others issue theirs
    from inside backward -- the ranks then disagree on the order of the all-reduces and the job deadlocks (seen on 2 GPUs)."""
    from saunet_b200.parallel import GradArena
    m = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 5, 1))
    arena = GradArena(m, bucket_mb=0.0001)
    assert len(arena.buckets) >= 2
    ps = list(m.parameters())
    arena.learn_schedule(10, {id(p): i for i, p in enumerate(ps)})
    first = arena.schedule_for(10)
    assert first is not None and sorted(b for bs in first.values() for b in bs) == list(range(len(arena.buckets)))
    arena.learn_schedule(7, {id(p): 0 for p in ps})           # another tape length (e.g. concurrency switched off)
    assert arena.schedule_for(10) == first and arena.schedule_for(7) is not None and arena.schedule_for(8) is None


class _DataSGD(torch.optim.Optimizer):
    """writes through p.data like the reference's radam.py:76 (no version bump)"""

    def __init__(self, params):
        super().__init__(params, dict(lr=0.1))

    def step(self, closure=None):
        for gr in self.param_groups:
            for p in gr["params"]:
                p.data.copy_(p.data * 0.5)


def test_pack_cache_follows_data_writes(monkeypatch):
    """`p.data.copy_()` does not bump `_version`; the packed-weight cache must still go stale after an
    optimizer step (global post-step hook -> generation counter) and re-pack IN PLACE (a captured CUDA graph keeps
    reading the old address).  Host logic only: the pack kernel is replaced by a recorder."""
    from saunet_b200 import _C, engine
    calls = []
    monkeypatch.setattr(_C, "call", lambda name, *a, **k: calls.append((name, a)))

    class TP:
        stream = 0
        repacked = []
    w = torch.nn.Parameter(torch.randn(8, 4, 3, 3))
    p0 = engine.packed(TP, w, 0)
    assert len(calls) == 1 and engine.packed(TP, w, 0) == p0 and len(calls) == 1          # cached
    v = w._version
    opt = _DataSGD([w])
    opt.step()
    assert w._version == v                                                                  # the trap
    assert engine.packed(TP, w, 0) == p0 and len(calls) == 2                                # stale -> re-packed, same buffer
    with torch.no_grad():
        w.mul_(2.0)                                                                         # version bump path
    assert engine.packed(TP, w, 0) == p0 and len(calls) == 3
    w.data.mul_(2.0)
    assert engine.packed(TP, w, 0) == p0 and len(calls) == 3                                # undetectable without a hint ...
    engine.invalidate_packed()
    assert engine.packed(TP, w, 0) == p0 and len(calls) == 4                                # ... which the API provides
    engine.FORCE_PACK = True
    try:
        tp = TP()
        tp.repacked = []
        assert engine.packed(tp, w, 0) == p0 and len(calls) == 5 and len(tp.repacked) == 1  # graph capture: always, once
        assert engine.packed(tp, w, 0) == p0 and len(calls) == 5
    finally:
        engine.FORCE_PACK = False
        engine.reset_capture_flags()
    assert all(not e[4] for e in engine._PACK.values())


def test_grad_arena_survives_zero_grad_set_to_none():
    """train.py:93 calls module.zero_grad() (set_to_none=True): the arena views must come back (and
    the arena must be cleared) before the backward kernels accumulate, or optimizer.step() skips every parameter."""
    from saunet_b200.parallel import GradArena
    m = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.BatchNorm2d(8))
    arena = GradArena(m)
    arena.flat.fill_(3.0)
    m.zero_grad()
    assert all(p.grad is None for p in m.parameters())
    arena.ensure_attached()
    assert float(arena.flat.abs().sum()) == 0.0
    for p in m.parameters():
        assert p.grad is not None and p.grad.data_ptr() == arena.ptr(p)
    arena.flat.fill_(2.0)
    arena.ensure_attached()                      # nothing detached: gradients are left to accumulate
    assert float(arena.flat.min()) == 2.0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference algorithm on the host CPU, the one place outside tests/ and smoke()
    that executes oracle/): one JSON line with the keys the driver reads, no GPU needed."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "slices/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "config" in d and "workload" in d["config"] and d["dtype"] == "f32"


def test_bf16_weight_image_sizes_and_precision_switch():
    """Host-side contract of the bf16 operand class (BASELINE configs[2]): a bf16 tile image holds two elements per float of
    storage (half the single-pass TF32 image, a quarter of the 3xTF32 one), for the plain and the chunk-major padded layout;
    `engine.set_precision` accepts exactly the four documented classes and maps them onto (tile passes, wgrad precision)."""
    from saunet_b200 import _C, engine
    lib = _C.load()
    for K, N, BN in ((288, 32, 32), (1152, 128, 128), (147, 64, 64)):
        one, three, bf = (lib.saunet_tc_packed_floats(K, N, BN, p) for p in (1, 3, 16))
        assert three == 2 * one and 2 * bf == one and bf > 0
    for taps, Cin, N, BN in ((9, 16, 16, 16), (9, 48, 64, 64), (9, 128, 32, 32)):
        one, three, bf = (lib.saunet_tc_packed_floats_cm(taps, Cin, N, BN, p) for p in (1, 3, 16))
        assert three == 2 * one and 2 * bf == one and bf == ((Cin + 31) // 32) * taps * ((N + BN - 1) // BN) * BN * 16
    prev = engine.get_precision()
    try:
        for name in ("fp32", "3xtf32", "tf32", "bf16"):
            engine.set_precision(name)
            assert engine.get_precision() == name
        with pytest.raises(ValueError):
            engine.set_precision("fp16")
    finally:
        engine.set_precision(prev)


def test_bench_defaults_follow_dtype():
    """`bench.py` with no flags runs BASELINE configs[1] (fp32 class, batch 16); `--dtype bf16` alone selects configs[2]'s
    batch of 32 per GPU; the config text names the arithmetic class."""
    import importlib
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in _sys.path:
        _sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    c16 = bench.workload_config(16, 1, "train")
    assert "batch 16/GPU" in c16["workload"] and "configs[1]" in c16["workload"] and "bf16" not in c16["workload"]
    c32 = bench.workload_config(32, 8, "train_loop", dtype="bf16")
    assert "configs[2]" in c32["workload"] and "bf16 tensor-core operands" in c32["workload"] and c32["global_batch"] == 256
    assert c32["parallelism"] == "dp8"
