#!/usr/bin/env python
"""bench.py -- SAUNet fwd + DualLoss + bwd throughput in 2-D slices/sec (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
                  [--workload train|train_loop|blocks|volume] [--optimizer radam|sgd|adam] [--torch-gpu-context]

--workload train (default)  BASELINE configs[1], the headline: fwd + DualLoss + bwd (+ gradient all-reduce when N>1).
           train_loop       BASELINE configs[2]-shaped loop: the same step PLUS the fused optimizer step (--optimizer),
                            batch --batch per GPU (the tensor-core arithmetic is the fp32-class 3xTF32 path; a bf16
                            operand path is not built -- DESIGN.md section 4).
           blocks           BASELINE configs[3]: DualAttBlock / GatedSpatialConv2d channel sweep at 128x128, batch 32.
           volume           BASELINE configs[4]: 16-slice 256x256 stacks, eval mode, z axis sharded over the ranks.

ours       one process per GPU (torchrun for N>1); each rank runs `--batch` 256x256 slices per step through the
           reference-facing API (models.SegmentationModule(crit, unet)(feed_dict, epoch) -> loss.backward()), all
           arithmetic in libsaunet_b200.so; gradients all-reduced over NCCL.  `value` = device-timed with inputs
           resident in HBM; `e2e` = same call with the step's inputs copied from pinned host memory and the loss
           read back every step.
reference  the reference's algorithm on the host CPU (oracle/ port, torch CPU ops, all host threads): the
           reported CPU baseline.  Rank 0 only.

A step = zero grads -> SAUNet forward (train-mode BatchNorm, on-device Canny) -> DualLoss -> backward
(+ gradient all-reduce when N>1).  Synthetic ACDC-shaped slices, deterministic random-init weights.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "shape-attentive-unet_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import datetime  # noqa: E402

import torch  # noqa: E402

METRIC = "2D slices/sec, SAUNet fwd+bwd 256x256x1->4-class (whole job, device-timed, max over ranks)"
UNIT = "slices/s"
FLOP_PER_SLICE = 216.13e9          # conv FLOPs fwd+bwd per 256x256 slice
SIZE = 256
REF_SAMPLE = 4                     # slices per CPU step: the bounded sample of the batch-16 workload the CPU arms time


ARITH = {"f32": "fp32-class 3xTF32 arithmetic",
         "bf16": "bf16 tensor-core operands (kind::f16) with fp32 accumulation for every forward / data-gradient convolution, "
                 "single-pass TF32 weight gradients; fp32 activations in HBM, fp32 BatchNorm statistics, loss, master weights"}


def workload_config(B, world, workload="train", extra=None, dtype="f32"):
    """`config` of the JSON line -- shared by both arms so that the driver compares like with like."""
    names = {"train": "SAUNet fwd+DualLoss+bwd, batch %d/GPU, 256x256x3 fp32 slices -> 4 classes, train-mode BN, on-device Canny "
                      "(BASELINE configs[1]%s)" % (B, "" if dtype == "f32" else "; " + ARITH[dtype]),
             "train_loop": "SAUNet training loop: fwd+DualLoss+bwd+gradient all-reduce+fused optimizer step, batch %d/GPU, 256x256x3 "
                           "slices -> 4 classes, train-mode BN (BASELINE configs[2]; %s)" % (B, ARITH[dtype])}
    c = {"workload": names[workload], "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
         "l2": "per-step working set (activations >> 126 MB L2) exceeds L2; no explicit flush"}
    if extra:
        c.update(extra)
    return c


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_ours(dev, batch):
    from models import SAUNet, SegmentationModule
    from loss import DualLoss
    from saunet_b200 import synth
    from saunet_b200.parallel import GradArena
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        unet = SAUNet(num_classes=4, pretrained=False)
    unet.load_state_dict(synth.synthetic_state_dict(unet.state_dict(), seed=0))
    unet = unet.to(dev).train()
    seg_mod = SegmentationModule(DualLoss(num_classes=4), unet, 4).to(dev).train()
    arena = GradArena(unet, bucket_mb=float(os.environ.get("SAUNET_BUCKET_MB", "32")))
    return seg_mod, unet, arena


def host_batch(batch, rank):
    from saunet_b200 import synth
    base = synth.synthetic_batch(min(batch, 4), SIZE, seed=304 + rank)
    rep = (batch + base["image"].shape[0] - 1) // base["image"].shape[0]
    d = {k: v.repeat((rep,) + (1,) * (v.dim() - 1))[:batch].contiguous().pin_memory() for k, v in base.items()}
    return d


def run_ours(args):
    import torch.distributed as dist
    from saunet_b200 import _C
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # (a collective that cannot complete must fail the run within minutes, not hang it until the driver's clock runs out)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _C.load()
    from saunet_b200 import engine
    engine.set_precision("bf16" if args.dtype == "bf16" else "3xtf32")
    B = args.batch
    seg_mod, unet, arena = build_ours(dev, B)
    hb = host_batch(B, rank)
    resident = {k: v.to(dev) for k, v in hb.items()}
    opt = None
    if args.workload == "train_loop":
        from saunet_b200.optim import create_fused_optimizer
        opt = create_fused_optimizer(unet, arena, args.optimizer, lr=1e-4)      # train.sh: --lr_encoder 1e-4

    def step(feed, reduce=True):
        arena.zero()
        loss, acc = seg_mod({"image": feed["image"], "mask": (feed["seg"], feed["edge"])}, 0)
        loss.backward()
        if reduce:
            arena.all_reduce()
        if opt is not None:
            opt.step()
        return loss

    def step_e2e():
        if e2e_mode == "pipelined":
            return loop.step(hb)
        if graphed is not None and e2e_mode == "cuda_graph":
            loss = graphed(hb)
            arena.all_reduce()
            if opt is not None:
                opt.step()
            return float(loss.item())
        feed = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
        return float(step(feed).item())

    def timed(fn, iters, finish=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        if finish is not None:
            finish()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step(resident)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _C.launch_count()
    ms = timed(lambda: step(resident), args.steps)
    launches = _C.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- per-kernel roofline pass (one extra step, every C-ABI call bracketed by CUDA events) ----
    roof = None
    cpu_base = None
    if rank == 0:
        # three profiled steps; per C-ABI call the MEDIAN of its three timings is kept, so that a single slow launch
        # (allocator growth, a clock dip) cannot pick the "dominant kernel"
        # (serialised: with the weight gradients / shape stream on side streams the bracketing events of one call
        #  would span kernels of other streams sharing the SMs)
        runs = []
        conc, engine.CONCURRENCY = engine.CONCURRENCY, False
        for _ in range(3):
            _C.PROFILE = []
            with arena.no_sync():                 # rank 0 only: NO collective may be issued in this pass
                step(resident, reduce=False)
            torch.cuda.synchronize()
            prof, _C.PROFILE = _C.PROFILE, None
            runs.append([(name, a.elapsed_time(b), fl, nb, kern) for name, a, b, fl, nb, _tag, kern in prof])
        engine.CONCURRENCY = conc
        agg, kagg = {}, {}
        if len({len(r) for r in runs}) == 1:
            calls = [(r0[0], sorted((r0[1], r1[1], r2[1]))[1], r0[2], r0[3], r0[4]) for r0, r1, r2 in zip(*runs)]
        else:
            calls = runs[-1]
        for name, t, fl, nb, kern in calls:
            e = agg.setdefault(name, [0.0, 0, 0.0, 0.0])
            e[0] += t; e[1] += 1; e[2] += fl; e[3] += nb
            e = kagg.setdefault(kern, [0.0, 0, 0.0, 0.0])          # by the kernel family the call dispatched to
            e[0] += t; e[1] += 1; e[2] += fl; e[3] += nb
        total = sum(e[0] for e in agg.values())
        pk = peaks()
        conv_ms = sum(agg[k][0] for k in agg if "conv2d" in k)
        conv_fl = sum(agg[k][2] for k in agg if "conv2d" in k)
        # dominant KERNEL of the step (actual kernel family, not the C-ABI entry point): algorithmic FLOPs (convs:
        # 2*M*K*N per launch) or bytes (BatchNorm backward: tensors read + written once) over its CUDA-event time
        kern, (t_ms, cnt, fl, nb) = max(kagg.items(), key=lambda kv: kv[1][0])
        # DRAM traffic of one launch from the committed `ncu --set full` captures (profiles/r0?_*.md), next to
        # the algorithmic bytes of the captured geometry: traffic ~ algorithmic means no wasted re-reads
        NCU = {"conv_halo_tma_kernel": ("profiles/r02_halo_tma_raw64_final.md", 487.9e6, "B16 256x256 64->64 3x3 (TMA-fed): 268.8 MB read + 219.1 MB written (part of the output still in L2) vs 537 MB algorithmic; 64.7 % tensor-pipe active"),
               "conv_pw_t_kernel": ("profiles/r02_pwt_bnbwd.md", 598.8e6, "B16 128x128 1x1 data gradient with the fused BatchNorm-backward epilogue: 470.1 MB read + 128.8 MB written"),
               "conv_wgrad_halo_kernel": ("profiles/r01_final_wgrad_halo.md", 176.2e6, "B16 128x128 128->32 3x3: 171.4 MB read + 4.8 MB written vs 168 MB algorithmic (Q 134 MB + dY 34 MB)"),
               "conv_halo_kernel": ("profiles/r01_final_halo_n32.md", 153.8e6, "B16 128x128 128->32 3x3: 134.6 MB read + 19.2 MB written vs 168 MB algorithmic"),
               "conv_halo_persist_kernel": ("profiles/r01_final_halo_persist.md", 73.9e6, "B16 64x64 256->128 3x3: 69.6 MB read + 4.3 MB written (output still in L2) vs 101 MB algorithmic; 68.8 % tensor-pipe active"),
               "conv_tc_kernel": ("profiles/r01_final_tc_1x1.md", 138.1e6, "B16 64x64 480->128 1x1: 126.4 MB read + 11.8 MB written (output still in L2) vs 159 MB algorithmic"),
               "conv_wgrad_pw_kernel": ("profiles/r01_final_wgrad_pw.md", 163.5e6, "B16 64x64 480->128 1x1: 159.8 MB read + 3.7 MB written vs 159 MB algorithmic"),
               "bn_bwd_apply4_kernel": ("profiles/r01_final_bn.md", 364.0e6, "C128 npix262144: 268.5 MB read + 95.6 MB written vs 403 MB algorithmic (dx partly still in L2)"),
               "bn_bwd_reduce4_kernel": ("profiles/r01_final_bn.md", 272.0e6, "C128 npix262144: 268.4 MB read vs 268 MB algorithmic")}
        ncu = NCU.get(kern)
        if fl > 0:
            ach = fl / (t_ms / 1e3) / 1e12
            roof = {"bound": "tensor", "kernel": kern, "achieved": round(ach, 2), "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": round(ach / pk["tflops"], 4), "traffic": ncu[1] if ncu else None, "launches": cnt,
                    "avg_launch_ms": round(t_ms / cnt, 4), "share_of_step": round(t_ms / total, 3), "peak_src": pk["src"],
                    "all_conv_tflops": round(conv_fl / (conv_ms / 1e3) / 1e12, 2),
                    "note": ("peak = measured dense bf16 cuBLAS (sustained); this path computes fp32-class 3xTF32: 3 tf32 MMAs per "
                             "product at half the bf16 rate, i.e. a ceiling of peak/6; `achieved` counts each product once")
                            if args.dtype == "f32" else
                            ("peak = measured dense bf16 cuBLAS (sustained); forward / data-gradient convolutions run bf16 MMAs "
                             "(ceiling = peak), weight gradients single-pass TF32 (ceiling = peak/2)")}
        else:
            ach = nb / (t_ms / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": kern, "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": round(ach / pk["hbm_gbs"], 4), "traffic": ncu[1] if ncu else None, "launches": cnt,
                    "avg_launch_ms": round(t_ms / cnt, 4), "share_of_step": round(t_ms / total, 3), "peak_src": pk["src"]}
        if ncu:
            roof["traffic_src"] = "%s (%s)" % (ncu[0], ncu[2])
        roof["by_call_ms"] = {k: round(v[0], 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:8]}
        roof["by_kernel"] = {k: {"ms": round(v[0], 2), "launches": v[1],
                                 ("tflops" if v[2] > 0 else "gbs"): round((v[2] / 1e12 if v[2] > 0 else v[3] / 1e9) / (v[0] / 1e3), 1)}
                             for k, v in sorted(kagg.items(), key=lambda kv: -kv[1][0])[:10]}
        if world == 1 and not args.no_cpu_baseline:
            cpu_base = cpu_baseline_sample(steps=2, batch=REF_SAMPLE)

    # end-to-end arm: the same step captured once as a CUDA graph (saunet_b200.graphs.GraphedStep, part of the public
    # API): per step = H2D copies of image/seg/edge from pinned memory into the graph's static inputs, one graph
    # launch (identical kernels), gradient all-reduce, loss.item().  Falls back to the eager call if capture fails.
    graphed = None
    if not args.no_graph:
        try:
            from saunet_b200.graphs import GraphedStep
            graphed = GraphedStep(seg_mod, arena, resident)
        except Exception as e:                      # noqa: BLE001
            print("CUDA-graph capture unavailable, e2e runs eagerly: %r" % (e,), file=sys.stderr)
            graphed = None

    if world > 1:               # the graph mode is offered only if EVERY rank captured (the ranks must agree on the path)
        t = torch.tensor([1 if graphed is not None else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 0:
            graphed = None
    ctx = None
    if args.torch_gpu_context and rank == 0:
        ctx = torch_gpu_context(dev, B)
    # three ways to drive the same step, all part of the public API: the plain eager call (copy, step, loss.item()), its
    # CUDA-graph replay (saunet_b200.graphs.GraphedStep: no host launch time, but fewer kernels of different streams in
    # flight) and the pipelined loop (saunet_b200.loop.TrainLoop: the next batch is copied on a copy stream while the
    # current step computes, and a step's loss is read after the NEXT step has been issued).  Every mode copies every
    # step's inputs from pinned host memory and reads every step's loss inside the timed region; the end-to-end arm uses
    # whichever is fastest on this box (1 untimed + 3 timed steps each decide).
    from saunet_b200.loop import TrainLoop
    loop = TrainLoop(seg_mod, arena, hb, optimizer=opt)
    modes = ["pipelined", "eager"] + (["cuda_graph"] if graphed is not None else [])
    if args.e2e_mode != "auto":
        modes = [args.e2e_mode]
    trial = {}
    for e2e_mode in modes:
        step_e2e()
        trial[e2e_mode] = timed(step_e2e, 3, finish=loop.flush if e2e_mode == "pipelined" else None)
    e2e_mode = min(trial, key=trial.get)
    if world > 1:                                       # every rank must take the same path
        t = torch.tensor([modes.index(e2e_mode)], device=dev)
        dist.broadcast(t, 0)
        e2e_mode = modes[int(t.item())]
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps, finish=loop.flush if e2e_mode == "pipelined" else None)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    if rank == 0:
        h2d = sum(v.numel() * v.element_size() for v in hb.values())
        out = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
               "config": workload_config(B, world, args.workload,
                                         {"optimizer": "fused " + args.optimizer} if args.workload == "train_loop" else None,
                                         dtype=args.dtype),
               "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                       "ms_per_step": round(ms_e2e / args.steps, 3),
                       "launch": e2e_mode, "trial_ms_per_step": {k: round(v / 3, 3) for k, v in trial.items()}},
               "gpu_launches": int(launches), "clocks": clocks,
               "tensor_pipe_fraction": round(FLOP_PER_SLICE * value / world / (peaks()["tflops"] * 1e12), 4),
               "roofline": roof, "cpu_baseline": cpu_base}
        if ctx is not None:
            out["torch_gpu_context"] = ctx
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def torch_gpu_context(dev, B):
    """CONTEXT ONLY (opt-in, never the product path): the oracle's stock torch ops (ATen / cuDNN) run on the GPU for the same
    step -- what `python train.py` of the reference would execute on this box ("the bar to beat is
    stock PyTorch").  allow_tf32 off = cuDNN fp32 (the precision class of this repo's 3xTF32 path), on = PyTorch's default
    for convolutions (single-pass TF32, ~1e-3 relative error: not parity-valid)."""
    from oracle import saunet_oracle as O
    from saunet_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import template_state_dict
    w = synth.synthetic_state_dict(template_state_dict(), seed=0)
    data = synth.synthetic_batch(min(B, 4), SIZE, seed=304)
    rep = (B + 3) // 4
    img = data["image"].repeat(rep, 1, 1, 1)[:B]
    canny = O.canny_map(img).to(dev)
    img, seg_t, edge_t = img.to(dev), data["seg"].repeat(rep, 1, 1)[:B].to(dev), data["edge"].repeat(rep, 1, 1, 1)[:B].to(dev)
    sd = {k: (v.to(dev).requires_grad_(v.is_floating_point() and "running" not in k and "_tmp" not in k and "_running" not in k)
              if v.is_floating_point() else v.to(dev)) for k, v in w.items()}
    res = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32

        def step():
            for v in sd.values():
                if v.is_floating_point() and v.grad is not None:
                    v.grad = None
            seg, edge = O.saunet_forward(sd, img, training=True, canny=canny)
            O.dual_loss(seg, edge, seg_t, edge_t).backward()
        try:
            ms = _timed(step, 5)
            res["allow_tf32_%s" % ("on" if tf32 else "off")] = {"ms_per_step": round(ms, 2), "slices_per_s": round(B / ms * 1e3, 1)}
        except Exception as e:                      # noqa: BLE001
            res["allow_tf32_%s" % ("on" if tf32 else "off")] = {"error": repr(e)[:200]}
    torch.backends.cudnn.allow_tf32 = True
    res["note"] = "oracle (reference algorithm as torch functional ops) on cuda:0 through stock ATen/cuDNN, batch %d; context only" % B
    return res


def cpu_baseline_sample(steps, batch, threads=None):
    """The reference algorithm (oracle/ port: torch CPU ops) on the host cores: fwd + DualLoss + bwd."""
    from oracle import saunet_oracle as O
    from saunet_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import template_state_dict
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    w = synth.synthetic_state_dict(template_state_dict(), seed=0)
    data = synth.synthetic_batch(batch, SIZE, seed=304)
    O.train_step(w, data["image"][:1], data["seg"][:1], data["edge"][:1])        # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(w, data["image"], data["seg"], data["edge"])
    dt = time.perf_counter() - t0
    return {"value": round(steps * batch / dt, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d steps of batch %d (256x256) fwd+DualLoss+bwd, torch %s CPU ops, %.1f s" %
                      (steps, batch, torch.__version__, dt)}


def run_reference(args):
    """The reference's algorithm on the host CPU (oracle/ port of the reference's torch ops; the reference itself is a set
    of scripts importing nibabel / skimage / matplotlib and cannot be installed or run on the GPU box).  Same metric and
    `config` as the product arm; each step processes a BOUNDED SAMPLE of the batch-16 workload -- REF_SAMPLE slices -- because one
    batch-16 CPU step takes ~10-25 s (24.8 s on 8 cores), i.e. 4-10 minutes for the driver's
    20 + 5 steps; per-slice CPU throughput does not improve with batch (0.65 slices/s at batch 16 vs 0.77 at 4)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import saunet_oracle as O
    from saunet_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import template_state_dict
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    batch = REF_SAMPLE
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = synth.synthetic_state_dict(template_state_dict(), seed=0)
    data = synth.synthetic_batch(batch, SIZE, seed=304)
    for _ in range(max(1, min(args.warmup, 2))):
        O.train_step(w, data["image"][:1], data["seg"][:1], data["edge"][:1])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.train_step(w, data["image"], data["seg"], data["edge"])
    dt = time.perf_counter() - t0
    v = round(args.steps * batch / dt, 3)
    sample = ("each step = %d slices of the batch-%d 256x256 workload (bounded sample: a full batch-16 CPU step takes 10-25 s), "
              "fwd+DualLoss+bwd in torch %s CPU ops, %d host threads" % (batch, args.batch, torch.__version__, threads))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args.batch, world, "train"),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def _timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_blocks(args):
    """BASELINE configs[3]: DualAttBlock + GatedSpatialConv2d at 128x128 output maps, batch 32, C in {64,128,256,512},
    through the public nn.Module API.  Reported against the algorithmic byte counts of SURVEY.md section 8(d):
    attention TAIL fwd = 3*C*HW*B*4 B ("2-pass": train-mode BN inside the spatial attention forces one extra read of
    `fused`), bwd = 2x; GSConv fwd = (4C+4)*HW*B*4 B (two batch-statistic barriers), bwd = 2x.  The tail is timed in
    isolation by summing the CUDA-event times of its own C-ABI calls (saunet_b200._C.SCOPE == 'tail')."""
    from models.attention_blocks import DualAttBlock
    from models.GSConv import GatedSpatialConv2d
    from saunet_b200 import _C, engine
    engine.set_precision("bf16" if args.dtype == "bf16" else "3xtf32")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    pk = peaks()
    B, S = 32, 128
    HW = S * S

    def nhwc(*shape):
        return torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)

    def scoped(fn, scope):
        """median over 3 runs of the summed event time of the calls made under `scope`"""
        vals = []
        for _ in range(3):
            _C.PROFILE = []
            fn()
            torch.cuda.synchronize()
            prof, _C.PROFILE = _C.PROFILE, None
            vals.append(sum(a.elapsed_time(b) for _n, a, b, _f, _nb, tag, _k in prof if tag.startswith(scope + ":")))
        return sorted(vals)[1]

    res = {"dualatt": [], "gsconv": []}
    for C in (64, 128, 256, 512):
        blk = DualAttBlock(inchannels=[C, C], outchannels=C).to(dev).train()
        lo, skip = nhwc(B, C, S // 2, S // 2).requires_grad_(True), nhwc(B, C, S, S).requires_grad_(True)

        def fwd():
            with torch.no_grad():
                blk([lo, skip])

        def fwdbwd():
            o, sp = blk([lo, skip])
            (o.sum() + sp.sum()).backward()
        f, fb = _timed(fwd, args.steps), _timed(fwdbwd, args.steps)
        tail_f, tail_fb = scoped(fwd, "tail"), scoped(fwdbwd, "tail")
        tail_bytes = 3.0 * C * HW * B * 4
        flops = 2.0 * B * HW * (9 * 2 * C * C + 4 * C * C + C * C // 4)        # c3x3rb + convT (4 taps/output) + spatial down
        res["dualatt"].append({"C": C, "block_fwd_ms": round(f, 3), "block_fwdbwd_ms": round(fb, 3),
                               "block_fwd_tflops": round(flops / f / 1e9, 1), "block_fwdbwd_tflops": round(3 * flops / fb / 1e9, 1),
                               "tail_fwd_ms": round(tail_f, 3), "tail_fwdbwd_ms": round(tail_fb, 3),
                               "tail_fwd_gbs": round(tail_bytes / tail_f / 1e6, 1), "tail_fwd_frac_hbm": round(tail_bytes / tail_f / 1e6 / pk["hbm_gbs"], 3),
                               "tail_fwdbwd_gbs": round(3 * tail_bytes / tail_fb / 1e6, 1),
                               "tail_fwdbwd_frac_hbm": round(3 * tail_bytes / tail_fb / 1e6 / pk["hbm_gbs"], 3)})
        gs = GatedSpatialConv2d(C, C).to(dev).train()
        x, g = nhwc(B, C, S, S).requires_grad_(True), nhwc(B, 1, S, S).requires_grad_(True)

        def gfwd():
            with torch.no_grad():
                gs(x, g)

        def gfb():
            o, a = gs(x, g)
            (o.sum() + a.sum()).backward()
        f, fb = _timed(gfwd, args.steps), _timed(gfb, args.steps)
        byt = (4.0 * C + 4) * HW * B * 4
        res["gsconv"].append({"C": C, "fwd_ms": round(f, 3), "fwdbwd_ms": round(fb, 3), "fwd_gbs": round(byt / f / 1e6, 1),
                              "fwd_frac_hbm": round(byt / f / 1e6 / pk["hbm_gbs"], 3), "fwdbwd_gbs": round(3 * byt / fb / 1e6, 1),
                              "fwdbwd_frac_hbm": round(3 * byt / fb / 1e6 / pk["hbm_gbs"], 3)})
        del blk, gs, lo, skip, x, g
        torch.cuda.empty_cache()
    best = max(r["tail_fwd_gbs"] for r in res["dualatt"])
    out = {"metric": "achieved HBM GB/s of algorithmic bytes, DualAttBlock attention tail / GatedSpatialConv2d, 128x128 maps, batch 32 "
                     "(BASELINE configs[3]); value = best attention-tail forward", "value": best, "unit": "GB/s", "n_gpus": 1,
           "steps": args.steps, "warmup": 3, "higher_is_better": True, "dtype": args.dtype, "data": "synthetic",
           "config": {"workload": "DualAttBlock([C,C]->C) and GatedSpatialConv2d(C,C), C in 64..512, 128x128 output, batch 32, train mode; "
                                  + ARITH[args.dtype]},
           "peak": pk["hbm_gbs"], "peak_src": pk["src"], "results": res}
    print(json.dumps(out), flush=True)


def run_volume(args):
    """BASELINE configs[4]: 16-slice 256x256 stacks, eval mode, z axis sharded r::world over the ranks, argmax on device,
    one all_gather of uint8 label maps per volume; also 8 volumes in flight as one call."""
    import torch.distributed as dist
    from models import SAUNet
    from saunet_b200 import synth
    from saunet_b200.inference import predict_volume
    from saunet_b200 import engine
    engine.set_precision("bf16" if args.dtype == "bf16" else "3xtf32")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # (a collective that cannot complete must fail the run within minutes, not hang it until the driver's clock runs out)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = SAUNet(num_classes=4, pretrained=False)
    m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), seed=0))
    m = m.to(dev).eval()
    base = synth.synthetic_batch(4, SIZE, seed=7)["image"]
    res = {}
    for nvol in (1, 8):
        vol = base.repeat(4 * nvol, 1, 1, 1).contiguous().to(dev)          # 16 * nvol slices
        ms = _timed(lambda: predict_volume(m, vol), args.steps)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        res["%d_volume%s" % (nvol, "" if nvol == 1 else "s")] = {"slices": 16 * nvol, "ms": round(ms, 3), "slices_per_s": round(16 * nvol / ms * 1e3, 1)}
    if rank == 0:
        out = {"metric": "2D slices/sec, SAUNet volume inference (eval-mode BN, argmax on device, z axis sharded over ranks)",
               "value": res["8_volumes"]["slices_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": 3,
               "higher_is_better": True, "scaling": "strong", "dtype": args.dtype, "data": "synthetic",
               "config": {"workload": "16-slice 256x256x3 stacks, one and eight volumes per call (BASELINE configs[4]); " + ARITH[args.dtype],
                          "parallelism": "z-shard x%d" % world}, "results": res}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="slices per GPU (default 16; 32 for --dtype bf16 = BASELINE configs[2])")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"],
                    help="arithmetic of the convolutions: f32 = 3xTF32 (fp32 class, the headline), bf16 = bf16 operands / fp32 accumulate")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the e2e arm eagerly instead of through a CUDA graph")
    ap.add_argument("--e2e-mode", default="auto", choices=["auto", "pipelined", "eager", "cuda_graph"])
    ap.add_argument("--workload", default="train", choices=["train", "train_loop", "blocks", "volume"])
    ap.add_argument("--optimizer", default="radam", choices=["radam", "sgd", "adam"], help="train_loop: fused optimizer (train.sh uses radam)")
    ap.add_argument("--torch-gpu-context", action="store_true",
                    help="also time the oracle's stock torch ops (cuDNN) on the GPU, allow_tf32 off and on: a context number, never the product path")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 32 if args.dtype == "bf16" else 16
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "blocks":
        run_blocks(args)
    elif args.workload == "volume":
        run_volume(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
