#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REAL reference.

Runs only in the build container (needs /root/reference, read-only).  The
reference is imported unmodified with the shims SURVEY.md section 8(c) lists:
  * stub ``matplotlib`` / ``matplotlib.cm`` (models/models.py:10 ->
    misc_functions.py:9 imports it, nothing on the path uses it);
  * ``Tensor.cuda`` / ``Module.cuda`` -> identity (models/models.py:363 and
    loss.py:130-156 hard-code .cuda());
  * ``SAUNet(pretrained=False)`` (no network).
Weights come from ``synth.synthetic_state_dict`` (values depend only on key
name/shape/seed) and inputs from ``synth.synthetic_batch``, so the GPU-box
tests can regenerate both bit-identically without the reference.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz|json
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
PKG = os.path.join(ROOT, "shape-attentive-unet_b200")


def import_reference():
    for name in ("matplotlib", "matplotlib.cm"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    import models.models as ref_models      # noqa: E402
    import models.attention_blocks as ref_att
    import models.GSConv as ref_gsc
    import models.resnet as ref_resnet
    import loss as ref_loss
    return ref_models, ref_att, ref_gsc, ref_resnet, ref_loss


def load_synth():
    import importlib.util
    spec = importlib.util.spec_from_file_location("saunet_synth", os.path.join(PKG, "saunet_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def grads_summary(model):
    names, l2, sm = [], [], []
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        l2.append(float(p.grad.double().norm()))
        sm.append(float(p.grad.double().sum()))
    return names, np.array(l2), np.array(sm)


FULL_GRAD_KEYS = ["final.weight", "final.bias", "gate1.weight", "cw.weight", "fuse.weight", "c3.weight",
                  "d0.weight", "res1.conv1.weight", "res1.bn1.weight", "gate1._gate_conv.1.weight",
                  "gate1._gate_conv.0.weight", "gate1._gate_conv.4.bias", "expand.0.weight", "expand.1.weight",
                  "dec2.spatialAttn.phi.weight", "dec2.spatialAttn.bn.weight", "dec2.channelAttn.fc1.weight",
                  "dec2.channelAttn.fc2.bias", "dec2.mrf.up.0.bias", "dec1.block.1.weight", "dec1.block.2.bias",
                  "dec0.0.weight", "center.0.bias", "encoder.features.conv0.weight", "encoder.features.norm0.weight",
                  "encoder.features.denseblock1.denselayer1.norm1.weight",
                  "encoder.features.denseblock1.denselayer6.conv2.weight",
                  "encoder.features.transition1.conv.weight", "encoder.features.norm5.bias",
                  "encoder.features.denseblock4.denselayer16.conv1.weight"]
BN_KEYS = ["encoder.features.norm0", "encoder.features.denseblock1.denselayer2.norm1",
           "encoder.features.denseblock3.denselayer24.norm2", "encoder.features.transition2.norm",
           "encoder.features.norm5", "res1.bn1", "res3.bn2", "gate2._gate_conv.0", "gate2._gate_conv.4",
           "expand.1", "center.1", "dec5.mrf.up.1", "dec4.c3x3rb.1", "dec3.spatialAttn.bn", "dec1.block.2",
           "dec0.1"]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref_models, ref_att, ref_gsc, ref_resnet, ref_loss = import_reference()
    synth = load_synth()

    # ---- 1. state_dict key contract ------------------------------------
    model = ref_models.SAUNet(num_classes=4, pretrained=False)
    sd0 = model.state_dict()
    groups = {}
    for k, v in sd0.items():
        groups.setdefault((v.data_ptr(), tuple(v.shape)) if v.numel() else (k,), []).append(k)
    keys = [{"key": k, "shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", "")} for k, v in sd0.items()]
    children = [n for n, _ in model.named_children()]
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump({"keys": keys, "children": children,
                   "param_names": [k for k, _ in model.named_parameters()],
                   "n_params": int(sum(p.numel() for p in model.parameters()))}, f)
    print("state_dict entries:", len(keys), "children:", len(children))

    weights = synth.synthetic_state_dict(sd0, seed=0)
    model.load_state_dict(weights)
    crit = ref_loss.DualLoss(num_classes=4)

    # ---- 2. whole-model goldens ----------------------------------------
    def run(batch, size, training, tag, probe_stride=1, with_maps=False):
        model.load_state_dict(weights)
        model.train(training)
        model.zero_grad()
        data = synth.synthetic_batch(batch, size, seed=304)
        x = data["image"].clone()
        maps = None
        if training:
            if with_maps:       # ONE forward (a second one would update the BatchNorm running statistics twice)
                seg, edge, maps = model(x, return_att=True)
            else:
                seg, edge = model(x)
            loss = crit((seg, edge), (data["seg"], data["edge"]))
            loss.backward()
        else:
            with torch.no_grad():
                if with_maps:
                    seg, edge, maps = model(x, return_att=True)
                else:
                    seg, edge = model(x)
                loss = crit((seg, edge), (data["seg"], data["edge"]))
        out = {"logits": np32(seg)[:, :, ::probe_stride, ::probe_stride],
               "edge": np32(edge)[:, :, ::probe_stride, ::probe_stride],
               "loss": np.float64(loss.item()),
               "probe_stride": np.int64(probe_stride),
               "image_sum": np.float64(data["image"].double().sum().item()),
               "seg_sum": np.int64(data["seg"].sum().item()),
               "edge_sum": np.float64(data["edge"].sum().item())}
        if maps is not None:
            # models/models.py:386-393: [att2, att3, att4, att5, g1, g2, g3], each [B,1,H,W]
            assert len(maps) == 7
            for i, t in enumerate(maps):
                out["map%d" % i] = np32(t)[:, :, ::probe_stride, ::probe_stride]
            # SegmentationModule's training-branch metrics (models/models.py:51-74,92) on the reference's own logits
            base = ref_models.SegmentationModuleBase()
            acc, jac = base.pixel_acc(torch.round(torch.nn.functional.softmax(seg.detach(), dim=1)).long(),
                                      data["seg"].long(), 4)
            out["acc"] = np.float64(float(acc))
            out["jaccard"] = np.array([float(j) for j in jac], dtype=np.float64)
        # the canny map the reference computed internally (models/models.py:359-362)
        import cv2
        im = np.mean(x.numpy(), axis=1).astype(np.uint8)
        out["canny_sum"] = np.float64(sum(float(cv2.Canny(im[i], 10, 100).sum()) for i in range(batch)))
        if training:
            names, l2, sm = grads_summary(model)
            out["grad_names"] = np.array(names)
            out["grad_l2"] = l2
            out["grad_sum"] = sm
            gp = dict(model.named_parameters())
            for k in FULL_GRAD_KEYS:
                out["grad/" + k] = np32(gp[k].grad)
            sd1 = model.state_dict()
            for k in BN_KEYS:
                out["bn/" + k + ".running_mean"] = np32(sd1[k + ".running_mean"])
                out["bn/" + k + ".running_var"] = np32(sd1[k + ".running_var"])
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
        print(tag, "loss", float(loss), "logits absmax", float(seg.abs().max()))

    only = set(sys.argv[1:])          # e.g. `make_golden.py saunet_train_b16_s256` regenerates just that fixture

    def want(tag):
        return not only or tag in only

    if want("saunet_train_b2_s64"):
        run(2, 64, True, "saunet_train_b2_s64")
    if want("saunet_eval_b2_s64"):
        run(2, 64, False, "saunet_eval_b2_s64")
    if want("saunet_train_b1_s256"):
        run(1, 256, True, "saunet_train_b1_s256", probe_stride=8)
    if want("saunet_eval_b1_s256"):
        run(1, 256, False, "saunet_eval_b1_s256", probe_stride=8)
    # BASELINE configs[1] itself: batch 16 at 256x256, train mode, with the attention maps and the metrics
    if want("saunet_train_b16_s256"):
        run(16, 256, True, "saunet_train_b16_s256", probe_stride=8, with_maps=True)
    if want("saunet_maps_b2_s64"):
        run(2, 64, True, "saunet_maps_b2_s64", with_maps=True)
    # the reference's OWN gradient noise floor at the headline size: its fp32 backward re-run with every weight
    # perturbed by ~1 ulp (3e-7 relative) / ~10 ulp (1e-6, the operand rounding of 3xTF32), worst of three draws.
    # ReLU / max-pool mask flips make deep gradients (norm0, conv0 ...) move by ~1 % elementwise under such a
    # perturbation, so the GPU test bounds the CUDA path by a multiple of THIS floor, not by a fixed number.
    if want("saunet_train_b16_s256_noise"):
        data = synth.synthetic_batch(16, 256, seed=304)

        def grads_of(w):
            model.load_state_dict(w)
            model.train(True)
            model.zero_grad()
            seg, edge = model(data["image"].clone())
            crit((seg, edge), (data["seg"], data["edge"])).backward()
            return {k: p.grad.detach().double().clone() for k, p in model.named_parameters() if p.grad is not None}

        base = grads_of(weights)
        names = list(base.keys())
        out = {"grad_names": np.array(names)}
        for eps in (3e-7, 1e-6):
            tag = "%g" % eps
            nrm = np.zeros(len(names))
            el = {k: 0.0 for k in FULL_GRAD_KEYS}
            l2 = {k: 0.0 for k in FULL_GRAD_KEYS}
            for seed in (1, 2, 3):
                gen = torch.Generator().manual_seed(seed)
                w2, done = {}, {}
                for k, v in weights.items():
                    if v.is_floating_point() and v.dim() >= 1 and "running" not in k:
                        key = (v.data_ptr(), tuple(v.shape))
                        if key not in done:        # aliased encoder tensors get ONE perturbation
                            done[key] = v * (1 + eps * torch.randn(v.shape, generator=gen))
                        w2[k] = done[key]
                    else:
                        w2[k] = v
                g1 = grads_of(w2)
                for i, k in enumerate(names):
                    n0 = max(float(base[k].norm()), 1e-12)
                    nrm[i] = max(nrm[i], abs(float(g1[k].norm()) - n0) / n0)
                for k in FULL_GRAD_KEYS:
                    d = g1[k] - base[k]
                    el[k] = max(el[k], float(d.abs().max() / base[k].abs().max().clamp_min(1e-30)))
                    l2[k] = max(l2[k], float(d.norm() / base[k].norm().clamp_min(1e-30)))
            out["noise_norm/" + tag] = nrm
            for k in FULL_GRAD_KEYS:
                out["noise_el/%s/%s" % (tag, k)] = np.float64(el[k])
                out["noise_l2/%s/%s" % (tag, k)] = np.float64(l2[k])
            print("noise eps", eps, "max norm err", nrm.max(), "max l2", max(l2.values()), max(l2, key=l2.get))
        np.savez_compressed(os.path.join(HERE, "saunet_train_b16_s256_noise.npz"), **out)
        model.load_state_dict(weights)
    # ---- optimizer golden: the REAL radam.RAdam (radam.py:15-78) on two tensors (a decayed conv weight, an un-decayed
    # bias), 12 steps (N_sma crosses 5 at step 6: both update branches), seeded gradients
    if want("optim_radam"):
        import warnings as _w
        from radam import RAdam
        gen = torch.Generator().manual_seed(77)
        w0 = torch.randn(8, 4, 3, 3, generator=gen)
        b0 = torch.randn(8, generator=gen)
        grads = [(torch.randn(8, 4, 3, 3, generator=gen), torch.randn(8, generator=gen)) for _ in range(12)]
        pw, pb = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(b0.clone())
        with _w.catch_warnings():
            _w.simplefilter("ignore")
            opt = RAdam([dict(params=[pw], weight_decay=1e-2), dict(params=[pb], weight_decay=0.0)], lr=1e-2)
            ws, bs = [], []
            for gw, gb in grads:
                pw.grad, pb.grad = gw.clone(), gb.clone()
                opt.step()
                ws.append(np32(pw)); bs.append(np32(pb))
        np.savez_compressed(os.path.join(HERE, "optim_radam.npz"), w0=np32(w0), b0=np32(b0),
                            gw=np.stack([np32(g[0]) for g in grads]), gb=np.stack([np32(g[1]) for g in grads]),
                            w=np.stack(ws), b=np.stack(bs), lr=np.float64(1e-2), wd=np.float64(1e-2))
        print("optim_radam: |w12 - w0| =", float((pw.detach() - w0).abs().max()))
    if only and not (only & {"blocks", "loss", "canny"}):
        return

    # ---- 3. block-level goldens (standalone modules, fwd + bwd) ---------
    def block_case(tag, module, inputs, call):
        sdm = module.state_dict()
        w = synth.synthetic_state_dict(sdm, seed=7)
        module.load_state_dict(w)
        module.train(True)
        ins = [t.clone().requires_grad_(True) for t in inputs]
        outs = call(module, ins)
        outs = outs if isinstance(outs, (tuple, list)) else (outs,)
        g = torch.Generator().manual_seed(11)
        cot = [torch.randn(o.shape, generator=g) for o in outs]
        sum((o * c).sum() for o, c in zip(outs, cot)).backward()
        out = {}
        for i, t in enumerate(inputs):
            out["in%d" % i] = np32(t)
            out["din%d" % i] = np32(ins[i].grad)
        for i, (o, c) in enumerate(zip(outs, cot)):
            out["out%d" % i] = np32(o)
            out["cot%d" % i] = np32(c)
        for k, p in module.named_parameters():
            out["grad/" + k] = np32(p.grad)
        for k, v in module.state_dict().items():
            if k.endswith(("running_mean", "running_var")):
                out["bn/" + k] = np32(v)
        out["keys"] = np.array(list(sdm.keys()))
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
        print(tag, [tuple(o.shape) for o in outs])

    g = torch.Generator().manual_seed(5)
    block_case("block_dualatt_c32_16", ref_att.DualAttBlock(inchannels=[32, 16], outchannels=32),
               [torch.randn(2, 32, 6, 5, generator=g), torch.randn(2, 16, 12, 10, generator=g)],
               lambda m, i: m([i[0], i[1]]))
    block_case("block_dualatt_c64_64", ref_att.DualAttBlock(inchannels=[64, 64], outchannels=64),
               [torch.randn(2, 64, 8, 8, generator=g), torch.randn(2, 64, 16, 16, generator=g)],
               lambda m, i: m([i[0], i[1]]))
    block_case("block_gsconv_c8", ref_gsc.GatedSpatialConv2d(8, 8),
               [torch.randn(2, 8, 12, 9, generator=g), torch.randn(2, 1, 12, 9, generator=g)],
               lambda m, i: m(i[0], i[1]))
    block_case("block_gsconv_c32", ref_gsc.GatedSpatialConv2d(32, 32),
               [torch.randn(2, 32, 16, 16, generator=g), torch.randn(2, 1, 16, 16, generator=g)],
               lambda m, i: m(i[0], i[1]))
    block_case("block_gsconv_c64", ref_gsc.GatedSpatialConv2d(64, 64),
               [torch.randn(1, 64, 8, 8, generator=g), torch.randn(1, 1, 8, 8, generator=g)],
               lambda m, i: m(i[0], i[1]))
    block_case("block_basic_c16", ref_resnet.BasicBlock(16, 16),
               [torch.randn(2, 16, 10, 12, generator=g)], lambda m, i: m(i[0]))
    block_case("block_decoder_64_48_32", ref_models.DecoderBlock(64, 48, 32),
               [torch.randn(2, 64, 6, 6, generator=g)], lambda m, i: m(i[0]))

    # ---- 4. loss goldens (loss.py:149-159) -------------------------------
    g = torch.Generator().manual_seed(21)
    seg = (2.0 * torch.randn(3, 4, 17, 13, generator=g)).requires_grad_(True)
    edge_p = torch.sigmoid(3.0 * torch.randn(3, 1, 17, 13, generator=g)).requires_grad_(True)
    seg_t = torch.randint(0, 4, (3, 17, 13), generator=g)
    edge_t = (torch.rand(3, 1, 17, 13, generator=g) > 0.8).float()
    loss = crit((seg, edge_p), (seg_t, edge_t))
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "loss_dual.npz"), seg=np32(seg), edge=np32(edge_p), seg_t=seg_t.numpy(),
                        edge_t=np32(edge_t), loss=np.float64(loss.item()), dseg=np32(seg.grad), dedge=np32(edge_p.grad),
                        dice=np.float64(ref_loss.dice_loss(seg_t, seg).item()))
    print("loss_dual", float(loss))

    # ---- 5. Canny goldens: the reference's own expression (models.py:359-362)
    import cv2
    data = synth.synthetic_batch(4, 256, seed=304)
    im = np.mean(data["image"].numpy(), axis=1).astype(np.uint8)
    canny = np.stack([cv2.Canny(im[i], 10, 100) for i in range(4)])
    rng = np.random.default_rng(3)
    extra_in, extra_out = [], []
    for t in range(6):
        h, w = int(rng.integers(8, 97)), int(rng.integers(8, 97))
        a = rng.integers(0, 256, (h, w)).astype(np.uint8)
        if t % 2:
            a = cv2.GaussianBlur(a, (7, 7), 2)
        extra_in.append(a)
        extra_out.append(cv2.Canny(a, 10, 100))
    np.savez_compressed(os.path.join(HERE, "canny_ref.npz"), im_u8=im, canny=canny,
                        **{"xin%d" % i: a for i, a in enumerate(extra_in)},
                        **{"xout%d" % i: a for i, a in enumerate(extra_out)})
    print("canny edge fraction", float((canny > 0).mean()))


if __name__ == "__main__":
    main()
