"""The oracle (oracle/saunet_oracle.py, oracle/canny_oracle.c) against outputs
of the real reference committed by tests/golden/make_golden.py.  CPU only."""
import numpy as np
import pytest
import torch

from helpers import load_golden, template_state_dict, rel_err, alias_map
from saunet_b200 import synth
from oracle import saunet_oracle as O
from oracle import canny as ocanny

torch.set_num_threads(8)


def _weights():
    return synth.synthetic_state_dict(template_state_dict(), seed=0)


@pytest.mark.parametrize("tag,batch,size,training", [
    ("saunet_train_b2_s64", 2, 64, True),
    ("saunet_eval_b2_s64", 2, 64, False),
    ("saunet_eval_b1_s256", 1, 256, False),
])
def test_forward_matches_reference(tag, batch, size, training):
    g = load_golden(tag)
    data = synth.synthetic_batch(batch, size, seed=304)
    assert abs(float(data["image"].double().sum()) - float(g["image_sum"])) < 1e-6
    assert int(data["seg"].sum()) == int(g["seg_sum"])
    assert float(data["edge"].sum()) == float(g["edge_sum"])
    sd = O.prepare_params(_weights())
    with torch.no_grad():
        seg, edge = O.saunet_forward(sd, data["image"], training=training)
        loss = O.dual_loss(seg, edge, data["seg"], data["edge"])
    s = int(g["probe_stride"])
    # same torch ops on the same machine class: tolerance is fp32 round-off only
    assert rel_err(seg[:, :, ::s, ::s], g["logits"]) < 2e-5
    assert rel_err(edge[:, :, ::s, ::s], g["edge"]) < 2e-5
    assert abs(float(loss) - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))


def test_train_step_grads_and_bn_buffers():
    g = load_golden("saunet_train_b2_s64")
    data = synth.synthetic_batch(2, 64, seed=304)
    r = O.train_step(_weights(), data["image"], data["seg"], data["edge"])
    assert abs(float(r["loss"]) - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    alias = alias_map()
    names = [str(n) for n in g["grad_names"]]
    l2 = dict(zip(names, g["grad_l2"]))
    checked = 0
    for k, gr in r["grads"].items():
        if k in alias or k not in l2:     # oracle is driven by canonical names
            continue
        ref = float(l2[k])
        got = float(gr.double().norm())
        assert abs(got - ref) <= 2e-3 * max(ref, 1e-6) + 1e-7, (k, got, ref)
        checked += 1
    assert checked == len(names) - 0 and checked > 500
    for k in g:
        if k.startswith("grad/"):
            name = k[5:]
            assert rel_err(r["grads"][name], g[k]) < 5e-4, name
        if k.startswith("bn/"):
            assert rel_err(r["bn_updates"][k[3:]], g[k]) < 1e-5, k


def test_maps_and_metrics_match_reference():
    """return_att maps (models/models.py:386-393) and SegmentationModule's pixel accuracy / Jaccard (:51-74,92)."""
    g = load_golden("saunet_maps_b2_s64")
    data = synth.synthetic_batch(2, 64, seed=304)
    r = O.train_step(_weights(), data["image"], data["seg"], data["edge"], return_att=True)
    assert rel_err(r["logits"], g["logits"]) < 2e-5
    assert len(r["maps"]) == 7
    for i, t in enumerate(r["maps"]):
        assert t.shape == (2, 1, 64, 64) and rel_err(t, g["map%d" % i]) < 2e-5, i
    acc, jac = O.pixel_metrics(r["logits"], data["seg"], 4)
    assert abs(float(acc) - float(g["acc"])) < 1e-6
    assert np.allclose([float(j) for j in jac], g["jaccard"], atol=1e-6)


def test_b16_fixture_is_the_headline_config():
    """tests/golden/saunet_train_b16_s256.npz = BASELINE configs[1] (batch 16, 256x256, train mode) minted from the real
    reference; here only its bookkeeping is checked (the oracle needs ~25 s per step at this size: the GPU test
    compares the CUDA path with it directly)."""
    g = load_golden("saunet_train_b16_s256")
    assert g["logits"].shape == (16, 4, 32, 32) and int(g["probe_stride"]) == 8
    assert all(("map%d" % i) in g and g["map%d" % i].shape == (16, 1, 32, 32) for i in range(7))
    assert len(g["grad_names"]) > 500 and g["jaccard"].shape == (3,)
    data = synth.synthetic_batch(16, 256, seed=304)
    assert abs(float(data["image"].double().sum()) - float(g["image_sum"])) < 1e-3
    assert int(data["seg"].sum()) == int(g["seg_sum"]) and float(data["edge"].sum()) == float(g["edge_sum"])
    acc, jac = O.pixel_metrics(torch.from_numpy(g["logits"]), data["seg"][:, ::8, ::8], 4)   # probes only: sanity, not parity
    assert 0.0 <= float(acc) <= 1.0


def test_radam_oracle_matches_reference():
    """oracle/optim_oracle.py (radam.py:15-78 restated) against 12 steps of the real radam.RAdam."""
    from oracle import optim_oracle as OO
    g = load_golden("optim_radam")
    w, b = g["w0"].copy(), g["b0"].copy()
    st = [np.zeros_like(w), np.zeros_like(w), np.zeros_like(b), np.zeros_like(b)]
    crossed = False
    for t in range(12):
        n = OO.radam_step(w, g["gw"][t], st[0], st[1], t, float(g["lr"]), weight_decay=float(g["wd"]))
        OO.radam_step(b, g["gb"][t], st[2], st[3], t, float(g["lr"]))
        crossed |= n >= 5
        assert np.abs(w - g["w"][t]).max() < 2e-6 and np.abs(b - g["b"][t]).max() < 2e-6, t
    assert crossed


def test_loss_matches_reference():
    g = load_golden("loss_dual")
    seg = torch.from_numpy(g["seg"]).requires_grad_(True)
    edge = torch.from_numpy(g["edge"]).requires_grad_(True)
    loss = O.dual_loss(seg, edge, torch.from_numpy(g["seg_t"]), torch.from_numpy(g["edge_t"]))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    assert rel_err(seg.grad, g["dseg"]) < 1e-6
    assert rel_err(edge.grad, g["dedge"]) < 1e-6
    assert abs(float(O.dice_loss(torch.from_numpy(g["seg_t"]), seg)) - float(g["dice"])) < 1e-6


def test_canny_oracle_bit_exact_vs_cv2_fixture():
    g = load_golden("canny_ref")
    for i in range(g["im_u8"].shape[0]):
        assert np.array_equal(ocanny.canny_u8(g["im_u8"][i], 10, 100), g["canny"][i])
    i = 0
    while "xin%d" % i in g:
        assert np.array_equal(ocanny.canny_u8(g["xin%d" % i], 10, 100), g["xout%d" % i])
        i += 1
    assert i == 6


def test_image_to_u8_matches_reference_expression():
    g = load_golden("canny_ref")
    data = synth.synthetic_batch(4, 256, seed=304)
    assert np.array_equal(O.image_to_u8(data["image"]), g["im_u8"])
    cm = O.canny_map(data["image"])
    assert np.array_equal(cm[:, 0].numpy().astype(np.uint8), g["canny"])


@pytest.mark.parametrize("tag,kind", [
    ("block_dualatt_c32_16", "dualatt"), ("block_dualatt_c64_64", "dualatt"),
    ("block_gsconv_c8", "gsconv"), ("block_gsconv_c32", "gsconv"), ("block_gsconv_c64", "gsconv"),
    ("block_basic_c16", "basic"), ("block_decoder_64_48_32", "decoder"),
])
def test_blocks_match_reference(tag, kind):
    g = load_golden(tag)
    keys = [str(k) for k in g["keys"]]
    tmpl = {}
    for k in keys:
        if "grad/" + k in g:
            tmpl[k] = torch.zeros(g["grad/" + k].shape)
    # shapes of buffers are implied by the matching weight
    for k in keys:
        if k not in tmpl:
            base = k.rsplit(".", 1)[0]
            leaf = k.rsplit(".", 1)[1]
            if leaf == "num_batches_tracked":
                tmpl[k] = torch.zeros((), dtype=torch.int64)
            elif leaf == "_running_iter":
                tmpl[k] = torch.zeros(1)
            else:
                tmpl[k] = torch.zeros(g["grad/" + base + ".weight"].shape)
    tmpl = {k: tmpl[k] for k in keys}
    sd = O.prepare_params(synth.synthetic_state_dict(tmpl, seed=7), requires_grad=True)
    sd = {"m." + k: v for k, v in sd.items()}
    ins = [torch.from_numpy(g["in%d" % i]).requires_grad_(True) for i in range(2) if "in%d" % i in g]
    rec = O.BNRecorder()
    if kind == "dualatt":
        outs = O.dual_att_block(sd, "m", ins[0], ins[1], True, rec)
    elif kind == "gsconv":
        outs = O.gated_spatial_conv(sd, "m", ins[0], ins[1], True, rec)
    elif kind == "basic":
        outs = (O.basic_block(sd, "m", ins[0], True, rec),)
    else:
        outs = (O.decoder_block(sd, "m", ins[0], True, rec),)
    sum((o * torch.from_numpy(g["cot%d" % i])).sum() for i, o in enumerate(outs)).backward()
    for i, o in enumerate(outs):
        assert rel_err(o, g["out%d" % i]) < 1e-5
    for i, t in enumerate(ins):
        assert rel_err(t.grad, g["din%d" % i]) < 1e-4
    for k in g:
        if k.startswith("grad/"):
            assert rel_err(sd["m." + k[5:]].grad, g[k]) < 2e-4, k
        if k.startswith("bn/") and "_tmp" not in k:
            assert rel_err(rec.updates["m." + k[3:]], g[k]) < 1e-5, k
