"""CPU: pin the oracle (oracle/darknet_oracle.py) against fixtures produced by running the reference itself
(oracle/gen_golden.py).  Tolerances are fp32 re-association only."""
import numpy as np
import pytest
import torch

from helpers import anchor_vecs, golden, module_defs, orc


def _state(name, requires_grad=False):
    import models
    from helpers import cfg_path
    m = models.Darknet(cfg_path(name))
    sd = orc.synth_state_dict(m.state_dict(), 0)
    if requires_grad:
        for k, v in sd.items():
            if v.dtype.is_floating_point and not k.endswith(('running_mean', 'running_var')):
                v.requires_grad_(True)
    return sd


@pytest.mark.parametrize("name,B,S,seed", [("yolov3-tiny", 2, 96, 3), ("yolov3", 2, 64, 0), ("yolov4", 2, 64, 0),
                                           ("yolov3-tiny", 1, 416, 0)])
def test_eval_forward_matches_reference(name, B, S, seed):
    g = golden("%s_%d_eval" % (name, S))
    x = orc.synth_images(B, S, S, seed=seed)
    with torch.no_grad():
        io, p = orc.darknet_forward(module_defs(name), _state(name), x, name)
    assert io.shape == g["io"].shape
    np.testing.assert_allclose(io.numpy(), g["io"], rtol=2e-4, atol=2e-4)
    for i, pi in enumerate(p):
        np.testing.assert_allclose(pi.numpy(), g["p%d" % i], rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("name", ["yolov3-tiny", "yolov3", "yolov4"])
def test_train_forward_loss_backward_matches_reference(name):
    g = golden("%s_128_train" % name)
    sd = _state(name, requires_grad=True)
    x = orc.synth_images(4, 128, 128, seed=0)
    t = orc.synth_targets(4, 6, 80, seed=1)
    p, stats = orc.darknet_forward(module_defs(name), sd, x, name, training=True)
    for i, pi in enumerate(p):
        np.testing.assert_allclose(pi.detach().numpy(), g["p%d" % i], rtol=1e-3, atol=1e-3)
    hyp = dict(orc.DEFAULT_HYP)
    loss, items = orc.compute_loss(p, t, anchor_vecs(name), hyp, 80, 1.0)
    np.testing.assert_allclose(items.numpy(), g["items"], rtol=1e-4)
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    norms = dict(zip(names, g["grad_norms"]))
    for k in names:
        if ("grad::" + k) in g.files:
            np.testing.assert_allclose(sd[k].grad.numpy(), g["grad::" + k], rtol=2e-3, atol=2e-4 * (1 + norms[k]))
    worst = max(abs(float(sd[k].grad.norm()) - norms[k]) / (norms[k] + 1e-6) for k in names if norms[k] > 0)
    assert worst < 5e-3, worst
    for k in g.files:
        if k.startswith("stat::"):
            np.testing.assert_allclose(stats[k[6:]].numpy(), g[k], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["yolov3", "yolov4"])
def test_eval_forward_640_matches_reference(name):
    """BASELINE size (640x640, batch 2): every 29th decoded row + float64 sums of the reference's own output
    (oracle/gen_golden_640.py)."""
    g = golden("%s_640_eval" % name)
    x = orc.synth_images(2, 640, 640, seed=0)
    with torch.no_grad():
        io, p = orc.darknet_forward(module_defs(name), _state(name), x, name)
    assert tuple(io.shape) == tuple(g["io_shape"])
    rows = g["rows"]
    np.testing.assert_allclose(io[:, rows].numpy(), g["io_rows"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(io.double().sum(dim=(1, 2)).numpy(), g["io_sum"], rtol=1e-5)
    np.testing.assert_allclose(io.double().abs().sum(dim=(1, 2)).numpy(), g["io_abs_sum"], rtol=1e-5)
    off = 0
    for i, pi in enumerate(p):
        flat = pi.reshape(2, -1, pi.shape[-1])
        sel = rows[(rows >= off) & (rows < off + flat.shape[1])] - off
        np.testing.assert_allclose(flat[:, sel].numpy(), g["p%d_rows" % i], rtol=2e-4, atol=2e-4)
        off += flat.shape[1]


def test_train_step_640_matches_reference():
    """The per-GPU slice of BASELINE configs[2]: yolov4.cfg, batch 8, 640x640, forward + compute_loss + backward of the
    reference itself (fixture) vs the oracle: sampled predictions, loss items, every gradient norm, all BatchNorm /
    bias gradients element-wise, all running statistics."""
    name = "yolov4"
    g = golden("yolov4_640_train")
    sd = _state(name, requires_grad=True)
    x = orc.synth_images(8, 640, 640, seed=0)
    t = orc.synth_targets(8, 8, 80, seed=1)
    p, stats = orc.darknet_forward(module_defs(name), sd, x, name, training=True)
    for i, pi in enumerate(p):
        flat = pi.detach().reshape(8, -1, pi.shape[-1])
        np.testing.assert_allclose(flat[:, ::29].numpy(), g["p%d_rows" % i], rtol=1e-3, atol=1e-3)
        np.testing.assert_allclose(float(pi.detach().double().abs().sum()), float(g["p%d_abs_sum" % i]), rtol=1e-4)
    loss, items = orc.compute_loss(p, t, anchor_vecs(name), dict(orc.DEFAULT_HYP), 80, 1.0)
    np.testing.assert_allclose(items.numpy(), g["items"], rtol=1e-4)
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    norms = dict(zip(names, g["grad_norms"]))
    worst = max(abs(float(sd[k].grad.norm()) - norms[k]) / (norms[k] + 1e-6) for k in names if norms[k] > 0)
    assert worst < 5e-3, worst
    for k in names:
        if ("grad::" + k) in g.files:
            np.testing.assert_allclose(sd[k].grad.numpy(), g["grad::" + k], rtol=5e-3, atol=5e-4 * (1 + norms[k]))
    for k in g.files:
        if k.startswith("stat::"):
            np.testing.assert_allclose(stats[k[6:]].numpy(), g[k], rtol=1e-4, atol=1e-5)


def test_loss_and_build_targets_match_reference():
    g = golden("loss_case")
    p = [torch.from_numpy(g["p%d" % i]).requires_grad_(True) for i in range(3)]
    av = [torch.from_numpy(g["anchor_vec%d" % i]) for i in range(3)]
    t = torch.from_numpy(g["targets"])
    loss, items = orc.compute_loss(p, t, av, dict(orc.DEFAULT_HYP), 80, 1.0)
    np.testing.assert_allclose(items.numpy(), g["items"], rtol=1e-5)
    loss.backward()
    for i in range(3):
        np.testing.assert_allclose(p[i].grad.numpy(), g["dp%d" % i], rtol=1e-4, atol=1e-6)
    tcls, tbox, idx, avs = orc.build_targets([x.detach() for x in p], t, av, 0.20)
    for i in range(3):
        assert np.array_equal(torch.stack(idx[i]).numpy(), g["idx%d" % i])       # int64, bit exact
        assert np.array_equal(tcls[i].numpy(), g["tcls%d" % i])
        np.testing.assert_array_equal(tbox[i].numpy(), g["tbox%d" % i])
        np.testing.assert_array_equal(avs[i].numpy(), g["av%d" % i])


def test_quantizer_matches_reference():
    g = golden("quant_case")
    x = torch.from_numpy(g["x"])
    best_i, _ = orc.cos_scale_search(x, 8)
    assert best_i == int(np.argmax(g["scale_list"]))
    scale = orc.pow2_scale(best_i - 5, 8)
    assert scale == float(g["scale"])
    np.testing.assert_array_equal(orc.fake_quant(x, scale, 8).numpy(), g["y_train"])
    np.testing.assert_array_equal(orc.fake_quant(x * 1.7, scale, 8).numpy(), g["y_eval"])
    np.testing.assert_array_equal(orc.round_half_away(torch.from_numpy(g["round_in"]).float()).numpy(), g["round_out"])


def test_bn_fold_and_mish_backward_match_reference():
    g = golden("misc_case")
    sd = _state("yolov3-tiny")
    pre = "module_list.2."
    wf, bf = orc.fold_bn(sd[pre + "Conv2d.weight"], None, sd[pre + "BatchNorm2d.weight"], sd[pre + "BatchNorm2d.bias"],
                         sd[pre + "BatchNorm2d.running_mean"], sd[pre + "BatchNorm2d.running_var"], 1e-5)
    np.testing.assert_allclose(wf.numpy(), g["fused_w"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(bf.numpy(), g["fused_b"], rtol=1e-5, atol=1e-6)
    x = torch.from_numpy(g["mish_x"])
    np.testing.assert_allclose(orc.activation(x, 'mish').numpy(), g["mish_y"], rtol=1e-6, atol=1e-7)
    dx = orc.mish_backward(x, torch.from_numpy(g["mish_g"]))
    np.testing.assert_allclose(dx.numpy(), g["mish_dx_formula"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dx.numpy(), g["mish_dx_autograd"], rtol=1e-4, atol=1e-5)


def test_metrics_oracle_matches_reference():
    """oracle/metrics_oracle.py (NMS, clip, TP matching, AP) against every intermediate the reference produced for the
    held batch (oracle/gen_golden_map.py -> map_case.npz): detections bit for bit, TP matrices, AP table, mAP."""
    from oracle import metrics_oracle as mo
    g = golden("map_case")
    pred = torch.from_numpy(g["inf_out"])
    S = int(g["size"])
    dets = mo.nms(pred, float(g["conf_thres"]), float(g["iou_thres"]))
    labels = []
    for i, d in enumerate(dets):
        mo.clip_boxes(d, S, S)
        ref = torch.from_numpy(g["det%d" % i])
        assert d.shape == ref.shape
        assert torch.equal(d[:, 5], ref[:, 5])
        assert torch.allclose(d, ref, rtol=0, atol=1e-4)
        lab = torch.from_numpy(g["labels%d" % i])
        labels.append(lab)
        assert np.array_equal(mo.match_image(d, lab).numpy(), g["correct%d" % i])
    m50, m, n = mo.mean_ap(pred, labels, float(g["conf_thres"]), float(g["iou_thres"]), S, S)
    assert n == sum(g["det%d" % i].shape[0] for i in range(len(dets)))
    assert abs(m50 - float(g["map50"])) < 1e-9 and abs(m - float(g["map"])) < 1e-9


def _mobilenet_case(kind):
    import json
    g = golden("yolov3-mobilenet_128_%s" % kind)
    defs = json.loads(str(g["defs_json"]))
    for d in defs:
        if "anchors" in d:
            d["anchors"] = np.asarray(d["anchors"])
    shapes = json.loads(str(g["shapes_json"]))
    state = orc.synth_state_dict({k: torch.zeros(v, dtype=torch.long if k.endswith("num_batches_tracked") else
                                                 torch.float32) for k, v in shapes.items()}, 0)
    return g, defs, state


def test_mobilenet_depthwise_se_forward_matches_reference():
    """SURVEY 8a rows a5 (depthwise blocks) and a8 (squeeze-excite): the oracle's branches against the reference's
    eval forward of yolov3-mobilenet (15 depthwise layers k3/k5 s1/s2, 8 SE blocks, relu6 / h_swish)."""
    g, defs, state = _mobilenet_case("eval")
    x = orc.synth_images(2, 128, 128, seed=0)
    with torch.no_grad():
        io, p = orc.darknet_forward(defs, state, x, "yolov3-mobilenet")
    assert io.shape == g["io"].shape
    assert torch.allclose(io, torch.from_numpy(g["io"]), rtol=1e-4, atol=1e-4)
    for i, pi in enumerate(p):
        assert torch.allclose(pi, torch.from_numpy(g["p%d" % i]), rtol=1e-4, atol=2e-4)


def test_mobilenet_training_step_matches_reference():
    g, defs, state = _mobilenet_case("train")
    for k, v in state.items():
        if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    x = orc.synth_images(4, 128, 128, seed=0)
    t = orc.synth_targets(4, 6, 80, seed=1)
    pred, _ = orc.darknet_forward(defs, state, x, "yolov3-mobilenet", training=True)
    for i, pi in enumerate(pred):
        assert torch.allclose(pi, torch.from_numpy(g["p%d" % i]), rtol=1e-3, atol=1e-3)
    anchors = []
    strides = orc.yolo_strides("yolov3-mobilenet", 3)
    for d in defs:
        if d["type"] == "yolo":
            anchors.append(torch.from_numpy(np.asarray(d["anchors"])[d["mask"]]).float() / strides[len(anchors)])
    loss, items = orc.compute_loss(pred, t, anchors, dict(orc.DEFAULT_HYP), 80, 1.0)
    assert np.allclose(items.detach().numpy(), g["items"], rtol=1e-3)
    loss.backward()
    ref = dict(zip([str(n) for n in g["grad_names"]], g["grad_norms"]))
    worst = max(abs(float(state[k].grad.norm()) - v) / (v + 1e-8) for k, v in ref.items())
    assert worst < 2e-2, worst


def test_preprocess_oracle_matches_reference():
    """oracle/preprocess_oracle.py (numpy restatement of OpenCV's fixed-point 8-bit INTER_LINEAR resize + the reference's
    letterbox logic) against the outputs of the reference's own letterbox (cv2) on seeded images."""
    import json
    import os
    import sys
    from helpers import ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import preprocess_oracle as po
    g = golden("letterbox_case")
    for i, (h, w, c, seed, kw) in enumerate(json.loads(str(g["cases"]))):
        if isinstance(kw.get("new_shape"), list):
            kw["new_shape"] = tuple(kw["new_shape"])
        out, ratio, pad = po.letterbox(po.synth_image(h, w, c, seed), **kw)
        assert np.array_equal(out, g["out%d" % i]), (i, kw)
        assert np.array_equal(np.array(ratio, np.float64), g["ratio%d" % i])
        assert np.array_equal(np.array(pad, np.float64), g["pad%d" % i])


def test_metrics_oracle_nms_modes_match_reference():
    """oracle/metrics_oracle.nms in the modes map_case.npz does not exercise (best-class, class-agnostic, `classes=` filter,
    single-class model, non-finite / out-of-range rows, an image without candidates) against the reference's own
    non_max_suppression outputs (oracle/gen_golden_nms.py -> nms_modes_case.npz).  The kept set, its order, scores and
    classes must be identical; merged boxes agree to the summation order of torch.mm (same library here: exact)."""
    from oracle import metrics_oracle as mo
    g = golden("nms_modes_case")
    modes = {
        "multi": dict(conf_thres=0.1, iou_thres=0.6),
        "best": dict(conf_thres=0.1, iou_thres=0.6, multi_label=False),
        "agnostic": dict(conf_thres=0.1, iou_thres=0.45, agnostic=True),
        "classes": dict(conf_thres=0.1, iou_thres=0.6, classes=[1, 4, 7]),
    }
    checked = 0
    for tag in ("nc10", "nc1"):
        pred = torch.from_numpy(g["pred_" + tag])
        for mode, kw in modes.items():
            if ("%s_%s_0" % (tag, mode)) not in g.files:
                continue
            dets = mo.nms(pred.clone(), **kw)
            for b, d in enumerate(dets):
                ref = g["%s_%s_%d" % (tag, mode, b)]
                if ref.shape[0] == 0:
                    assert d is None, (tag, mode, b)
                    continue
                assert d is not None and tuple(d.shape) == ref.shape, (tag, mode, b)
                assert np.array_equal(d[:, 4:].numpy(), ref[:, 4:]), (tag, mode, b)
                np.testing.assert_allclose(d[:, :4].numpy(), ref[:, :4], rtol=1e-6, atol=1e-5)
                checked += 1
    assert checked >= 12
