"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (imported from /root/reference, CPU) on seeded
inputs.  Run here (the GPU box has no /root/reference); the outputs are committed.

    python oracle/gen_golden.py

The fixtures pin oracle/darknet_oracle.py (tests/test_oracle_golden.py) and are compared against the CUDA path on
the GPU box (tests/test_gpu_model.py).
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("B2Y_REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "mpl_stub"))
os.chdir(REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import models as ref_models  # noqa: E402  (the reference)
from utils import utils as ref_utils  # noqa: E402
from utils.quantized import quantized_ptq_cos as ref_ptq  # noqa: E402

sys.path.insert(0, HERE)
import darknet_oracle as orc  # noqa: E402

CFGS = {
    "yolov3-tiny": "cfg/yolov3tiny/yolov3-tiny.cfg",
    "yolov3": "cfg/yolov3/yolov3.cfg",
    "yolov4": "cfg/yolov4/yolov4.cfg",
}
HYP = dict(orc.DEFAULT_HYP)


def build(name, seed=0):
    m = ref_models.Darknet(CFGS[name])
    m.load_state_dict(orc.synth_state_dict(m.state_dict(), seed))
    return m


def attach_hyp(m, nc=80):
    m.nc, m.hyp, m.gr = nc, dict(HYP), 1.0
    m.hyp['cls'] = m.hyp['cls'] * nc / 80.0


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(GOLD, exist_ok=True)
    snap = {}
    for name in CFGS:
        m = build(name)
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "w.weights")
            ref_models.save_weights(m, p)
            wsize = os.path.getsize(p)
        snap[name] = {
            "modules": len(m.module_list),
            "convs": sum(1 for d in m.module_defs if d['type'] == 'convolutional'),
            "params": int(sum(p.numel() for p in m.parameters())),
            "yolo_layers": list(m.yolo_layers),
            "strides": [int(m.module_list[j].stride) for j in m.yolo_layers],
            "weights_bytes": wsize,
            "state_keys_head": list(m.state_dict().keys())[:8],
            "n_state_keys": len(m.state_dict()),
        }
    with open(os.path.join(GOLD, "cfg_snapshot.json"), "w") as f:
        json.dump(snap, f, indent=1, sort_keys=True)

    # ---- eval forwards -------------------------------------------------------------------------------------
    for name, (B, S) in {"yolov3-tiny": (1, 416), "yolov3": (2, 64), "yolov4": (2, 64)}.items():
        m = build(name).eval()
        x = orc.synth_images(B, S, S, seed=0)
        with torch.no_grad():
            io, p, _ = m(x)
        tag = "%s_%d_eval" % (name, S)
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), io=io.numpy(),
                            **{"p%d" % i: t.numpy() for i, t in enumerate(p)})
        print(tag, tuple(io.shape), float(io.abs().mean()))
    # a small tiny case too (cheap for every test run)
    m = build("yolov3-tiny").eval()
    with torch.no_grad():
        io, p, _ = m(orc.synth_images(2, 96, 96, seed=3))
    np.savez_compressed(os.path.join(GOLD, "yolov3-tiny_96_eval.npz"), io=io.numpy(),
                        **{"p%d" % i: t.numpy() for i, t in enumerate(p)})

    # ---- train forward + loss + backward ---------------------------------------------------------------------
    # (batch 4 @128x128: the deepest BatchNorm still sees 4*4*4 = 64 samples per channel, so the fp16 engine can be
    #  compared against these fp32 fixtures without the ill-conditioning of 8-sample batch statistics)
    for name, (B, S) in {"yolov3-tiny": (4, 128), "yolov3": (4, 128), "yolov4": (4, 128)}.items():
        m = build(name).train()
        attach_hyp(m)
        x = orc.synth_images(B, S, S, seed=0)
        t = orc.synth_targets(B, 6, 80, seed=1)
        pred, _ = m(x)
        loss, items = ref_utils.compute_loss(pred, t, m)
        loss.backward()
        sd = m.state_dict()
        out = {"loss": loss.detach().numpy(), "items": items.numpy()}
        for i, pi in enumerate(pred):
            out["p%d" % i] = pi.detach().numpy()
        # gradients: full tensors for a few layers + L2 norms for all (keeps the fixture small)
        names, norms = [], []
        for k, prm in m.named_parameters():
            names.append(k)
            norms.append(float(prm.grad.norm()) if prm.grad is not None else -1.0)
        out["grad_names"] = np.array(names)
        out["grad_norms"] = np.array(norms, dtype=np.float64)
        keep = [k for k in names if k.startswith("module_list.0.") or k.startswith("module_list.1.")]
        keep += names[-4:]
        for k in keep:
            out["grad::" + k] = dict(m.named_parameters())[k].grad.numpy()
        bn_keys = [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]
        for k in bn_keys[:4] + bn_keys[-2:]:
            out["stat::" + k] = sd[k].numpy()
        np.savez_compressed(os.path.join(GOLD, "%s_%d_train.npz" % (name, S)), **out)
        print(name, "train loss", float(loss), items.tolist())

    # ---- loss on random predictions (grids 2/4/8 at B=3) --------------------------------------------------
    m = build("yolov3")
    attach_hyp(m)
    g = torch.Generator().manual_seed(7)
    B = 3
    preds = [torch.randn(B, 3, s, s, 85, generator=g).requires_grad_(True) for s in (2, 4, 8)]
    for j, s in zip(m.yolo_layers, (2, 4, 8)):
        m.module_list[j].create_grids((s, s))
    t = orc.synth_targets(B, 10, 80, seed=5)
    loss, items = ref_utils.compute_loss(preds, t, m)
    loss.backward()
    tcls, tbox, indices, av = ref_utils.build_targets(preds, t, m)
    out = {"targets": t.numpy(), "loss": loss.detach().numpy(), "items": items.numpy()}
    for i in range(3):
        out["p%d" % i] = preds[i].detach().numpy()
        out["dp%d" % i] = preds[i].grad.numpy()
        out["tcls%d" % i] = tcls[i].numpy()
        out["tbox%d" % i] = tbox[i].numpy()
        out["idx%d" % i] = torch.stack(indices[i]).numpy()
        out["av%d" % i] = av[i].numpy()
        out["anchor_vec%d" % i] = m.module_list[m.yolo_layers[i]].anchor_vec.numpy()
    np.savez_compressed(os.path.join(GOLD, "loss_case.npz"), **out)
    print("loss case", float(loss), [len(c) for c in tcls])

    # ---- quantiser (COSPTQ, eval formula + calibration search) ---------------------------------------------
    g = torch.Generator().manual_seed(11)
    xq = torch.randn(4, 16, 9, 9, generator=g) * 3.0
    q = ref_ptq.Quantizer(bits=8, out_channels=-1)
    q.train()
    y_train = q(xq)
    scale_after = float(q.scale)
    q.eval()
    y_eval = q(xq * 1.7)
    np.savez_compressed(os.path.join(GOLD, "quant_case.npz"), x=xq.numpy(), y_train=y_train.numpy(),
                        scale=np.float32(scale_after), y_eval=y_eval.numpy(),
                        scale_list=np.array(q.scale_list), round_in=np.array([-2.5, -1.5, -0.5, 0.5, 1.5, 2.5, 0.49]),
                        round_out=ref_ptq.Round.apply(torch.tensor([-2.5, -1.5, -0.5, 0.5, 1.5, 2.5, 0.49])).numpy())
    print("quant scale", scale_after)

    # ---- fuse_conv_and_bn + Mish backward -----------------------------------------------------------------
    m = build("yolov3-tiny")
    blk = m.module_list[2]
    fused = ref_models.torch_utils.fuse_conv_and_bn(blk[0], blk[1])
    xm = torch.randn(2000, generator=g) * 5
    xm.requires_grad_(True)
    from utils.layers import Mish, MishImplementation
    ym = Mish()(xm)
    gm = torch.randn(2000, generator=g)
    ym.backward(gm)
    g2 = MishImplementation.backward(type("c", (), {"saved_tensors": (xm.detach(),)})(), gm)
    np.savez_compressed(os.path.join(GOLD, "misc_case.npz"), fused_w=fused.weight.detach().numpy(),
                        fused_b=fused.bias.detach().numpy(), mish_x=xm.detach().numpy(), mish_y=ym.detach().numpy(),
                        mish_g=gm.numpy(), mish_dx_autograd=xm.grad.numpy(), mish_dx_formula=g2.numpy())
    print("done ->", GOLD)


if __name__ == "__main__":
    main()
