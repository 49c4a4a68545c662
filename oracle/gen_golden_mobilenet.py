"""Generate tests/golden/yolov3-mobilenet_{eval,train}.npz by RUNNING THE REFERENCE (CPU) on the MobileNetV3 backbone
config (BASELINE configs[3]: depthwise-separable convolutions + squeeze-excite).  Pins the
depthwise / SE branches of oracle/darknet_oracle.py ahead of the CUDA kernels for those rows (next round).

    python oracle/gen_golden_mobilenet.py

The fixture carries the parsed layer list (derived data, not the cfg text) and the parameter shapes, so the CPU test
needs neither the reference checkout nor its cfg files.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("B2Y_REFERENCE_ROOT", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "mpl_stub"))
os.chdir(REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import models as ref_models  # noqa: E402  (the reference)
from utils import utils as ru  # noqa: E402

sys.path.insert(0, HERE)
import darknet_oracle as orc  # noqa: E402

CFG = "cfg/yolov3-mobilenet/yolov3-mobilenet-coco.cfg"
NAME = "yolov3-mobilenet"


def jsonable(defs):
    out = []
    for d in defs:
        e = {}
        for k, v in d.items():
            e[k] = v.tolist() if isinstance(v, np.ndarray) else v
        out.append(e)
    return out


def main():
    model = ref_models.Darknet(CFG, img_size=(128, 128))
    sd = orc.synth_state_dict(model.state_dict(), 0)
    model.load_state_dict(sd)
    defs = jsonable(model.module_defs)
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    common = {"defs_json": np.array(json.dumps(defs)), "shapes_json": np.array(json.dumps(shapes))}

    # eval forward
    model.eval()
    x = orc.synth_images(2, 128, 128, seed=0)
    with torch.no_grad():
        io, p = model(x)[:2]
    out = dict(common, io=io.numpy())
    for i, pi in enumerate(p):
        out["p%d" % i] = pi.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", NAME + "_128_eval.npz"), **out)

    # training forward + loss + backward (batch 4)
    model.load_state_dict(sd)
    model.train()
    model.nc, model.gr, model.hyp = 80, 1.0, dict(orc.DEFAULT_HYP)
    x = orc.synth_images(4, 128, 128, seed=0)
    t = orc.synth_targets(4, 6, 80, seed=1)
    model.zero_grad()
    pred = model(x)[0]
    loss, items = ru.compute_loss(pred, t, model)
    loss.backward()
    names = [k for k, p_ in model.named_parameters() if p_.grad is not None]
    out = dict(common, items=items.numpy(), grad_names=np.array(names),
               grad_norms=np.array([float(dict(model.named_parameters())[k].grad.norm()) for k in names]))
    for i, pi in enumerate(pred):
        out["p%d" % i] = pi.detach().numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", NAME + "_128_train.npz"), **out)
    print("wrote fixtures: %d layers, %d parameters tensors, loss items %s" % (len(defs), len(shapes), items.tolist()))


if __name__ == "__main__":
    main()
