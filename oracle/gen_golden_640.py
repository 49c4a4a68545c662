"""Reference-generated fixtures at the BASELINE sizes (640x640), by RUNNING THE REFERENCE ITSELF on CPU (fp32).

    python oracle/gen_golden_640.py [eval] [train]

  tests/golden/yolov3_640_eval.npz   yolov3.cfg  eval forward, batch 2 (BASELINE configs[1] geometry)
  tests/golden/yolov4_640_eval.npz   yolov4.cfg  eval forward, batch 2
  tests/golden/yolov4_640_train.npz  yolov4.cfg  training step (forward + compute_loss + backward), batch 8
                                     = the per-GPU slice of BASELINE configs[2]

The full outputs are tens of MB, so the fixtures keep every ROW_STRIDE-th decoded row (row = (anchor, y, x) of
io.view(B, -1, 85), all 85 columns), float64 sums of every head, the loss items, the L2 norm of every parameter
gradient, a few full gradient tensors and BatchNorm running statistics.  They pin oracle/darknet_oracle.py at these sizes
(tests/test_oracle_golden.py) and are compared with the CUDA path on the GPU box (tests/test_gpu_baseline_sizes.py).
"""
import os
import sys

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

sys.path.insert(0, HERE)
import darknet_oracle as orc  # noqa: E402

CFGS = {"yolov3": "cfg/yolov3/yolov3.cfg", "yolov4": "cfg/yolov4/yolov4.cfg"}
ROW_STRIDE = 29
SIZE = 640


def build(name, seed=0):
    m = ref_models.Darknet(CFGS[name], img_size=(SIZE, SIZE))
    m.load_state_dict(orc.synth_state_dict(m.state_dict(), seed))
    return m


def gen_eval(name, B=2):
    m = build(name).eval()
    x = orc.synth_images(B, SIZE, SIZE, seed=0)
    with torch.no_grad():
        io, p, _ = m(x)
    rows = np.arange(0, io.shape[1], ROW_STRIDE)
    out = {"rows": rows, "io_rows": io[:, rows].numpy(), "io_shape": np.array(io.shape),
           "io_sum": io.double().sum(dim=(1, 2)).numpy(), "io_abs_sum": io.double().abs().sum(dim=(1, 2)).numpy()}
    off = 0
    for i, pi in enumerate(p):
        flat = pi.reshape(B, -1, pi.shape[-1])
        n = flat.shape[1]
        sel = rows[(rows >= off) & (rows < off + n)] - off
        out["p%d_rows" % i] = flat[:, sel].numpy()
        out["p%d_sum" % i] = pi.double().sum().numpy()
        off += n
    np.savez_compressed(os.path.join(GOLD, "%s_%d_eval.npz" % (name, SIZE)), **out)
    print(name, "eval", tuple(io.shape), out["io_sum"])


def gen_train(name="yolov4", B=8):
    m = build(name).train()
    m.nc, m.hyp, m.gr = 80, dict(orc.DEFAULT_HYP), 1.0
    x = orc.synth_images(B, SIZE, SIZE, seed=0)
    t = orc.synth_targets(B, 8, 80, seed=1)
    pred, _ = m(x)
    loss, items = ref_utils.compute_loss(pred, t, m)
    loss.backward()
    out = {"loss": loss.detach().numpy(), "items": items.numpy()}
    for i, pi in enumerate(pred):
        flat = pi.detach().reshape(B, -1, pi.shape[-1])
        sel = np.arange(0, flat.shape[1], ROW_STRIDE)
        out["p%d_rows" % i] = flat[:, sel].numpy()
        out["p%d_sum" % i] = pi.detach().double().sum().numpy()
        out["p%d_abs_sum" % i] = pi.detach().double().abs().sum().numpy()
    named = dict(m.named_parameters())
    names = list(named)
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([float(named[k].grad.norm()) for k in names], dtype=np.float64)
    out["grad_absmax"] = np.array([float(named[k].grad.abs().max()) for k in names], dtype=np.float64)
    # full tensors: everything small (BatchNorm / bias vectors of all layers) + the first two and the last conv weights
    keep = [k for k in names if named[k].numel() <= 4096]
    keep += [k for k in names if k.startswith(("module_list.0.Conv2d", "module_list.1.Conv2d"))] + [names[-2]]
    for k in dict.fromkeys(keep):
        out["grad::" + k] = named[k].grad.numpy()
    sd = m.state_dict()
    bn_keys = [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]
    for k in bn_keys:
        out["stat::" + k] = sd[k].numpy()
    np.savez_compressed(os.path.join(GOLD, "%s_%d_train.npz" % (name, SIZE)), **out)
    print(name, "train loss", float(loss), items.tolist())


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    what = sys.argv[1:] or ["eval", "train"]
    if "eval" in what:
        for n in CFGS:
            gen_eval(n)
    if "train" in what:
        gen_train()
