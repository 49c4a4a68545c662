"""Generate tests/golden/kd_case.npz by RUNNING THE REFERENCE's knowledge-distillation losses (utils/utils.py:435-520:
compute_lost_KD, compute_lost_KD2, compute_lost_KD3) on seeded student / teacher head outputs, with autograd gradients
w.r.t. the student.  Run here (the GPU box has no /root/reference); the output is committed.

    python oracle/gen_golden_kd.py
"""
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


def main():
    torch.manual_seed(0)
    m = ref_models.Darknet("cfg/yolov3/yolov3.cfg")
    m.nc, m.hyp, m.gr = 80, dict(orc.DEFAULT_HYP), 1.0
    g = torch.Generator().manual_seed(21)
    B, grids = 3, (2, 4, 8)
    for j, s in zip(m.yolo_layers, grids):
        m.module_list[j].create_grids((s, s))
    stu = [torch.randn(B, 3, s, s, 85, generator=g) for s in grids]
    tea = [(p + 0.7 * torch.randn(p.shape, generator=g)) for p in stu]          # a teacher near the student
    t = orc.synth_targets(B, 10, 80, seed=5)
    t = torch.cat((t, t[:4]), 0)                                                # duplicate labels: cells matched twice
    out = {"targets": t.numpy()}
    for i in range(3):
        out["s%d" % i] = stu[i].numpy()
        out["t%d" % i] = tea[i].numpy()
        out["anchor_vec%d" % i] = m.module_list[m.yolo_layers[i]].anchor_vec.numpy()

    def run(tag, fn):
        ps = [p.clone().requires_grad_(True) for p in stu]
        pt = [p.clone() for p in tea]
        res = fn(ps, pt)
        extra = None
        if isinstance(res, tuple):
            res, extra = res
        res.sum().backward()
        out[tag + "_loss"] = res.detach().numpy()
        if extra is not None:
            out[tag + "_ratio"] = np.float64(extra)
        for i in range(3):
            out["%s_ds%d" % (tag, i)] = ps[i].grad.numpy()
        print(tag, res.detach().reshape(-1).tolist(), extra)

    run("kd1", lambda ps, pt: ru.compute_lost_KD(ps, pt, m.nc, B))
    run("kd2", lambda ps, pt: ru.compute_lost_KD2(m, t, ps, pt))
    run("kd3", lambda ps, pt: ru.compute_lost_KD3(m, t, ps, pt))
    path = os.path.join(ROOT, "tests", "golden", "kd_case.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
