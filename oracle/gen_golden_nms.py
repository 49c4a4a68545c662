"""Generate tests/golden/nms_modes_case.npz by RUNNING THE REFERENCE's non_max_suppression (utils/utils.py:782-860) on a
seeded synthetic prediction in the modes the yolov3-tiny fixture (map_case.npz) does not cover: best-class
(multi_label=False), class-agnostic, the `classes=` filter, a single-class model, score ties, rows rejected by the
size / finite filters and an image without candidates.  Pins oracle/metrics_oracle.nms in those modes
(tests/test_oracle_golden.py); the device path is compared against that oracle on the GPU (tests/test_gpu_nms.py).

    python oracle/gen_golden_nms.py
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

from utils import utils as ru  # noqa: E402  (the reference)

MODES = {
    "multi": dict(conf_thres=0.1, iou_thres=0.6),
    "best": dict(conf_thres=0.1, iou_thres=0.6, multi_label=False),
    "agnostic": dict(conf_thres=0.1, iou_thres=0.45, agnostic=True),
    "classes": dict(conf_thres=0.1, iou_thres=0.6, classes=[1, 4, 7]),
}


def synth(B, R, nc, seed):
    g = torch.Generator().manual_seed(seed)
    p = torch.zeros(B, R, 5 + nc)
    p[..., 0:2] = torch.rand(B, R, 2, generator=g) * 320
    p[..., 2:4] = torch.exp(torch.rand(B, R, 2, generator=g) * 3.5 + 1.0)
    p[..., 4] = torch.rand(B, R, generator=g) * 0.05
    p[..., 5:] = torch.rand(B, R, nc, generator=g) * 0.2
    levels = torch.tensor([0.25, 0.5, 0.75, 1.0])
    for b in range(B):
        rows = torch.randperm(R, generator=g)
        k = 0
        for _ in range(25):
            c = torch.rand(2, generator=g) * 240 + 40
            wh = torch.exp(torch.rand(2, generator=g) * 2.0 + 2.5)
            cls = int(torch.randint(0, nc, (1,), generator=g))
            for _ in range(8):
                r = int(rows[k]); k += 1
                p[b, r, 0:2] = c + torch.randn(2, generator=g) * 0.06 * wh
                p[b, r, 2:4] = wh * torch.exp(torch.randn(2, generator=g) * 0.08)
                p[b, r, 4] = levels[int(torch.randint(1, 4, (1,), generator=g))]
                p[b, r, 5:] = torch.rand(nc, generator=g) * 0.05
                p[b, r, 5 + cls] = levels[int(torch.randint(0, 4, (1,), generator=g))]
                if nc > 1:
                    p[b, r, 5 + (cls + 1) % nc] = levels[int(torch.randint(0, 3, (1,), generator=g))]
    p[0, 3, 0] = float("nan"); p[0, 3, 2:5] = torch.tensor([20.0, 20.0, 0.9])
    p[0, 4, 2:5] = torch.tensor([20.0, 20.0, 0.9]); p[0, 4, 5] = float("inf")
    p[0, 5, 2:5] = torch.tensor([5000.0, 20.0, 0.9])
    p[0, 6, 2:5] = torch.tensor([2.0, 20.0, 0.9])
    p[B - 1, :, 4] = 0.0                                   # an image without candidates
    return p


def main():
    out = {}
    for tag, nc in (("nc10", 10), ("nc1", 1)):
        pred = synth(3, 600, nc, seed=5 + nc)
        out["pred_" + tag] = pred.numpy()
        for mode, kw in MODES.items():
            if nc == 1 and mode == "classes":
                continue
            dets = ru.non_max_suppression(pred.clone(), **kw)
            for b, d in enumerate(dets):
                out["%s_%s_%d" % (tag, mode, b)] = np.zeros((0, 6), np.float32) if d is None else d.numpy()
            print(tag, mode, [0 if d is None else d.shape[0] for d in dets])
    path = os.path.join(ROOT, "tests", "golden", "nms_modes_case.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
