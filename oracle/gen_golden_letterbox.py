"""Generate tests/golden/letterbox_case.npz by RUNNING THE REFERENCE's letterbox (utils/datasets.py:611-646, which calls
cv2.resize / cv2.copyMakeBorder) on seeded uint8 images, followed by the loaders' BGR -> RGB / HWC -> CHW shuffle
(datasets.py:113).  Only the outputs are stored; the inputs are regenerated from the seeds (preprocess_oracle.synth_image,
integer arithmetic only).  Run here (the GPU box has no /root/reference); the output is committed.

    python oracle/gen_golden_letterbox.py
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

import numpy as np  # noqa: E402

sys.path.insert(0, HERE)
from preprocess_oracle import synth_image  # noqa: E402

CASES = [
    # height, width, channels, seed, letterbox kwargs
    (120, 160, 3, 1, dict(new_shape=416)),                                   # up-scale, auto (minimum rectangle)
    (375, 500, 3, 2, dict(new_shape=320)),                                   # down-scale, auto
    (233, 317, 3, 3, dict(new_shape=(256, 320), auto=False)),                # rectangular target, odd sizes
    (97, 131, 3, 4, dict(new_shape=224, auto=False, scaleup=False)),         # no resize: pure border + shuffle
    (60, 200, 3, 5, dict(new_shape=192, auto=False, scaleFill=True)),        # stretch
    (150, 150, 3, 6, dict(new_shape=300, auto=False)),                       # exact 2x up
    (480, 640, 3, 7, dict(new_shape=320, auto=False)),                       # exact 2x down
    (7, 5, 3, 8, dict(new_shape=64, auto=False)),                            # tiny source: border rows / columns dominate
]


def main():
    os.chdir(REF)
    np.int = int                               # datasets.py:278 still uses the alias removed from numpy
    from utils.datasets import letterbox      # noqa: E402  (the reference)
    out = {"cases": np.array(json.dumps(CASES))}
    for i, (h, w, c, seed, kw) in enumerate(CASES):
        img0 = synth_image(h, w, c, seed)
        img, ratio, pad = letterbox(img0, **kw)
        img = np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1))
        out["out%d" % i] = img
        out["ratio%d" % i] = np.array(ratio, np.float64)
        out["pad%d" % i] = np.array(pad, np.float64)
        print(i, (h, w, c), kw, "->", img.shape, ratio, pad)
    path = os.path.join(ROOT, "tests", "golden", "letterbox_case.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
