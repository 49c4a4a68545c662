"""GPU: device letterbox (inference slice) -- b2y_letterbox_u8 through the C ABI and
b200yolo.preprocess -- against the reference's own letterbox outputs (tests/golden/letterbox_case.npz: datasets.letterbox ->
cv2.resize INTER_LINEAR + cv2.copyMakeBorder, then BGR -> RGB / HWC -> CHW).  Bit-exact."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from helpers import ROOT, build_model, golden, orc

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from preprocess_oracle import synth_image  # noqa: E402

pytestmark = pytest.mark.gpu


def _cases():
    g = golden("letterbox_case")
    cases = json.loads(str(g["cases"]))
    for c in cases:
        if isinstance(c[4].get("new_shape"), list):
            c[4]["new_shape"] = tuple(c[4]["new_shape"])
    return g, cases


def test_letterbox_bit_exact_vs_reference():
    from b200yolo import preprocess
    g, cases = _cases()
    for i, (h, w, c, seed, kw) in enumerate(cases):
        img0 = synth_image(h, w, c, seed)
        src = img0 if i % 2 else torch.from_numpy(img0).cuda()          # host arrays and device tensors both
        out, ratio, pad = preprocess.letterbox(src, **kw)
        assert out.is_cuda and out.dtype == torch.uint8
        ref = g["out%d" % i]
        got = out.cpu().numpy()
        assert got.shape == ref.shape, (i, got.shape, ref.shape)
        assert np.array_equal(got, ref), "case %d %s: %d bytes differ from cv2" % (i, kw, int((got != ref).sum()))
        assert np.array_equal(np.array(ratio, np.float64), g["ratio%d" % i])
        assert np.array_equal(np.array(pad, np.float64), g["pad%d" % i])


def test_letterbox_batch_feeds_the_model():
    """[B, 3, H, W] uint8 batch written slot by slot == the per-image results; the model consumes it as it is (the stem
    kernel applies the reference's "/ 256") and gives the same output as the float tensor the reference would build."""
    from b200yolo import preprocess
    imgs = [synth_image(h, w, 3, s) for h, w, s in ((120, 160, 1), (233, 317, 3), (97, 131, 4))]
    batch, ratios, pads = preprocess.letterbox_batch(imgs, new_shape=128)
    assert batch.shape == (3, 3, 128, 128) and batch.dtype == torch.uint8
    for b, im in enumerate(imgs):
        single, r, p = preprocess.letterbox(im, 128, auto=False)
        assert torch.equal(batch[b], single) and r == ratios[b] and p == pads[b]
    model = build_model("yolov3-tiny", device="cuda").eval()
    with torch.no_grad():
        io_u8 = model(batch)[0].clone()
        io_f = model(batch.float() / 256.0)[0]
    assert torch.equal(io_u8, io_f)


def test_letterbox_rejects_bad_input():
    from b200yolo import preprocess
    with pytest.raises(ValueError):
        preprocess.letterbox(np.zeros((10, 10, 3), np.float32), 64)
    with pytest.raises(NotImplementedError):
        preprocess.letterbox(np.zeros((10, 10, 3), np.uint8), 64, color=(1, 2, 3))
