"""CPU: the device letterbox's per-pixel code (csrc/preprocess_core.h, shared by the CUDA kernel and this host harness)
and its host-side parameter logic against the reference's own letterbox outputs (tests/golden/letterbox_case.npz,
oracle/gen_golden_letterbox.py: datasets.letterbox -> cv2.resize / cv2.copyMakeBorder, then BGR -> RGB, HWC -> CHW).
Bit-exact.  The GPU test (tests/test_gpu_preprocess.py) runs the same cases through the kernel and the C ABI."""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from helpers import PKG, ROOT, golden

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from preprocess_oracle import synth_image  # noqa: E402  (seeded inputs of the fixture; integer arithmetic only)


def _cases():
    g = golden("letterbox_case")
    cases = json.loads(str(g["cases"]))
    for c in cases:
        if isinstance(c[4].get("new_shape"), list):
            c[4]["new_shape"] = tuple(c[4]["new_shape"])
    return g, cases


@pytest.fixture(scope="module")
def host_lib():
    d = tempfile.mkdtemp(prefix="b2y_lbhost_")
    so = os.path.join(d, "liblbhost.so")
    src = os.path.join(ROOT, "tests", "host", "letterbox_host.cpp")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(PKG, "csrc"), src, "-o", so],
                   check=True)
    return C.CDLL(so)


def test_letterbox_params_match_reference():
    from b200yolo.preprocess import letterbox_params
    g, cases = _cases()
    for i, (h, w, c, seed, kw) in enumerate(cases):
        (rw, rh), ratio, pad, (top, bottom, left, right) = letterbox_params((h, w), **kw)
        assert (c, rh + top + bottom, rw + left + right) == g["out%d" % i].shape
        assert np.array_equal(np.array(ratio, np.float64), g["ratio%d" % i])
        assert np.array_equal(np.array(pad, np.float64), g["pad%d" % i])


def test_device_pixel_code_on_host_matches_reference(host_lib):
    from b200yolo.preprocess import letterbox_params
    g, cases = _cases()
    for i, (h, w, c, seed, kw) in enumerate(cases):
        img0 = synth_image(h, w, c, seed)
        (rw, rh), _, _, (top, bottom, left, right) = letterbox_params((h, w), **kw)
        dh, dw = rh + top + bottom, rw + left + right
        out = np.zeros((c, dh, dw), np.uint8)
        rc = host_lib.b2y_letterbox_u8_host(img0.ctypes.data_as(C.c_void_p), h, w, c, C.c_longlong(w * c), rh, rw, top,
                                            left, out.ctypes.data_as(C.c_void_p), dh, dw, 1, 114)
        assert rc == 0
        assert np.array_equal(out, g["out%d" % i]), "case %d %s differs from cv2 in %d bytes" % (
            i, kw, int((out != g["out%d" % i]).sum()))


def test_resize_arithmetic_matches_cv2_on_random_sizes(host_lib):
    """Property check beyond the fixture: for 40 seeded (source, target) size pairs -- strong up- and down-scaling,
    1-pixel-wide sources, exact integer ratios, 1 and 3 channels -- the device per-pixel code (host harness) and the numpy
    oracle both reproduce cv2.resize(..., INTER_LINEAR) byte for byte.  OpenCV is what the reference's letterbox calls;
    it ships in this image (skipped where it does not)."""
    cv2 = pytest.importorskip("cv2")
    import preprocess_oracle as po
    rng = np.random.default_rng(123)
    pairs = [(1, 37, 64, 64), (37, 1, 64, 64), (2, 2, 96, 160), (480, 640, 480, 640), (300, 300, 100, 100),
             (300, 300, 900, 900), (511, 13, 64, 257), (1080, 1920, 135, 240)]
    while len(pairs) < 40:
        pairs.append(tuple(int(v) for v in (rng.integers(1, 400), rng.integers(1, 400), rng.integers(1, 400),
                                            rng.integers(1, 400))))
    for k, (h, w, dh, dw) in enumerate(pairs):
        c = 1 if k % 5 == 4 else 3
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        ref = cv2.resize(img if c == 3 else img[:, :, 0], (dw, dh), interpolation=cv2.INTER_LINEAR).reshape(dh, dw, c)
        if (h, w) == (dh, dw):
            assert np.array_equal(ref, img)
            continue
        assert np.array_equal(po.resize_linear_u8(img, dw, dh), ref), ("oracle", h, w, dh, dw, c)
        out = np.zeros((c, dh, dw), np.uint8)
        host_lib.b2y_letterbox_u8_host(img.ctypes.data_as(C.c_void_p), h, w, c, C.c_longlong(w * c), dh, dw, 0, 0,
                                       out.ctypes.data_as(C.c_void_p), dh, dw, 0, 114)
        assert np.array_equal(out.transpose(1, 2, 0), ref), ("device code", h, w, dh, dw, c)
