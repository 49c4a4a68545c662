"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (numpy, integer arithmetic) of the image path of the reference's inference loaders:

  * resize_linear_u8()  <- cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR) for uint8 images.  The arithmetic lives in
                           OpenCV (third-party dependency of the reference, requirements.txt `opencv-python`, not vendored;
                           this container: 4.13.0): modules/imgproc/src/resize.cpp, fixed point with 11 coefficient bits --
                           fx = (float)((dx + 0.5) * scale_x - 0.5); columns clamp fx / sx at the borders, rows clamp only
                           the row index; coefficients cvRound(c * 2048); horizontal pass in int32; vertical pass
                           (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
  * letterbox()         <- utils/datasets.py:611-646 (ratio, new_unpad, padding split, border colour 114) followed by
                           `img[:, :, ::-1].transpose(2, 0, 1)` of datasets.py:113, 184, 254.

Pinned against outputs of the reference's own letterbox (which calls cv2) on seeded images: tests/golden/letterbox_case.npz
(oracle/gen_golden_letterbox.py), tests/test_oracle_golden.py::test_preprocess_oracle_matches_reference.
"""
import numpy as np


def _coeffs(src, dst, clamp):
    scale = 1.0 / (dst / src)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo, hi = s < 0, s >= src - 1
        f = np.where(lo | hi, np.float32(0), f)
        s = np.where(lo, 0, np.where(hi, src - 1, s))
    a0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)       # cvRound: ties to even
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return s, a0, a1


def resize_linear_u8(img, width, height):
    """img HWC uint8 -> [height, width, C] uint8, bit-identical to cv2.resize(..., INTER_LINEAR)."""
    H, W = img.shape[:2]
    s = img.reshape(H, W, -1).astype(np.int64)
    sx, a0, a1 = _coeffs(W, width, True)
    sy, b0, b1 = _coeffs(H, height, False)
    x1 = np.minimum(sx + 1, W - 1)
    rows = s[:, sx, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
    y0, y1 = np.clip(sy, 0, H - 1), np.clip(sy + 1, 0, H - 1)
    out = (((b0[:, None, None] * (rows[y0] >> 4)) >> 16) + ((b1[:, None, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox(img, new_shape=(416, 416), color=114, auto=True, scaleFill=False, scaleup=True):
    """-> (uint8 [C, H, W] in reversed channel order (BGR -> RGB), ratio, (dw, dh))."""
    shape = img.shape[:2]
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:
        r = min(r, 1.0)
    ratio = r, r
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, 64), np.mod(dh, 64)
    elif scaleFill:
        dw, dh = 0.0, 0.0
        new_unpad = new_shape
        ratio = new_shape[0] / shape[1], new_shape[1] / shape[0]
    dw /= 2
    dh /= 2
    im = img.reshape(shape[0], shape[1], -1)
    if shape[::-1] != tuple(new_unpad):
        im = resize_linear_u8(im, new_unpad[0], new_unpad[1])
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    out = np.full((im.shape[0] + top + bottom, im.shape[1] + left + right, im.shape[2]), color, np.uint8)
    out[top:top + im.shape[0], left:left + im.shape[1]] = im
    return np.ascontiguousarray(out[:, :, ::-1].transpose(2, 0, 1)), ratio, (dw, dh)


def synth_image(h, w, c, seed):
    """Deterministic uint8 HWC image: smooth ramps + a pseudo-random texture (LCG, integer arithmetic only)."""
    y, x = np.mgrid[0:h, 0:w].astype(np.uint32)
    out = np.zeros((h, w, c), np.uint8)
    for k in range(c):
        n = (x * np.uint32(1103515245) + y * np.uint32(12345) + np.uint32(seed * 7919 + k * 104729)) & np.uint32(0xffffffff)
        n = (n * np.uint32(1664525) + np.uint32(1013904223)) & np.uint32(0xffffffff)
        ramp = (x * np.uint32(3 + k) + y * np.uint32(5 - k)) & np.uint32(255)
        out[:, :, k] = ((ramp + ((n >> np.uint32(24)) & np.uint32(63))) & np.uint32(255)).astype(np.uint8)
    return out
