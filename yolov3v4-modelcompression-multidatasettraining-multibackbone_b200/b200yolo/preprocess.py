"""Input pipeline on the device, inference slice: the reference's `letterbox`
(utils/datasets.py:611-646) and the BGR -> RGB / HWC -> CHW shuffle of its loaders (datasets.py:108-118, 180-186, 250-256)
as ONE kernel (csrc/preprocess.cu) that writes the uint8 NCHW tensor the model consumes (the stem kernel applies the
reference's `/ 256`).  Bit-exact against cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder, which is what the reference calls.

    img, ratio, pad = letterbox(img0, new_shape=416)                # img0: HWC uint8 BGR (cv2.imread) on the device
    batch = letterbox_batch([img0, img1, ...], new_shape=640)       # [B, 3, 640, 640] uint8, auto=False
    pred = model(batch)                                             # uint8 in: "/ 256" happens inside the first conv

Host images (numpy arrays / CPU tensors) are copied to the device as they are (raw BGR bytes, the smallest form of the
image); resizing, padding, channel shuffle and the float conversion never touch the CPU.
"""
import ctypes as C

import numpy as np
import torch

from . import lib
from .lib import call, ptr, stream_ptr


def letterbox_params(shape, new_shape=(416, 416), auto=True, scaleFill=False, scaleup=True):
    """The host-side integer logic of datasets.py:614-640 for a source of `shape` = (height, width):
    returns (new_unpad (w, h), ratio, (dw, dh), (top, bottom, left, right))."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:                                  # only scale down (better test mAP)
        r = min(r, 1.0)
    ratio = r, r
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:                                         # minimum rectangle
        dw, dh = np.mod(dw, 64), np.mod(dh, 64)
    elif scaleFill:                                  # stretch (the reference passes (h, w) to cv2 here; kept)
        dw, dh = 0.0, 0.0
        new_unpad = new_shape
        ratio = new_shape[0] / shape[1], new_shape[1] / shape[0]
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad, ratio, (dw, dh), (top, bottom, left, right)


def _device_image(img0, device):
    if isinstance(img0, np.ndarray):
        img0 = torch.from_numpy(np.ascontiguousarray(img0))
    if img0.dim() == 2:
        img0 = img0[:, :, None]
    if img0.dtype != torch.uint8 or img0.dim() != 3:
        raise ValueError("letterbox expects an HWC uint8 image (cv2.imread layout)")
    if not img0.is_cuda:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        img0 = img0.contiguous().to(device, non_blocking=True)
    return img0.contiguous()


def letterbox(img0, new_shape=(416, 416), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True,
              is_gray_scale=False, out=None, device=None, to_rgb=True):
    """Drop-in for datasets.letterbox + the loaders' `img[:, :, ::-1].transpose(2, 0, 1)`: returns
    (uint8 [C, H, W] device tensor in RGB order, ratio, (dw, dh)).  `out` = a preallocated [C, H, W] uint8 device view
    (e.g. one slot of a batch) to write into."""
    if not torch.cuda.is_available():
        raise lib.B2YError("letterbox runs on the CUDA device only (no CPU fallback)")
    c0 = color[0] if isinstance(color, (tuple, list)) else color
    if isinstance(color, (tuple, list)) and any(int(c) != int(c0) for c in color):
        raise NotImplementedError("letterbox border colour must be the same on all channels (the reference uses 114)")
    src = _device_image(img0, device if out is None else out.device)
    H, W, Cn = src.shape
    (rw, rh), ratio, pad, (top, bottom, left, right) = letterbox_params((H, W), new_shape, auto, scaleFill, scaleup)
    dh, dw = rh + top + bottom, rw + left + right
    if out is None:
        out = torch.empty((Cn, dh, dw), dtype=torch.uint8, device=src.device)
    if tuple(out.shape) != (Cn, dh, dw) or out.dtype != torch.uint8 or not out.is_contiguous():
        raise ValueError("letterbox: `out` must be a contiguous uint8 [%d, %d, %d] tensor" % (Cn, dh, dw))
    with torch.cuda.device(src.device):
        call("b2y_letterbox_u8", ptr(src), H, W, Cn, C.c_longlong(W * Cn), rh, rw, top, left, ptr(out), dh, dw,
             1 if (to_rgb and Cn == 3) else 0, int(c0), stream_ptr())
    return out, ratio, pad


def letterbox_batch(images, new_shape=640, scaleup=True, device=None):
    """Fixed-size batch (auto=False, what LoadImagesAndLabels / LoadStreams(rect=False) use): list of HWC uint8 BGR
    images -> (uint8 [B, 3, H, W] RGB batch on the device, [ratio], [(dw, dh)])."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    first = _device_image(images[0], device)
    batch = torch.empty((len(images), first.shape[2], new_shape[0], new_shape[1]), dtype=torch.uint8, device=first.device)
    ratios, pads = [], []
    for b, im in enumerate(images):
        _, r, p = letterbox(first if b == 0 else im, new_shape, auto=False, scaleup=scaleup, out=batch[b])
        ratios.append(r)
        pads.append(p)
    return batch, ratios, pads
