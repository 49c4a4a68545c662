"""Detection post-processing on the device, csrc/nms.cu behind the reference's names.

    non_max_suppression(prediction, conf_thres, iou_thres, multi_label, classes, agnostic)   utils/utils.py:782-860
    match_batch(...)                                                                         test.py:123-175

The reference loops over the images in Python and, per image, issues ~25 ATen launches with several host
synchronisations (boolean-mask indexing, nonzero) plus torchvision's NMS with an n x n/64 bit matrix in HBM.  Here the
whole batch is processed by six launches and two small device-to-host reads (the candidate count, which sizes the
buffers, and the number of kept boxes per image).  Results: kept indices, scores and classes are bit-identical to the
reference (same candidate order, stable descending sort, fp32 IoU, threshold compared in double like torchvision);
merged boxes agree to fp32 summation order (the reference's torch.mm).  CUDA tensors only: there is no CPU fallback.
"""
import ctypes as C

import torch

from . import lib
from .lib import call, ptr, stream_ptr


class PackedDetections:
    """All detections of a batch in one buffer: image b owns rows offsets[b] .. offsets[b] + counts[b] of `det`
    ([x1, y1, x2, y2, conf, cls], descending score)."""

    def __init__(self, det, offsets, counts, offsets_dev, counts_dev):
        self.det, self.offsets, self.counts = det, offsets, counts
        self.offsets_dev, self.counts_dev = offsets_dev, counts_dev

    def __len__(self):
        return len(self.counts)

    def image(self, b):
        """[n, 6] view of image b or None (the reference leaves None for images without detections)."""
        n = self.counts[b]
        return self.det[self.offsets[b]:self.offsets[b] + n] if n else None

    def to_list(self):
        return [self.image(b) for b in range(len(self.counts))]


def _bytes(n, device):
    return torch.empty((max(int(n), 16),), dtype=torch.uint8, device=device)


def nms_packed(prediction, conf_thres=0.1, iou_thres=0.6, multi_label=True, classes=None, agnostic=False):
    """prediction [B, rows, 5 + nc] (xywh px, obj, class probabilities) on the device -> PackedDetections."""
    if not torch.is_tensor(prediction):
        prediction = torch.stack(list(prediction), 0)
    if not prediction.is_cuda:
        raise lib.B2YError("non_max_suppression runs on the CUDA device only (no CPU fallback)")
    assert prediction.dim() == 3 and prediction.shape[2] > 5, "expected [batch, rows, 5 + nc]"
    pred = prediction.detach().float().contiguous()
    B, R, no = pred.shape
    nc = no - 5
    dev = pred.device
    allow = None
    if classes:
        allow = torch.zeros(nc, dtype=torch.uint8)
        for c in classes:
            if 0 <= int(c) < nc:
                allow[int(c)] = 1
        allow = allow.to(dev)
    with torch.cuda.device(dev):
        row_off = torch.empty((B * R + 1,), dtype=torch.int32, device=dev)
        img_off = torch.empty((B + 1,), dtype=torch.int32, device=dev)
        ws_bytes = int(lib.raw().b2y_nms_count_workspace_bytes(B, R))
        if ws_bytes == 0:
            raise lib.B2YError("non_max_suppression: batch x rows out of range")
        ws = _bytes(ws_bytes, dev)
        call("b2y_nms_count", ptr(pred), B, R, nc, float(conf_thres), int(bool(multi_label)), ptr(allow), ptr(row_off),
             ptr(img_off), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
        offsets = img_off.cpu().tolist()                 # sync 1: the candidate count sizes everything below
        total = offsets[-1]
        det = torch.empty((max(total, 1), 6), dtype=torch.float32, device=dev)
        counts_dev = torch.empty((B,), dtype=torch.int32, device=dev)
        ws2_bytes = int(lib.raw().b2y_nms_run_workspace_bytes(total)) if total else 16
        if total and ws2_bytes == 0:
            raise lib.B2YError("non_max_suppression: %d candidates exceed the supported range" % total)
        ws2 = _bytes(ws2_bytes, dev)
        call("b2y_nms_run", ptr(pred), B, R, nc, float(conf_thres), C.c_double(float(iou_thres)),
             int(bool(multi_label)), int(bool(agnostic)), ptr(allow), ptr(row_off), ptr(img_off), C.c_longlong(total),
             ptr(ws2), C.c_size_t(ws2.numel()), ptr(det), ptr(counts_dev), stream_ptr())
        counts = counts_dev.cpu().tolist()               # sync 2
    return PackedDetections(det, offsets[:-1], counts, img_off, counts_dev)


def non_max_suppression(prediction, conf_thres=0.1, iou_thres=0.6, multi_label=True, classes=None, agnostic=False):
    """Drop-in for utils.utils.non_max_suppression: list (one entry per image) of [n, 6] tensors
    (x1, y1, x2, y2, conf, cls) or None."""
    return nms_packed(prediction, conf_thres, iou_thres, multi_label, classes, agnostic).to_list()


def match_labels(packed, tcls, tbox, lab_off, iouv, clip_w=0.0, clip_h=0.0):
    """Matching kernel on prepared labels: tcls fp32 [m], tbox fp32 [m, 4] (xyxy px), lab_off int32 [B + 1] (image b
    owns labels lab_off[b] .. lab_off[b + 1]).  clip_w > 0: boxes are first clamped in place to [0, clip_w] x [0, clip_h].
    Returns one bool [n, niou] matrix per image (None where the image has no detections)."""
    det = packed.det
    dev = det.device
    B = len(packed)
    iouv = iouv.to(device=dev, dtype=torch.float32).contiguous()
    niou = int(iouv.numel())
    tcls = tcls.to(device=dev, dtype=torch.float32).contiguous()
    tbox = tbox.to(device=dev, dtype=torch.float32).contiguous()
    lab_off = lab_off.to(device=dev, dtype=torch.int32).contiguous()
    assert lab_off.numel() == B + 1 and tbox.shape == (tcls.numel(), 4)
    with torch.cuda.device(dev):
        n_det, n_lab = int(det.shape[0]), int(tcls.numel())
        correct = torch.zeros((n_det, niou), dtype=torch.uint8, device=dev)
        ws = _bytes(lib.raw().b2y_tp_match_workspace_bytes(n_det, n_lab), dev)
        call("b2y_tp_match", ptr(det), ptr(packed.offsets_dev), ptr(packed.counts_dev), C.c_longlong(n_det), ptr(tcls),
             ptr(tbox), ptr(lab_off), C.c_longlong(n_lab), ptr(iouv), niou, B, float(clip_w), float(clip_h), ptr(ws),
             C.c_size_t(ws.numel()), ptr(correct), stream_ptr())
    out = []
    for b in range(B):
        n = packed.counts[b]
        out.append(correct[packed.offsets[b]:packed.offsets[b] + n].bool() if n else None)
    return out


def prepare_labels(targets, batch, width, height, device):
    """targets [nT, 6] (image, cls, xywh normalised) -> (tcls [nT], tbox [nT, 4] xyxy px, lab_off int32 [batch + 1]) with the
    labels grouped by image in their given order, exactly what test.py:124, 145-148 builds per image:
    labels = targets[targets[:, 0] == si, 1:];  tbox = xywh2xyxy(labels[:, 1:5]) * whwh.  Device-agnostic torch ops."""
    if targets is None or targets.numel() == 0:
        return (torch.zeros((0,), dtype=torch.float32, device=device),
                torch.zeros((0, 4), dtype=torch.float32, device=device),
                torch.zeros((batch + 1,), dtype=torch.int32, device=device))
    t = targets.to(device).float()
    img = t[:, 0].long()
    order = torch.sort(img, stable=True)[1]       # labels = targets[targets[:, 0] == si] keeps the given order
    t = t[order]
    whwh = torch.tensor([width, height, width, height], dtype=torch.float32, device=device)
    xywh = t[:, 2:6]
    tbox = torch.empty_like(xywh)                 # xywh2xyxy(labels[:, 1:5]) * whwh   (test.py:148)
    tbox[:, 0] = xywh[:, 0] - xywh[:, 2] / 2
    tbox[:, 1] = xywh[:, 1] - xywh[:, 3] / 2
    tbox[:, 2] = xywh[:, 0] + xywh[:, 2] / 2
    tbox[:, 3] = xywh[:, 1] + xywh[:, 3] / 2
    tbox = tbox * whwh
    cnt = torch.bincount(img.clamp(0, batch - 1), minlength=batch)[:batch]
    lab_off = torch.zeros((batch + 1,), dtype=torch.int32, device=device)
    lab_off[1:] = torch.cumsum(cnt, 0).to(torch.int32)
    return t[:, 1], tbox, lab_off


def match_batch(packed, targets, width, height, iouv, clip=True):
    """test.py:123-175 for the whole batch.  packed: PackedDetections; targets [nT, 6] (image, cls, xywh normalised) on
    the device; iouv: the IoU thresholds (torch.linspace(0.5, 0.95, 10)).  Boxes are clipped in place to the image like
    test.py:137 (clip=False skips it).  Returns a list with one bool [n, niou] `correct` matrix per image (None where
    the image has no detections): a prediction is a true positive at threshold q when it is the first prediction, in
    score order, whose best same-class target is that target and its IoU exceeds iouv[q]."""
    dev = packed.det.device
    with torch.cuda.device(dev):
        tcls, tbox, lab_off = prepare_labels(targets, len(packed), width, height, dev)
    cw, ch = (float(width), float(height)) if clip else (0.0, 0.0)
    return match_labels(packed, tcls, tbox, lab_off, iouv, cw, ch)
