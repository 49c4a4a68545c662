#!/usr/bin/env python
"""The reference's detect.py flow (detect.py:60-140) with every step after the image decode on the device:

    cv2.imread -> raw BGR bytes to HBM -> letterbox kernel (resize + border + BGR->RGB, CHW)   datasets.py:108-118, 611-646
               -> model(uint8 batch)  ("/ 256" inside the first conv)                           detect.py:99-112
               -> non_max_suppression on the device (multi_label=False like detect.py:115)      utils.py:782-860
               -> scale_coords back to the source image, rounded                                detect.py:130

    python tools/detect_device.py --cfg yolov3 [--weights file.weights|file.pt] [--img-size 640] image1.jpg image2.jpg ...

Without --weights the model keeps its random initialisation (useful only to exercise the pipeline); without image
arguments a synthetic 480 x 640 image is used.  Prints one line per image: the detections as (x1, y1, x2, y2, conf, cls)
in source-image pixels.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "yolov3v4-modelcompression-multidatasettraining-multibackbone_b200")
sys.path.insert(0, PKG)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def scale_coords(img1_shape, coords, img0_shape):
    """utils.py:137-151: boxes from the letterboxed shape back to the source image (gain / pad from the two shapes)."""
    gain = max(img1_shape) / max(img0_shape)
    pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    coords[:, 0].clamp_(0, img0_shape[1])
    coords[:, 1].clamp_(0, img0_shape[0])
    coords[:, 2].clamp_(0, img0_shape[1])
    coords[:, 3].clamp_(0, img0_shape[0])
    return coords


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="yolov3", help="a generated cfg name (b200yolo.cfggen) or a path to a .cfg file")
    ap.add_argument("--weights", default="")
    ap.add_argument("--img-size", type=int, default=640)
    ap.add_argument("--conf-thres", type=float, default=0.3)
    ap.add_argument("--iou-thres", type=float, default=0.6)
    ap.add_argument("images", nargs="*")
    args = ap.parse_args()
    import models
    from b200yolo import cfggen, preprocess
    from utils import utils as U
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = args.cfg if os.path.isfile(args.cfg) else cfggen.write_cfg(args.cfg, "/tmp/b2y_cfg_detect_%d" % os.getpid())
    model = models.Darknet(cfg, img_size=(args.img_size, args.img_size))
    if args.weights.endswith(".pt"):
        model.load_state_dict(torch.load(args.weights, map_location="cpu")["model"])
    elif args.weights:
        models.load_darknet_weights(model, args.weights)
    model = model.to(dev).eval()
    sources = []
    if args.images:
        import cv2
        for path in args.images:
            im0 = cv2.imread(path)                      # HWC BGR uint8, like LoadImages (datasets.py:104)
            assert im0 is not None, "Image Not Found " + path
            sources.append((path, im0))
    else:
        rng = np.random.default_rng(0)
        sources.append(("<synthetic 480x640>", rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)))
    with torch.no_grad():
        for path, im0 in sources:
            img, ratio, pad = preprocess.letterbox(im0, new_shape=args.img_size)       # uint8 [3, H, W] RGB on the device
            pred = model(img[None])[0]                                               # uint8 in: "/ 256" in the stem kernel
            det = U.non_max_suppression(pred, args.conf_thres, args.iou_thres, multi_label=False)[0]
            if det is None:
                print("%s: %dx%d no detections" % (path, img.shape[1], img.shape[2]))
                continue
            det = det.clone()
            det[:, :4] = scale_coords(img.shape[1:], det[:, :4], im0.shape).round()
            rows = ["(%d, %d, %d, %d, %.3f, %d)" % tuple(r) for r in det[:20].cpu().tolist()]
            print("%s: %dx%d %d detections%s: %s" % (path, img.shape[1], img.shape[2], det.shape[0],
                                                     " (first 20)" if det.shape[0] > 20 else "", ", ".join(rows)))


if __name__ == "__main__":
    main()
