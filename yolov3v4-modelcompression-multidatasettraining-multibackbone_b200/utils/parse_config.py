"""Darknet .cfg / .data parsers -- same observable behaviour as the reference's utils/parse_config.py:6-70.

parse_model_cfg(path) -> list of block dicts (element 0 is the [net] block)
  * path may omit the '.cfg' suffix and/or the 'cfg/' directory prefix            (reference :8-11)
  * blank lines and lines starting with '#' are dropped, fringe whitespace stripped (:15-16)
  * '[convolutional]' blocks are pre-populated with batch_normalize=0 (NOT depthwise) (:22-23)
  * anchors -> float64 ndarray (n, 2); from/layers/mask (and comma-valued size) -> list[int]  (:28-31)
  * other values: str.isnumeric() -> int, anything else stays a *string* ('0.1', '-1', '.7')  (:33-37)
  * every key of every non-[net] block must be in SUPPORTED_KEYS (assert)                  (:40-49)
"""
import os

import numpy as np

SUPPORTED_KEYS = (
    'type', 'batch_normalize', 'filters', 'size', 'stride', 'pad', 'activation', 'layers', 'groups', 'reduction',
    'from', 'mask', 'anchors', 'classes', 'num', 'jitter', 'ignore_thresh', 'truth_thresh', 'random', 'stride_x',
    'stride_y', 'weights_type', 'weights_normalization', 'scale_x_y', 'beta_nms', 'nms_kind', 'iou_loss',
    'iou_normalizer', 'cls_normalizer', 'iou_thresh', 'group_id', 'resize')

_INT_LIST_KEYS = ('from', 'layers', 'mask')


def _resolve(path, suffix, folder):
    if suffix and not path.endswith(suffix):
        path = path + suffix
    if not os.path.exists(path):
        alt = folder + os.sep + path
        if os.path.exists(alt):
            path = alt
    return path


def _convert_value(key, raw):
    if key == 'anchors':
        return np.array([float(v) for v in raw.split(',')]).reshape((-1, 2))
    if key in _INT_LIST_KEYS or (key == 'size' and ',' in raw):
        return [int(v) for v in raw.split(',')]
    raw = raw.strip()
    if raw.isnumeric():
        as_int = int(raw)
        return as_int if (as_int - float(raw)) == 0 else float(raw)
    return raw


def parse_model_cfg_text(text, origin='<string>'):
    """Parse cfg *text* (used by the in-repo cfg generator); same rules as parse_model_cfg."""
    blocks = []
    for line in text.split('\n'):
        if not line or line.startswith('#'):
            continue
        line = line.strip()
        if not line:
            continue
        if line.startswith('['):
            block = {'type': line[1:-1].rstrip()}
            if block['type'] == 'convolutional':
                block['batch_normalize'] = 0
            blocks.append(block)
            continue
        key, raw = line.split('=')
        key = key.rstrip()
        blocks[-1][key] = _convert_value(key, raw)

    seen = []
    for block in blocks[1:]:
        for key in block:
            if key not in seen:
                seen.append(key)
    unsupported = [k for k in seen if k not in SUPPORTED_KEYS]
    assert not any(unsupported), "Unsupported fields %s in %s. See https://github.com/ultralytics/yolov3/issues/631" % (
        unsupported, origin)
    return blocks


def parse_model_cfg(path):
    path = _resolve(path, '.cfg', 'cfg')
    with open(path, 'r') as f:
        text = f.read()
    return parse_model_cfg_text(text, path)


def parse_data_cfg(path):
    path = _resolve(path, '', 'data')
    options = dict()
    with open(path, 'r') as f:
        for line in f.readlines():
            line = line.strip()
            if line == '' or line.startswith('#'):
                continue
            key, val = line.split('=')
            options[key.strip()] = val.strip()
    return options
