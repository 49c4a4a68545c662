"""`utils` package of the drop-in (same import paths as the reference: utils.parse_config, utils.layers,
utils.utils, utils.torch_utils, utils.quantized.*).

Overlay: when B2Y_REFERENCE_ROOT points at a checkout of the reference, its own utils/ directory is appended to
this package's search path, so modules that are *outside* the accelerated hot path (utils.datasets,
utils.prune_utils, utils.google_utils, ...) keep resolving to the reference's files while every hot-path module
resolves to this package first.  See INTEGRATION.md.
"""
import os as _os

_ref = _os.environ.get("B2Y_REFERENCE_ROOT", "")
if _ref:
    _ref_utils = _os.path.join(_ref, "utils")
    if _os.path.isdir(_ref_utils) and _ref_utils not in __path__:
        __path__.append(_ref_utils)
