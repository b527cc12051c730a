"""The predict-path piece of ``utils/raster_tools.py``: the per-side-buffer chip grid.

``raster_tools.generate_chip_indices(H, W, buff, kernel)`` (``utils/raster_tools.py:23-46``) differs from the
``prediction_tools`` generator (``utils/prediction_tools.py:87-109``): ``buff`` is the margin trimmed from EACH
side (chip side = kernel + 2*buff) and the range end is inclusive (``H - (kernel + buff) + 1``), so the grid reaches
the last position whose chip still fits.  Everything else of that module (rasterio / GDAL windows, COG writing)
is out of scope.
"""
from __future__ import annotations

import numpy as np

from . import prediction_tools as _pt


def generate_chip_indices(H, W, buff=128, kernel=256):
    """``utils/raster_tools.py:23-46`` -- identical: row-major (y, x) upper-left corners of the kept cores."""
    x_buff = y_buff = buff
    y_indices = list(range(y_buff, H - (kernel + buff) + 1, kernel))
    x_indices = list(range(x_buff, W - (kernel + buff) + 1, kernel))
    return [(y_index, x_index) for y_index in y_indices for x_index in x_indices]


def predict_chips(arr, m, buff=128, kernel=256, norm=None, channel=0, template=None):
    """Tiled prediction over the per-side-buffer grid: every (kernel + 2*buff)^2 chip goes through the engine and
    its kernel^2 core is accumulated into ``template`` (float64 zeros by default), like
    ``prediction_tools.predict_chips`` does for its own grid."""
    raw = arr.raw if hasattr(arr, 'raw') else np.asarray(arr)
    H, W = raw.shape[:2]
    if template is None:
        template = np.zeros((H, W))
    idx = generate_chip_indices(H, W, buff, kernel)
    return _pt.predict_chips(arr, idx, template, m, kernel=kernel, buff=2 * buff, norm=norm, channel=channel)
