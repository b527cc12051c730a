"""B200-native tiled U-Net segmentation inference (drop-in for the predict path of
mjevans26/Satellite_ComputerVision: ``utils/model_tools.py`` U-Net builders +
``utils/prediction_tools.py`` tiled prediction + ``utils/processing.py`` normalisers).

All arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI of
``include/scv.h`` (``libscv.so``, bound with ctypes in ``_lib.py``).  There is no
CPU fallback: without the built library or without a CUDA device the compute
entry points raise.
"""
from . import model_tools, prediction_tools, processing  # noqa: F401
from ._lib import ScvError, lib_path, load_library  # noqa: F401

__version__ = '0.1.0'
