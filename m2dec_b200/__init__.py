"""m2dec_b200 — B200-native reconstruction back end for the m2dec video decoder.

The product is the C-ABI shared library ``libm2dec_b200.so`` (``include/m2d_plugin.h``,
``include/b2_recon.h``).  This package is the thin Python host side: it loads the library
with ctypes and mirrors the reference application's driver class (``M2Decoder``,
reference src/app/m2decoder.h) on top of the plugin table.

There is no CPU reconstruction in here: if the library is missing, importing fails.
"""
from ._native import lib, LIB_PATH, FrameT, InfoT, FuncTable, M2Picture, M2MB  # noqa: F401
from .decoder import M2Decoder, decode_file, frame_md5, crop_frame  # noqa: F401
from .replay import Replay  # noqa: F401

__all__ = ["lib", "LIB_PATH", "M2Decoder", "decode_file", "frame_md5", "crop_frame"]
