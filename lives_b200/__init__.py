"""lives_b200 -- B200-native per-frame pixel engine for the LiVES hot path
(palette conversion -> resize / letterbox -> effect blend / composite -> gamma).

The product is lives_b200/libpe_b200.so (hand-written sm_100a CUDA behind the C ABI of include/pixel_engine.h);
this package is the thin host-side mirror of the reference's interface.  No CPU fallback exists.
"""
from . import _capi
from ._capi import PixelEngineError, PixelEngineUnavailable  # noqa: F401
from .engine import *  # noqa: F401,F403

__version__ = "0.1.0"
