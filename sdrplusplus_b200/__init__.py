"""sdrplusplus_b200 -- B200-native (sm_100a CUDA) implementation of SDR++'s streaming DSP hot path.

The product is the C-ABI shared library ``libb200dsp.so`` (include/b200dsp.h); this package is the thin
ctypes binding the tests and bench.py use, plus the host-side mirrors of the reference's front-end /
block interfaces.  There is no CPU fallback: every compute call fails loudly without a CUDA device.
"""
from .lib import load, B200Error  # noqa: F401
from .frontend import FrontEnd, VfoConfig, Block, SpectrumHandler, RdsDemod  # noqa: F401
from . import lib as _lib  # noqa: F401

__all__ = ["load", "B200Error", "FrontEnd", "VfoConfig", "Block", "SpectrumHandler", "RdsDemod"]
