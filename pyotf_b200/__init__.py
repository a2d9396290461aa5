"""pyotf_b200 -- B200-native drop-in for pyotf's PSF/OTF + phase-retrieval hot path.

Mirrors the reference's module layout (``otf``, ``phaseretrieval``, ``zernike``, ``utils``);
the numerical core is hand-written sm_100a CUDA behind the C ABI in include/pyotf_b200.h.
"""
from ._engine import get_precision, set_precision  # noqa: F401

__version__ = "0.1.0"
