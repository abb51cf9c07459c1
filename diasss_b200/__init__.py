"""diasss_b200 -- B200-native (sm_100a CUDA) front end of halajun/diasss: ORB extraction + pairwise matching.

Layout: csrc/ (CUDA kernels + C ABI -> libdiasss_b200.so), binding.py (ctypes), frontend.py (host-side mirror of
the reference's ORBextractor / Frame::DetectFeature / FEAmatcher interface), synth.py (seeded synthetic surveys).
"""
from . import binding  # noqa: F401
from .binding import KP_DTYPE, DsxError  # noqa: F401

__all__ = ["binding", "KP_DTYPE", "DsxError"]
