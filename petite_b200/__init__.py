"""petite_b200: B200-native engine for PETITE's shower-stepping hot path.

``Shower`` / ``DarkShower`` / ``Particle`` keep the reference's Python API; the stepping runs as hand-written
sm_100a CUDA behind the C ABI in ``include/petite_b200.h``.  Importing the package does not load the CUDA
library; constructing a ``Shower`` does, and fails loudly without it (no CPU fallback).
"""
__version__ = "0.1.0"
from .particle import Particle, mass_dict, meson_decay_dict  # noqa: F401


def __getattr__(name):
    if name == "Shower":
        from .shower import Shower
        return Shower
    if name == "DarkShower":
        from .dark_shower import DarkShower
        return DarkShower
    raise AttributeError(name)
