"""r2n2_b200: B200-native (sm_100a) implementation of 3D-R2N2's recurrent reconstruction hot
path behind the reference's lib/layers.py, models/*.py and lib/solver.py surface.

The package directory is named `3d-r2n2_b200` (not an importable identifier); import it as
`r2n2_b200` through the shim `r2n2_b200.py` at the repository root.
"""
from . import _lib  # noqa: F401

__all__ = ['_lib']
