"""psitools_public_b200 -- B200-native (sm_100a) implementation of the psitools D(omega) hot path.

Importing the package does not touch the GPU; the first PSIDispersion / PSIMode construction loads
``libpsi_b200.so`` (built in-tree by ``__graft_entry__.build()``) and fails loudly when the library
or a CUDA device is missing -- there is no CPU fallback.
"""
from . import (tanhsinh, sizedensity, power_bump, psi_dispersion, complex_roots, psi_mode,  # noqa: F401
               psi_mode_mpi, complex_roots_mpi, psi_grid_refine, terminalvelocitysolver)
from .tanhsinh import TanhSinh, TanhSinhNoDeepCopy  # noqa: F401
from .sizedensity import SizeDensity  # noqa: F401
from .power_bump import PowerBump, PowerBumpTail  # noqa: F401
from .psi_dispersion import PSIDispersion  # noqa: F401
from .psi_mode import PSIMode  # noqa: F401
from .complex_roots import ClosedPath, Rectangle, RationalApproximation, RootFollower  # noqa: F401
from .terminalvelocitysolver import PSIModeTV, TerminalVelocitySolver  # noqa: F401
from .psi_grid_refine import PSIGridRefiner  # noqa: F401

__all__ = ["TanhSinh", "TanhSinhNoDeepCopy", "SizeDensity", "PowerBump", "PowerBumpTail",
           "PSIDispersion", "PSIMode", "ClosedPath", "Rectangle", "RationalApproximation",
           "PSIGridRefiner", "RootFollower", "PSIModeTV", "TerminalVelocitySolver"]
