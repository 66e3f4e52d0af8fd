"""thirring2d_b200 — B200-native (sm_100a) Dirac-apply + CG hot path of rantahar/Thirring2D.

The product is the C-ABI shared library ``libthirring_b200.so`` (CUDA, include/thirring_b200.h) and the
reference-signature shim ``libthirring_hmc.so`` (include/thirring_hmc_abi.h).  This package is the thin
host-side mirror used by tests and bench.py: it loads the library with ctypes and exposes the reference's
function names (fm_mul, fm_conjugate_mul, fmdm_invert_cg, fm_invert_cg) over batched numpy / torch buffers.
There is no CPU fallback: importing works anywhere, creating a ``Context`` needs a CUDA device.
"""
from .lib import load_library, library_path, TBError  # noqa: F401
from .api import (  # noqa: F401
    Context,
    MODE_REF_COMPAT,
    MODE_ADJOINT,
    OP_M,
    OP_MDAG,
    OP_MCONJ,
    OP_MDM,
    CG_CONVERGED,
    CG_MAXITER,
    CG_DIVERGED,
    CG_ZERO_SOURCE,
    BC_ANTISYMMETRIC,
    BC_SYMMETRIC,
    BC_OPENX,
)

__all__ = ["Context", "load_library", "library_path", "TBError"]
