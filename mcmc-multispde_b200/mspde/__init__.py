"""mspde -- host-side mirror of the reference's mcmc/ package surface for the pCN hot path, backed by the
sm_100a CUDA library libmspde.so through a ctypes C-ABI shim (see include/mspde.h)."""
from . import _ffi  # noqa: F401
