"""koopfit — B200-native Ksysid EDMD fit (lift -> Gram/cross-covariance -> solve).

Host-side mirror of the reference's `Ksysid` class over the C ABI in
include/koopfit.h (libkoopfit.so: hand-written sm_100a CUDA).  Import as
`import koopfit` through the shim at the repository root.
"""
from . import _abi  # noqa: F401
from ._abi import Basis, KoopfitError  # noqa: F401
from .fitter import Fitter, MultiFitter  # noqa: F401
