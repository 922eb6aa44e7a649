"""nmfk_b200: host-side mirror of the NMFk.jl hot path over the B200 C ABI.

Same function names, argument meaning and error behaviour as the reference's Julia API for
this path (`execute`, `execute_run`, `execute_singlerun`, `NMFmultiplicative`,
`clustersolutions`-through-`robustness`, `getk`, `signalorder`); all numerics run in
libnmfk_b200.so (hand-written CUDA for sm_100a).  The Julia shim in ../julia/NMFkB200 binds
the same entry points with `ccall`; this package exists because no Julia runtime is available
in the build/test environment."""
from ._lib import F32, F64, LIB_PATH, NegativeEntriesError, NMFkError, Params, load
from .api import (Batch, Context, NMFmultiplicative, NMFmultiplicative_darray, NMFsparsity, default_params, execute, execute_k, execute_run,
                  execute_singlerun, getk, kmeanspp_seeds, robustkmeans, signalorder, trace)

__all__ = ["F32", "F64", "LIB_PATH", "NMFkError", "NegativeEntriesError", "Params", "load", "Batch", "Context",
           "NMFmultiplicative", "NMFmultiplicative_darray", "NMFsparsity", "default_params", "execute", "execute_k", "execute_run", "execute_singlerun", "getk",
           "signalorder", "trace", "robustkmeans", "kmeanspp_seeds"]
