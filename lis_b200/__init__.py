"""lis_b200 -- a B200-native (sm_100a) implementation of the Lis SpMV/Krylov hot path.

The product is the C library ``lis_b200/_lib/liblis_b200.so`` (host C behind the ``lis.h``
API + hand-written CUDA kernels, see ``include/``).  This Python package is only the
harness-side loader used by ``tests/``, ``bench.py`` and ``__graft_entry__.py``:

* :func:`load_library`  -- ``ctypes`` handle of the product library (fails loudly if it has
  not been built: there is no Python/CPU fallback for any compute entry point);
* :class:`Shim`         -- the shared test driver ``tests/shim/lis_shim.c`` (public Lis API
  only) compiled against this library; the same source compiled against the reference gives
  the oracle side of a parity test.
"""
from .capi import (LIB_DIR, REPO_ROOT, FMT, Shim, build, device_available, load_kernels,
                   load_library, load_shim)

__all__ = ["LIB_DIR", "REPO_ROOT", "FMT", "Shim", "build", "device_available", "load_kernels",
           "load_library", "load_shim"]
