"""icp-4dradar_b200 — B200 (sm_100a) registration hot path of ICP-4DRadar behind a C ABI.

The product is ``libicp4r_cuda.so`` (csrc/, built by ``__graft_entry__.build()``); this package is the thin
Python harness over it (ctypes) used by the tests and bench, plus the synthetic-scene generators.  The C++
adapters that mirror the reference's call shapes live in ``adapters/icp4r``.

The directory name carries a hyphen, so import it through the loader at the repo root::

    from icp4r_loader import pkg      # -> this package, registered as ``icp4dradar_b200``
"""
from .api import (  # noqa: F401
    DEVICE, HOST, P2LINE, P2PLANE_KNN, P2PLANE_3PT, P2P_GN, P2P_SVD, GICP, ACC_LEN, Icp4r, Icp4rError, Opts, Result,
    lib_path, load_library, default_opts,
)
from . import synth  # noqa: F401
from . import shard  # noqa: F401
from . import pipeline  # noqa: F401
from . import io  # noqa: F401
