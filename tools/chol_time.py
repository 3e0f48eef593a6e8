"""Calls the dense FP64 Cholesky solve for a few sizes (run under `ncu --metrics gpu__time_duration.sum` for a launch list)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sat_bundleadjust_b200 import _lib  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(0)
for n in [int(a) for a in sys.argv[1:]] or [300, 1800]:
    A = rng.standard_normal((n, n + 5))
    S = np.asfortranarray(A @ A.T + 0.1 * np.eye(n))
    b = rng.standard_normal(n)
    x = np.linalg.solve(S, b)
    info = ctypes.c_int32(-1)
    _lib.check(lib.sba_cholesky_solve(_lib.dptr(S), _lib.dptr(b), n, ctypes.byref(info)))
    print(n, info.value, np.abs(b - x).max() / np.abs(x).max())
    if os.environ.get("CHOL_TIMED", "1") == "1":
        S2 = np.asfortranarray(A @ A.T + 0.1 * np.eye(n))
        b2 = rng.standard_normal(n)
        xo, ms = np.empty(n), ctypes.c_double(0.0)
        lib.sba_cholesky_solve_timed.restype = ctypes.c_int
        _lib.check(lib.sba_cholesky_solve_timed(_lib.dptr(S2), _lib.dptr(b2), ctypes.c_int32(n), ctypes.c_int32(20),
                                                _lib.dptr(xo), ctypes.byref(ms)))
        print("   timed: n = %d  %.1f us per factor+solve (L2-warm, CUDA events), err %.2e"
              % (n, ms.value * 1e3, np.abs(xo - np.linalg.solve(S2, b2)).max()))
