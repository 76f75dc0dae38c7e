"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference ``/root/reference/multigrid.py`` so that its own functions can be
executed on synthetic hierarchies.  The reference imports ``dolfinx`` and ``ufl`` at module level
(multigrid.py:2-3) but only uses them inside ``res_calculator`` / ``err_calculator``
(multigrid.py:203-218), which the V-cycle never calls, so two empty stub modules are enough.

``/root/reference`` exists only in the development container.  On the GPU box this module raises
``ReferenceUnavailable``; tests that need it are skipped there and rely on the golden fixtures in
``tests/golden/`` that ``tests/golden/gen_golden.py`` produced with this very module.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_DIR = os.environ.get("MGB_REFERENCE_DIR", "/root/reference")


class ReferenceUnavailable(RuntimeError):
    pass


_cached = None


def load_reference():
    """Return the reference ``multigrid`` module (fresh module object, cached)."""
    global _cached
    if _cached is not None:
        return _cached
    path = os.path.join(REFERENCE_DIR, "multigrid.py")
    if not os.path.exists(path):
        raise ReferenceUnavailable(f"{path} not present (GPU box?)")
    for name in ("dolfinx", "ufl"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("_reference_multigrid", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod


def available():
    return os.path.exists(os.path.join(REFERENCE_DIR, "multigrid.py"))


def run_reference_vcycles(H, ncycles, v0=None, f=None, test_tuple=False):
    """Drive the reference exactly as Multigrid_prototype.py:135-143 does, but calling
    ``V_cycle_scheme`` directly at the finest level.

    Returns (list of v after each cycle, debug tuple of the last cycle or None)."""
    import numpy as np
    ref = load_reference()
    H.A_jacobi_sp_dict = {l: ref.getJacobiMatrices(H.A_sp_dict[l]) for l in H.levels()}   # proto:135-136
    ref.initialize_problem(H)                                                                # proto:138-140
    lf = H.finest_level
    n = H.A_sp_dict[lf][0].shape[0]
    v = np.zeros((n, 1)) if v0 is None else np.array(v0, dtype=np.float64).reshape(n, 1)
    f = H.b_dict[lf] if f is None else np.array(f, dtype=np.float64).reshape(n, 1)
    out = []
    dbg = None
    for _ in range(ncycles):
        if test_tuple:
            v, f2h, v2h, errh = ref.V_cycle_scheme(H.A_jacobi_sp_dict[lf], v, f, True)
            dbg = (f2h, v2h, errh)
        else:
            v = ref.V_cycle_scheme(H.A_jacobi_sp_dict[lf], v, f)
        out.append(v.copy())
    return out, dbg
