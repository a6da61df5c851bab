"""heatsim2 - drop-in package: the reference's module names on the B200 backend.

Put ``<repo>/dropin`` in front of ``sys.path`` (``PYTHONPATH=<repo>/dropin:<repo>``) and an unmodified heatsim2
script - ``import heatsim2``; ``heatsim2.setup(...)``; ``heatsim2.run_adi_steps(...)`` - runs on the GPU.

What is what (north_star: "the Python API that defines grids, materials and boundaries
stays unchanged"):

* REPLACED, the hot path (heatsim2/__init__.py:15-17,37,39 import these names):
  ``heatsim2.alternatingdirection_c_pyx`` (+ alias ``heatsim2.alternatingdirection``), ``heatsim2.crank_nicolson``,
  ``heatsim2.tridiag``, ``heatsim2.setup``, ``heatsim2.run_adi_steps`` -> heatsim2_b200 (CUDA through the C ABI).
* UNCHANGED, the definition API: ``expression``, ``boundary_conducting``, ``boundary_conducting_anisotropic``,
  ``boundary_insulating``, ``boundary_thininsulatinglayer``, ``surface_temperature``, ``hs2_indexing``.  When a
  reference checkout is available (environment variable ``HEATSIM2_REFERENCE`` = the directory that holds
  ``expression.py``, or an installed ``heatsim2`` further down ``sys.path``) those files are loaded FROM THERE,
  unmodified, under their own names, so ``from heatsim2.expression import group`` inside a plug-in keeps meaning the
  reference's engine; ``setup`` converts what the plug-ins return (crank_nicolson.from_foreign_expression).
  Without a reference tree the same names resolve to heatsim2_b200's own implementations of the same interfaces.
* grid builders / constants of heatsim2/__init__.py (:42-197): same names and return tuples, from heatsim2_b200.
"""
import importlib.util
import os
import sys

import heatsim2_b200 as _b200
from heatsim2_b200 import (TEMPERATURE_COMPUTE, TEMPERATURE_FIXED, NO_SOURCE, IMPULSE_SOURCE, STEPPED_SOURCE,  # noqa: F401
                           IMPULSE_POINT_SOURCE_JOULES, SPATIALLY_Z_DECAYING_TEMPORAL_IMPULSE,
                           build_grid, build_grid_min_step, build_grid_min_step_edge, zero_elements)

__version__ = "b200-" + _b200.__version__

_UNCHANGED = ("expression", "boundary_conducting", "boundary_conducting_anisotropic", "boundary_insulating",
              "boundary_thininsulatinglayer", "surface_temperature", "hs2_indexing")


def _find_reference():
    """directory holding the reference's pure-python modules, or None"""
    here = os.path.dirname(os.path.abspath(__file__))
    cands = []
    env = os.environ.get("HEATSIM2_REFERENCE")
    if env:
        cands += [env, os.path.join(env, "heatsim2")]
    for p in sys.path:
        d = os.path.join(p or ".", "heatsim2")
        if os.path.abspath(d) != here:
            cands.append(d)
    for d in cands:
        if os.path.isfile(os.path.join(d, "expression.py")) and os.path.isfile(os.path.join(d, "boundary_conducting.py")):
            return os.path.abspath(d)
    return None


reference_dir = _find_reference()


def _load(name):
    full = __name__ + "." + name
    path = os.path.join(reference_dir, name + ".py") if reference_dir else None
    if path and os.path.isfile(path):
        spec = importlib.util.spec_from_file_location(full, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod                  # before exec: the plug-ins import heatsim2.expression themselves
        spec.loader.exec_module(mod)
    else:
        mod = importlib.import_module("heatsim2_b200." + name)
        sys.modules[full] = mod
    globals()[name] = mod
    return mod


for _name in _UNCHANGED:
    try:
        _load(_name)
    except Exception:                             # e.g. the reference's hs2_indexing is unused there; never fatal
        if _name in ("hs2_indexing",):
            sys.modules[__name__ + "." + _name] = importlib.import_module("heatsim2_b200." + _name)
            globals()[_name] = sys.modules[__name__ + "." + _name]
        else:
            raise

from . import tridiag                                    # noqa: E402,F401
from . import alternatingdirection_c_pyx                 # noqa: E402
from . import alternatingdirection_c_pyx as alternatingdirection   # noqa: E402,F401
from . import crank_nicolson                             # noqa: E402

setup = crank_nicolson.setup
run_adi_steps = alternatingdirection_c_pyx.run_adi_steps
run_adi_steps_n = alternatingdirection_c_pyx.run_adi_steps_n      # extension: device-resident loop
