"""Import the unmodified reference (built by oracle/build_ref.py into
oracle/_ref/) under its own name ``heatsim2``.  TEST INFRASTRUCTURE ONLY."""
import contextlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    sys.path.insert(0, HERE)
    try:
        import build_ref
        return build_ref.have_ref()
    finally:
        sys.path.remove(HERE)


def load():
    """Return the reference ``heatsim2`` module or None when it is not built."""
    if not available():
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import heatsim2
    return heatsim2


def quiet_setup(ref, *args):
    """reference setup() prints two lines per grid row (crank_nicolson.pyx:273,275)"""
    with contextlib.redirect_stdout(io.StringIO()):
        return ref.setup(*args)


def run(ref, problem_dict, nsteps=None, record=None):
    import numpy as np
    P, S = quiet_setup(ref, *problem_dict["setup_args"])
    T = np.array(problem_dict["T0"], dtype=np.float64)
    dt, t0 = problem_dict["dt"], problem_dict["t0"]
    n = problem_dict["nsteps"] if nsteps is None else nsteps
    rec = {}
    for it in range(n):
        T = ref.run_adi_steps(P, S, t0 + dt * it, dt, T, problem_dict["volumetric_elements"], problem_dict["volumetric"])
        if record is not None and (it + 1) in record:
            rec[it + 1] = T.copy()
    return (T, rec) if record is not None else T
