"""heatsim2.crank_nicolson on the B200 backend (reference: heatsim2/crank_nicolson.pyx)."""
from heatsim2_b200.crank_nicolson import (setup, shift_expression, subst_thermal_conductivity, compile_problem,  # noqa: F401
                                          evaluate_boundaries, from_foreign_expression)
