"""heatsim2.alternatingdirection_c_pyx on the B200 backend (reference: heatsim2/alternatingdirection_c_pyx.pyx)."""
from heatsim2_b200.alternatingdirection_c_pyx import (adi_params, pyadi_step, adi_setup, adi_expressions,  # noqa: F401
                                                       add_equation_to_adi_matrices, run_adi_steps, run_adi_steps_n,
                                                       stage_dicts, parse_stage_dict, class_coefficients)
