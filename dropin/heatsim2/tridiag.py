"""heatsim2.tridiag on the B200 backend (reference: heatsim2/tridiag.pyx)."""
from heatsim2_b200.tridiag import *        # noqa: F401,F403
from heatsim2_b200.tridiag import tridiaglu, tridiagsolve      # noqa: F401
