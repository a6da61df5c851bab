"""heatsim2_b200 - B200-native drop-in for heatsim2's ADI Crank-Nicolson path.

Same public surface as the reference package (``heatsim2/__init__.py``):
grid builders (:53, :102, :146), ``zero_elements`` (:191), the material /
source constants (:42-50), the boundary plug-in modules, and the two entry
points ``setup`` (:37) and ``run_adi_steps`` (:39).  Problem definition stays
on the host in numpy exactly as before; ``setup`` compiles it into a device
plan and ``run_adi_steps`` advances the field with hand-written sm_100a CUDA
kernels reached through the C ABI in ``include/hs2_b200.h``.  There is no CPU
implementation of the time step in this package.

    import heatsim2_b200 as heatsim2      # drop-in
"""
import numpy as np

from . import boundary_conducting
from . import boundary_conducting_anisotropic
from . import boundary_insulating
from . import boundary_thininsulatinglayer
from . import expression
from . import surface_temperature

__version__ = "0.1.0"

# material[0]
TEMPERATURE_COMPUTE = 0
TEMPERATURE_FIXED = 1

# volumetric[0]
NO_SOURCE = 0
IMPULSE_SOURCE = 1
STEPPED_SOURCE = 2
IMPULSE_POINT_SOURCE_JOULES = 3
SPATIALLY_Z_DECAYING_TEMPORAL_IMPULSE = 4


def _grid_tuple(z_bnd, y_bnd, x_bnd, dz, dy, dx):
    """Centres, centre meshgrids, face meshgrids and radii from face coords."""
    z = z_bnd[:-1] + dz / 2.0
    y = y_bnd[:-1] + dy / 2.0
    x = x_bnd[:-1] + dx / 2.0
    zfaces = np.meshgrid(z_bnd, y, x, indexing="ij")
    yfaces = np.meshgrid(z, y_bnd, x, indexing="ij")
    xfaces = np.meshgrid(z, y, x_bnd, indexing="ij")
    zgrid, ygrid, xgrid = np.meshgrid(z, y, x, indexing="ij")
    r2d = np.sqrt(xgrid ** 2 + ygrid ** 2)
    r3d = np.sqrt(xgrid ** 2 + ygrid ** 2 + zgrid ** 2)
    return (z, y, x, zgrid, ygrid, xgrid, z_bnd, y_bnd, x_bnd) + \
        tuple(zfaces) + tuple(yfaces) + tuple(xfaces) + (r3d, r2d)


def build_grid(minz, maxz, nz, miny, maxy, ny, minx, maxx, nx):
    """Grid whose outer FACES are at min/max; returns
    ``(dz,dy,dx, z,y,x, zgrid,ygrid,xgrid, z_bnd,y_bnd,x_bnd, <9 face
    meshgrids>, r3d, r2d)`` like the reference (``__init__.py:53-99``)."""
    z_bnd, dz = np.linspace(minz, maxz, num=nz + 1, retstep=True)
    y_bnd, dy = np.linspace(miny, maxy, num=ny + 1, retstep=True)
    x_bnd, dx = np.linspace(minx, maxx, num=nx + 1, retstep=True)
    return (dz, dy, dx) + _grid_tuple(z_bnd, y_bnd, x_bnd, dz, dy, dx)


def build_grid_min_step(minzcenter, dz, nz, minycenter, dy, ny, minxcenter, dx, nx):
    """Grid from the first cell CENTRE and the step (``__init__.py:102-143``)."""
    z_bnd = minzcenter - dz / 2.0 + np.arange(nz + 1, dtype="d") * dz
    y_bnd = minycenter - dy / 2.0 + np.arange(ny + 1, dtype="d") * dy
    x_bnd = minxcenter - dx / 2.0 + np.arange(nx + 1, dtype="d") * dx
    return _grid_tuple(z_bnd, y_bnd, x_bnd, dz, dy, dx)


def build_grid_min_step_edge(minzedge, dz, nz, minyedge, dy, ny, minxedge, dx, nx):
    """Grid from the first cell FACE and the step (``__init__.py:146-186``)."""
    z_bnd = minzedge + np.arange(nz + 1, dtype="d") * dz
    y_bnd = minyedge + np.arange(ny + 1, dtype="d") * dy
    x_bnd = minxedge + np.arange(nx + 1, dtype="d") * dx
    return _grid_tuple(z_bnd, y_bnd, x_bnd, dz, dy, dx)


def zero_elements(nz, ny, nx):
    """All-zero u8 class arrays: materials, z/y/x faces, volumetric sources
    (``__init__.py:191-197``)."""
    return (np.zeros((nz, ny, nx), dtype="u1"),
            np.zeros((nz + 1, ny, nx), dtype="u1"),
            np.zeros((nz, ny + 1, nx), dtype="u1"),
            np.zeros((nz, ny, nx + 1), dtype="u1"),
            np.zeros((nz, ny, nx), dtype="u1"))


from . import tridiag                                   # noqa: E402
from . import alternatingdirection_c_pyx                # noqa: E402
from . import alternatingdirection_c_pyx as alternatingdirection  # noqa: E402
from . import crank_nicolson                            # noqa: E402

setup = crank_nicolson.setup
run_adi_steps = alternatingdirection_c_pyx.run_adi_steps
run_adi_steps_n = alternatingdirection_c_pyx.run_adi_steps_n      # extension: device-resident loop + on-device observation
