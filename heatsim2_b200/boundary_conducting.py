"""Isotropic conducting face: q = -k_face * dT/dn with the arithmetic mean of
the two cell conductivities (reference ``heatsim2/boundary_conducting.py:18-25``).

The normal difference is wrapped in ``group`` so that the ADI stage builder can
find it and apply the Crank-Nicolson average to that direction only."""
from .expression import group


def _normal_flux(k_minus, k_plus, T_minus, T_plus, d):
    k_face = (k_minus + k_plus) * 0.5
    return -k_face * group((T_plus - T_minus) * (1.0 / d))


def qz(kmatm55, kmatp55, dz, dy, dx, Tm55, Tp55, *others):
    return _normal_flux(kmatm55, kmatp55, Tm55, Tp55, dz)


def qy(kmat5m5, kmat5p5, dz, dy, dx, T5m5, T5p5, *others):
    return _normal_flux(kmat5m5, kmat5p5, T5m5, T5p5, dy)


def qx(kmat55m, kmat55p, dz, dy, dx, T55m, T55p, *others):
    return _normal_flux(kmat55m, kmat55p, T55m, T55p, dx)
