"""Conducting face with a 3x3 conductivity tensor K (index order z,y,x):
q_n = -sum_a K_face[n,a] * dT/da, K_face the mean of the two cells' tensors.
Tangential derivatives are centred differences averaged over the two sides of
the face (reference ``heatsim2/boundary_conducting_anisotropic.py:18-24``).

With K diagonal in the grid axes the tangential groups are multiplied by zero
and vanish in ``fullreduce``; the three-stage ADI handles exactly that case.

Operand order after (k_minus, k_plus, dz, dy, dx) follows the reference:
(minus, plus) pairs at the face centre, then shifted by -1/+1 along the first
and then the second tangential axis; remaining corner operands are unused."""
from .expression import group


def _tangential(plus_hi, plus_lo, minus_hi, minus_lo, d):
    return group((plus_hi - plus_lo) * (0.25 / d)) + group((minus_hi - minus_lo) * (0.25 / d))


def qz(kmatm55, kmatp55, dz, dy, dx, Tm55, Tp55,
       Tm45, Tp45, Tm65, Tp65, Tm54, Tp54, Tm56, Tp56, *corners):
    K = (kmatm55 + kmatp55) * 0.5
    dTdz = group((Tp55 - Tm55) * (1.0 / dz))
    dTdy = _tangential(Tp65, Tp45, Tm65, Tm45, dy)
    dTdx = _tangential(Tp56, Tp54, Tm56, Tm54, dx)
    return -K[0, 0] * dTdz - K[0, 1] * dTdy - K[0, 2] * dTdx


def qy(kmat5m5, kmat5p5, dz, dy, dx, T5m5, T5p5,
       T4m5, T4p5, T6m5, T6p5, T5m4, T5p4, T5m6, T5p6, *corners):
    K = (kmat5m5 + kmat5p5) * 0.5
    dTdz = _tangential(T6p5, T4p5, T6m5, T4m5, dz)
    dTdy = group((T5p5 - T5m5) * (1.0 / dy))
    dTdx = _tangential(T5p6, T5p4, T5m6, T5m4, dx)
    return -K[1, 0] * dTdz - K[1, 1] * dTdy - K[1, 2] * dTdx


def qx(kmat55m, kmat55p, dz, dy, dx, T55m, T55p,
       T45m, T45p, T65m, T65p, T54m, T54p, T56m, T56p, *corners):
    K = (kmat55m + kmat55p) * 0.5
    dTdz = _tangential(T65p, T45p, T65m, T45m, dz)
    # the reference's x-face dT/dy mixes T55m with T54m (boundary_conducting_anisotropic.py:63);
    # the symmetric stencil is used here.  Identical whenever K is axis-aligned
    # (the only case the three-stage ADI accepts), because the term is then * 0.
    dTdy = _tangential(T56p, T54p, T56m, T54m, dy)
    dTdx = group((T55p - T55m) * (1.0 / dx))
    return -K[2, 0] * dTdz - K[2, 1] * dTdy - K[2, 2] * dTdx
