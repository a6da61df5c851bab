"""Thin resistive layer on a face: q = -h * (T_plus - T_minus), with ``h`` the
conduction coefficient normal to the face in W/(m^2 K), given as the extra
element of the boundary tuple ``(boundary_thininsulatinglayer, h)``
(reference ``heatsim2/boundary_thininsulatinglayer.py:9-23``)."""
from .expression import group


def _layer_flux(T_minus, T_plus, h):
    return -h * group(T_plus - T_minus)


def qz(kmatm55, kmatp55, dz, dy, dx, Tm55, Tp55,
       Tm45, Tp45, Tm65, Tp65, Tm54, Tp54, Tm56, Tp56, Tm46, Tp46, Tm64, Tp64,
       conductioncoefficient):
    return _layer_flux(Tm55, Tp55, conductioncoefficient)


def qy(kmat5m5, kmat5p5, dz, dy, dx, T5m5, T5p5,
       T4m5, T4p5, T6m5, T6p5, T5m4, T5p4, T5m6, T5p6, T4m6, T4p6, T6m4, T6p4,
       conductioncoefficient):
    return _layer_flux(T5m5, T5p5, conductioncoefficient)


def qx(kmat55m, kmat55p, dz, dy, dx, T55m, T55p,
       T45m, T45p, T65m, T65p, T54m, T54p, T56m, T56p, T46m, T46p, T64m, T64p,
       conductioncoefficient):
    return _layer_flux(T55m, T55p, conductioncoefficient)
