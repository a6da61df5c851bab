"""Insulating (zero-flux) face.  Plug-in interface of the reference's
``heatsim2/boundary_insulating.py``: module-level ``qz``, ``qy``, ``qx`` taking
the two face conductivities, the three cell sizes and the 14 neighbouring
temperature operands, returning the flux through the face (W/m^2)."""


def _no_flux(*operands):
    return 0.0


qz = qy = qx = _no_flux
