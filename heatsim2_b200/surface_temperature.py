"""Surface-temperature estimate for an insulated z-min face
(reference ``heatsim2/surface_temperature.py:4-36``).  Works on numpy arrays
and on torch tensors alike (slicing + arithmetic only), so it can run on the
device-resident field without a copy."""


def insulating_z_min_surface_temperature(T, dz):
    """T is indexed [z,y,x]; layers 0 and 1 sit at z=dz/2 and 3dz/2.

    Returns the mean of two extrapolations to z=0: an even parabola
    T = a z^2 + c (zero slope at the insulated wall) and a straight line."""
    step = T[1] - T[0]
    curvature = step / (2.0 * dz ** 2.0)
    parabola_at_wall = T[0] - 0.25 * curvature * dz ** 2.0
    line_at_wall = T[0] - step / 2.0
    return (line_at_wall + parabola_at_wall) / 2.0
