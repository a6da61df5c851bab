"""Index arithmetic for z-dependent lateral coarsening ("compressed geometry").

Same four function names as the reference's ``heatsim2/hs2_indexing.py`` (marked
"NOT CURRENTLY USED" there: nothing in heatsim2 imports it, and as shipped its
functions refer to undefined names - ``compressed_index`` asserts on itself
(:11), the others use ``uncompressed_index`` without defining it (:31,:49,:58)).
These are working versions of what that file describes: layer ``k`` of a grid
of shape ``(nz, ny, nx)`` is stored with its y and x axes coarsened by
``2**xy_scaling_log2[k]``, layers laid end to end in one flat array.  Not on the
ADI path; kept so that ``heatsim2.hs2_indexing`` keeps importing when the
reference tree is absent."""

# example table of the reference (indexed by z position)
xy_scaling_log2 = [0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2]


def compressed_index(uncompressed_index, xy_scaling_log2):
    """(k, j, i) of the fine grid -> (k, j', i') inside its coarsened layer."""
    assert len(uncompressed_index) == 3
    k = uncompressed_index[0]
    s = xy_scaling_log2[k]
    return (k, uncompressed_index[1] >> s, uncompressed_index[2] >> s)


def _layer_shape(k, xy_scaling_log2, uncompressed_shape):
    s = xy_scaling_log2[k]
    return (uncompressed_shape[1] >> s, uncompressed_shape[2] >> s)


def compressed_single_index(compressed_index, xy_scaling_log2, uncompressed_shape):
    """(k, j', i') -> position in the flat array of all coarsened layers."""
    k, j, i = compressed_index
    index = 0
    for zpos in range(k):
        ly, lx = _layer_shape(zpos, xy_scaling_log2, uncompressed_shape)
        index += ly * lx
    return index + j * _layer_shape(k, xy_scaling_log2, uncompressed_shape)[1] + i


def compressed_index_from_single(single_index, xy_scaling_log2, uncompressed_shape):
    """inverse of :func:`compressed_single_index`"""
    start = 0
    for k in range(uncompressed_shape[0]):
        ly, lx = _layer_shape(k, xy_scaling_log2, uncompressed_shape)
        if single_index < start + ly * lx:
            rem = single_index - start
            return (k, rem // lx, rem % lx)
        start += ly * lx
    raise IndexError("single index %d beyond the compressed grid" % single_index)


def uncompressed_index_range(compressed_index, xy_scaling_log2):
    """(k, j', i') -> ((k, k+1), (j0, j1), (i0, i1)): the fine cells one coarse cell covers."""
    k = compressed_index[0]
    f = 1 << xy_scaling_log2[k]
    return ((k, k + 1), (compressed_index[1] * f, (compressed_index[1] + 1) * f),
            (compressed_index[2] * f, (compressed_index[2] + 1) * f))
