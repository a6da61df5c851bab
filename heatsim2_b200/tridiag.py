"""Tridiagonal LU and solve - drop-in for the reference's
``heatsim2/tridiag.pyx`` (``tridiaglu`` :10, ``tridiagsolve`` :47).

``Amat`` is ``n x 3`` (sub-diagonal, diagonal, super-diagonal per row; the
first sub- and last super-diagonal entries must be 0).  ``tridiaglu`` returns
``(Lmat, Umat)`` in the reference's layout (``Lmat[:,0]`` sub-diagonal,
``Lmat[:,1]`` pivots, ``Umat[:,1]`` ones, ``Umat[:,2]`` scaled super-diagonal)
and ``tridiagsolve`` solves ``L U x = b``.

Both run on the GPU through the C ABI (``hs2_tridiag_lu`` /
``hs2_tridiag_solve``): the two first-order recurrences of the solve and the
continued-fraction recurrence of the factorisation are evaluated as parallel
scans over the whole chain.  numpy in -> numpy out, CUDA tensor in -> CUDA
tensor out.  ``tridiaglu_host`` is a small numpy helper used only to
materialise ``pyadi_step.Lmat/Umat`` for inspection.
"""
import ctypes

import numpy as np
import torch

from . import _cabi


def _dev_tensor(a, name, ncols=None):
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    if t.dtype != torch.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double' for %s" % name)
    if ncols is not None and (t.dim() != 2 or t.shape[1] != ncols):
        raise ValueError("%s must be n x %d" % (name, ncols))
    if ncols is None and t.dim() != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % t.dim())
    if not torch.cuda.is_available():
        raise RuntimeError("heatsim2_b200.tridiag needs a CUDA device; there is no CPU path")
    return t.cuda().contiguous(), isinstance(a, torch.Tensor) and a.is_cuda


def _scratch(n, dev):
    nbytes = _cabi.lib().hs2_tridiag_scratch_bytes(n)
    return torch.empty(max(1, (nbytes + 7) // 8), dtype=torch.float64, device=dev)


def tridiaglu(Amat):
    A, on_dev = _dev_tensor(Amat, "Amat", 3)
    n = A.shape[0]
    if n == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    first, last = float(A[0, 0]), float(A[-1, 2])
    assert first == 0.0
    assert last == 0.0
    L = torch.zeros_like(A)
    U = torch.empty_like(A)
    scratch = _scratch(n, A.device)
    with torch.cuda.device(A.device):
        stream = torch.cuda.current_stream().cuda_stream
        _cabi.check(_cabi.lib().hs2_tridiag_lu(n, A.data_ptr(), L.data_ptr(), U.data_ptr(), scratch.data_ptr(),
                                               ctypes.c_void_p(stream)))
    if on_dev:
        return (L, U)
    return (L.cpu().numpy(), U.cpu().numpy())


def tridiagsolve(Lmat, Umat, bvec):
    L, _ = _dev_tensor(Lmat, "Lmat", 3)
    U, _ = _dev_tensor(Umat, "Umat", 3)
    b, on_dev = _dev_tensor(bvec, "bvec")
    n = b.shape[0]
    if L.shape[0] != n or U.shape[0] != n:
        raise ValueError("Lmat, Umat and bvec must have the same number of rows")
    x = torch.empty_like(b)
    scratch = _scratch(n, b.device)
    with torch.cuda.device(b.device):
        stream = torch.cuda.current_stream().cuda_stream
        _cabi.check(_cabi.lib().hs2_tridiag_solve(n, L.data_ptr(), U.data_ptr(), b.data_ptr(), x.data_ptr(),
                                                  scratch.data_ptr(), ctypes.c_void_p(stream)))
    if on_dev:
        return x
    return x.cpu().numpy()


def tridiaglu_host(Amat):
    """numpy Thomas factorisation in the reference's (Lmat, Umat) layout, for
    host-side inspection of small systems (not used by any time step)."""
    A = np.asarray(Amat, dtype=np.float64)
    n = A.shape[0]
    L = np.zeros((n, 3))
    U = np.zeros((n, 3))
    U[:, 1] = 1.0
    piv = A[0, 1]
    for r in range(n):
        L[r, 1] = piv
        U[r, 2] = A[r, 2] / piv
        if r < n - 1:
            L[r + 1, 0] = A[r + 1, 0]
            piv = A[r + 1, 1] - U[r, 2] * A[r + 1, 0]
    return (L, U)
