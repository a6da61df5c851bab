"""heatsim2_b200.tridiag (GPU scans) against the reference algorithm
(oracle restatement of heatsim2/tridiag.pyx) - includes the reference's own
self-check (tridiag.pyx:71-112: random 4x4, |LU-A|/|A| < 1e-10, solve vs inv)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_system(n, rng, dominant=True):
    A = rng.random((n, 3))
    if dominant:
        A[:, 1] += 2.0
    A[0, 0] = 0.0
    A[-1, 2] = 0.0
    return A


def _dense(A):
    n = A.shape[0]
    M = np.diag(A[:, 1])
    M += np.diag(A[1:, 0], -1) + np.diag(A[:-1, 2], 1)
    return M


def test_reference_selfcheck_4x4():
    from heatsim2_b200 import tridiag
    rng = np.random.default_rng(0)
    A = _random_system(4, rng, dominant=False)
    A[:, 1] += 1.0
    L, U = tridiag.tridiaglu(A)
    Ld = np.diag(L[:, 1]) + np.diag(L[1:, 0], -1)
    Ud = np.diag(U[:, 1]) + np.diag(U[:-1, 2], 1)
    Ad = _dense(A)
    assert np.linalg.norm(Ld @ Ud - Ad) / np.linalg.norm(Ad) < 1e-10
    b = rng.random(4)
    x = tridiag.tridiagsolve(L, U, b)
    assert np.linalg.norm(x - np.linalg.inv(Ad) @ b) / np.linalg.norm(x) < 1e-10


@pytest.mark.parametrize("n", [1, 2, 7, 2048, 2049, 100003, 1 << 21])
def test_lu_and_solve_match_reference_algorithm(n):
    import adi_oracle
    from heatsim2_b200 import tridiag
    rng = np.random.default_rng(n)
    A = _random_system(n, rng)
    if n > 10:          # independent blocks like the ADI use (zeros break the chain)
        A[::97, 0] = 0.0
        A[96::97, 2] = 0.0
    b = rng.random(n)
    L, U = tridiag.tridiaglu(A)
    if n <= 200000:
        Lo, Uo = adi_oracle.tridiaglu(A)
        xo = adi_oracle.tridiagsolve(Lo, Uo, b)
    else:
        import heatsim2_b200.tridiag as t
        Lo, Uo = t.tridiaglu_host(A)
        xo = None
    assert np.abs(L - Lo).max() <= 1e-13 * np.abs(Lo).max()
    assert np.abs(U - Uo).max() <= 1e-13 * np.abs(Uo).max()
    x = tridiag.tridiagsolve(L, U, b)
    if xo is not None:
        assert np.abs(x - xo).max() <= 1e-13 * np.abs(xo).max()
    # residual check at any size
    r = A[:, 1] * x
    r[1:] += A[1:, 0] * x[:-1]
    r[:-1] += A[:-1, 2] * x[1:]
    assert np.abs(r - b).max() <= 1e-12


def test_device_tensors_stay_on_device():
    import torch
    from heatsim2_b200 import tridiag
    A = torch.from_numpy(_random_system(1000, np.random.default_rng(1))).cuda()
    L, U = tridiag.tridiaglu(A)
    x = tridiag.tridiagsolve(L, U, torch.ones(1000, dtype=torch.float64, device="cuda"))
    assert L.is_cuda and U.is_cuda and x.is_cuda


def test_rejects_open_chain():
    from heatsim2_b200 import tridiag
    A = np.ones((5, 3))
    with pytest.raises(AssertionError):
        tridiag.tridiaglu(A)
