"""Zernike polynomials and index conventions (mirrors pyotf/zernike.py).

The index maps and name tables are tiny host-side integer arithmetic; ``zernike`` itself is
evaluated on the GPU (csrc/zernike.cuh, C ABI ``pyotf_b200_zernike``) in float64.
"""
from typing import Tuple

import numpy as np

from .utils import cart2pol  # noqa: F401  (re-exported like the reference module does)

# classical aberration names keyed by (radial degree n, azimuthal degree m)
# (pyotf/zernike.py:25-76; see https://en.wikipedia.org/wiki/Zernike_polynomials)
degrees2name = {
    (0, 0): "piston",
    (1, -1): "tilt",
    (1, 1): "tip",
    (2, -2): "oblique astigmatism",
    (2, 0): "defocus",
    (2, 2): "vertical astigmatism",
    (3, -3): "vertical trefoil",
    (3, -1): "vertical coma",
    (3, 1): "horizontal coma",
    (3, 3): "oblique trefoil",
    (4, -4): "oblique quadrafoil",
    (4, -2): "oblique secondary astigmatism",
    (4, 0): "primary spherical",
    (4, 2): "vertical secondary astigmatism",
    (4, 4): "vertical quadrafoil",
    (5, -3): "vertical secondary trefoil",
    (5, -1): "vertical secondary coma",
    (5, 1): "horizontal seconday coma",
    (5, 3): "horizontal secondary trefoil",
    (6, -4): "oblique secondary quadrafoil",
    (6, -2): "oblique tertiary astigmatism",
    (6, 0): "secondary spherical",
    (6, 2): "vertical tertiary astigmatism",
    (6, 4): "vertical secondary quadrafoil",
    (7, -3): "vertical tertiary trefoil",
    (7, -1): "vertical tertiary coma",
    (7, 1): "horizontal tertiary coma",
    (7, 3): "horizontal tertiary trefoil",
    (8, -2): "oblique quaternary astigmatism",
    (8, 0): "tertiary spherical",
    (8, 2): "vertical quaternary astigmatism",
    (9, -3): "vertical quaternary trefoil",
    (9, -1): "vertical quaternary coma",
    (9, 1): "horizontal quaternary coma",
    (9, 3): "horizontal quaternary trefoil",
    (10, 0): "quaternary spherical",
}
name2degrees = {name: deg for deg, name in degrees2name.items()}
assert len(name2degrees) == len(degrees2name)


def _check_degrees(n, m):
    """Validate (n, m): n >= |m|, n - m even, both signed-integer typed (zernike.py:84-98)."""
    n, m = np.asarray(n), np.asarray(m)
    if np.any((n < abs(m)) | ((n - m) % 2 == 1)):
        raise ValueError(f"Invalid combination of ({n}, {m}) in Noll indexing")
    for what, arr in (("Radial", n), ("Azimuthal", m)):
        if not np.issubdtype(arr.dtype, np.signedinteger):
            raise ValueError(f"{what} degree is not integer, input = {arr}")
    return n, m


def _check_index(j):
    """Validate a sequential index: positive and signed-integer typed (zernike.py:146-155)."""
    j = np.asarray(j)
    if not (j > 0).all():
        raise ValueError("Invalid Noll index")
    if not np.issubdtype(j.dtype, np.signedinteger):
        raise ValueError(f"Index is not integer, input = {j}")
    return j


def degrees2osa(n: int, m: int) -> int:
    """(n, m) -> OSA/ANSI index j = (n(n+2) + m) / 2 (zernike.py:101-111).  Vectorised."""
    n, m = _check_degrees(n, m)
    return ((n * (n + 2) + m) / 2).astype(int)


def degrees2fringe(n: int, m: int) -> int:
    """(n, m) -> Fringe (University of Arizona) index (zernike.py:114-121).  Vectorised."""
    n, m = _check_degrees(n, m)
    am = abs(m)
    return ((1 + (n + am) / 2) ** 2 - 2 * am + (1 - np.sign(m)) / 2).astype(int)


def degrees2noll(n: int, m: int) -> int:
    """(n, m) -> Noll index (zernike.py:124-143).  Vectorised.

    j = n(n+1)/2 + |m| + p with p in {0, 1} chosen so that even j <-> m > 0 (cosine terms).
    """
    n, m = _check_degrees(n, m)
    low = (n % 4) < 2
    p = np.where(low, np.where(m > 0, 0, 1), np.where(m < 0, 0, 1))
    return (n * (n + 1) / 2 + abs(m) + p).astype(int)


def osa2degrees(j: int) -> Tuple[int, int]:
    """OSA/ANSI index -> (n, m) (zernike.py:158-170).  Vectorised."""
    j = _check_index(j)
    n = np.ceil((-3 + np.sqrt(9 + 8 * j)) / 2).astype(int)
    m = 2 * j - n * (n + 2)
    return n.astype(int), m.astype(int)


def noll2degrees(j: int) -> Tuple[int, int]:
    """Noll index -> (n, m) (zernike.py:173-220).  Vectorised."""
    j = _check_index(j)
    n = np.ceil((-3 + np.sqrt(1 + 8 * j)) / 2).astype(int)
    jr = j - (n * (n + 1) / 2).astype(int)
    low = (n % 4) < 2
    cand_pos = np.where(low, jr, jr - 1)
    cand_neg = np.where(low, -(jr - 1), -jr)
    m = np.where((n - cand_pos) % 2 == 0, cand_pos, cand_neg)
    return n, m


def fringe2degrees(j: int):
    """Not provided by the reference either (zernike.py:223-228)."""
    raise NotImplementedError


def _name_table(index_of):
    table = {index_of(n, m): name for (n, m), name in degrees2name.items()}
    table = dict(sorted(table.items()))
    return table, {name: idx for idx, name in table.items()}


noll2name, name2noll = _name_table(degrees2noll)
osa2name, name2osa = _name_table(degrees2osa)
fringe2name, name2fringe = _name_table(degrees2fringe)


def zernike(r: float, theta: float, n: int, m: int, norm: bool = True) -> float:
    """Zernike polynomials Z_n^m(r, theta) on the unit disk (zernike.py:251-365).

    ``n`` / ``m`` may be scalars or 1-D sequences; the result is stacked over modes and
    squeezed like the reference.  Values are zero for r > 1 and when n - m is odd.
    Evaluated on the GPU in float64.

    >>> x = np.linspace(-1, 1, 512)
    >>> xx, yy = np.meshgrid(x, x)
    >>> r, theta = cart2pol(yy, xx)
    >>> zern = zernike(r, theta, 4, 0)  # doctest: +SKIP
    """
    n, m = np.asarray(n), np.asarray(m)
    if n.ndim > 1:
        raise ValueError("Radial degree has the wrong shape")
    if m.ndim > 1:
        raise ValueError("Azimuthal degree has the wrong shape")
    if n.shape != m.shape:
        raise ValueError("Radial and Azimuthal degrees have different shapes")
    r = np.asarray(r, dtype=float)
    theta = np.asarray(theta, dtype=float)
    if not (r >= 0).all():
        raise ValueError("r must always be greater or equal to 0")
    if r.ndim > 2:
        raise ValueError("Input rho and theta cannot have more than two dimensions")
    n, m = n.ravel(), m.ravel()
    if not (n >= abs(m)).all():
        raise ValueError("n must always be greater or equal to m")
    from . import _engine

    r_b, theta_b = np.broadcast_arrays(r, theta)
    out = _engine.zernike_eval(r_b, theta_b, n, m, bool(norm))
    return out.reshape((len(n),) + r_b.shape).squeeze()
