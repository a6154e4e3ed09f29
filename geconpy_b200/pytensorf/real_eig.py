"""``gEconpy.pytensorf.real_eig`` (real_eig.py:10-140).

``real_eig(M) -> (re, im)`` returns eigenvalues sorted by modulus; on the hot path it exists only to count
``|lambda| > 1`` (perturbation.py:499-505).  The B200 kernel ``gecon_bk_count_*`` produces that count directly with a
matrix-sign iteration, so no device eigen-solver exists; ``count_outside_unit_circle`` exposes the count for a general
square matrix M (the pencil (I, M))."""

from __future__ import annotations

import numpy as np

from .. import batched


def count_outside_unit_circle(M):
    """#{|eig(M)| > 1} for square M (numpy, optional leading batch axis), on the GPU.

    Uses the Blanchard-Kahn kernel on the pencil whose matrix is M:  with A = M, B = -I (so that G = I + 1e-8 I) and
    C = 0, no lead columns, the kernel's M-matrix is (1 + 1e-8)^-1 M."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    n = M.shape[-1]
    I = np.broadcast_to(np.eye(n), M.shape)
    nu, _st = batched.bk_count(M, -np.ascontiguousarray(I), np.zeros_like(M), np.zeros(0, dtype=np.int32))
    return nu


def real_eig(M):
    raise NotImplementedError(
        "no device eigen-solver: the hot path needs only the count of eigenvalues outside the unit circle "
        "(count_outside_unit_circle / check_bk_condition_pt)"
    )
