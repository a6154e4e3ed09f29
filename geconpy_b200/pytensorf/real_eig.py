"""``gEconpy.pytensorf.real_eig`` on B200 (real_eig.py:10-140).

``real_eig(M) -> (re, im)``: eigenvalues of a real general matrix as two real arrays, sorted by ascending modulus.  Numeric
inputs (numpy / torch CUDA, optionally with a leading batch axis) run ``gecon_real_eig_*`` -- balancing, Householder
Hessenberg reduction and Francis double-shift QR with the matrix resident in shared memory, one warp per matrix.  Symbolic
inputs go through the ``RealEig`` Op, whose ``perform`` makes the same call (one launch for the whole batch under Blockwise).

On the likelihood path the reference only needs the COUNT of ``|lambda| > 1`` (perturbation.py:499-505); the kernel behind
``count_outside_unit_circle`` / ``check_bk_condition_pt`` gets it from a matrix-sign iteration without forming eigenvalues.
The eigenvalues themselves are for diagnostics (``check_bk_condition(return_value="dataframe")``) and to cross-check the count.
"""

from __future__ import annotations

import numpy as np

from .. import batched
from ..solvers._pt import HAVE_PYTENSOR, Apply, Op, pt, require_pytensor


def count_outside_unit_circle(M):
    """#{|eig(M)| > 1} for square M (numpy, optional leading batch axis), on the GPU, WITHOUT eigenvalues.

    Uses the Blanchard-Kahn kernel on the pencil whose matrix is M:  with A = M, B = -I (so that G = I + 1e-8 I) and
    C = 0, no lead columns, the kernel's M-matrix is (1 + 1e-8)^-1 M."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    n = M.shape[-1]
    I = np.broadcast_to(np.eye(n), M.shape)
    nu, _st = batched.bk_count(M, -np.ascontiguousarray(I), np.zeros_like(M), np.zeros(0, dtype=np.int32))
    return nu


class RealEig(Op):
    """The reference's Op contract (real_eig.py:10-36): ``(m,m)->(m),(m)``, both outputs real, ascending modulus.  The
    pullback of the reference recomputes eigenVECTORS with ``pt.linalg.eig`` (real_eig.py:38-62); the estimation graph
    detaches the eigenvalues before use (perturbation.py:612-616), so no gradient ever flows here and none is provided."""

    __props__ = ()
    gufunc_signature = "(m,m)->(m),(m)"

    def __init__(self):
        require_pytensor("RealEig")
        super().__init__()

    def make_node(self, M):
        M = pt.as_tensor_variable(M)
        if M.type.ndim != 2:
            raise ValueError(f"RealEig requires a 2-d matrix, got ndim={M.type.ndim}")
        n = M.type.shape[0]
        outputs = [pt.vector(dtype=M.type.dtype, shape=(n,)), pt.vector(dtype=M.type.dtype, shape=(n,))]
        return Apply(self, [M], outputs)

    def infer_shape(self, fgraph, node, input_shapes):
        return [(input_shapes[0][0],), (input_shapes[0][0],)]

    def perform(self, node, inputs, outputs):
        (M,) = inputs
        M = np.ascontiguousarray(M, dtype=np.float64)
        lead, m = M.shape[:-2], M.shape[-1]
        re, im, _st = batched.real_eig(M.reshape(-1, m, m))
        dt = node.outputs[0].type.dtype
        outputs[0][0] = np.asarray(re, dtype=dt).reshape(*lead, m)
        outputs[1][0] = np.asarray(im, dtype=dt).reshape(*lead, m)


def real_eig(M):
    """``(re, im)`` sorted by ascending modulus (real_eig.py:65-100)."""
    if HAVE_PYTENSOR and hasattr(M, "owner") and hasattr(M, "type"):
        return RealEig()(M)
    re, im, _st = batched.real_eig(M)
    return re, im


__all__ = ["RealEig", "real_eig", "count_outside_unit_circle"]
