"""``gEconpy.solvers.gensys`` on B200.

Reference: gEconpy/solvers/gensys.py -- ``determine_n_unstable`` (:52-95), ``split_matrix_on_eigen_stability``
(:98-118), ``gensys`` (:398-521), ``interpret_gensys_output`` (:524-565), ``_gensys_setup`` (:568-614),
``solve_policy_function_with_gensys`` (:617-631), ``GensysWrapper`` / ``gensys_pt`` (:634-683).

Scope: Sims' complex sorted QZ is NOT re-implemented on the GPU.  What the estimation
path consumes from gensys is ``T = G1[:n, :n]`` and ``success = (eu[0] == 1 and eu[1] == 1)``; the reference itself
asserts that this T equals the cycle-reduction T to 1e-8 (tests/model/test_perturbation.py:166-206).  Here

* T, R come from the batched cycle-reduction kernel (``gecon_cr_solve_*``),
* existence / uniqueness come from the Blanchard-Kahn eigenvalue count of the same Sims pencil
  (``gecon_bk_count_*``): n_unstable == n_forward <=> eu = [1, 1, 0]; fewer unstable roots than forward-looking
  variables = indeterminacy (eu = [1, 0, n_forward - n_unstable]); more = no stable solution (eu = [0, 1, 0]).

The pure bookkeeping helpers of the module (pencil assembly, eigenvalue classification, return-code messages) are
host code and are provided with the reference's signatures.  ``gensys`` / ``build_u_v_d`` (the QZ + SVD machinery with
its f_mat / f_wt / y_wt / gev outputs) stay CPU-only in the reference.  ``gensys`` itself is provided for the pencils
``_gensys_setup`` assembles (the only ones its callers build); a general pencil and ``build_u_v_d`` raise ``NotImplementedError``.
"""

from __future__ import annotations

import numpy as np

from .. import _lib as L
from .. import batched
from ._pt import HAVE_PYTENSOR, Apply, Op, pt, require_pytensor


def determine_n_unstable(alpha, beta, div, realsmall):
    """Classify generalized eigenvalues beta_i / alpha_i: ``(div, n_unstable, zxz)`` (gensys.py:26-95)."""
    infer = div is None
    cut = 1.01 if infer else float(div)
    count = 0
    zxz = False
    for a, b in zip(np.abs(np.asarray(alpha)), np.abs(np.asarray(beta))):
        if infer and a > 0:
            ratio = b / a
            if 1 + realsmall < ratio <= cut:
                cut = 0.5 * (1 + ratio)
        if b > cut * a:
            count += 1
        zxz = bool(a < realsmall and b < realsmall)
    return float(cut), int(count), bool(zxz)


def split_matrix_on_eigen_stability(A, n_unstable: int):
    """Rows of the stable block first, the trailing ``n_unstable`` rows second (gensys.py:98-118)."""
    A = np.asarray(A)
    cut = A.shape[0] - int(n_unstable)
    return A[:cut], A[cut:]


def _gensys_setup(A, B, C, D, tol: float = 1e-8):
    """Sims pencil of the linearised model: ``(G0, G1, const, Psi, Pi)`` with ``G0 = -Gamma0`` restricted to the
    equations plus the numerically non-zero lead columns of C (gensys.py:568-614)."""
    A, B, C, D = (np.asarray(x, dtype=np.float64) for x in (A, B, C, D))
    n, k = A.shape[0], D.shape[1]
    lead = np.flatnonzero(np.abs(C).sum(axis=0) > tol)
    keep = np.concatenate([np.arange(n), n + lead])
    O, I = np.zeros((n, n)), np.eye(n)
    gamma0 = np.block([[B, C], [-I, O]])[np.ix_(keep, keep)]
    gamma1 = np.block([[A, O], [O, I]])[np.ix_(keep, keep)]
    psi = np.vstack([D, np.zeros((n, k))])[keep]
    pi = np.vstack([O, I])[keep][:, lead]
    return -gamma0, gamma1, np.zeros((keep.size, 1)), psi, pi


def interpret_gensys_output(eu):
    """Human-readable meaning of gensys' ``eu`` codes (gensys.py:524-565)."""
    head = f"Gensys return codes: {' '.join(map(str, eu))}, with the following meaning:\n"
    e0, e1 = eu[0], eu[1]
    if e0 == -2 and e1 == -2:
        body = "Coincident zeros.  Indeterminacy and/or nonexistence. Check that your system is correctly defined."
    elif e0 == -1:
        body = f"System is indeterminate. There are {eu[2]} loose endogenous variables."
    elif e1 == -1:
        body = "Solution exists, but it is not unique -- sunspots."
    elif e0 == 0 and e1 == 0:
        body = "Solution does not exist."
    elif e0 == 1 and e1 == 0:
        body = "Solution exists, but is not unique."
    elif e0 == 1 and e1 == 1:
        body = "Gensys found a unique solution."
    else:
        body = "Unknown return code. Check the gensys documentation."
    return (head + body).strip()


def _structural_lead_idx(C, tol):
    return np.flatnonzero(np.abs(np.asarray(C)).sum(axis=-2).reshape(-1, np.shape(C)[-1]).max(axis=0) > tol).astype(np.int32)


def _eu_from_count(n_unstable: int, n_forward: int, status: int):
    if status & L.ST_BK_INCONCLUSIVE or n_unstable < 0:
        return [-2, -2, 0]
    if n_unstable == n_forward:
        return [1, 1, 0]
    if n_unstable < n_forward:
        return [1, 0, int(n_forward - n_unstable)]
    return [0, 1, 0]


def solve_policy_function_with_gensys(A, B, C, D, tol: float = 1e-8, return_all_matrices: bool = True):
    """Same 9-tuple layout as the reference (``G_1, constant, impact, f_mat, f_wt, y_wt, gev, eu, loose``) so that
    ``Model._solve_with_gensys``'s slicing ``T = G_1[:n, :n]``, ``R = impact[:n]`` keeps working (model.py:1682-1709).
    ``G_1`` is the n x n policy matrix itself and ``impact`` is R; the forward-solution matrices are not computed
    (None).  On failure ``G_1`` and ``impact`` are None, as in gensys.py:515-516."""
    A, B, C, D = (np.ascontiguousarray(x, dtype=np.float64) for x in (A, B, C, D))
    if A.ndim != 2:
        raise ValueError("solve_policy_function_with_gensys takes single (n, n) matrices; see gensys_batched for batches")
    lead = np.flatnonzero(np.abs(C).sum(axis=0) > tol).astype(np.int32)
    out = batched.cr_solve(A, B, C, D, max_iter=1000, tol=min(tol, 1e-9))
    nu, st = batched.bk_count(A, B, C, lead)
    eu = _eu_from_count(int(nu), int(lead.size), int(st))
    ok = eu[0] == 1 and eu[1] == 1 and bool(out.converged) and not (int(out.status) & L.ST_SINGULAR)
    if not ok:
        if eu[0] == 1 and eu[1] == 1:  # the count is fine but the iteration failed: report non-existence
            eu = [0, 0, 0]
        return (None, None, None, None, None, None, None, eu, None) if return_all_matrices else (None, None, None, eu)
    G_1, impact = np.ascontiguousarray(out.T), np.ascontiguousarray(out.R)
    constant = np.zeros((A.shape[0], 1))
    if return_all_matrices:
        return G_1, constant, impact, None, None, None, None, eu, None
    return G_1, constant, impact, eu


def gensys_batched(A, B, C, D, lead_idx=None, tol: float = 1e-8, max_iter: int = 1000):
    """Batched ``(T, R, success)`` over a leading draw axis: what ``gensys_pt`` returns, for a whole population.
    ``lead_idx``: structural lead-variable columns (defaults to the numerically non-zero columns of C)."""
    lead = _structural_lead_idx(C, tol) if lead_idx is None else np.asarray(lead_idx, dtype=np.int32)
    out = batched.cr_solve(A, B, C, D, max_iter=max_iter, tol=min(tol, 1e-9))
    nu, st = batched.bk_count(A, B, C, lead)
    success = ((out.status | st) & (L.ST_CR_NOT_CONVERGED | L.ST_SINGULAR | L.ST_BK | L.ST_BK_INCONCLUSIVE)) == 0
    return out.T, out.R, success


def _pencil_to_model(g0, g1, psi, pi, tol):
    """Recognise a Sims pencil assembled by ``_gensys_setup`` (gensys.py:568-614) and recover ``(A, B, C, D, lead)``; None if the
    pencil does not have that block structure."""
    g0, g1, psi, pi = (np.asarray(x, dtype=np.float64) for x in (g0, g1, psi, pi))
    m = g0.shape[0]
    nl = pi.shape[1] if pi.ndim == 2 else 0
    n = m - nl
    if g0.shape != (m, m) or g1.shape != (m, m) or n < 1 or psi.shape[0] != m or pi.shape[0] != m:
        return None
    sel = g0[n:, :n]
    ok = (
        np.array_equal(g1[n:, n:], np.eye(nl)) and not g1[:n, n:].any() and not g1[n:, :n].any() and not g0[n:, n:].any()
        and not psi[n:].any() and not pi[:n].any() and np.array_equal(pi[n:], np.eye(nl))
        and np.array_equal(sel.sum(axis=1), np.ones(nl)) and np.isin(sel, (0.0, 1.0)).all()
    )  # fmt: skip
    if not ok:
        return None
    lead = sel.argmax(axis=1).astype(np.int32)
    C = np.zeros((n, n))
    C[:, lead] = -g0[:n, n:]
    return np.ascontiguousarray(g1[:n, :n]), np.ascontiguousarray(-g0[:n, :n]), C, np.ascontiguousarray(psi[:n]), lead


def gensys(g0, g1, c, psi, pi, div=None, tol=1e-8, return_all_matrices=True):
    """``gensys(g0, g1, c, psi, pi, div, tol, return_all_matrices)`` (gensys.py:398-521) for the pencils this package's callers
    build: ``g0, g1, c, psi, pi = _gensys_setup(A, B, C, D)``.  The block structure is recognised, the model matrices are
    recovered and the system is solved by cycle reduction + the Blanchard-Kahn count; the result is embedded
    in the pencil's coordinates ``y_t = [x_t ; E_t x_{t+1, lead}]``:

        G_1 = [[T, 0], [(T T)[lead], 0]],   impact = [[R], [(T R)[lead]]],   C = 0,   eu as in ``solve_policy_function_with_gensys``.

    ``f_mat, f_wt, y_wt, gev, loose`` (by-products of the complex QZ + SVD machinery) are None.  A pencil WITHOUT that structure
    needs Sims' general algorithm, which is not on the B200 path: ``NotImplementedError``."""
    rec = _pencil_to_model(g0, g1, psi, pi, tol)
    if rec is None:
        raise NotImplementedError(
            "gensys on a general pencil needs the complex QZ + SVD algorithm, which is outside the B200 hot path (SURVEY.md 8a, "
            "row a7). Supported: pencils assembled by _gensys_setup(A, B, C, D); or call solve_policy_function_with_gensys / "
            "gensys_batched with the model matrices."
        )
    A, B, C, D, lead = rec
    n, m = A.shape[0], np.shape(g0)[0]
    out = solve_policy_function_with_gensys(A, B, C, D, tol=tol, return_all_matrices=True)
    T, R, eu = out[0], out[2], out[7]
    if T is None:
        return (None, None, None, None, None, None, None, eu, None) if return_all_matrices else (None, None, None, eu)
    TT, TR = batched.gemm(T, T), pt_free_matmul(T, R)
    G_1 = np.zeros((m, m))
    G_1[:n, :n], G_1[n:, :n] = T, TT[lead]
    impact = np.vstack([R, TR[lead]])
    const = np.zeros((m, 1))
    if return_all_matrices:
        return G_1, const, impact, None, None, None, None, eu, None
    return G_1, const, impact, eu


def pt_free_matmul(T, R):
    """(n, n) @ (n, k) on the GPU through the square-product kernel (R zero-padded to n columns)."""
    n, k = R.shape
    Rp = np.zeros((n, n))
    Rp[:, :k] = R
    return batched.gemm(T, Rp)[:, :k]


def build_u_v_d(eta, realsmall=2.2204460492503131e-16):
    raise NotImplementedError(
        "build_u_v_d (the rank-revealing SVD of Q2 Pi inside Sims' gensys, gensys.py:128-172) belongs to the CPU-only QZ/SVD "
        "machinery; the B200 path decides existence / uniqueness from the Blanchard-Kahn count (SURVEY.md 8a, row a7)"
    )


# ---------------------------------------------------------------------------------------------------- pytensor layer
class GensysWrapper(Op):
    """pytensor Op with the reference's contract (gensys.py:634-676): ``(n,n),(n,n),(n,n),(n,k)->(n,n),()`` returning
    the policy matrix and a boolean ``success``; one kernel pair per call, batched over leading axes."""

    __props__ = ("tol",)
    gufunc_signature = "(n,n),(n,n),(n,n),(n,k)->(n,n),()"

    def __init__(self, tol=1e-8):
        require_pytensor("GensysWrapper")
        self.tol = tol
        super().__init__()

    def make_node(self, A, B, C, D):
        from .cycle_reduction import linalg_output_dtype

        inputs = list(map(pt.as_tensor, [A, B, C, D]))
        o_dtype = linalg_output_dtype(*(inp.type.dtype for inp in inputs))
        outputs = [pt.tensor("T", dtype=o_dtype, shape=inputs[0].type.shape), pt.scalar("success", dtype="bool")]
        return Apply(self, inputs, outputs)

    def infer_shape(self, fgraph, node, input_shapes):
        return [input_shapes[0], ()]

    def perform(self, node, inputs, outputs):
        A, B, C, D = (np.ascontiguousarray(x, dtype=np.float64) for x in inputs)
        lead, n, k = A.shape[:-2], A.shape[-1], D.shape[-1]
        T, _R, ok = gensys_batched(A.reshape(-1, n, n), B.reshape(-1, n, n), C.reshape(-1, n, n), D.reshape(-1, n, k), tol=self.tol)
        outputs[0][0] = np.asarray(T, dtype=node.outputs[0].type.dtype).reshape(*lead, n, n)
        outputs[1][0] = np.asarray(ok).reshape(lead) if lead else np.bool_(np.asarray(ok).reshape(-1)[0])

    def pullback(self, inputs, outputs, cotangents):
        from .cycle_reduction import _linear_policy_jvp

        A_bar, B_bar, C_bar = _linear_policy_jvp(inputs, outputs, cotangents)
        return [A_bar, B_bar, C_bar, pt.zeros_like(inputs[3])]


def gensys_pt(A, B, C, D, tol=1e-8):
    """Symbolic ``(T, R, success)`` (gensys.py:679-683)."""
    from .shared import pt_compute_selection_matrix

    T, success = GensysWrapper(tol=tol)(A, B, C, D)
    return T, pt_compute_selection_matrix(B, C, D, T), success


__all__ = [
    "GensysWrapper", "determine_n_unstable", "split_matrix_on_eigen_stability", "interpret_gensys_output", "_gensys_setup",
    "solve_policy_function_with_gensys", "gensys_batched", "gensys", "gensys_pt", "build_u_v_d", "HAVE_PYTENSOR",
]  # fmt: skip
