"""``gEconpy.model.perturbation`` on B200: linearisation entry point and Blanchard-Kahn checks.

Reference: gEconpy/model/perturbation.py -- ``linearize_model`` (:29-198), ``check_bk_condition`` (:508-583),
``check_bk_condition_pt`` (:586-625), ``compute_bk_eigenvalues(_pt)`` (:412-505).
"""

from __future__ import annotations

import logging

import numpy as np

from .. import _lib as L
from .. import batched
from .compiled import CompiledModel

_log = logging.getLogger(__name__)
_FLOAT_ZERO_TOL = 1e-8


def _names(xs):
    return [x if isinstance(x, str) else getattr(x, "base_name", None) or getattr(x, "name", None) or str(x) for x in xs]


def linearize_model(variables, equations=None, shocks=None, cache=None, loglin_variables=None, order: int = 1, eq_order=None, var_order=None,
                    log_linearize: bool = True, not_loglin_variables=()):
    """The reference's entry point and argument list (perturbation.py:29-38):
    ``linearize_model(variables, equations, shocks, cache, loglin_variables, order, eq_order, var_order) ->
    ([A, B, C, D], ss_nodes, eq_order, var_order)``.

    * ``variables`` / ``shocks``: names (or symbols carrying ``base_name`` / ``name``); ``equations``: one expression per
      variable, sympy or text, written in this repo's time notation ``<name>__tm1 | __t | __tp1 | __ss``
      (``tests/golden/make_models.py`` converts the front-end's TimeAwareSymbols to it);
    * ``cache``: where the reference keeps its sympy -> pytensor node cache.  Here it carries the rest of the model that the
      generated kernel needs: ``{"name", "free_params", "deterministic_params", "steady_state", "assumptions", "linear"}``;
    * ``loglin_variables``: the variables to log-linearise (None = all), as in the reference;
    * ``eq_order`` / ``var_order``: permutations to check against the ones computed from the incidence structure.

    Returns the four Jacobians as ``sympy`` matrices (rows ``eq_order``, columns ``var_order``; entries in the steady-state
    symbols returned as ``ss_nodes`` and the parameter symbols; log-linear column scaling NOT applied: it is evaluated per
    draw by the generated kernel, perturbation.py:178-190), the steady-state symbols, and the two permutations.  The
    numerical counterpart -- the same four matrices for a whole batch of draws from one generated CUDA kernel -- is
    ``CompiledModel(spec).jacobian(theta)``; calling ``linearize_model(spec)`` with a spec (dict / path / name) returns
    ``(CompiledModel, eq_order, var_order)``."""
    if order != 1:
        raise NotImplementedError("Only order = 1 linearization is currently implemented.")
    if equations is None and shocks is None:  # a model spec: the batched evaluator
        cm = CompiledModel(variables, log_linearize=log_linearize, not_loglin_variables=not_loglin_variables)
        return cm, cm.eq_order, cm.var_order
    import sympy as sp

    from .codegen import LinearizedModel

    var_names, shock_names = _names(variables), _names(shocks or [])
    extra = dict(cache or {})
    missing = [k_ for k_ in ("free_params", "steady_state") if k_ not in extra]
    if missing:
        raise ValueError(f"linearize_model: cache must carry {missing} (parameters and the analytic steady state of the model)")
    spec = {
        "name": extra.get("name", "model"), "variables": var_names, "shocks": shock_names,
        "equations": [e if isinstance(e, str) else sp.sstr(e) for e in equations], "free_params": dict(extra["free_params"]),
        "deterministic_params": dict(extra.get("deterministic_params", {})), "steady_state": dict(extra["steady_state"]),
        "assumptions": dict(extra.get("assumptions", {})), "linear": bool(extra.get("linear", False)),
    }  # fmt: skip
    not_ll = tuple(not_loglin_variables) if loglin_variables is None else tuple(v for v in var_names if v not in set(_names(loglin_variables)))
    lin = LinearizedModel(spec, log_linearize=log_linearize, not_loglin_variables=not_ll)
    for given, mine, what in ((eq_order, lin.eq_order, "eq_order"), (var_order, lin.var_order, "var_order")):
        if given is not None and not np.array_equal(np.asarray(given), mine):
            raise ValueError(f"linearize_model: the supplied {what} is not the [static | lag | mixed | lead] order of this model")
    jac = [sp.Matrix(lin.entries[m]) if lin.entries[m] and lin.entries[m][0] else sp.zeros(lin.n, 0) for m in "ABCD"]
    return jac, [lin.ss_sym[v] for v in lin.vars_perm], lin.eq_order, lin.var_order


def check_bk_condition_pt(A, B, C, D, lead_var_idx):
    """``(bk_ok, n_forward, n_unstable)`` on the regularised pencil of the estimation graph (perturbation.py:586-625).
    Numeric arrays (numpy / torch CUDA), optionally with a leading draw axis; ``lead_var_idx`` are the (permuted)
    column positions of the structural lead variables (statespace.py:224-233, 769)."""
    lead = np.asarray(lead_var_idx, dtype=np.int32)
    nu, st = batched.bk_count(A, B, C, lead)
    ok = (st & (L.ST_BK | L.ST_BK_INCONCLUSIVE)) == 0
    return ok, int(lead.size), nu


def _bk_pencil(A, B, C, lead):
    """``(G, Gamma1_sel)`` with ``G = -Gamma0_sel + 1e-8 I`` of the Sims pencil (perturbation.py:480-505, gensys.py:568-614), batched and
    assembled with ``pytensorf.block`` (on the device for torch CUDA inputs).  The reference's matrix is ``M = G^-1 Gamma1_sel``."""
    from ..pytensorf.block import block

    lead = np.asarray(lead, dtype=np.int64)
    n = A.shape[-1]
    I, O = np.eye(n), np.zeros((n, n))
    keep = np.concatenate([np.arange(n), n + lead])
    g0 = block([[B, C], [-I, O]])
    g1 = block([[A, O], [O, I]])
    if hasattr(g0, "index_select"):  # torch
        import torch

        kk = torch.as_tensor(keep, device=g0.device)
        g0 = g0.index_select(-2, kk).index_select(-1, kk).contiguous()
        g1 = g1.index_select(-2, kk).index_select(-1, kk).contiguous()
        G = -g0 + _FLOAT_ZERO_TOL * torch.eye(keep.size, dtype=g0.dtype, device=g0.device)
    else:
        g0 = np.ascontiguousarray(g0[..., keep, :][..., :, keep])
        g1 = np.ascontiguousarray(g1[..., keep, :][..., :, keep])
        G = -g0 + _FLOAT_ZERO_TOL * np.eye(keep.size)
    return G, g1


def _bk_matrix(A, B, C, lead):
    """``M = (-Gamma0_sel + 1e-8 I)^-1 Gamma1_sel`` by the batched pivoted solve."""
    G, g1 = _bk_pencil(A, B, C, lead)
    M, _st = batched.solve(G, g1)
    return M


def compute_bk_eigenvalues_pt(A, B, C, _D, lead_var_idx):
    """``(eigvals_real, eigvals_imag)`` of the regularised pencil matrix ``M = G^-1 Gamma1``, sorted by modulus
    (perturbation.py:448-505): what ``check_bk_condition_pt`` counts.  Numeric arrays (numpy / torch CUDA), optionally with a
    leading draw axis.

    ``M`` itself has entries of order 1e8 (the regularised infinite eigenvalues), so any backward-stable eigen-solver applied to
    it -- ``numpy.linalg.eig`` included -- returns the eigenvalues near the unit circle only to ``eps * 1e8 * cond`` ~ 1e-4.  The
    eigenvalue kernel is therefore run on the Cayley transform ``N = (Gamma1 - G)^-1 (Gamma1 + G) = (M - I)^-1 (M + I)``, whose
    entries are O(1) unless an eigenvalue sits at +1: ``mu = (lambda + 1) / (lambda - 1)`` maps back by the same formula,
    ``lambda = (mu + 1) / (mu - 1)``.  The finite eigenvalues then agree with the QZ eigenvalues of the pencil to ~1e-10; the
    regularised infinite ones (mu within 1e-8 of 1) come back as 1e7...inf instead of exactly ~1e8 -- far outside the unit circle
    either way.  A draw whose ``Gamma1 - G`` is singular (an eigenvalue exactly 1) falls back to the kernel on ``M``."""
    G, g1 = _bk_pencil(A, B, C, lead_var_idx)
    squeeze = G.ndim == 2
    if squeeze:
        G, g1 = G[None], g1[None]
    Ncay, st_solve = batched.solve(g1 - G, g1 + G)
    mu_re, mu_im, st_eig = batched.real_eig(Ncay, sort=False)
    is_torch = hasattr(mu_re, "device") and not isinstance(mu_re, np.ndarray)
    if is_torch:
        import torch as xp
    else:
        xp = np
    # lambda = (mu + 1) / (mu - 1) = ((a + 1)(a - 1) + b^2 - 2 b i) / ((a - 1)^2 + b^2),  mu = a + b i
    den = (mu_re - 1.0) ** 2 + mu_im**2
    with np.errstate(divide="ignore", invalid="ignore"):
        re = ((mu_re + 1.0) * (mu_re - 1.0) + mu_im**2) / den
        im = -2.0 * mu_im / den
    inf = float("inf")
    zero_den = den == 0
    re = xp.where(zero_den, xp.full_like(re, inf), re)
    im = xp.where(zero_den, xp.zeros_like(im), im)
    bad = (st_solve != 0) | (st_eig != 0)
    if bool(bad.any()):  # an eigenvalue at +1 (or a failed QR sweep): the direct route for those draws
        Mb, _ = batched.solve(G[bad], g1[bad])
        rb, ib, _ = batched.real_eig(Mb, sort=False)
        re[bad], im[bad] = rb, ib
    if is_torch:
        idx = xp.argsort(xp.hypot(re, im), dim=1, stable=True)
        re, im = xp.gather(re, 1, idx), xp.gather(im, 1, idx)
    else:
        idx = np.argsort(np.hypot(re, im), axis=1, kind="stable")
        re, im = np.take_along_axis(re, idx, 1), np.take_along_axis(im, idx, 1)
    if squeeze:
        return re[0], im[0]
    return re, im


def compute_bk_eigenvalues(A, B, C, D, tol: float = 1e-8):
    """``(eigvals_real, eigvals_imag, n_forward)`` with the reference's signature (perturbation.py:412-445).

    The reference takes the generalized eigenvalues ``beta / (alpha + tol)`` of an ordered QZ decomposition of the pencil;
    here they are the eigenvalues of ``(-Gamma0 + tol I)^-1 Gamma1``, the regularisation the reference's own estimation
    graph uses (perturbation.py:499-505): finite eigenvalues agree to O(tol), infinite ones come out as O(1 / tol) instead
    of ``beta / tol`` -- both far outside the unit circle, so the count is the same."""
    A, B, C = (np.ascontiguousarray(x, dtype=np.float64) for x in (A, B, C))
    lead = np.flatnonzero(np.abs(C).sum(axis=-2).reshape(-1, C.shape[-1]).max(axis=0) > tol)
    re, im = compute_bk_eigenvalues_pt(A, B, C, D, lead)
    return re, im, int(lead.size)


def check_bk_condition(A, B, C, D, tol: float = 1e-8, verbose: bool = True, on_failure: str = "ignore", return_value="dataframe"):
    """Blanchard-Kahn check with the reference's signature and return values (perturbation.py:508-583):
    ``'dataframe'`` -> columns ``Modulus``, ``Real``, ``Imaginary`` (one row per eigenvalue, ascending modulus),
    ``'bool'`` -> satisfied, ``None`` -> nothing.  The decision itself comes from the eigenvalue-free count kernel
    (``gecon_bk_count_*``); the eigenvalue kernel fills the table, and the two counts are cross-checked."""
    if return_value not in ["dataframe", "bool", None]:
        raise ValueError(f'Unknown return type "{return_value}"')
    A, B, C = (np.ascontiguousarray(x, dtype=np.float64) for x in (A, B, C))
    lead = np.flatnonzero(np.abs(C).sum(axis=0) > tol).astype(np.int32)
    nu, st = batched.bk_count(A, B, C, lead)
    n_forward, n_unstable = int(lead.size), int(nu)
    satisfied = (int(st) & (L.ST_BK | L.ST_BK_INCONCLUSIVE)) == 0
    re = im = None
    if return_value == "dataframe" or int(st) & L.ST_BK_INCONCLUSIVE:
        re, im = compute_bk_eigenvalues_pt(A, B, C, D, lead)
        modulus = np.hypot(re, im)
        if int(st) & L.ST_BK_INCONCLUSIVE and np.isfinite(modulus).all():  # an eigenvalue on the unit circle: count them as the reference does
            n_unstable = int((modulus > 1).sum())
            satisfied = n_unstable == n_forward
    message = (
        f"Model solution has {n_unstable} eigenvalues greater than one in modulus and {n_forward} forward-looking variables."
        f"\nBlanchard-Kahn condition is{'' if satisfied else ' NOT'} satisfied."
    )
    if not satisfied:
        if n_unstable > n_forward:
            message += " No stable solution (more unstable eigenvalues than forward-looking variables)."
        else:
            message += " No unique solution (more forward-looking variables than unstable eigenvalues)."
    if not satisfied and on_failure == "raise":
        raise ValueError(message)
    if verbose:
        _log.info(message)
    if return_value is None:
        return None
    if return_value == "bool":
        return bool(satisfied)
    import pandas as pd

    return pd.DataFrame({"Modulus": np.hypot(re, im), "Real": re, "Imaginary": im})
