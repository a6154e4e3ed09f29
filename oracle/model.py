"""Oracle stage 1: theta -> deterministic parameters -> analytic steady state -> A, B, C, D.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Restates, with sympy + ``lambdify``:

* ``gEconpy/model/parameters.py:11-69``   compile_param_dict_func   (free -> deterministic parameters)
* ``gEconpy/model/steady_state.py:315-357`` compile_known_ss        (analytic steady state)
* ``gEconpy/model/perturbation.py:97-198`` linearize_model          (incidence -> eq_order / var_order, Jacobian
                                                                     entries, log-linear column scaling)
* ``gEconpy/model/compile.py:196-202`` + ``gEconpy/utilities.py:42-48``  (entry = d eq / d x evaluated with every
                                                                     time-indexed variable at its steady state
                                                                     and every shock at zero)
* ``gEconpy/model/perturbation.py:201-284`` make_not_loglin_flags   (the ``Model`` path's static log-lin flags)

The model comes from a JSON spec (``tests/golden/models/*.json``; format in ``tests/golden/make_models.py``).
No common-subexpression elimination is done here on purpose: the product's code generator does its own CSE
and this independent evaluation is what it is checked against.
"""

from __future__ import annotations

import json

from pathlib import Path

import numpy as np
import sympy as sp

_FLOAT_ZERO_TOL = 1e-8  # gEconpy/model/perturbation.py:26

_T_SUFFIX = {-1: "__tm1", 0: "__t", 1: "__tp1", "ss": "__ss"}


def load_spec(path_or_name) -> dict:
    p = Path(path_or_name)
    if not p.exists():
        p = Path(__file__).resolve().parent.parent / "tests" / "golden" / "models" / f"{path_or_name}.json"
    return json.loads(p.read_text())


class OracleModel:
    """Numeric evaluation of x_ss(theta) and the permuted Jacobians for one model spec."""

    def __init__(self, spec: dict | str):
        if not isinstance(spec, dict):
            spec = load_spec(spec)
        self.spec = spec
        self.name = spec["name"]
        self.var_names = list(spec["variables"])
        self.shock_names = list(spec["shocks"])
        self.param_names = list(spec["free_params"])
        self.n = len(self.var_names)
        self.k = len(self.shock_names)
        self.defaults = dict(spec["free_params"])

        names = {}
        for v in self.var_names:
            for suf in _T_SUFFIX.values():
                names[v + suf] = sp.Symbol(v + suf)
        for s in self.shock_names:
            for suf in _T_SUFFIX.values():
                names[s + suf] = sp.Symbol(s + suf)
        for p in list(spec["free_params"]) + list(spec["deterministic_params"]):
            names[p] = sp.Symbol(p)
        self._ns = names

        def parse(s):
            return sp.sympify(s, locals=names)

        self.free_syms = [names[p] for p in self.param_names]
        self.det_exprs = {names[k]: parse(v) for k, v in spec["deterministic_params"].items()}
        self.equations = [parse(e) for e in spec["equations"]]
        self.ss_syms = [names[v + "__ss"] for v in self.var_names]
        ss = spec["steady_state"]
        self.analytic_ss = all(ss[v] is not None for v in self.var_names)
        self.ss_exprs = [parse(ss[v]) if ss[v] is not None else None for v in self.var_names]

        # ---- structural incidence and the [S|L|E|B] x [s|p|m|f] permutations (perturbation.py:112-158)
        lag = [names[v + "__tm1"] for v in self.var_names]
        now = [names[v + "__t"] for v in self.var_names]
        lead = [names[v + "__tp1"] for v in self.var_names]
        shocks_t = [names[s + "__t"] for s in self.shock_names]
        n_eq = len(self.equations)
        eq_has_lag = np.zeros(n_eq, bool)
        eq_has_lead = np.zeros(n_eq, bool)
        var_has_lag = np.zeros(self.n, bool)
        var_has_lead = np.zeros(self.n, bool)
        for i, eq in enumerate(self.equations):
            atoms = eq.free_symbols
            for j in range(self.n):
                if lag[j] in atoms:
                    eq_has_lag[i] = var_has_lag[j] = True
                if lead[j] in atoms:
                    eq_has_lead[i] = var_has_lead[j] = True
        self.var_has_lag, self.var_has_lead = var_has_lag, var_has_lead
        self.eq_order = np.concatenate(
            [
                np.where(~eq_has_lag & ~eq_has_lead)[0],
                np.where(eq_has_lag & ~eq_has_lead)[0],
                np.where(~eq_has_lag & eq_has_lead)[0],
                np.where(eq_has_lag & eq_has_lead)[0],
            ]
        )
        self.var_order = np.concatenate(
            [
                np.where(~var_has_lag & ~var_has_lead)[0],
                np.where(var_has_lag & ~var_has_lead)[0],
                np.where(var_has_lag & var_has_lead)[0],
                np.where(~var_has_lag & var_has_lead)[0],
            ]
        )
        self.inv_var_order = np.argsort(self.var_order)
        self.inv_eq_order = np.argsort(self.eq_order)
        # structural lead variables in ORIGINAL variable positions (statespace.py:224-233)
        self.lead_var_idx = np.flatnonzero(var_has_lead)

        # ---- Jacobian entries at the steady state, in ORIGINAL equation x variable order
        to_ss = {}
        for v in self.var_names:
            for t in (-1, 0, 1):
                to_ss[names[v + _T_SUFFIX[t]]] = names[v + "__ss"]
        shock_zero = {}
        for s in self.shock_names:
            for suf in _T_SUFFIX.values():
                shock_zero[names[s + suf]] = sp.Float(0.0)

        def entry(eq, x):
            return eq.diff(x).xreplace(to_ss).xreplace(shock_zero)

        grids = []
        for wrt in (lag, now, lead, shocks_t):
            grids.append([[entry(eq, x) for x in wrt] for eq in self.equations])
        args = self.ss_syms + self.free_syms + list(self.det_exprs)
        self._jac_fn = [sp.lambdify(args, sp.Matrix(g) if g and g[0] else sp.zeros(n_eq, 0), modules="numpy") for g in grids]
        self._det_fn = sp.lambdify(self.free_syms, list(self.det_exprs.values()), modules="numpy") if self.det_exprs else None
        if self.analytic_ss:
            self._ss_fn = sp.lambdify(self.free_syms + list(self.det_exprs), self.ss_exprs, modules="numpy")

    # ------------------------------------------------------------------ parameters / steady state
    def theta_vector(self, **updates) -> np.ndarray:
        d = dict(self.defaults)
        unknown = set(updates) - set(d)
        if unknown:
            raise KeyError(f"unknown parameters {sorted(unknown)}")
        d.update(updates)
        return np.array([d[p] for p in self.param_names], dtype=np.float64)

    def _det_values(self, theta):
        if self._det_fn is None:
            return []
        return [float(x) for x in self._det_fn(*theta)]

    def steady_state(self, theta) -> np.ndarray:
        if not self.analytic_ss:
            raise NotImplementedError("numeric steady states are outside the estimation path (build.py:658-659)")
        theta = np.asarray(theta, dtype=np.float64)
        with np.errstate(all="ignore"):
            vals = self._ss_fn(*theta, *self._det_values(theta))
        return np.array([complex(v).real if np.iscomplexobj(v) else v for v in vals], dtype=np.float64)

    # ------------------------------------------------------------------ log-linearisation column scale
    def column_scale(self, x_ss, mode="statespace", log_linearize=True, not_loglin_variables=(), loglin_negative_ss=False):
        """Scale factor per variable (ORIGINAL order).

        mode="statespace": per-draw ``switch(x_ss > 0, x_ss, 1)`` unless the sign is declared
        (perturbation.py:178-190; the estimation path, build.py:672-675).
        mode="model": static flags of ``make_not_loglin_flags`` (perturbation.py:260-284): level if listed, if
        |x_ss| < 1e-8, or if x_ss < 0 (unless ``loglin_negative_ss``); otherwise scaled as in the statespace path.
        """
        if self.spec.get("linear", False):
            log_linearize = False
        scale = np.ones(self.n)
        if not log_linearize:
            return scale
        for j, v in enumerate(self.var_names):
            if v in not_loglin_variables:
                continue
            if mode == "model":
                if abs(x_ss[j]) < _FLOAT_ZERO_TOL:
                    continue
                if x_ss[j] < 0 and not loglin_negative_ss:
                    continue
            assum = self.spec["assumptions"].get(v, {})
            if assum.get("negative", False):
                continue
            if assum.get("positive", False):
                scale[j] = x_ss[j]
            else:
                scale[j] = x_ss[j] if x_ss[j] > 0 else 1.0
        return scale

    # ------------------------------------------------------------------ Jacobians
    def jacobians(self, theta, x_ss=None, permuted=True, **loglin_kwargs):
        """A, B, C (n x n), D (n x k).  ``permuted=True``: rows in eq_order, columns of A,B,C in var_order."""
        theta = np.asarray(theta, dtype=np.float64)
        if x_ss is None:
            x_ss = self.steady_state(theta)
        det = self._det_values(theta)
        with np.errstate(all="ignore"):
            mats = [np.array(f(*x_ss, *theta, *det), dtype=np.float64).reshape(len(self.equations), -1) for f in self._jac_fn]
        scale = self.column_scale(x_ss, **loglin_kwargs)
        A, B, C, D = mats
        A, B, C = A * scale, B * scale, C * scale
        if permuted:
            eo, vo = self.eq_order, self.var_order
            A, B, C, D = A[eo][:, vo], B[eo][:, vo], C[eo][:, vo], D[eo]
        return [np.ascontiguousarray(M) for M in (A, B, C, D)]

    def unpermute_policy(self, T, R):
        """statespace.py:217-220."""
        inv = self.inv_var_order
        return T[inv][:, inv], R[inv]

    @property
    def permuted_lead_var_idx(self):
        """statespace.py:769: lead variables translated to permuted column positions."""
        return self.inv_var_order[self.lead_var_idx]
