"""Model spec -> CUDA device code for  theta -> deterministic parameters -> x_ss(theta) -> A, B, C, D.

This is the B200 replacement for the compiled pytensor function ``f(*ss, *params) -> [A, B, C, D]`` that the
reference builds in ``statespace_from_gcn`` (gEconpy/model/build.py:640-695) from

* ``compile_param_dict_func``  (gEconpy/model/parameters.py:11-69)    free -> deterministic parameters,
* ``compile_known_ss``         (gEconpy/model/steady_state.py:315-357) analytic steady state,
* ``linearize_model``          (gEconpy/model/perturbation.py:97-198)  incidence -> eq_order / var_order, entries
                                                                       d eq_i / d x_j at the steady state with
                                                                       shocks = 0, log-linear column scaling,
* ``build_symbolic_jacobians`` (gEconpy/model/compile.py:163-222)      one shared ``sp.cse`` over all four matrices;
                                                                       constant entries baked in, symbolic scattered.

The contract is the same -- same entries, same orderings, one shared CSE -- but the expression DAG is printed as a
CUDA ``__global__`` function (one thread per parameter draw) instead of being handed to a tracing compiler, and is
compiled by nvcc for sm_100a at model-build time (``geconpy_b200.build.build_model``).

The input is this repo's model-spec format (``tests/golden/models/*.json``: variables, shocks, equations with
``<name>__tm1 | __t | __tp1 | __ss`` symbols, free / deterministic parameters, analytic steady state, sign
assumptions).  Parsing GCN files and deriving first-order conditions is out of scope.
"""

from __future__ import annotations

import json
import re

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np
import sympy as sp

_TIME_SUFFIX = {-1: "__tm1", 0: "__t", 1: "__tp1"}


def load_spec(path_or_dict) -> dict:
    if isinstance(path_or_dict, dict):
        return path_or_dict
    return json.loads(Path(path_or_dict).read_text())


def _c_ident(prefix: str, name: str) -> str:
    return prefix + re.sub(r"[^0-9A-Za-z_]", "_", name)


@dataclass
class LinearizedModel:
    """Symbolic linearisation of one model spec plus everything the kernels need to know about its structure."""

    spec: dict
    log_linearize: bool = True
    not_loglin_variables: tuple = ()
    name: str = field(init=False)

    def __post_init__(self):
        spec = self.spec
        self.name = spec["name"]
        self.var_names = list(spec["variables"])
        self.shock_names = list(spec["shocks"])
        self.param_names = list(spec["free_params"])
        self.defaults = {k: float(v) for k, v in spec["free_params"].items()}
        self.det_names = list(spec.get("deterministic_params", {}))
        self.n = len(self.var_names)
        self.k = len(self.shock_names)
        self.n_theta = len(self.param_names)
        if spec.get("calibrated_params"):
            raise NotImplementedError("calibrated parameters are not supported on the estimation path (build.py:637-638)")
        if any(spec["steady_state"].get(v) is None for v in self.var_names):
            raise NotImplementedError("a complete analytic steady state is required on the estimation path (build.py:658-659)")
        unknown = set(self.not_loglin_variables) - set(self.var_names)
        if unknown:
            raise ValueError(f"unknown variables in not_loglin_variables: {sorted(unknown)}")
        if len(spec["equations"]) != self.n:
            raise ValueError(f"{self.name}: {len(spec['equations'])} equations for {self.n} variables")

        # ---- symbols: C-safe names, looked up from the spec's names while parsing
        ns = {}
        self.p_sym = {p: sp.Symbol(_c_ident("p_", p)) for p in self.param_names + self.det_names}
        self.ss_sym = {v: sp.Symbol(_c_ident("ss_", v)) for v in self.var_names}
        ns.update(self.p_sym)
        self.t_sym = {}
        for v in self.var_names:
            ns[v + "__ss"] = self.ss_sym[v]
            for t, suf in _TIME_SUFFIX.items():
                self.t_sym[(v, t)] = ns[v + suf] = sp.Symbol(_c_ident(f"x{t + 1}_", v))
        self.e_sym = {}
        for s in self.shock_names:
            for t, suf in _TIME_SUFFIX.items():
                self.e_sym[(s, t)] = ns[s + suf] = sp.Symbol(_c_ident(f"e{t + 1}_", s))
            ns[s + "__ss"] = sp.Float(0.0)

        def parse(text):
            return sp.sympify(text, locals=ns)

        self.det_exprs = [parse(spec["deterministic_params"][d]) for d in self.det_names]
        self.ss_exprs = [parse(spec["steady_state"][v]) for v in self.var_names]
        equations = [parse(e) for e in spec["equations"]]

        # ---- incidence and the [S|L|E|B] x [s|p|m|f] permutations (perturbation.py:112-158)
        n = self.n
        eq_lag = np.zeros(n, bool)
        eq_lead = np.zeros(n, bool)
        var_lag = np.zeros(n, bool)
        var_lead = np.zeros(n, bool)
        for i, eq in enumerate(equations):
            free = eq.free_symbols
            for j, v in enumerate(self.var_names):
                if self.t_sym[(v, -1)] in free:
                    eq_lag[i] = var_lag[j] = True
                if self.t_sym[(v, 1)] in free:
                    eq_lead[i] = var_lead[j] = True
        groups_e = (~eq_lag & ~eq_lead, eq_lag & ~eq_lead, ~eq_lag & eq_lead, eq_lag & eq_lead)
        groups_v = (~var_lag & ~var_lead, var_lag & ~var_lead, var_lag & var_lead, ~var_lag & var_lead)
        self.eq_order = np.concatenate([np.flatnonzero(g) for g in groups_e]).astype(np.int32)
        self.var_order = np.concatenate([np.flatnonzero(g) for g in groups_v]).astype(np.int32)
        self.inv_var_order = np.argsort(self.var_order).astype(np.int32)
        self.inv_eq_order = np.argsort(self.eq_order).astype(np.int32)
        self.var_has_lag, self.var_has_lead = var_lag, var_lead
        # structural lead variables (statespace.py:224-233), translated to permuted positions (statespace.py:769)
        self.lead_var_idx = np.flatnonzero(var_lead).astype(np.int32)
        self.permuted_lead_var_idx = self.inv_var_order[self.lead_var_idx].astype(np.int32)
        # state (lagged) variables occupy one contiguous column block in solver order
        self.state_var_idx = np.flatnonzero(var_lag).astype(np.int32)
        # ... and so do the lead variables: solver order is [static | lag only | lag and lead | lead only].  These are the
        # structural non-zero column ranges of A and C handed to the solver kernel (gecon_cr_args.lag_lo ... lead_hi).
        n_static, n_lagonly, n_mixed = (int(g_.sum()) for g_ in groups_v[:3])
        self.col_ranges = (n_static, n_static + n_lagonly + n_mixed, n_static + n_lagonly, n)

        # ---- entries in solver order: rows eq_order, columns var_order; vars -> ss, shocks -> 0 (compile.py:196-202)
        to_ss = {}
        for v in self.var_names:
            for t in _TIME_SUFFIX:
                to_ss[self.t_sym[(v, t)]] = self.ss_sym[v]
        for s in self.shock_names:
            for t in _TIME_SUFFIX:
                to_ss[self.e_sym[(s, t)]] = sp.Integer(0)
        eqs_perm = [equations[i] for i in self.eq_order]
        vars_perm = [self.var_names[j] for j in self.var_order]

        def grid(wrt):
            return [[eq.diff(x).xreplace(to_ss) for x in wrt] for eq in eqs_perm]

        self.entries = {
            "A": grid([self.t_sym[(v, -1)] for v in vars_perm]),
            "B": grid([self.t_sym[(v, 0)] for v in vars_perm]),
            "C": grid([self.t_sym[(v, 1)] for v in vars_perm]),
            "D": grid([self.e_sym[(s, 0)] for s in self.shock_names]),
        }

        # ---- log-linear column scale per variable, in solver column order (perturbation.py:178-190)
        linear = bool(spec.get("linear", False))
        self.scale_kind = []  # "one" | "ss" | "switch"
        for v in vars_perm:
            assum = spec.get("assumptions", {}).get(v, {})
            if linear or not self.log_linearize or v in self.not_loglin_variables or assum.get("negative", False):
                self.scale_kind.append("one")
            elif assum.get("positive", False):
                self.scale_kind.append("ss")
            else:
                self.scale_kind.append("switch")
        self.vars_perm = vars_perm

    # ------------------------------------------------------------------------------------------------ statistics
    def dag_stats(self) -> dict:
        flat = [e for m in "ABCD" for row in self.entries[m] for e in row]
        sym = [e for e in flat if not e.is_number]
        subs, red = sp.cse(sym, optimizations="basic")
        return {
            "symbolic_entries": len(sym),
            "constant_nonzero_entries": sum(1 for e in flat if e.is_number and e != 0),
            "cse_temporaries": len(subs),
            "ops": int(sum(sp.count_ops(r) for _, r in subs) + sum(sp.count_ops(r) for r in red)),
        }

    # ------------------------------------------------------------------------------------------------ observation equations
    def log_linearized_variables(self) -> set:
        """Names the state-space model treats as log-linearised (build.py:675, statespace.py:154)."""
        return (set(self.var_names) - set(self.not_loglin_variables)) if self.log_linearize else set()

    def parse_observation_equation(self, name: str, expr_str: str):
        """GCN-syntax observation equation -> sympy, in this model's symbols (statespace.py:390-444): ``v[]`` is a
        contemporaneous model variable, ``v[-k]`` a lag, ``v[ss]`` a steady-state value, bare names are parameters.
        Leads, unknown variables and unknown symbols raise ``ValueError`` with the reference's wording."""
        ns = dict(self.p_sym)
        seen = {}

        def ref(mm):
            v, idx = mm.group(1), mm.group(2).replace(" ", "")
            if idx == "ss":
                key = f"{v}__ss"
            else:
                t = 0 if idx == "" else int(idx)
                if t > 0:
                    raise ValueError(
                        f"Observation equation {name!r} contains a lead reference {v}[{idx}]. Only contemporaneous and lagged "
                        "model variables are allowed."
                    )
                key = f"{v}__lag{-t}"
            if v not in self.var_names:
                raise ValueError(f"Observation equation {name!r} references unknown model variable {v!r}. Known: {sorted(self.var_names)}")
            if key not in seen:
                seen[key] = self.ss_sym[v] if idx == "ss" else sp.Symbol(_c_ident("obsx_", key), real=True)
            return f" __OBSREF_{key}__ "

        text = re.sub(r"([A-Za-z_][A-Za-z_0-9]*)\[([^\]]*)\]", ref, expr_str).replace("^", "**")
        for key, sym in seen.items():
            ns[f"__OBSREF_{key}__"] = sym
        try:
            expr = sp.sympify(text, locals=ns)
        except (sp.SympifyError, SyntaxError, TypeError) as e:
            raise ValueError(f"cannot parse observation equation {name!r}: {expr_str!r}") from e
        known = set(ns.values())
        for fs in expr.free_symbols:
            if fs not in known:
                raise ValueError(
                    f"Observation equation {name!r} references unknown symbol {fs.name!r}: not a model variable, parameter, or hyperparameter."
                )
        refs = {}
        for key, sym in seen.items():
            if not key.endswith("__ss"):
                v, lag = key.rsplit("__lag", 1)
                refs[sym] = (v, -int(lag))
        return expr, refs

    def linearize_observation_equation(self, expr, refs):
        """First-order linearisation around the steady state (statespace.py:446-507): every reference v_{t+k} becomes
        v_ss exp(v~) (log-linearised) or v_ss + v~; intercept = value at v~ = 0, coefficient = d/dv~ there.  Returns
        (intercept, {(variable, lag): coefficient}) in steady-state and parameter symbols."""
        loglin = self.log_linearized_variables()
        forward, tildes = {}, {}
        for sym, (v, lag) in refs.items():
            tl = sp.Symbol(f"_tilde_{v}_{'0' if lag == 0 else f'm{-lag}'}", real=True)
            tildes[(v, lag)] = tl
            forward[sym] = self.ss_sym[v] * sp.exp(tl) if v in loglin else self.ss_sym[v] + tl
        g = expr.xreplace(forward)
        zero = {tl: sp.Integer(0) for tl in tildes.values()}
        return g.xreplace(zero), {key: sp.diff(g, tl).xreplace(zero) for key, tl in tildes.items()}

    def _ss_prelude(self):
        """Text of: free parameters -> deterministic parameters -> analytic steady state (one CSE pass)."""
        cc = lambda e: sp.ccode(e, strict=True)  # noqa: E731
        out = [f"    const double {self.p_sym[p].name} = th[{i}];" for i, p in enumerate(self.param_names)]
        for d, e in zip(self.det_names, self.det_exprs):
            out.append(f"    const double {self.p_sym[d].name} = {cc(e)};")
        ss_subs, ss_red = sp.cse(self.ss_exprs, symbols=sp.numbered_symbols("s_tmp_"), optimizations="basic")
        for sym, e in ss_subs:
            out.append(f"    const double {sym.name} = {cc(e)};")
        for v, e in zip(self.var_names, ss_red):
            out.append(f"    const double {self.ss_sym[v].name} = {cc(e)};")
        return out

    def obs_source(self, z_cells: dict, d_cells: dict, tag: str) -> str:
        """CUDA source of the per-draw observation kernel: writes the parameter-dependent cells of the design matrix
        (``z_cells``: {flat index into a draw's [p][k_states] block: sympy expr}) and of the observation intercept
        (``d_cells``: {row: expr}); every other cell of Z and d is left as the caller initialised it."""
        cc = lambda e: sp.ccode(e, strict=True)  # noqa: E731
        exprs = [sp.sympify(e) for e in list(z_cells.values()) + list(d_cells.values())]
        subs, red = sp.cse(exprs, symbols=sp.numbered_symbols("o_tmp_"), optimizations="basic") if exprs else ([], [])
        lines = self._ss_prelude()
        for sym, e in subs:
            lines.append(f"    const double {sym.name} = {cc(e)};")
        nz = len(z_cells)
        for idx, e in zip(z_cells, red[:nz]):
            lines.append(f"    Z[{int(idx)}] = {cc(e)};")
        for row, e in zip(d_cells, red[nz:]):
            lines.append(f"    d[{int(row)}] = {cc(e)};")
        # reverse mode over the same program: theta_bar += sum_cells Zb[cell] dZ[cell]/dtheta + sum_rows db[row] dd[row]/dtheta
        stmts = [(self.p_sym[d], e) for d, e in zip(self.det_names, self.det_exprs)]
        ss_subs, ss_red = sp.cse(self.ss_exprs, symbols=sp.numbered_symbols("s_tmp_"), optimizations="basic")
        stmts += list(ss_subs) + [(self.ss_sym[v], e) for v, e in zip(self.var_names, ss_red)] + list(subs)
        fwd = [ln for ln in lines if " Z[" not in ln and " d[" not in ln]
        rev = [f"    double b_{self.p_sym[p].name} = 0.0;" for p in self.param_names] + [f"    double b_{sym.name} = 0.0;" for sym, _ in stmts]
        rev.append("    double g;")

        def push(expr, seed):
            for sv in sorted(expr.free_symbols, key=lambda x: x.name):
                dv = sp.diff(expr, sv)
                if dv != 0:
                    rev.append(f"    b_{sv.name} += {seed} * ({cc(dv)});")

        for idx, e in zip(z_cells, red[:nz]):
            if sp.sympify(e).free_symbols:
                rev.append(f"    g = Zb[{int(idx)}];")
                push(sp.sympify(e), "g")
        for row, e in zip(d_cells, red[nz:]):
            if sp.sympify(e).free_symbols:
                rev.append(f"    g = db[{int(row)}];")
                push(sp.sympify(e), "g")
        for sym, e in reversed(stmts):
            push(e, f"b_{sym.name}")
        for i, p in enumerate(self.param_names):
            rev.append(f"    thb[{i}] += b_{self.p_sym[p].name};")
        ident = re.sub(r"[^0-9A-Za-z_]", "_", f"{self.name}_{tag}")
        return _OBS_TEMPLATE.format(name=ident, n_theta=self.n_theta, body="\n".join(lines), vjp_body="\n".join(fwd + rev))

    # ------------------------------------------------------------------------------------------------ reverse mode
    def vjp_body(self) -> str:
        """Device code of the vector-Jacobian product  theta_bar = sum_M <M_bar, dM/dtheta> (+ <xss_bar, dxss/dtheta>):
        source-to-source reverse mode over the SAME straight-line program the forward kernel evaluates (deterministic
        parameters -> steady state -> scales -> shared-CSE temporaries -> entries).  Every statement ``s = f(operands)``
        contributes ``b_operand += b_s * df/doperand`` in reverse order; the partial derivatives are taken symbolically
        per statement (the statements are CSE-sized, so they stay small).  The last stage of the gradient path
        (SURVEY 8f rank 3): what pytensor's autodiff does to the compiled [A, B, C, D] graph in the reference."""
        n, k = self.n, self.k
        cc = lambda e: sp.ccode(e, strict=True)  # noqa: E731
        fwd, stmts = [], []  # forward text; (symbol name, expr) in evaluation order
        for i, p in enumerate(self.param_names):
            fwd.append(f"    const double {self.p_sym[p].name} = th[{i}];")
        for d, e in zip(self.det_names, self.det_exprs):
            stmts.append((self.p_sym[d], e))
        ss_subs, ss_red = sp.cse(self.ss_exprs, symbols=sp.numbered_symbols("s_tmp_"), optimizations="basic")
        stmts += list(ss_subs)
        stmts += [(self.ss_sym[v], e) for v, e in zip(self.var_names, ss_red)]
        for sym, e in stmts:
            fwd.append(f"    const double {sym.name} = {cc(e)};")
        for j, (v, kind) in enumerate(zip(self.vars_perm, self.scale_kind)):
            ss = self.ss_sym[v].name
            rhs = {"one": "1.0", "ss": ss, "switch": f"(({ss} > 0.0) ? {ss} : 1.0)"}[kind]
            fwd.append(f"    const double sc{j} = {rhs};")
        sym_entries, where, const_entries = [], [], []
        for m in "ABCD":
            for i, row in enumerate(self.entries[m]):
                for j, e in enumerate(row):
                    if e.is_number:
                        if e != 0:
                            const_entries.append((m, i, j, float(e)))
                    else:
                        sym_entries.append(e)
                        where.append((m, i, j))
        subs, red = sp.cse(sym_entries, symbols=sp.numbered_symbols("j_tmp_"), optimizations="basic") if sym_entries else ([], [])
        for sym, e in subs:
            fwd.append(f"    const double {sym.name} = {cc(e)};")
        stmts_all = stmts + list(subs)
        rev = []
        names = [self.p_sym[p].name for p in self.param_names] + [sym.name for sym, _ in stmts_all]
        for nm in names:
            rev.append(f"    double b_{nm} = 0.0;")
        for j, kind in enumerate(self.scale_kind):
            if kind != "one":
                rev.append(f"    double b_sc{j} = 0.0;")
        rev.append("    double g;")
        width = {"A": n, "B": n, "C": n, "D": k}

        def push(target_expr, seed):
            """b_s += seed * d target / d s for every operand s"""
            for sv in sorted(target_expr.free_symbols, key=lambda x: x.name):
                dv = sp.diff(target_expr, sv)
                if dv != 0:
                    rev.append(f"    b_{sv.name} += {seed} * ({cc(dv)});")

        for (m, i, j), e in zip(where, red):
            scaled = m != "D" and self.scale_kind[j] != "one"
            rev.append(f"    g = {m}b[{i * width[m] + j}];")
            push(e, f"g * sc{j}" if scaled else "g")
            if scaled:
                rev.append(f"    b_sc{j} += g * ({cc(e)});")
        for m, i, j, val in const_entries:
            if m != "D" and self.scale_kind[j] != "one":
                rev.append(f"    b_sc{j} += {m}b[{i * width[m] + j}] * {val!r};")
        rev.append("    if (xssb) {")
        for i, v in enumerate(self.var_names):
            rev.append(f"        b_{self.ss_sym[v].name} += xssb[{i}];")
        rev.append("    }")
        for j, (v, kind) in enumerate(zip(self.vars_perm, self.scale_kind)):
            ss = self.ss_sym[v].name
            if kind == "ss":
                rev.append(f"    b_{ss} += b_sc{j};")
            elif kind == "switch":
                rev.append(f"    b_{ss} += ({ss} > 0.0) ? b_sc{j} : 0.0;")
        for sym, e in reversed(stmts_all):
            push(e, f"b_{sym.name}")
        for i, p in enumerate(self.param_names):
            rev.append(f"    thb[{i}] = b_{self.p_sym[p].name};")
        return "\n".join(fwd + rev)

    # ------------------------------------------------------------------------------------------------ code
    def cuda_source(self) -> str:
        """One translation unit: the per-draw device function, the batched kernel and the C-ABI launchers."""
        n, k = self.n, self.k
        lines = []
        emit = lines.append
        cc = lambda e: sp.ccode(e, strict=True)  # noqa: E731

        emit("    // free parameters")
        for i, p in enumerate(self.param_names):
            emit(f"    const double {self.p_sym[p].name} = th[{i}];")
        if self.det_names:
            emit("    // deterministic parameters (parameters.py:11-69)")
            for d, e in zip(self.det_names, self.det_exprs):
                emit(f"    const double {self.p_sym[d].name} = {cc(e)};")
        emit("    // analytic steady state (steady_state.py:315-357), one CSE pass")
        ss_subs, ss_red = sp.cse(self.ss_exprs, symbols=sp.numbered_symbols("s_tmp_"), optimizations="basic")
        for s, e in ss_subs:
            emit(f"    const double {s.name} = {cc(e)};")
        for v, e in zip(self.var_names, ss_red):
            emit(f"    const double {self.ss_sym[v].name} = {cc(e)};")
        emit("    if (xss) {")
        for i, v in enumerate(self.var_names):
            emit(f"        xss[{i}] = {self.ss_sym[v].name};")
        emit("    }")
        emit("    // log-linearisation column scales in solver column order (perturbation.py:178-190)")
        for j, (v, kind) in enumerate(zip(self.vars_perm, self.scale_kind)):
            ss = self.ss_sym[v].name
            rhs = {"one": "1.0", "ss": ss, "switch": f"(({ss} > 0.0) ? {ss} : 1.0)"}[kind]
            emit(f"    const double sc{j} = {rhs};")

        # one shared CSE over the symbolic entries of all four matrices (compile.py:207-212)
        sym_entries, where = [], []
        const_entries = []
        for m in "ABCD":
            for i, row in enumerate(self.entries[m]):
                for j, e in enumerate(row):
                    if e.is_number:
                        if e != 0:
                            const_entries.append((m, i, j, float(e)))
                    else:
                        sym_entries.append(e)
                        where.append((m, i, j))
        subs, red = sp.cse(sym_entries, symbols=sp.numbered_symbols("j_tmp_"), optimizations="basic") if sym_entries else ([], [])
        emit("    // Jacobian entries: shared CSE temporaries, then scatter (rows eq_order, columns var_order).  GECON_ST stores either")
        emit("    // into the dense matrices or into the compact vector of structural non-zeros (entries grouped A, B, C, D; row-major)")
        for s, e in subs:
            emit(f"    const double {s.name} = {cc(e)};")
        emit("    double v; bool fin = true;")
        width = {"A": n, "B": n, "C": n, "D": k}
        nz = self.nonzero_structure()
        slot = {key: ci for ci, key in enumerate(nz)}
        for (m, i, j), e in zip(where, red):
            scale = f" * sc{j}" if (m != "D" and self.scale_kind[j] != "one") else ""
            emit(f"    v = ({cc(e)}){scale}; fin = fin && (fabs(v) <= 1.7e308); GECON_ST({m}, {i * width[m] + j}, {slot[(m, i, j)]}, v);")
        emit("    // constant entries")
        for m, i, j, val in const_entries:
            scale = f" * sc{j}" if (m != "D" and self.scale_kind[j] != "one") else ""
            if scale:
                emit(f"    v = {val!r}{scale}; fin = fin && (fabs(v) <= 1.7e308); GECON_ST({m}, {i * width[m] + j}, {slot[(m, i, j)]}, v);")
            else:
                emit(f"    GECON_ST({m}, {i * width[m] + j}, {slot[(m, i, j)]}, {val!r});")
        body = "\n".join(lines)
        ident = re.sub(r"[^0-9A-Za-z_]", "_", self.name)
        offs = [sum(1 for key in nz if "ABCD".index(key[0]) < q) for q in range(5)]
        table = ", ".join(str((i << 16) | j) for (_m, i, j) in nz) or "0"
        return _TEMPLATE.format(name=ident, n=n, k=k, n_theta=self.n_theta, body=body, vjp_body=self.vjp_body(), nnz=len(nz),
                                cr_spec=self.cr_spec_block(offs),
                                nnz_alloc=max(1, len(nz)), table=table, offs=", ".join(map(str, offs)),
                                lag_lo=self.col_ranges[0], lag_hi=self.col_ranges[1], lead_lo=self.col_ranges[2], lead_hi=self.col_ranges[3],
                                n_lead=len(self.permuted_lead_var_idx), lead_list=", ".join(map(str, self.permuted_lead_var_idx)) or "0")

    def cr_spec_block(self, offs) -> str:
        """The per-model build of the warp-per-draw solver (csrc/cr_warp_spec.cu): number of variables, packed lag / lead column
        ranges and packed width as compile-time constants, for systems the warp kernel covers (the eligibility rule of
        ``cr_warp_eligible`` in csrc/cr_solve.cu: n <= 32, a forward-looking block, packed blocks no wider than the padded matrix).
        Empty for the other models: they keep the generic kernels of the core library."""
        n, k = len(self.var_names), len(self.shock_names)
        lag_lo, lag_hi, lead_lo, lead_hi = self.col_ranges
        hint = lag_hi > lag_lo or lead_hi > lead_lo
        o0 = (lag_lo & ~1) if hint else 0
        w0 = (lag_hi - o0) if hint else n
        o2 = (lead_lo & ~1) if hint else 0
        w2 = (lead_hi - o2) if hint else n
        c = max(1, -(-max(w0, w2, k, len(self.permuted_lead_var_idx)) // 8))
        np_ = -(-n // 8) * 8
        if np_ > 32 or offs[3] <= offs[2] or 8 * c > np_:
            return ""
        return (f"// ---- the solver compiled for this model (one warp per draw; csrc/cr_warp_spec.cu, found through -I <csrc>)\n"
                f"#define GECON_CW_SPEC_N {n}\n#define GECON_CW_SPEC_O0 {o0}\n#define GECON_CW_SPEC_W0 {w0}\n"
                f"#define GECON_CW_SPEC_O2 {o2}\n#define GECON_CW_SPEC_W2 {w2}\n#define GECON_CW_SPEC_C {c}\n"
                f'#include "cr_warp_spec.cu"\n')

    def nonzero_structure(self):
        """Structural non-zeros of A, B, C, D in solver order as a list of (matrix, row, col), grouped by matrix and row-major inside
        each: the layout of the compact Jacobian vector (``gecon_model_jacobian_compact``) the fused pipeline consumes."""
        return [(m, i, j) for m in "ABCD" for i, row in enumerate(self.entries[m]) for j, e in enumerate(row) if not (e.is_number and e == 0)]


_TEMPLATE = r"""// GENERATED by geconpy_b200/model/codegen.py for model "{name}" -- do not edit.
// theta -> deterministic parameters -> analytic steady state -> A, B, C (n x n), D (n x k) in solver order.
// One thread per parameter draw; the matrices are zero-filled by the launcher and only non-zero entries are written.
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef GECON_HOST_CHECK
// Test-only build (g++, no CUDA): the generated expression code is compiled as a host function so that the CPU test
// suite can check the code generator against the oracle without a GPU.  Never used by the product.
#define __device__
#define __forceinline__ inline
#define __restrict__
#else
#include <cuda_runtime.h>
#include "gecon_b200.h"
#endif

#define GECON_MODEL_N {n}
#define GECON_MODEL_K {k}
#define GECON_MODEL_NTHETA {n_theta}
#ifndef GECON_ST_JAC_NONFINITE
#define GECON_ST_JAC_NONFINITE 0x200
#endif

#define GECON_MODEL_NNZ {nnz}
// structural non-zeros of A, B, C, D (row << 16 | col), grouped by matrix: entries of matrix q are [off[q], off[q + 1])
static const int32_t gecon_nz_table_h[{nnz_alloc}] = {{{table}}};
static const int32_t gecon_nz_off_h[5] = {{{offs}}};
static const int32_t gecon_lead_idx_h[{n_lead} > 0 ? {n_lead} : 1] = {{{lead_list}}};

// COMPACT: store into vals[GECON_MODEL_NNZ] (A, B, C, D unused) instead of the dense, zero-filled matrices
#define GECON_ST(M, DI, CI, V) do {{ if (COMPACT) vals[CI] = (V); else M[DI] = (V); }} while (0)
template <bool COMPACT>
__device__ __forceinline__ bool gecon_model_eval_t(const double* __restrict__ th, double* __restrict__ A, double* __restrict__ B,
                                                   double* __restrict__ C, double* __restrict__ D, double* __restrict__ xss,
                                                   double* __restrict__ vals) {{
{body}
    return fin;
}}
#undef GECON_ST

__device__ __forceinline__ bool gecon_model_eval(const double* __restrict__ th, double* __restrict__ A, double* __restrict__ B,
                                                 double* __restrict__ C, double* __restrict__ D, double* __restrict__ xss) {{
    return gecon_model_eval_t<false>(th, A, B, C, D, xss, nullptr);
}}

// theta_bar = sum over the entries of <M_bar, dM/dtheta> (+ <xss_bar, dx_ss/dtheta>): reverse mode over the same program
__device__ __forceinline__ void gecon_model_vjp(const double* __restrict__ th, const double* __restrict__ Ab,
                                                const double* __restrict__ Bb, const double* __restrict__ Cb,
                                                const double* __restrict__ Db, const double* __restrict__ xssb,
                                                double* __restrict__ thb) {{
{vjp_body}
}}

#ifdef GECON_HOST_CHECK
extern "C" int gecon_model_vjp_host_check(const double* theta, int64_t N, const double* Ab, const double* Bb, const double* Cb,
                                          const double* Db, const double* xssb, double* theta_bar) {{
    const size_t nn = (size_t)GECON_MODEL_N * GECON_MODEL_N;
    for (int64_t i = 0; i < N; ++i)
        gecon_model_vjp(theta + (size_t)i * GECON_MODEL_NTHETA, Ab + i * nn, Bb + i * nn, Cb + i * nn,
                        Db + (size_t)i * GECON_MODEL_N * GECON_MODEL_K, xssb ? xssb + (size_t)i * GECON_MODEL_N : nullptr,
                        theta_bar + (size_t)i * GECON_MODEL_NTHETA);
    return 0;
}}

extern "C" int gecon_model_eval_host_check(const double* theta, int64_t N, double* A, double* B, double* C, double* D, double* xss,
                                           int32_t* status) {{
    const size_t nn = (size_t)GECON_MODEL_N * GECON_MODEL_N;
    for (int64_t i = 0; i < N; ++i) {{
        const bool fin = gecon_model_eval(theta + (size_t)i * GECON_MODEL_NTHETA, A + i * nn, B + i * nn, C + i * nn,
                                          D + (size_t)i * GECON_MODEL_N * GECON_MODEL_K, xss ? xss + (size_t)i * GECON_MODEL_N : nullptr);
        if (status) status[i] = fin ? 0 : GECON_ST_JAC_NONFINITE;
    }}
    return 0;
}}
#else
__global__ void __launch_bounds__(128) gecon_model_jacobian_kernel(const double* __restrict__ theta, long long N, double* __restrict__ A,
                                                                   double* __restrict__ B, double* __restrict__ C,
                                                                   double* __restrict__ D, double* __restrict__ xss,
                                                                   int* __restrict__ status) {{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {{
        const size_t nn = (size_t)GECON_MODEL_N * GECON_MODEL_N;
        const bool fin = gecon_model_eval(theta + (size_t)i * GECON_MODEL_NTHETA, A + i * nn, B + i * nn, C + i * nn,
                                          D + (size_t)i * GECON_MODEL_N * GECON_MODEL_K, xss ? xss + (size_t)i * GECON_MODEL_N : nullptr);
        if (status) status[i] = fin ? 0 : GECON_ST_JAC_NONFINITE;
    }}
}}

__global__ void __launch_bounds__(128) gecon_model_vjp_kernel(const double* __restrict__ theta, long long N, const double* __restrict__ Ab,
                                                              const double* __restrict__ Bb, const double* __restrict__ Cb,
                                                              const double* __restrict__ Db, const double* __restrict__ xssb,
                                                              double* __restrict__ theta_bar) {{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {{
        const size_t nn = (size_t)GECON_MODEL_N * GECON_MODEL_N;
        gecon_model_vjp(theta + (size_t)i * GECON_MODEL_NTHETA, Ab + i * nn, Bb + i * nn, Cb + i * nn,
                        Db + (size_t)i * GECON_MODEL_N * GECON_MODEL_K, xssb ? xssb + (size_t)i * GECON_MODEL_N : nullptr,
                        theta_bar + (size_t)i * GECON_MODEL_NTHETA);
    }}
}}

__global__ void __launch_bounds__(128) gecon_model_jacobian_compact_kernel(const double* __restrict__ theta, long long theta_stride, long long N,
                                                                           double* __restrict__ vals, double* __restrict__ xss,
                                                                           int* __restrict__ status) {{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {{
        const bool fin = gecon_model_eval_t<true>(theta + (size_t)i * theta_stride, nullptr, nullptr, nullptr, nullptr,
                                                  xss ? xss + (size_t)i * GECON_MODEL_N : nullptr, vals + (size_t)i * GECON_MODEL_NNZ);
        if (status) status[i] = fin ? 0 : GECON_ST_JAC_NONFINITE;
    }}
}}

// DEVICE pointers.  The compact Jacobian: vals[N][GECON_MODEL_NNZ], the structural non-zeros of A, B, C, D only (layout:
// gecon_model_structure).  theta rows may be strided (theta_stride >= n_theta doubles: the free parameters are the leading
// columns of a wider parameter vector).  No memsets, one kernel on `stream`.
extern "C" int gecon_model_jacobian_compact(const double* theta, int64_t theta_stride, int64_t N, double* vals, double* xss, int32_t* status,
                                            void* stream) {{
    if (N <= 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (N + 127) / 128;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    gecon_model_jacobian_compact_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(theta, theta_stride, N, vals, xss, status);
    return (int)cudaGetLastError();
}}

// Structure of the model as the solver kernels need it (HOST arrays owned by the library): number of structural non-zeros, their
// (row << 16 | col) table grouped by matrix with offsets off[5], the contiguous lag / lead column ranges
// {{lag_lo, lag_hi, lead_lo, lead_hi}} (solver order) and the positions of the structural lead variables.
extern "C" int gecon_model_structure(int32_t* nnz, const int32_t** table, const int32_t** off, int32_t* col_ranges, int32_t* n_lead,
                                     const int32_t** lead_idx) {{
    if (nnz) *nnz = GECON_MODEL_NNZ;
    if (table) *table = gecon_nz_table_h;
    if (off) *off = gecon_nz_off_h;
    if (col_ranges) {{
        col_ranges[0] = {lag_lo};
        col_ranges[1] = {lag_hi};
        col_ranges[2] = {lead_lo};
        col_ranges[3] = {lead_hi};
    }}
    if (n_lead) *n_lead = {n_lead};
    if (lead_idx) *lead_idx = gecon_lead_idx_h;
    return 0;
}}

{cr_spec}
// The fused theta -> log-likelihood entry point of this model (include/gecon_b200.h, gecon_pipeline_args): fills in the model's
// own kernel and structure tables and runs the core library's pipeline.
extern "C" int gecon_model_loglik(gecon_pipeline_args* a, void* stream) {{
    if (!a) return -1;
#ifdef GECON_CW_SPEC_N
    if (a->struct_size >= offsetof(gecon_pipeline_args, cr_solve) + sizeof(a->cr_solve)) a->cr_solve = gecon_model_cr_solve;
#endif
    a->jacobian = gecon_model_jacobian_compact;
    a->nz_table = gecon_nz_table_h;
    a->nz_off = gecon_nz_off_h;
    a->nnz = GECON_MODEL_NNZ;
    a->n = GECON_MODEL_N;
    a->k = GECON_MODEL_K;
    a->n_theta = GECON_MODEL_NTHETA;
    a->col_ranges[0] = {lag_lo};
    a->col_ranges[1] = {lag_hi};
    a->col_ranges[2] = {lead_lo};
    a->col_ranges[3] = {lead_hi};
    if (a->check_bk && !a->lead_idx) {{
        a->lead_idx = gecon_lead_idx_h;
        a->n_lead = {n_lead};
    }}
    return gecon_loglik_pipeline(a, stream);
}}

// DEVICE pointers: theta_bar[N][n_theta] = vector-Jacobian product of (A_bar, B_bar, C_bar, D_bar, xss_bar or NULL)
extern "C" int gecon_model_vjp_batched(const double* theta, int64_t N, const double* Ab, const double* Bb, const double* Cb,
                                       const double* Db, const double* xssb, double* theta_bar, void* stream) {{
    if (N <= 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (N + 127) / 128;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    gecon_model_vjp_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(theta, N, Ab, Bb, Cb, Db, xssb, theta_bar);
    return (int)cudaGetLastError();
}}

extern "C" int gecon_model_info(int32_t* n, int32_t* k, int32_t* n_theta) {{
    if (n) *n = GECON_MODEL_N;
    if (k) *k = GECON_MODEL_K;
    if (n_theta) *n_theta = GECON_MODEL_NTHETA;
    return 0;
}}

// DEVICE pointers; returns 0 or a cudaError_t.  Two memsets + one kernel on `stream`.
extern "C" int gecon_model_jacobian_batched(const double* theta, int64_t N, double* A, double* B, double* C, double* D, double* xss,
                                            int32_t* status, void* stream) {{
    if (N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nn = (size_t)N * GECON_MODEL_N * GECON_MODEL_N * sizeof(double);
    cudaError_t e;
    if ((e = cudaMemsetAsync(A, 0, nn, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMemsetAsync(B, 0, nn, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMemsetAsync(C, 0, nn, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMemsetAsync(D, 0, (size_t)N * GECON_MODEL_N * GECON_MODEL_K * sizeof(double), st)) != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (N + 127) / 128;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    gecon_model_jacobian_kernel<<<(int)blocks, 128, 0, st>>>(theta, N, A, B, C, D, xss, status);
    return (int)cudaGetLastError();
}}

// HOST pointers: copies theta in, runs, copies A, B, C, D (and xss, status if non-null) out, synchronises.
extern "C" int gecon_model_jacobian_host(const double* theta, int64_t N, double* A, double* B, double* C, double* D, double* xss,
                                         int32_t* status) {{
    if (N <= 0) return 0;
    const size_t nn = (size_t)N * GECON_MODEL_N * GECON_MODEL_N * sizeof(double);
    const size_t nd = (size_t)N * GECON_MODEL_N * GECON_MODEL_K * sizeof(double);
    const size_t nt = (size_t)N * GECON_MODEL_NTHETA * sizeof(double);
    const size_t nx = (size_t)N * GECON_MODEL_N * sizeof(double);
    double *dth = nullptr, *dA = nullptr, *dB = nullptr, *dC = nullptr, *dD = nullptr, *dx = nullptr;
    int32_t* ds = nullptr;
    cudaError_t e = cudaSuccess;
#define TRY(x) if (e == cudaSuccess) e = (x)
    TRY(cudaMalloc(&dth, nt));
    TRY(cudaMalloc(&dA, nn));
    TRY(cudaMalloc(&dB, nn));
    TRY(cudaMalloc(&dC, nn));
    TRY(cudaMalloc(&dD, nd ? nd : 8));
    TRY(cudaMalloc(&dx, nx));
    TRY(cudaMalloc(&ds, (size_t)N * sizeof(int32_t)));
    TRY(cudaMemcpy(dth, theta, nt, cudaMemcpyHostToDevice));
    if (e == cudaSuccess) e = (cudaError_t)gecon_model_jacobian_batched(dth, N, dA, dB, dC, dD, dx, ds, nullptr);
    TRY(cudaMemcpy(A, dA, nn, cudaMemcpyDeviceToHost));
    TRY(cudaMemcpy(B, dB, nn, cudaMemcpyDeviceToHost));
    TRY(cudaMemcpy(C, dC, nn, cudaMemcpyDeviceToHost));
    if (nd) TRY(cudaMemcpy(D, dD, nd, cudaMemcpyDeviceToHost));
    if (xss) TRY(cudaMemcpy(xss, dx, nx, cudaMemcpyDeviceToHost));
    if (status) TRY(cudaMemcpy(status, ds, (size_t)N * sizeof(int32_t), cudaMemcpyDeviceToHost));
#undef TRY
    cudaFree(dth); cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dD); cudaFree(dx); cudaFree(ds);
    return (int)e;
}}
#endif  // GECON_HOST_CHECK
"""


_OBS_TEMPLATE = r"""// GENERATED by geconpy_b200/model/codegen.py (observation equations of "{name}") -- do not edit.
// theta -> steady state -> parameter-dependent cells of the design matrix Z and of the observation intercept d
// (gEconpy/model/statespace.py:299-331, 363-388, 446-507).  One thread per draw.
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef GECON_HOST_CHECK
#define __device__
#define __forceinline__ inline
#define __restrict__
#else
#include <cuda_runtime.h>
#endif
#define GECON_MODEL_NTHETA {n_theta}

__device__ __forceinline__ void gecon_obs_eval(const double* __restrict__ th, double* __restrict__ Z, double* __restrict__ d) {{
{body}
}}

// theta_bar += <(Z_bar, d_bar), d(Z, d)/dtheta>: reverse mode over the same program (the adjoint of the observation equations)
__device__ __forceinline__ void gecon_obs_vjp(const double* __restrict__ th, const double* __restrict__ Zb, const double* __restrict__ db,
                                              double* __restrict__ thb) {{
{vjp_body}
}}

#ifdef GECON_HOST_CHECK
extern "C" int gecon_obs_vjp_host_check(const double* theta, int64_t N, const double* Zb, int64_t z_stride, const double* db, int64_t d_stride,
                                        double* theta_bar) {{
    for (int64_t i = 0; i < N; ++i)
        gecon_obs_vjp(theta + (size_t)i * GECON_MODEL_NTHETA, Zb + (size_t)i * z_stride, db + (size_t)i * d_stride,
                      theta_bar + (size_t)i * GECON_MODEL_NTHETA);
    return 0;
}}

extern "C" int gecon_obs_host_check(const double* theta, int64_t N, double* Z, int64_t z_stride, double* d, int64_t d_stride) {{
    for (int64_t i = 0; i < N; ++i) gecon_obs_eval(theta + (size_t)i * GECON_MODEL_NTHETA, Z + (size_t)i * z_stride, d + (size_t)i * d_stride);
    return 0;
}}
#else
__global__ void __launch_bounds__(128) gecon_obs_kernel(const double* __restrict__ theta, long long N, double* __restrict__ Z,
                                                        long long z_stride, double* __restrict__ d, long long d_stride) {{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        gecon_obs_eval(theta + (size_t)i * GECON_MODEL_NTHETA, Z + (size_t)i * z_stride, d + (size_t)i * d_stride);
}}

__global__ void __launch_bounds__(128) gecon_obs_vjp_kernel(const double* __restrict__ theta, long long N, const double* __restrict__ Zb,
                                                            long long z_stride, const double* __restrict__ db, long long d_stride,
                                                            double* __restrict__ theta_bar) {{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        gecon_obs_vjp(theta + (size_t)i * GECON_MODEL_NTHETA, Zb + (size_t)i * z_stride, db + (size_t)i * d_stride,
                      theta_bar + (size_t)i * GECON_MODEL_NTHETA);
}}

// DEVICE pointers: theta_bar[N][n_theta] += vector-Jacobian product of (Z_bar [N][p][k_states], d_bar [N][p])
extern "C" int gecon_obs_vjp_batched(const double* theta, int64_t N, const double* Zb, int64_t z_stride, const double* db, int64_t d_stride,
                                     double* theta_bar, void* stream) {{
    if (N <= 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (N + 127) / 128;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    gecon_obs_vjp_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(theta, N, Zb, z_stride, db, d_stride, theta_bar);
    return (int)cudaGetLastError();
}}

// DEVICE pointers: Z is [N][p][k_states] (z_stride doubles per draw), d is [N][p] (d_stride); returns 0 or a cudaError_t
extern "C" int gecon_obs_batched(const double* theta, int64_t N, double* Z, int64_t z_stride, double* d, int64_t d_stride, void* stream) {{
    if (N <= 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (N + 127) / 128;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    gecon_obs_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(theta, N, Z, z_stride, d, d_stride);
    return (int)cudaGetLastError();
}}
#endif
"""
