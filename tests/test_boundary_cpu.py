"""Host-side boundary of SURVEY 8(b) that needs no GPU: the Op contracts (through the pytensor shim), ``pytensorf.block``, the
``linearize_model`` argument list, ``configure`` option handling, the structured ``gensys`` entry point's argument checks."""

from __future__ import annotations

import json

import numpy as np
import pytest

import pt_shim

from helpers import model


@pytest.fixture
def ops(monkeypatch):
    mods = pt_shim.install(monkeypatch)
    yield mods
    monkeypatch.undo()
    pt_shim.uninstall()


def test_op_contracts_match_the_reference(ops):
    """``__props__``, gufunc signatures, output dtypes (``linalg_output_dtype``) and shapes of every solver Op
    (cycle_reduction.py:186-213, gensys.py:634-676, real_eig.py:10-36)."""
    cr, gs, re_ = ops["cycle_reduction"], ops["gensys"], ops["real_eig"]
    op = cr.CycleReductionWrapper(max_iter=77, tol=1e-6)
    assert op.__props__ == ("max_iter", "tol") and op.gufunc_signature == "(n,n),(n,n),(n,n)->(n,n)"
    assert (op.max_iter, op.tol) == (77, 1e-6)
    A64, A32 = np.zeros((5, 5)), np.zeros((5, 5), dtype=np.float32)
    node = op.make_node(A64, A64, A32)
    assert node.outputs[0].type.dtype == "float64" and node.outputs[0].type.shape == (5, 5) and len(node.inputs) == 3
    assert op.make_node(A32, A32, A32).outputs[0].type.dtype == "float32"
    assert op.infer_shape(None, node, [(5, 5)] * 3) == [(5, 5)]
    g = gs.GensysWrapper(tol=1e-7)
    assert g.__props__ == ("tol",) and g.gufunc_signature == "(n,n),(n,n),(n,n),(n,k)->(n,n),()"
    node = g.make_node(A64, A64, A64, np.zeros((5, 2)))
    assert [o.type.dtype for o in node.outputs] == ["float64", "bool"] and node.outputs[1].type.shape == ()
    assert g.infer_shape(None, node, [(5, 5)] * 3 + [(5, 2)]) == [(5, 5), ()]
    s = cr.ScanCycleReduction(max_iter=50, tol=1e-7)
    node = s.make_node(A64, A64, A64)
    assert [o.type.dtype for o in node.outputs] == ["float64", "int32"] and s.gufunc_signature == "(n,n),(n,n),(n,n)->(n,n),()"
    e = re_.RealEig()
    assert e.__props__ == () and e.gufunc_signature == "(m,m)->(m),(m)"
    node = e.make_node(np.zeros((7, 7)))
    assert [o.type.shape for o in node.outputs] == [(7,), (7,)]
    with pytest.raises(ValueError, match="2-d matrix"):
        e.make_node(np.zeros((2, 7, 7)))
    pa = cr.PolicyAdjoint()
    assert pa.gufunc_signature == "(n,n),(n,n),(n,n),(n,n),(n,n)->(n,n),(n,n),(n,n)"
    # the pullback of every solver Op goes through the adjoint KERNEL's Op, not the n^2 x n^2 Kronecker graph
    T_bar = pt_shim.pt.as_tensor(np.zeros((5, 5)))
    outs = op.pullback([pt_shim.pt.as_tensor(A64)] * 3, [pt_shim.pt.as_tensor(A64)], [T_bar])
    assert len(outs) == 3 and all(isinstance(o.owner.op, cr.PolicyAdjoint) for o in outs)
    outs = g.pullback([pt_shim.pt.as_tensor(A64)] * 3 + [pt_shim.pt.as_tensor(np.zeros((5, 2)))], [pt_shim.pt.as_tensor(A64), None], [T_bar, None])
    assert len(outs) == 4 and outs[3].type.shape == (5, 2)


def test_symbolic_entry_points_build_nodes(ops):
    cr, gs = ops["cycle_reduction"], ops["gensys"]
    A, D = np.zeros((4, 4)), np.zeros((4, 1))
    T, n_steps = cr.ScanCycleReduction()(A, A, A)
    assert T.owner is n_steps.owner and n_steps.type.dtype == "int32"
    T, ok = gs.GensysWrapper()(A, A, A, D)
    assert T.owner.op.tol == 1e-8 and ok.type.dtype == "bool"


def test_without_pytensor_the_op_layer_says_so():
    from geconpy_b200.solvers import cycle_reduction as cr
    from geconpy_b200.solvers._pt import HAVE_PYTENSOR

    if HAVE_PYTENSOR:
        pytest.skip("pytensor is installed here")
    with pytest.raises(ImportError, match="pytensor"):
        cr.CycleReductionWrapper()
    with pytest.raises(ImportError, match="pytensor"):
        cr.scan_cycle_reduction(None, None, None, None)


def test_block_matches_numpy_block():
    """gEconpy/pytensorf/block.py:53 (the reference's own doctest + numpy.block on nested lists, batch axes broadcast)."""
    import torch

    from geconpy_b200.pytensorf.block import block

    A, B, C, D = np.array([[1, 2], [3, 4]]), np.array([[5], [6]]), np.array([[7, 8]]), np.array([[9]])
    assert np.array_equal(block([[A, B], [C, D]]), [[1, 2, 5], [3, 4, 6], [7, 8, 9]])
    rng = np.random.default_rng(0)
    X, I, O = rng.random((4, 3, 3)), np.eye(3), np.zeros((3, 3))
    G = block([[X, O], [-I, I]])
    assert G.shape == (4, 6, 6)
    for i in range(4):
        assert np.array_equal(G[i], np.block([[X[i], O], [-I, I]]))
    Gt = block([[torch.as_tensor(X), O], [-I, I]])
    assert isinstance(Gt, torch.Tensor) and np.array_equal(Gt.numpy(), G)
    assert np.array_equal(block([1, 2, 3]), [1, 2, 3]) and block(np.float64(2.0)).shape == (1,)
    with pytest.raises(ValueError, match="same nesting depth"):
        block([[A], [A, [B]]])
    with pytest.raises(ValueError, match="empty list"):
        block([[]])
    with pytest.raises(TypeError, match="tuples"):
        block(([A],))


def test_linearize_model_takes_the_reference_argument_list():
    """perturbation.py:29-38: (variables, equations, shocks, cache, loglin_variables, order, eq_order, var_order) ->
    ([A, B, C, D], ss_nodes, eq_order, var_order); entries checked against the oracle's Jacobians at the default parameters."""
    import sympy as sp

    from geconpy_b200.model.perturbation import linearize_model

    mod = model("rbc")
    spec = mod.spec
    cache = {k: spec[k] for k in ("name", "free_params", "deterministic_params", "steady_state", "assumptions", "linear")}
    jac, ss_nodes, eq_order, var_order = linearize_model(spec["variables"], spec["equations"], spec["shocks"], cache, None, 1)
    assert [j.shape for j in jac] == [(9, 9), (9, 9), (9, 9), (9, 1)] and len(ss_nodes) == 9
    assert np.array_equal(eq_order, mod.eq_order) and np.array_equal(var_order, mod.var_order)
    th = mod.theta_vector()
    xss = mod.steady_state(th)
    vals = {sp.Symbol(f"p_{p}"): v for p, v in zip(mod.param_names, th)}
    vals.update({sp.Symbol(f"ss_{v}"): x for v, x in zip(mod.var_names, xss)})
    A, B, C, D = mod.jacobians(th, mode="statespace")
    scale = mod.column_scale(xss)[mod.var_order] if hasattr(mod, "column_scale") else None
    for sym, num in zip(jac[:3], (A, B, C)):
        got = np.array(sym.xreplace(vals), dtype=np.float64)
        if scale is not None:
            got = got * scale[None, :]
        np.testing.assert_allclose(got, num, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.array(jac[3].xreplace(vals), dtype=np.float64), D, rtol=1e-10, atol=1e-12)
    with pytest.raises(NotImplementedError, match="order = 1"):
        linearize_model(spec["variables"], spec["equations"], spec["shocks"], cache, None, 2)
    with pytest.raises(ValueError, match="eq_order"):
        linearize_model(spec["variables"], spec["equations"], spec["shocks"], cache, None, 1, eq_order=np.arange(9)[::-1])
    with pytest.raises(ValueError, match="cache must carry"):
        linearize_model(spec["variables"], spec["equations"], spec["shocks"], {})


def test_configure_options_of_the_reference():
    """statespace.py:822-839: solver names, constant_params (names / "auto"), the pass-through options and the two gates."""
    from geconpy_b200 import _lib as L
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    cm = CompiledModel("rbc")
    ss = BatchedStateSpace(cm)
    ss.configure(observed_states=["Y"], solver="scan_cycle_reduction", constant_params=["alpha", "delta"], mode="FAST_RUN",
                 use_adjoint_gradients=True, use_direct_lyapunov=True, verbose=False)  # fmt: skip
    assert ss.solver == "scan_cycle_reduction" and ss.constant_params == ["alpha", "delta"]
    assert ss.param_names == ["beta", "rho_A", "sigma_C", "sigma_L", "sigma_epsilon_A"] and ss.n_param == 5
    assert ss.gate_mask & L.ST_BK and ss.gate_mask & L.ST_RESID
    ss.configure(observed_states=["Y"], constant_params="auto")  # every RBC parameter carries a prior -> nothing is frozen
    assert ss.constant_params == [] and ss.n_param == cm.n_theta + 1
    ss.configure(observed_states=["Y"], add_bk_check=False, add_solver_success_check=False)  # the reference's default graph
    assert not ss.gate_mask & (L.ST_BK | L.ST_RESID | L.ST_CR_NOT_CONVERGED) and ss.gate_mask & L.ST_JAC_NONFINITE and not ss.check_bk
    with pytest.raises(ValueError, match="unknown constant_params"):
        ss.configure(observed_states=["Y"], constant_params=["nope"])
    with pytest.raises(NotImplementedError, match="solver"):
        ss.configure(observed_states=["Y"], solver="qz")
    with pytest.raises(ValueError, match="backward_direct"):
        ss.configure(observed_states=["Y"], solver="backward_direct")


def test_gensys_general_pencils_are_declared_unsupported():
    from geconpy_b200.solvers import gensys as gs

    g0 = np.eye(3)
    with pytest.raises(NotImplementedError, match="_gensys_setup"):  # not a pencil of a linearised model
        gs.gensys(g0, g0 * 0.5, np.zeros((3, 1)), np.ones((3, 1)), np.ones((3, 1)))
    with pytest.raises(NotImplementedError):
        gs.build_u_v_d(np.eye(2))
