"""CPU oracle for the gEconpy estimation hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/scipy/sympy restatement of the reference's per-draw path

    theta -> x_ss(theta) -> A,B,C,D -> cycle reduction / gensys -> T,R -> (BK flag) -> P0 -> Kalman log-likelihood

used as the *checker* for the CUDA path in ``geconpy_b200``.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product package never does
(``tests/test_no_oracle_in_product.py`` enforces that) and the product fails loudly when its CUDA library is
missing instead of falling back to anything here.

Pinning status:

* Jacobians A,B,C,D ........ PINNED  against the reference goldens ``tests/_resources/expected_matrices.py``
* T, R (gensys, CR) ......... PINNED  against the reference's Dynare goldens ``tests/_resources/dynare_outputs/*.mat``
                                       and against the reference's own ``cycle_reduction_numpy`` executed from
                                       ``/root/reference`` (tests/golden/make_goldens.py)
* gensys ``eu`` codes ....... PINNED  (``pert_fails.gcn`` -> [1, 0, 2], tests/model/test_model.py:501-529)
* BK eigenvalue count ....... PINNED  as a count (tests/model/test_model.py:593-651)
* Kalman log-likelihood ..... PARITY UNPINNED by the reference: its filter lives in the third-party package
                                       ``pymc_extras`` (>= 0.12.0, pyproject.toml:45) which is neither under
                                       ``/root/reference`` nor installed here, and no reference test holds a numeric
                                       log-likelihood.  ``oracle.statespace.kalman_loglik`` restates the published
                                       ``StandardFilter`` algorithm and is pinned by
                                       independent known-answer tests instead (closed-form AR(1) likelihood,
                                       scipy multivariate-normal joint density, sequential-update filter).
"""
