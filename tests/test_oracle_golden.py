"""CPU: the C oracle (oracle/kreg_oracle.c) against the golden vectors produced by the reference's
own compiled library (tools/make_golden.py), and against that library itself when oracle/_ref is
present.  This is what pins the oracle (task rule 3)."""
import numpy as np
import pytest


def test_index_map_and_descriptors_bit_exact(golden, oracle):
    nat = golden["natoms"]
    assert np.array_equal(oracle.ac2d(nat), golden["ac2d"])
    xeq = oracle.xeq(golden["eq"])
    assert np.array_equal(xeq, golden["xeq"])
    assert np.array_equal(oracle.descriptors(golden["xyz_train"], xeq), golden["X_train"])
    assert np.array_equal(oracle.descriptors(golden["xyz_test"], xeq), golden["X_test"])


def test_kernel_matrix_literal_bit_exact(golden, oracle):
    k = oracle.kernel_matrix(golden["xyz_train"], golden["X_train"], golden["sigma"], golden["ntrval"], golden["ntrgr"])
    assert np.array_equal(k, golden["K"])
    assert np.array_equal(k, k.T)
    if golden["ntrval"]:
        assert np.all(np.diag(k)[: golden["ntrval"]] == 1.0)


def test_kernel_matrix_factored(golden, oracle):
    k = oracle.kernel_matrix(golden["xyz_train"], golden["X_train"], golden["sigma"], golden["ntrval"], golden["ntrgr"], factored=True)
    scale = np.abs(golden["K"]).max()
    assert np.abs(k - golden["K"]).max() <= 1e-14 * scale


def test_solve_residual(golden, oracle):
    g = golden
    alpha = oracle.solve(g["K"], g["ytrain"], g["ntrval"], g["ntrgr"], g["lambdav"], g["lambdagradxyz"])
    a = g["K"].copy()
    a[np.diag_indices_from(a)] += np.r_[np.full(g["ntrval"], g["lambdav"]), np.full(g["ntrgr"], g["lambdagradxyz"])]
    res = np.linalg.norm(a @ alpha - g["ytrain"]) / np.linalg.norm(g["ytrain"])
    assert res < 1e-9
    # same LAPACK (OpenBLAS) as the fixture generator: alpha reproduces to rounding
    assert np.abs(alpha - g["alpha"]).max() <= 1e-9 * np.abs(g["alpha"]).max()


def test_predict_literal_bit_exact(golden, oracle):
    g = golden
    e, gr = oracle.predict_kregf90(g["xyz_train"], g["X_train"], g["alpha"], g["sigma"], g["ntrval"], g["ntrgr"], g["xyz_test"], g["X_test"])
    assert np.array_equal(e, g["E_test"])
    assert np.array_equal(gr, g["G_test"])


def test_hessian_literal_bit_exact(golden, oracle):
    """YhessXYZest_KRR (KREG.f90:393-420): values part of alpha only; zero for gradient-only models."""
    g = golden
    h = oracle.hessian_kregf90(g["X_train"], g["alpha"], g["sigma"], g["ntrval"], g["xyz_test"], g["X_test"])
    assert h.shape == g["H_test"].shape == (g["ntest"], 3 * g["natoms"], 3 * g["natoms"])
    assert np.array_equal(h, g["H_test"])
    if g["ntrval"] == 0:
        assert not h.any()


def test_predict_mlatomf_and_factored(golden, oracle):
    g = golden
    nv = g["ntrval"]
    e, gr = oracle.predict_mlatomf(g["xyz_train"], g["X_train"], g["alpha"], g["sigma"], g["prior"], nv, g["ntrgr"], g["xyz_test"], g["X_test"])
    av = g["alpha"][:nv] if nv else None
    v = oracle.precompute_v(g["xyz_train"], g["X_train"], g["alpha"][nv:]) if g["ntrgr"] else None
    ef, gf, sc = oracle.predict_factored(g["X_train"], av, v, g["sigma"], g["prior"], g["xyz_test"], g["X_test"], want_scale=True)
    ref_e = g["prior"] + g["E_test"]
    for ee, gg in ((e, gr), (ef, gf)):
        assert np.all(np.abs(ee - ref_e) <= 1e-13 * np.maximum(sc, np.abs(ref_e)))
        assert np.abs(gg - g["G_test"]).max() <= 1e-11 * max(1.0, np.abs(g["G_test"]).max())


def test_oracle_against_live_reference(oracle):
    from oracle import ref_kreg

    if not ref_kreg.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from mlatom_b200 import synthetic as S

    nat = 7
    eq = S.equilibrium(nat)
    xtr, xte = S.geometries(eq, 21, 11), S.geometries(eq, 9, 12)
    xeq = ref_kreg.get_xeq(eq)
    assert np.array_equal(xeq, oracle.xeq(eq))
    X, Xte = ref_kreg.calc_re_descriptors(xtr, xeq), ref_kreg.calc_re_descriptors(xte, xeq)
    assert np.array_equal(X, oracle.descriptors(xtr, xeq))
    for nv, ng in ((21, 0), (21, 21 * 21), (0, 21 * 21)):
        k = ref_kreg.calc_kernel_matrix(xtr, X, 1.4, nv, ng)
        assert np.array_equal(k, oracle.kernel_matrix(xtr, X, 1.4, nv, ng))
        alpha = np.random.default_rng(5).standard_normal(nv + ng)
        e, g = ref_kreg.predict(xtr, X, xte, Xte, alpha, 1.4, nv, ng)
        eo, go = oracle.predict_kregf90(xtr, X, alpha, 1.4, nv, ng, xte, Xte)
        assert np.array_equal(e, eo) and np.array_equal(g, go)


def test_other_kernel_functions_closed_forms(oracle):
    """The oracle's restatement of MLatomF's Kfunk (A_KRR_kernel.f90:806-851) against the closed forms of the Matern
    family (n = 0: exp(-x); 1: (1 + x) exp(-x); 2: (1 + x + x^2/3) exp(-x); 3: (1 + x + 2x^2/5 + x^3/15) exp(-x), x = r/sigma)
    and the definitions of the exponential and Laplacian kernels."""
    rng = np.random.default_rng(3)
    a, b, sigma = rng.random((6, 11)), rng.random((5, 11)), 0.8
    r = np.linalg.norm(a[:, None] - b[None], axis=-1)
    x = r / sigma
    forms = {0: np.exp(-x), 1: (1 + x) * np.exp(-x), 2: (1 + x + x * x / 3) * np.exp(-x),
             3: (1 + x + 2 * x * x / 5 + x ** 3 / 15) * np.exp(-x)}
    for nn, want in forms.items():
        assert np.abs(oracle.kernel_block_generic("matern", a, b, sigma, nn=nn) - want).max() <= 1e-15
    assert np.abs(oracle.kernel_block_generic("exponential", a, b, sigma) - np.exp(-x)).max() <= 1e-15
    l1 = np.abs(a[:, None] - b[None]).sum(-1)
    assert np.abs(oracle.kernel_block_generic("laplacian", a, b, sigma) - np.exp(-l1 / sigma)).max() <= 1e-15
    assert np.abs(oracle.kernel_block_generic("gaussian", a, b, sigma) - np.exp(-r * r / (2 * sigma ** 2))).max() <= 1e-15
