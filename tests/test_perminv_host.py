"""Permutationally invariant kernel, CPU side: the host logic of mlatom_b200/perminv.py (option parsing, atom maps,
descriptor-index tables) against the oracle's literal permute-then-describe restatement (molDescr.f90:128-160, 197-220),
and the oracle's own consistency (the restated first-order derivative branch, A_KRR_kernel.f90:414-503, against finite
differences of the restated energy; invariance of the kernel on a symmetric molecule).  No GPU, no compute calls into the
CUDA library.  Parity UNPINNED for this path (no MLatomF binary in the reference checkout)."""
import numpy as np
import pytest

from mlatom_b200 import perminv
from mlatom_b200 import synthetic as S


def test_option_string_and_atom_maps():
    groups = perminv.parse_perm_inv_nuclei("4-5.6-7-8")
    assert groups == [[3, 4], [5, 6, 7]]
    pmap = perminv.atom_maps(9, groups)
    assert pmap.shape == (12, 9) and np.array_equal(pmap[0], np.arange(9))
    assert len({tuple(r) for r in pmap.tolist()}) == 12  # all distinct
    for r in pmap:  # atoms outside the groups never move; members stay inside their group
        assert np.array_equal(r[:3], [0, 1, 2]) and r[8] == 8
        assert sorted(r[3:5]) == [3, 4] and sorted(r[5:8]) == [5, 6, 7]
    with pytest.raises(ValueError):
        perminv.atom_maps(9, [[1, 2], [2, 3]])
    with pytest.raises(ValueError):
        perminv.atom_maps(4, [[3, 4]])


@pytest.mark.parametrize("nat,spec", [(9, "4-5.6-7-8"), (5, "2-3-4-5"), (21, "20-21"), (3, "1-2-3")])
def test_descriptor_index_tables_reproduce_the_permuted_descriptor(oracle, nat, spec):
    """Block P of the reference's permuted descriptor (atoms AND equilibrium distances permuted, then described) is the
    plain descriptor with its entries taken through pidx[P] -- bit for bit; pinv undoes it."""
    eq = S.equilibrium(nat)
    pmap = perminv.atom_maps(nat, perminv.parse_perm_inv_nuclei(spec))
    assert np.array_equal(pmap, oracle.perm_atom_maps(nat, perminv.parse_perm_inv_nuclei(spec)))
    pidx, pinv = perminv.descriptor_permutations(pmap)
    d = nat * (nat - 1) // 2
    x = S.geometries(eq, 6, 4, 0.1)
    xp = oracle.perm_descriptors(x, eq, pmap).reshape(6, pmap.shape[0], d)
    plain = oracle.descriptors(x, oracle.xeq(eq))
    assert np.array_equal(xp[:, 0], plain)
    for p in range(pmap.shape[0]):
        assert np.array_equal(xp[:, p], plain[:, pidx[p]])
        assert np.array_equal(plain[:, pinv[p]][:, pidx[p]], plain)


def test_oracle_gradient_branch_against_finite_differences(oracle):
    nat, sigma = 9, 1.1
    eq = S.equilibrium(nat)
    pmap = perminv.atom_maps(nat, [[3, 4], [5, 6, 7]])
    xtr, xte = S.geometries(eq, 25, 1, 0.1), S.geometries(eq, 4, 2, 0.1)
    xptr = oracle.perm_descriptors(xtr, eq, pmap)
    alpha = np.random.default_rng(0).standard_normal(25)
    e, g = oracle.perm_predict(xptr, alpha, sigma, 0.3, xte, oracle.perm_descriptors(xte, eq, pmap), pmap)
    h = 1e-5
    for a, t in [(0, 0), (3, 1), (6, 2), (8, 0)]:
        xp, xm = xte.copy(), xte.copy()
        xp[:, a, t] += h
        xm[:, a, t] -= h
        ep, _ = oracle.perm_predict(xptr, alpha, sigma, 0.3, xp, oracle.perm_descriptors(xp, eq, pmap), pmap, False)
        em, _ = oracle.perm_predict(xptr, alpha, sigma, 0.3, xm, oracle.perm_descriptors(xm, eq, pmap), pmap, False)
        assert np.abs((ep - em) / (2 * h) - g[:, a, t]).max() <= 1e-8


def test_oracle_kernel_is_invariant_on_a_symmetric_molecule(oracle):
    eq = np.array([[0, 0, 0], [1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=np.float64) * 0.63
    pmap = perminv.atom_maps(5, [[1, 2, 3, 4]])
    xa, xb = S.geometries(eq, 8, 1, 0.1), S.geometries(eq, 5, 2, 0.1)
    k = oracle.perm_kernel_block(oracle.perm_descriptors(xa, eq, pmap), oracle.perm_descriptors(xb, eq, pmap), 24, 0.9)
    ks = oracle.perm_kernel_block(oracle.perm_descriptors(xa[:, pmap[7]], eq, pmap), oracle.perm_descriptors(xb, eq, pmap),
                                  24, 0.9)
    assert np.abs(k - ks).max() <= 1e-14
    kd = oracle.perm_kernel_block(oracle.perm_descriptors(xa, eq, pmap), oracle.perm_descriptors(xa, eq, pmap), 24, 0.9)
    assert np.abs(np.diag(kd) - 1.0).max() <= 4e-16 and np.abs(kd - kd.T).max() <= 1e-15


@pytest.mark.parametrize("kind,nn,tol", [("gaussian", 0, 1e-8), ("matern", 1, 1e-8), ("matern", 3, 1e-8),
                                         ("exponential", 0, 1e-6), ("laplacian", 0, 1e-6)])
def test_oracle_generic_kernel_prediction_is_self_consistent(oracle, kind, nn, tol):
    """The restated generic prediction path (Yest_KRR + YgradXYZest_KRR's generic branch with dKdMiat_unpermuted /
    dKdXid, A_KRR_kernel.f90:731-744, 1407-1447, 1178-1233): its gradient against central differences of its own energy,
    and for the Gaussian kernel against the KREG.so-pinned MLatomF path."""
    nat, sigma = 9, 1.3
    eq = S.equilibrium(nat)
    xeq = oracle.xeq(eq)
    xtr, xte = S.geometries(eq, 30, 1, 0.1), S.geometries(eq, 4, 2, 0.1)
    xdtr, xdte = oracle.descriptors(xtr, xeq), oracle.descriptors(xte, xeq)
    alpha = np.random.default_rng(0).standard_normal(30)
    e, g = oracle.predict_generic(kind, xdtr, alpha, sigma, 0.5, xte, xdte, nn=nn)
    if kind == "gaussian":
        e1, g1 = oracle.predict_mlatomf(xtr, xdtr, alpha, sigma, 0.5, 30, 0, xte, xdte, True, True)
        assert np.abs(e - e1).max() <= 1e-14 and np.abs(g - g1).max() <= 1e-13
    h = 1e-5
    for a, t in [(0, 0), (4, 1), (8, 2)]:
        xp, xm = xte.copy(), xte.copy()
        xp[:, a, t] += h
        xm[:, a, t] -= h
        ep, _ = oracle.predict_generic(kind, xdtr, alpha, sigma, 0.5, xp, oracle.descriptors(xp, xeq), nn=nn, calc_grad=False)
        em, _ = oracle.predict_generic(kind, xdtr, alpha, sigma, 0.5, xm, oracle.descriptors(xm, xeq), nn=nn, calc_grad=False)
        assert np.abs((ep - em) / (2 * h) - g[:, a, t]).max() <= tol


def test_model_class_reads_the_mlatomf_arguments():
    """`additional_mlatomf_args` (mlatom/models.py:494-495) select the kernel function / the permutationally invariant
    kernel exactly as MLatomF's command line would (no compute: the decision logic only)."""
    from mlatom_b200 import models

    m = models.kreg(ml_program="MLatomF")
    assert m._other_kernel() is None and m._perm_inv_groups() is None
    m.additional_mlatomf_args = ["kernel=Matern"]
    assert m._other_kernel() == ("Matern", 2)  # MLatomF's default order (optionsModule.f90:105)
    m.additional_mlatomf_args = ["nn=4", "kernel=matern"]
    assert m._other_kernel() == ("matern", 4)
    m.additional_mlatomf_args = ["kernel=Gaussian"]
    assert m._other_kernel() is None
    m.additional_mlatomf_args = ["kernel=polynomial"]
    with pytest.raises(NotImplementedError):
        m._other_kernel()
    m.additional_mlatomf_args = ["permInvKernel", "permInvNuclei=4-5.6-7-8", "molDescrType=permuted"]
    assert m._perm_inv_groups() == [[3, 4], [5, 6, 7]]
    m.additional_mlatomf_args = ["permInvKernel"]
    with pytest.raises(ValueError):
        m._perm_inv_groups()
    m.additional_mlatomf_args = ["permInvKernel", "permInvGroups=1,2-3,4"]
    with pytest.raises(NotImplementedError):
        m._perm_inv_groups()
    for unsupported in (["molDescrType=permuted", "permInvNuclei=1-2"], ["molDescriptor=CM"], ["selectPerm"]):
        m.additional_mlatomf_args = unsupported  # models this package does not build must not be dropped silently
        with pytest.raises(NotImplementedError):
            m._perm_inv_groups()
    m.additional_mlatomf_args = ["molDescriptor=RE", "molDescrType=unsorted", "benchmark"]
    assert m._perm_inv_groups() is None and m._other_kernel() is None
