"""GPU parity tests: the CUDA path (through the C ABI, via mlatom_b200.engine) against the golden
vectors of the reference's compiled library and against the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): energies 1e-10 relative, gradients 1e-8; kernel-matrix
elements 1e-13 relative (SURVEY.md 8c protocol A); descriptors bit-exact.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from mlatom_b200 import synthetic as S  # noqa: E402

E_RTOL = 1e-10
G_TOL = 1e-8
H_TOL = 1e-9  # Hessian: same arithmetic as the gradient sums, one more contraction
K_RTOL = 1e-13


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mlatom_b200 import engine

    engine.device_check()
    return engine


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def lower(a):
    return np.tril(a)


# ---- golden vectors ---------------------------------------------------------------------------

def test_golden_descriptors_bit_exact(golden, eng):
    xeq = eng.xeq(dev(golden["eq"]))
    assert np.array_equal(xeq.cpu().numpy(), golden["xeq"])
    for key_xyz, key_x in (("xyz_train", "X_train"), ("xyz_test", "X_test")):
        x = eng.descriptors(dev(golden[key_xyz]), xeq)
        assert np.array_equal(x.cpu().numpy(), golden[key_x])


@pytest.mark.parametrize("fill", ["lower", "full"])
def test_golden_kernel_matrix(golden, eng, fill):
    g = golden
    xeq = eng.xeq(dev(g["eq"]))
    from mlatom_b200._lib import FILL_FULL, FILL_LOWER

    k = eng.build_kernel_matrix(dev(g["xyz_train"]), xeq, g["sigma"], 0.0, 0.0, use_values=g["ntrval"] > 0,
                                use_gradients=g["ntrgr"] > 0, fill=FILL_FULL if fill == "full" else FILL_LOWER)
    k = k.cpu().numpy()
    ref = g["K"]
    if fill == "lower":
        k, ref = lower(k), lower(ref)
    scale = np.abs(g["K"]).max()
    assert np.abs(k - ref).max() <= K_RTOL * scale
    if g["ntrval"]:
        assert np.all(np.diag(k)[: g["ntrval"]] == 1.0)


def test_golden_lambda_on_diagonal(golden, eng):
    g = golden
    xeq = eng.xeq(dev(g["eq"]))
    k = eng.build_kernel_matrix(dev(g["xyz_train"]), xeq, g["sigma"], 0.25, 0.5, use_values=g["ntrval"] > 0,
                                use_gradients=g["ntrgr"] > 0).cpu().numpy()
    d = np.diag(k) - np.diag(g["K"])
    assert np.abs(d[: g["ntrval"]] - 0.25).max(initial=0.0) <= 1e-13
    assert np.abs(d[g["ntrval"]:] - 0.5).max(initial=0.0) <= 1e-12


def test_golden_predict_from_reference_alpha(golden, eng):
    """Protocol B: E and dE/dxyz from the REFERENCE's alpha must match the reference's predictions."""
    g = golden
    model = eng.PackedModel.from_training_set(dev(g["xyz_train"]), eng.xeq(dev(g["eq"])), dev(g["alpha"]), g["sigma"],
                                              g["prior"], g["ntrval"], g["ntrgr"])
    e, gr = model.predict(dev(g["xyz_test"]))
    e, gr = e.cpu().numpy(), gr.cpu().numpy()
    ref_e = g["prior"] + g["E_test"]
    assert np.abs(e - ref_e).max() <= E_RTOL * np.abs(ref_e).max()
    assert np.abs(gr - g["G_test"]).max() <= G_TOL * max(1.0, np.abs(g["G_test"]).max())
    # energy-only passes skip the second GEMM (row sums of W instead of the ones column): same numbers to rounding
    e2, none = model.predict(dev(g["xyz_test"]), energy=True, gradient=False)
    assert none is None and np.abs(e2.cpu().numpy() - ref_e).max() <= E_RTOL * np.abs(ref_e).max()
    assert np.abs(e2.cpu().numpy() - e).max() <= 1e-12 * np.abs(ref_e).max()
    none, g2 = model.predict(dev(g["xyz_test"]), energy=False, gradient=True)
    assert none is None and np.array_equal(g2.cpu().numpy(), gr)


def test_golden_hessian(golden, eng):
    """kreg_hessian against the reference's YhessXYZest (golden H_test, KREG.f90:393-420) at the reference's alpha."""
    g = golden
    nv = int(g["ntrval"])
    xeq = eng.xeq(dev(g["eq"]))
    x_tr = eng.descriptors(dev(g["xyz_train"][:max(nv, 1)]), xeq)[:nv]
    h = eng.hessian(x_tr, dev(g["alpha"][:nv]) if nv else None, xeq, g["sigma"], dev(g["xyz_test"])).cpu().numpy()
    ref = g["H_test"]
    assert h.shape == ref.shape
    assert np.abs(h - ref).max() <= H_TOL * max(1.0, np.abs(ref).max())
    assert np.abs(h - h.transpose(0, 2, 1)).max() <= H_TOL * max(1.0, np.abs(ref).max())
    if nv == 0:
        assert not h.any()


@pytest.mark.parametrize("nat,ntr,nte", [(2, 5, 3), (3, 33, 4), (12, 100, 7), (16, 64, 3), (21, 40, 2), (24, 37, 3),
                                         (25, 40, 2), (32, 37, 3), (48, 9, 2)])
def test_hessian_vs_oracle(eng, oracle, nat, ntr, nte):
    eq = S.equilibrium(nat)
    xtr, xte = S.geometries(eq, ntr, 11), S.geometries(eq, nte, 12)
    xeq_o = oracle.xeq(eq)
    X_tr, X_te = oracle.descriptors(xtr, xeq_o), oracle.descriptors(xte, xeq_o)
    alpha = np.random.default_rng(5).standard_normal(ntr)
    sigma = 0.7 * np.sqrt(nat)
    ref = oracle.hessian_kregf90(X_tr, alpha, sigma, ntr, xte, X_te)
    h = eng.hessian(dev(X_tr), dev(alpha), eng.xeq(dev(eq)), sigma, dev(xte)).cpu().numpy()
    assert np.abs(h - ref).max() <= H_TOL * max(1.0, np.abs(ref).max())


def test_golden_train_residual_and_prediction(golden, eng):
    """Protocol C: alpha from cuSOLVER: residual of the shifted system + predictions vs the reference."""
    g = golden
    xeq = eng.xeq(dev(g["eq"]))
    xyz = dev(g["xyz_train"])
    k = eng.build_kernel_matrix(xyz, xeq, g["sigma"], g["lambdav"], g["lambdagradxyz"], use_values=g["ntrval"] > 0,
                                use_gradients=g["ntrgr"] > 0)
    alpha = eng.solve(k, dev(g["ytrain"])).cpu().numpy()
    a = g["K"].copy()
    a[np.diag_indices_from(a)] += np.r_[np.full(g["ntrval"], g["lambdav"]), np.full(g["ntrgr"], g["lambdagradxyz"])]
    res = np.linalg.norm(a @ alpha - g["ytrain"]) / np.linalg.norm(g["ytrain"])
    assert res < 1e-9
    model = eng.PackedModel.from_training_set(xyz, xeq, dev(alpha), g["sigma"], g["prior"], g["ntrval"], g["ntrgr"])
    e, gr = model.predict(dev(g["xyz_test"]))
    ref_e = g["prior"] + g["E_test"]
    # cond(K + lambda) * eps bounds how far two Cholesky libraries may differ
    cond = np.linalg.cond(a)
    tol = max(E_RTOL, 50 * cond * 2.2e-16)
    assert np.abs(e.cpu().numpy() - ref_e).max() <= tol * np.abs(ref_e).max()


# ---- seeded comparisons with the oracle at larger sizes -------------------------------------

@pytest.mark.parametrize("nat,ntr,nte,sigma", [(9, 777, 1500, 1.0), (9, 64, 300, 0.03125), (3, 50, 257, 0.7),
                                                (12, 130, 200, 1.5), (21, 200, 333, 2.0), (2, 9, 5, 1.0),
                                                (24, 210, 301, 2.2), (22, 64, 100, 2.0)])
def test_predict_values_model_vs_oracle(eng, oracle, nat, ntr, nte, sigma):
    eq = S.equilibrium(nat)
    xtr, xte = S.geometries(eq, ntr, 1), S.geometries(eq, nte, 2)
    xeq = oracle.xeq(eq)
    X, Xte = oracle.descriptors(xtr, xeq), oracle.descriptors(xte, xeq)
    alpha = np.random.default_rng(3).standard_normal(ntr)
    prior = -12.5
    e_ref, g_ref, sc = oracle.predict_factored(X, alpha, None, sigma, prior, xte, Xte, want_scale=True)
    model = eng.PackedModel.from_training_set(dev(xtr), dev(xeq), dev(alpha), sigma, prior, ntr, 0)
    e, g = model.predict(dev(xte))
    e, g = e.cpu().numpy(), g.cpu().numpy()
    assert np.all(np.abs(e - e_ref) <= 1e-13 * np.maximum(sc, 1.0) + E_RTOL * np.abs(e_ref))
    assert np.abs(g - g_ref).max() <= G_TOL * max(1.0, np.abs(g_ref).max())


@pytest.mark.parametrize("nat,ntr,nte,with_values", [(9, 40, 130, True), (9, 33, 70, False), (5, 61, 90, True),
                                                     (21, 17, 40, True), (24, 19, 70, True), (23, 9, 33, False)])
def test_predict_gradient_model_vs_oracle(eng, oracle, nat, ntr, nte, with_values):
    eq = S.equilibrium(nat)
    xtr, xte = S.geometries(eq, ntr, 1), S.geometries(eq, nte, 2)
    xeq = oracle.xeq(eq)
    X, Xte = oracle.descriptors(xtr, xeq), oracle.descriptors(xte, xeq)
    nv, ng = (ntr if with_values else 0), 3 * nat * ntr
    alpha = np.random.default_rng(4).standard_normal(nv + ng)
    sigma, prior = 1.2, 3.0
    v = oracle.precompute_v(xtr, X, alpha[nv:])
    e_ref, g_ref, sc = oracle.predict_factored(X, alpha[:nv] if nv else None, v, sigma, prior, xte, Xte, want_scale=True)
    model = eng.PackedModel.from_training_set(dev(xtr), dev(xeq), dev(alpha), sigma, prior, nv, ng)
    e, g = model.predict(dev(xte))
    e, g = e.cpu().numpy(), g.cpu().numpy()
    assert np.all(np.abs(e - e_ref) <= 1e-13 * np.maximum(sc, 1.0) + E_RTOL * np.abs(e_ref))
    assert np.abs(g - g_ref).max() <= G_TOL * max(1.0, np.abs(g_ref).max())
    # v itself
    vg = eng.precompute_v(dev(xtr), dev(xeq), dev(alpha[nv:]).reshape(ntr, nat, 3)).cpu().numpy()
    assert np.abs(vg - v).max() <= 1e-13 * np.abs(v).max()


@pytest.mark.parametrize("nat,ntr,sigma", [(9, 1000, 1.0), (9, 300, 0.05), (21, 150, 2.0), (4, 129, 0.5), (12, 257, 1.0)])
def test_kernel_matrix_values_vs_oracle(eng, oracle, nat, ntr, sigma):
    eq = S.equilibrium(nat)
    xtr = S.geometries(eq, ntr, 1)
    xeq = oracle.xeq(eq)
    X = oracle.descriptors(xtr, xeq)
    ref = oracle.kernel_matrix(xtr, X, sigma, ntr, 0)
    from mlatom_b200._lib import FILL_FULL

    k = eng.build_kernel_matrix(dev(xtr), dev(xeq), sigma, 1e-6, 0.0, fill=FILL_FULL).cpu().numpy()
    assert np.array_equal(k, k.T)
    ref[np.diag_indices_from(ref)] += 1e-6
    assert np.abs(k - ref).max() <= K_RTOL
    kl = eng.build_kernel_matrix(dev(xtr), dev(xeq), sigma, 1e-6, 0.0).cpu().numpy()
    assert np.array_equal(np.tril(kl), np.tril(k))


@pytest.mark.parametrize("nat,ntr,with_values", [(9, 30, True), (9, 25, False), (5, 40, True), (21, 6, True)])
def test_kernel_matrix_gradient_blocks_vs_oracle(eng, oracle, nat, ntr, with_values):
    eq = S.equilibrium(nat)
    xtr = S.geometries(eq, ntr, 1)
    xeq = oracle.xeq(eq)
    X = oracle.descriptors(xtr, xeq)
    nv, ng = (ntr if with_values else 0), 3 * nat * ntr
    ref = oracle.kernel_matrix(xtr, X, 1.3, nv, ng, factored=True)
    from mlatom_b200._lib import FILL_FULL

    k = eng.build_kernel_matrix(dev(xtr), dev(xeq), 1.3, 0.0, 0.0, use_values=with_values, use_gradients=True,
                                fill=FILL_FULL).cpu().numpy()
    assert np.abs(k - ref).max() <= K_RTOL * np.abs(ref).max()
    kl = eng.build_kernel_matrix(dev(xtr), dev(xeq), 1.3, 0.0, 0.0, use_values=with_values,
                                 use_gradients=True).cpu().numpy()
    assert np.abs(np.tril(kl) - np.tril(ref)).max() <= K_RTOL * np.abs(ref).max()


def test_kernel_matrix_row_slabs_equal_whole(eng):
    """Multi-GPU sharding unit: disjoint row slabs assembled separately give the same bytes."""
    nat, ntr = 9, 300
    eq = S.equilibrium(nat)
    xtr, xeq = dev(S.geometries(eq, ntr, 1)), eng.xeq(dev(eq))
    for use_g in (False, True):
        n = 40 if use_g else ntr
        whole = eng.build_kernel_matrix(xtr[:n], xeq, 1.0, 1e-6, 1e-5, use_gradients=use_g)
        m = whole.shape[0]
        out = torch.full_like(whole, float("nan"))
        cuts = [0, m // 3, m // 3 + 17, m]
        for r0, r1 in zip(cuts[:-1], cuts[1:]):
            eng.build_kernel_matrix(xtr[:n], xeq, 1.0, 1e-6, 1e-5, use_gradients=use_g, rows=(r0, r1), out=out)
        assert torch.equal(torch.tril(out), torch.tril(whole))


def test_kernel_block_vs_oracle(eng, oracle):
    nat = 9
    eq = S.equilibrium(nat)
    xa, xb = S.geometries(eq, 150, 5), S.geometries(eq, 333, 6)
    xeq = oracle.xeq(eq)
    Xa, Xb = oracle.descriptors(xa, xeq), oracle.descriptors(xb, xeq)
    d2 = ((Xa[:, None, :] - Xb[None, :, :]) ** 2).sum(-1)
    ref = np.exp(-d2 / (2 * 0.9 ** 2))
    k = eng.kernel_block(dev(xa), dev(xb), dev(xeq), 0.9).cpu().numpy()
    assert np.abs(k - ref).max() <= K_RTOL


# ---- properties at full size (no oracle needed) -------------------------------------------

def test_full_size_properties(eng):
    """BASELINE configs[1] at full size (Nat=9, Ntrain=10k, 1M test geometries): finite-difference consistency of E
    and dE, permutation equivariance over the batch, and host-pipeline == device path."""
    nat, ntr, nte = 9, 10000, 1000000
    eq = S.equilibrium(nat)
    xtr = S.geometries(eq, ntr, 1)
    alpha = np.random.default_rng(0).standard_normal(ntr) * 1e-2
    model = eng.PackedModel.from_training_set(dev(xtr), eng.xeq(dev(eq)), dev(alpha), 1.0, -100.0, ntr, 0)
    xte = S.geometries(eq, nte, 2)
    e, g = model.predict(dev(xte))
    assert torch.isfinite(e).all() and torch.isfinite(g).all()
    # batch-order equivariance, bit for bit
    perm = np.random.default_rng(1).permutation(nte)
    e2, g2 = model.predict(dev(xte[perm]))
    assert torch.equal(e2, e[torch.as_tensor(perm).cuda()]) and torch.equal(g2, g[torch.as_tensor(perm).cuda()])
    # central finite difference along a random direction, first 64 geometries
    h = 1e-5
    dirn = np.random.default_rng(2).standard_normal((64, nat, 3))
    ep, _ = model.predict(dev(xte[:64] + h * dirn), gradient=False)
    em, _ = model.predict(dev(xte[:64] - h * dirn), gradient=False)
    fd = ((ep - em) / (2 * h)).cpu().numpy()
    an = (g[:64].cpu().numpy() * dirn).sum((1, 2))
    assert np.abs(fd - an).max() <= 1e-6 * max(1.0, np.abs(an).max())
    # host pipeline: whole-wave chunks reproduce the device path bit for bit; small chunks take the split
    # (latency) mode, whose fixed-order slice reduction differs from the unsplit sum only by rounding
    eh, gh = model.predict_host(xte)
    torch.cuda.synchronize()
    assert torch.equal(eh, e.cpu()) and torch.equal(gh, g.cpu())
    eh, gh = model.predict_host(xte, chunk=8192)
    torch.cuda.synchronize()
    assert (eh - e.cpu()).abs().max() <= 1e-12 * e.abs().max().item() and (gh - g.cpu()).abs().max() <= 1e-11


@pytest.mark.parametrize("nat,ntr,use_grads", [(9, 2000, False), (9, 300, True), (21, 600, False), (5, 100, False),
                                               (24, 500, False), (24, 40, True)])
def test_small_batch_latency_mode_vs_oracle(eng, oracle, nat, ntr, use_grads):
    """Single-geometry / small-batch predictions (MD, geometry optimisation) split the training set over the chip;
    they must agree with the oracle and with the same geometries predicted inside a large batch."""
    eq = S.equilibrium(nat)
    xtr, xte = S.geometries(eq, ntr, 1), S.geometries(eq, 40000, 2)
    xeq = oracle.xeq(eq)
    X = oracle.descriptors(xtr, xeq)
    nv, ng = ntr, (3 * nat * ntr if use_grads else 0)
    alpha = np.random.default_rng(8).standard_normal(nv + ng) * 0.1
    model = eng.PackedModel.from_training_set(dev(xtr), dev(xeq), dev(alpha), 1.3, 2.0, nv, ng)
    d_xte = dev(xte)
    e_big, g_big = model.predict(d_xte)  # 40000 geometries: unsplit
    v = oracle.precompute_v(xtr, X, alpha[nv:]) if use_grads else None
    for n in (1, 7, 300, 5000):
        e, g = model.predict(d_xte[:n].contiguous())
        e_ref, g_ref, sc = oracle.predict_factored(X, alpha[:nv], v, 1.3, 2.0, xte[:n], oracle.descriptors(xte[:n], xeq),
                                                   want_scale=True)
        assert np.all(np.abs(e.cpu().numpy() - e_ref) <= 1e-13 * np.maximum(sc, 1.0) + E_RTOL * np.abs(e_ref))
        assert np.abs(g.cpu().numpy() - g_ref).max() <= G_TOL * max(1.0, np.abs(g_ref).max())
        assert (e - e_big[:n]).abs().max().item() <= 1e-12 * max(1.0, np.abs(sc).max())
        assert (g - g_big[:n]).abs().max().item() <= 1e-11 * max(1.0, np.abs(g_ref).max())
        e_only, _ = model.predict(d_xte[:n].contiguous(), gradient=False)
        assert (e_only - e).abs().max().item() <= 1e-12 * max(1.0, np.abs(sc).max())


@pytest.mark.parametrize("nat", list(range(2, 25)) + [25, 26, 27, 29, 32, 33, 40, 48, 57, 64])
def test_every_supported_molecule_size(eng, oracle, nat):
    """Natoms = 2..64 (KREG_MIN/MAX_NATOMS; above 24 the run-time-Natoms kernels and the tiled prediction path): K blocks
    and E/dE for a gradient-trained model against the oracle."""
    ntr, nte = 11, 23
    eq = S.equilibrium(nat)
    xtr, xte = S.geometries(eq, ntr, 1), S.geometries(eq, nte, 2)
    xeq = oracle.xeq(eq)
    X, Xte = oracle.descriptors(xtr, xeq), oracle.descriptors(xte, xeq)
    nv, ng = ntr, 3 * nat * ntr
    sigma = 0.8 + 0.06 * nat
    ref = oracle.kernel_matrix(xtr, X, sigma, nv, ng, factored=True)
    k = eng.build_kernel_matrix(dev(xtr), dev(xeq), sigma, 0.0, 0.0, use_gradients=True).cpu().numpy()
    assert np.abs(np.tril(k) - np.tril(ref)).max() <= K_RTOL * np.abs(ref).max()
    alpha = np.random.default_rng(nat).standard_normal(nv + ng) * 0.3
    v = oracle.precompute_v(xtr, X, alpha[nv:])
    for use_g in (False, True):
        a = alpha if use_g else alpha[:nv]
        e_ref, g_ref, sc = oracle.predict_factored(X, alpha[:nv], v if use_g else None, sigma, 1.0, xte, Xte, want_scale=True)
        model = eng.PackedModel.from_training_set(dev(xtr), dev(xeq), dev(a), sigma, 1.0, nv, ng if use_g else 0)
        e, g = model.predict(dev(xte))
        assert np.all(np.abs(e.cpu().numpy() - e_ref) <= 1e-13 * np.maximum(sc, 1.0) + E_RTOL * np.abs(e_ref))
        assert np.abs(g.cpu().numpy() - g_ref).max() <= G_TOL * max(1.0, np.abs(g_ref).max())


def test_edge_cases(eng):
    nat = 9
    eq = S.equilibrium(nat)
    xeq = eng.xeq(dev(eq))
    model = eng.PackedModel.from_training_set(dev(S.geometries(eq, 5, 1)), xeq, dev(np.ones(5)), 1.0, 0.0, 5, 0)
    # empty batch
    e, g = model.predict(torch.empty((0, nat, 3), dtype=torch.float64, device="cuda"))
    assert e.shape == (0,) and g.shape == (0, nat, 3)
    # a test geometry identical to a training geometry: k = 1 exactly for that pair
    x1 = S.geometries(eq, 1, 1)
    m1 = eng.PackedModel.from_training_set(dev(x1), xeq, dev(np.array([2.0])), 1.0, 0.0, 1, 0)
    e, g = m1.predict(dev(x1))
    assert abs(e.item() - 2.0) <= 1e-14 and g.abs().max().item() <= 1e-13
    # unsupported molecule size fails loudly, never silently
    from mlatom_b200._lib import KregError

    with pytest.raises(KregError):
        eng.xeq(dev(np.random.default_rng(0).standard_normal((65, 3))))


def test_full_size_kernel_matrices(eng, oracle):
    """BASELINE configs[2] and [4] at full size: 50k x 50k values K (20 GB buffer) and the 56 000-dim
    gradient-augmented K (25 GB).  Checked through size-independent properties: sampled rows against the independent
    rectangular-block path, exact 1 + lambda diagonal, and sampled geometry-pair blocks against the CPU oracle run on
    just those two geometries."""
    nat = 9
    eq = S.equilibrium(nat)
    xeq_h = oracle.xeq(eq)
    xeq = dev(xeq_h)
    rng = np.random.default_rng(11)
    # ---- cfg3: values block, Ntrain = 50 000 ----
    ntr, lam = 50000, 1e-6
    xyz_h = S.geometries(eq, ntr, 1)
    xyz = dev(xyz_h)
    k = eng.build_kernel_matrix(xyz, xeq, 1.0, lam, lam)
    diag = torch.diagonal(k)
    assert torch.all(diag == 1.0 + lam)
    rows = np.sort(rng.choice(ntr, size=48, replace=False))
    blk = eng.kernel_block(xyz[torch.as_tensor(rows).cuda()].contiguous(), xyz, xeq, 1.0)
    for n, i in enumerate(rows):
        a, b = k[i, :i], blk[n, :i]
        assert (a - b).abs().max().item() <= K_RTOL if i else True
    # against the oracle on a few rows (independent arithmetic)
    X = oracle.descriptors(xyz_h[rows[:4]], xeq_h)
    Xall = oracle.descriptors(xyz_h[:2000], xeq_h)
    ref = np.exp(-((X[:, None, :] - Xall[None, :, :]) ** 2).sum(-1) / 2.0)
    got = blk[:4, :2000].cpu().numpy()
    mask = np.ones_like(ref, dtype=bool)
    for n, i in enumerate(rows[:4]):
        if i < 2000:
            mask[n, i] = False
    assert np.abs(got - ref)[mask].max() <= K_RTOL
    del k, blk
    torch.cuda.empty_cache()
    # ---- cfg5: gradient-augmented, 2000 geometries -> M = 56 000 ----
    ntr, g3 = 2000, 3 * nat
    xyz_h = S.geometries(eq, ntr, 1)
    k = eng.build_kernel_matrix(dev(xyz_h), xeq, 1.0, lam, 2 * lam, use_gradients=True)
    m = ntr + g3 * ntr
    assert k.shape == (m, m)
    d = torch.diagonal(k)
    assert torch.all(d[:ntr] == 1.0 + lam)
    for _ in range(6):
        j, i = np.sort(rng.choice(ntr, size=2, replace=False))
        pair = xyz_h[[j, i]]
        Xp = oracle.descriptors(pair, xeq_h)
        ref = oracle.kernel_matrix(pair, Xp, 1.0, 2, 2 * g3)  # rows/cols: [v_j, v_i, g_j(27), g_i(27)]
        gi = slice(ntr + i * g3, ntr + (i + 1) * g3)
        gj = slice(ntr + j * g3, ntr + (j + 1) * g3)
        scale = np.abs(ref).max()
        assert abs(k[i, j].item() - ref[1, 0]) <= K_RTOL
        assert np.abs(k[gi, j].cpu().numpy() - ref[2 + g3:, 0]).max() <= K_RTOL * scale          # d k(i,j) / d M_i
        assert np.abs(k[gj, i].cpu().numpy() - ref[2:2 + g3, 1]).max() <= K_RTOL * scale         # d k(j,i) / d M_j
        assert np.abs(k[gi, gj].cpu().numpy() - ref[2 + g3:, 2:2 + g3]).max() <= K_RTOL * scale  # grad-grad (i, j)
        own = k[gi, gi].cpu().numpy() - 2 * lam * np.eye(g3)
        assert np.abs(np.tril(own) - np.tril(ref[2 + g3:, 2 + g3:])).max() <= K_RTOL * scale      # diagonal block
