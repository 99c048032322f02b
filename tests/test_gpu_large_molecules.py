"""Molecules above 24 atoms (VERDICT r01 missing #1): the reference has no size limit on the path
(KREG.f90:34-52, 241-271, 293-346, 534-596 take Natoms at run time).  Prediction runs the tiled three-kernel path
(kreg_predict_big.cu), the gradient blocks of K the run-time-Natoms kernels; both against the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from mlatom_b200 import synthetic as S  # noqa: E402

E_RTOL, G_TOL, K_RTOL = 1e-10, 1e-8, 1e-13


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mlatom_b200 import engine

    engine.device_check()
    return engine


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.mark.parametrize("nat,ntr,nte,use_grads", [(32, 700, 300, False), (32, 73, 129, True), (28, 1000, 5000, False),
                                                   (25, 36, 1, False), (40, 37, 260, True), (48, 150, 64, False)])
def test_prediction_vs_oracle(eng, oracle, nat, ntr, nte, use_grads):
    """Ragged sizes on purpose: Ntrain not a multiple of the 36-wide W chunks or the 64-wide column tiles, test batches
    not a multiple of the 128-row tiles, a single geometry; values- and gradient-trained; energy-only requests."""
    eq = S.equilibrium(nat)
    xtr, xte = S.geometries(eq, ntr, 1), S.geometries(eq, nte, 2)
    xeq = oracle.xeq(eq)
    X, Xte = oracle.descriptors(xtr, xeq), oracle.descriptors(xte, xeq)
    nv, ng = ntr, (3 * nat * ntr if use_grads else 0)
    sigma = 0.8 + 0.06 * nat
    alpha = np.random.default_rng(nat).standard_normal(nv + ng) * 0.3
    v = oracle.precompute_v(xtr, X, alpha[nv:]) if use_grads else None
    e_ref, g_ref, sc = oracle.predict_factored(X, alpha[:nv], v, sigma, 1.5, xte, Xte, want_scale=True)
    model = eng.PackedModel.from_training_set(dev(xtr), dev(xeq), dev(alpha), sigma, 1.5, nv, ng)
    e, g = model.predict(dev(xte))
    assert np.all(np.abs(e.cpu().numpy() - e_ref) <= 1e-13 * np.maximum(sc, 1.0) + E_RTOL * np.abs(e_ref))
    assert np.abs(g.cpu().numpy() - g_ref).max() <= G_TOL * max(1.0, np.abs(g_ref).max())
    e_only, none = model.predict(dev(xte), gradient=False)
    assert none is None and (e_only - e).abs().max().item() <= 1e-12 * max(1.0, np.abs(sc).max())
    none, g_only = model.predict(dev(xte), energy=False)
    assert none is None and torch.equal(g_only, g)
    # a sub-batch gives the same numbers as the same geometries inside the large batch (no dependence on tiling)
    n = min(nte, 5)
    e5, g5 = model.predict(dev(xte[:n]))
    assert (e5 - e[:n]).abs().max().item() <= 1e-12 * max(1.0, np.abs(sc).max())
    assert (g5 - g[:n]).abs().max().item() <= 1e-11 * max(1.0, np.abs(g_ref).max())
    # host pipeline (pinned staging, three streams) on top of the tiled path
    eh, gh = model.predict_host(xte)
    torch.cuda.synchronize()
    assert torch.equal(eh, e.cpu()) and torch.equal(gh, g.cpu())


def test_descriptor_model_and_graph(eng, oracle):
    """A model given by descriptors (MLatomF .unf files) and the CUDA-graphed single-geometry step, 30 atoms."""
    nat, ntr = 30, 200
    eq = S.equilibrium(nat)
    xtr, xte = S.geometries(eq, ntr, 1), S.geometries(eq, 3, 2)
    xeq = oracle.xeq(eq)
    X, Xte = oracle.descriptors(xtr, xeq), oracle.descriptors(xte, xeq)
    alpha = np.random.default_rng(2).standard_normal(ntr)
    e_ref, g_ref = oracle.predict_factored(X, alpha, None, 2.5, -1.0, xte, Xte)
    m1 = eng.PackedModel.from_descriptors(dev(X), dev(xeq), dev(alpha), 2.5, -1.0, nat)
    e, g = m1.predict(dev(xte))
    assert np.abs(e.cpu().numpy() - e_ref).max() <= E_RTOL * np.abs(e_ref).max()
    assert np.abs(g.cpu().numpy() - g_ref).max() <= G_TOL
    gp = eng.GraphPredictor(m1, 1)
    for i in range(3):
        eg, gg = gp(xte[i:i + 1])
        assert np.abs(eg - e_ref[i]).max() <= E_RTOL * np.abs(e_ref).max() and np.abs(gg - g_ref[i]).max() <= G_TOL


@pytest.mark.parametrize("nat,ntr", [(27, 12), (32, 40), (36, 9)])
def test_kernel_matrix_slabs_and_full_fill(eng, oracle, nat, ntr):
    """Gradient-augmented K of a large molecule: lower and full fill, and row slabs == whole matrix (the sharded path)."""
    from mlatom_b200._lib import FILL_FULL

    eq = S.equilibrium(nat)
    xtr = S.geometries(eq, ntr, 1)
    xeq = oracle.xeq(eq)
    X = oracle.descriptors(xtr, xeq)
    nv, ng = ntr, 3 * nat * ntr
    m = nv + ng
    ref = oracle.kernel_matrix(xtr, X, 1.7, nv, ng, factored=True)
    ref[np.diag_indices(m)] += np.r_[np.full(nv, 0.25), np.full(ng, 0.5)]
    scale = np.abs(ref).max()
    d_x, d_q = dev(xtr), dev(xeq)
    k = eng.build_kernel_matrix(d_x, d_q, 1.7, 0.25, 0.5, use_gradients=True)
    assert np.abs(np.tril(k.cpu().numpy()) - np.tril(ref)).max() <= K_RTOL * scale
    kf = eng.build_kernel_matrix(d_x, d_q, 1.7, 0.25, 0.5, use_gradients=True, fill=FILL_FULL).cpu().numpy()
    assert np.abs(kf - ref).max() <= K_RTOL * scale
    cuts = [0, nv // 2, nv + 3 * nat * (ntr // 3), nv + 3 * nat * (ntr // 3) + 7, m]
    ks = torch.full((m, m), float("nan"), dtype=torch.float64, device="cuda")
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        eng.build_kernel_matrix(d_x, d_q, 1.7, 0.25, 0.5, use_gradients=True, rows=(r0, r1), out=ks)
    assert torch.equal(torch.tril(ks), torch.tril(k))
    # gradients only (NtrVal = 0)
    kg = eng.build_kernel_matrix(d_x, d_q, 1.7, 0.0, 0.5, use_values=False, use_gradients=True).cpu().numpy()
    assert np.abs(np.tril(kg) - np.tril(ref[nv:, nv:])).max() <= K_RTOL * scale


def test_train_and_predict_through_the_api(eng, oracle):
    """End to end through KREG_API on a 26-atom molecule: train with values + gradients, predict, compare with the
    oracle at the trained alpha."""
    from mlatom_b200.kreg_api import KREG_API

    nat, ntr = 26, 30
    eq = S.equilibrium(nat)
    xtr, xte = S.geometries(eq, ntr, 1), S.geometries(eq, 70, 2)
    etr, gtr = S.pair_potential(xtr, eq)
    api = KREG_API().train_arrays(2.0, 1e-8, 1e-8, xtr, eq, y=etr, ygrad=gtr, prior="mean")
    a = api.alpha.reshape(-1)
    e, g = api.predict_arrays(xte)
    xeq = oracle.xeq(eq)
    X, Xte = oracle.descriptors(xtr, xeq), oracle.descriptors(xte, xeq)
    v = oracle.precompute_v(xtr, X, a[ntr:])
    e_ref, g_ref = oracle.predict_factored(X, a[:ntr], v, 2.0, float(api.prior), xte, Xte)
    assert np.abs(e - e_ref).max() <= E_RTOL * np.abs(e_ref).max()
    assert np.abs(g - g_ref).max() <= G_TOL * max(1.0, np.abs(g_ref).max())
    # the trained model interpolates its own training set: (K + lambda) alpha = y with lambda = 1e-8
    # (K + lambda) alpha = y  =>  the model evaluated on its training set returns y - lambda alpha
    e_fit, g_fit = api.predict_arrays(xtr)
    assert np.abs(e_fit - (etr - 1e-8 * a[:ntr])).max() <= 1e-8
    assert np.abs(g_fit.reshape(-1) - (gtr.reshape(-1) - 1e-8 * a[ntr:])).max() <= 1e-8
