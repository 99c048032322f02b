"""GPU: the drop-in Python API (models.kreg / KREG_API) end to end on the device, against the oracle."""
import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from mlatom_b200 import data as D, synthetic as S  # noqa: E402


@pytest.fixture(scope="module")
def M():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mlatom_b200 import models

    return models


def make_db(nat, n, seed, labels=True):
    eq = S.equilibrium(nat)
    xyz = S.geometries(eq, n, seed)
    e, g = S.pair_potential(xyz, eq)
    if labels:
        return D.molecular_database.from_arrays(S.atomic_numbers(nat), xyz, energy=e, energy_gradients=g), xyz, e, g
    return D.molecular_database.from_arrays(S.atomic_numbers(nat), xyz), xyz, e, g


@pytest.mark.parametrize("nat,ntr,use_grads,sigma,lam", [(9, 200, False, 1.0, 1e-6), (9, 30, True, 1.0, 1e-6),
                                                           (5, 50, True, 0.8, 1e-5), (21, 60, False, 2.0, 1e-6)])
def test_kreg_train_predict_vs_oracle(M, oracle, tmp_path, nat, ntr, use_grads, sigma, lam):
    db, xyz, e, g = make_db(nat, ntr, 1)
    tdb, xte, e_true, g_true = make_db(nat, 64, 2, labels=False)
    model = M.kreg(model_file=str(tmp_path / "model"), hyperparameters={"sigma": sigma, "lambda": lam})
    model.train(molecular_database=db, property_to_learn="energy",
                xyz_derivative_property_to_learn="energy_gradients" if use_grads else None, prior="mean")
    model.predict(molecular_database=tdb, calculate_energy=True, calculate_energy_gradients=True)
    e_gpu = tdb.get_properties("energy")
    g_gpu = tdb.get_xyz_vectorial_properties("energy_gradients")
    # oracle: same training through the CPU restatement (LAPACK Cholesky), same predictions
    xeq = oracle.xeq(xyz[int(np.argmin(e))])
    X, Xte = oracle.descriptors(xyz, xeq), oracle.descriptors(xte, xeq)
    nv, ng = ntr, (3 * nat * ntr if use_grads else 0)
    prior = float(e.mean())
    k = oracle.kernel_matrix(xyz, X, sigma, nv, ng, factored=True)
    alpha = oracle.solve(k, oracle.ytrain(e, g if use_grads else None, prior), nv, ng, lam, lam)
    v = oracle.precompute_v(xyz, X, alpha[nv:]) if use_grads else None
    e_ref, g_ref = oracle.predict_factored(X, alpha[:nv], v, sigma, prior, xte, Xte)
    # trained-alpha path: two Cholesky libraries differ by cond*eps; state the bound
    a = k.copy()
    a[np.diag_indices_from(a)] += lam
    tol = max(1e-10, 100 * np.linalg.cond(a) * 2.2e-16)
    assert np.abs(e_gpu - e_ref).max() <= tol * np.abs(e_ref).max()
    assert np.abs(g_gpu - g_ref).max() <= max(1e-8, tol) * max(1.0, np.abs(g_ref).max())
    # the model actually learned the potential
    if nat <= 9:
        assert np.sqrt(np.mean((e_gpu - e_true) ** 2)) < np.std(e_true)
    # fixed-alpha protocol through the API: load the file and predict with the oracle's alpha
    api = model.kreg_api
    api.alpha = alpha.reshape(1, -1)
    api._model = None
    e2, g2 = api.predict_arrays(xte)
    assert np.abs(e2 - e_ref).max() <= 1e-10 * np.abs(e_ref).max()
    assert np.abs(g2 - g_ref).max() <= 1e-8


def test_model_file_roundtrip_and_tensor_path(M, oracle, tmp_path):
    db, xyz, e, g = make_db(9, 80, 1)
    model = M.kreg(model_file=str(tmp_path / "rt.npz"), hyperparameters={"sigma": 1.0, "lambda": 1e-6})
    model.train(molecular_database=db, property_to_learn="energy", xyz_derivative_property_to_learn="energy_gradients")
    xte = S.geometries(S.equilibrium(9), 1000, 5)
    e1, g1 = model.predict_xyz(torch.from_numpy(xte).cuda())
    again = M.kreg(model_file=str(tmp_path / "rt.npz"))
    e2, g2 = again.predict_xyz(torch.from_numpy(xte).cuda())
    assert torch.equal(e1, e2) and torch.equal(g1, g2)
    e3, g3 = again.predict_xyz(xte)  # numpy in -> streamed host pipeline -> numpy out
    assert np.array_equal(e3, e1.cpu().numpy()) and np.array_equal(g3, g1.cpu().numpy())
    mol = D.molecule.from_arrays(S.atomic_numbers(9), xte[3])
    again.predict(molecule=mol, calculate_energy=True, calculate_energy_gradients=True)
    assert mol.energy == e3[3] and np.array_equal(mol.get_energy_gradients(), g3[3])
    assert mol.atoms[0].energy_gradients.shape == (3,)
    # Hessian as the reference defines it (KREG.f90:393-420), stored per molecule like kreg_api.py:130-131
    again.predict(molecule=mol, calculate_hessian=True)
    api = again.kreg_api
    X = np.asarray(api.X)
    xq = oracle.descriptors(xte[3:4], oracle.xeq(np.asarray(api.Req)))
    h_ref = oracle.hessian_kregf90(X, np.asarray(api.alpha).reshape(-1), float(api.sigma), int(api.NtrVal), xte[3:4], xq)[0]
    assert mol.hessian.shape == (27, 27)
    assert np.abs(mol.hessian - h_ref).max() <= 1e-9 * max(1.0, np.abs(h_ref).max())


def test_grid_search_fused_equals_pointwise_on_device(M, tmp_path):
    sub, _, _, _ = make_db(9, 60, 1)
    val, _, _, _ = make_db(9, 40, 3)
    res = []
    for fused in (True, False):
        model = M.kreg(model_file=str(tmp_path / f"grid{fused}"))
        model.hyperparameters["lambda"].minval, model.hyperparameters["lambda"].maxval = 1e-9, 1e-2
        model.hyperparameters["sigma"].minval, model.hyperparameters["sigma"].maxval = 0.25, 16.0
        model.optimize_hyperparameters(
            hyperparameters=["lambda", "sigma"], optimization_algorithm="grid",
            optimization_algorithm_kwargs={"grid_size": 4},
            training_kwargs={"property_to_learn": "energy", "xyz_derivative_property_to_learn": "energy_gradients"},
            prediction_kwargs={"property_to_predict": "estimated_energy",
                               "xyz_derivative_property_to_predict": "estimated_energy_gradients"},
            subtraining_molecular_database=sub, validation_molecular_database=val, fused=fused)
        res.append((model.hyperparameters["lambda"].value, model.hyperparameters["sigma"].value, model.validation_loss))
        # as in the reference the last validation pass ends with reset() (model_cls.py:590): the file is gone,
        # the trained in-memory backend remains (AL then calls model.kreg_api.save_model, al_utils.py:1105)
        assert not os.path.exists(model.model_file) and model.kreg_api.alpha is not None
    assert res[0][:2] == res[1][:2]
    assert abs(res[0][2] - res[1][2]) <= 1e-8 * res[1][2]


def test_log2_grid_search_on_device(M, tmp_path):
    """MLatomF's recursive log2 grid (A_KRR.f90:1049-1376) on the fused device path.  With an odd number of points
    every zoomed level contains the previous winner, so deeper searches can only lower the loss and the result is the
    minimum of everything visited (with an even number the reference -- and this -- may move off the previous best);
    the winner retrains to the loss the search reported."""
    sub, _, _, _ = make_db(9, 60, 1)
    val, _, _, _ = make_db(9, 40, 3)
    kw = dict(hyperparameters=["lambda", "sigma"], optimization_algorithm="log2grid",
              training_kwargs={"property_to_learn": "energy", "xyz_derivative_property_to_learn": "energy_gradients"},
              subtraining_molecular_database=sub, validation_molecular_database=val)
    losses = []
    for depth in (1, 3):
        model = M.kreg(model_file=str(tmp_path / f"lg{depth}"))
        model.hyperparameters["lambda"].minval, model.hyperparameters["lambda"].maxval = 2.0 ** -35, 2.0 ** -6
        model.hyperparameters["sigma"].minval, model.hyperparameters["sigma"].maxval = 0.25, 16.0
        model.optimize_hyperparameters(optimization_algorithm_kwargs={"depth": depth, "points": {"lambda": 3, "sigma": 5}}, **kw)
        best = min(t[2] for t in model.grid_trace)
        assert abs(model.validation_loss - best) <= 1e-8 * best
        assert 0.25 / 4 <= model.hyperparameters["sigma"].value <= 64.0
        losses.append(model.validation_loss)
        assert len(model.grid_trace) == depth * 5 * depth * 3
    assert losses[1] <= losses[0] * (1 + 1e-12)


def test_indefinite_system_falls_back_instead_of_dying(M, tmp_path):
    """The reference falls back to Bunch-Kaufman when dpotrf fails (mathUtils.f90:24-26) -- or kills the
    interpreter on other errors; here: a warning and a pivoted solve."""
    db, _, _, _ = make_db(5, 30, 1)
    model = M.kreg(model_file=str(tmp_path / "neg"), hyperparameters={"sigma": 1.0, "lambda": -0.5})
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        model.train(molecular_database=db, property_to_learn="energy")
    assert any("not positive definite" in str(x.message) for x in w)
    assert np.isfinite(model.kreg_api.alpha).all()


def test_graph_predictor_matches_direct_path(M):
    """CUDA-graph replay for fixed small batches (the MD step): same numbers as the direct call, also through the
    molecule API where KREG_API caches one graph per batch shape."""
    from mlatom_b200 import engine

    db, xyz, e, g = make_db(9, 120, 1)
    model = M.kreg(hyperparameters={"sigma": 1.0, "lambda": 1e-6})
    model.train(molecular_database=db, property_to_learn="energy", xyz_derivative_property_to_learn="energy_gradients",
                save_model=False)
    pm = model.kreg_api._ensure_model()
    xte = S.geometries(S.equilibrium(9), 5, 7)
    e_ref, g_ref = pm.predict(torch.from_numpy(xte).cuda())
    gp = engine.GraphPredictor(pm, 5)
    for _ in range(3):
        e1, g1 = gp(xte)
        assert np.array_equal(e1, e_ref.cpu().numpy()) and np.array_equal(g1, g_ref.cpu().numpy())
    # through the public API, one molecule at a time (md.py:198)
    for k in range(3):
        mol = D.molecule.from_arrays(S.atomic_numbers(9), xte[k])
        model.predict(molecule=mol, calculate_energy=True, calculate_energy_gradients=True)
        e_one, g_one = pm.predict(torch.from_numpy(xte[k:k + 1]).cuda())
        assert mol.energy == e_one.item() and np.array_equal(mol.get_energy_gradients(), g_one[0].cpu().numpy())
    assert len(model.kreg_api._graphs) == 1


@pytest.mark.parametrize("use_grads", [False, True])
def test_unf_model_file_roundtrip_on_device(M, tmp_path, use_grads):
    """.unf written from a trained model, read into a fresh backend: identical predictions (values-only files carry
    descriptors, not geometries -> the pack-from-descriptors entry point)."""
    from mlatom_b200.kreg_api import KREG_API

    db, xyz, e, g = make_db(9, 60, 1)
    model = M.kreg(hyperparameters={"sigma": 1.0, "lambda": 1e-6})
    model.train(molecular_database=db, property_to_learn="energy",
                xyz_derivative_property_to_learn="energy_gradients" if use_grads else None, prior="mean", save_model=False)
    xte = S.geometries(S.equilibrium(9), 500, 9)
    e1, g1 = model.predict_xyz(torch.from_numpy(xte).cuda())
    path = str(tmp_path / "model.unf")
    model.kreg_api.save_unf(path, atomic_numbers=S.atomic_numbers(9))
    api = KREG_API().load_unf(path)
    assert (api.XYZ is None) == (not use_grads)
    e2, g2 = api.predict_arrays(torch.from_numpy(xte).cuda())
    assert (e1 - e2).abs().max().item() <= 1e-12 * e1.abs().max().item()
    assert (g1 - g2).abs().max().item() <= 1e-11


def test_batched_md_on_device_conserves_energy(M):
    """Velocity Verlet for 64 trajectories entirely on the GPU (CUDA-graphed step): total energy is conserved, i.e.
    the predicted gradients are the derivative of the predicted energies; graph replay == eager stepping."""
    from mlatom_b200 import engine, md

    nat, ntr = 9, 400
    eq = S.equilibrium(nat)
    xtr = S.geometries(eq, ntr, 1, spread=0.08)
    e, g = S.pair_potential(xtr, eq)
    api_model = M.kreg(hyperparameters={"sigma": 1.0, "lambda": 1e-8})
    db = D.molecular_database.from_arrays(S.atomic_numbers(nat), xtr, energy=e, energy_gradients=g)
    api_model.train(molecular_database=db, property_to_learn="energy", xyz_derivative_property_to_learn="energy_gradients",
                    save_model=False)
    pm = api_model.kreg_api._ensure_model()
    b = 64
    x0 = torch.from_numpy(S.geometries(eq, b, 5, spread=0.02)).cuda()
    v0 = torch.from_numpy(np.random.default_rng(6).standard_normal((b, nat, 3)) * 0.02).cuda()
    masses = torch.from_numpy(np.where(S.atomic_numbers(nat) == 1, 1.0, 12.0)).cuda()
    runs = {}
    for use_graph in (True, False):
        sim = md.BatchedNVE(pm, x0, v0, masses, dt=0.01, use_graph=use_graph)
        e0 = (sim.epot + sim.ekin).clone()
        out = sim.run(300, record_every=50)
        etot = out["energies"].sum(1)                       # (records, B)
        drift = (etot - e0).abs().max().item()
        assert drift < 2e-4 * max(1.0, sim.ekin.abs().max().item() * 50), drift
        runs[use_graph] = (out["xyz"].clone(), etot.clone())
    assert torch.equal(runs[True][0], runs[False][0]) and torch.equal(runs[True][1], runs[False][1])


def test_active_learning_sampler_on_device(M):
    """al_utils.py:246-322, 1409-1455 on device tensors: main (E + gradients) and auxiliary (E only) models, uq =
    |E - E_aux|, threshold = median + 3 MAD of validation deviations; trajectories freeze at their first uncertain
    geometry; graph replay == eager stepping."""
    from mlatom_b200 import al, md

    nat, ntr = 9, 150
    eq = S.equilibrium(nat)
    xtr = S.geometries(eq, ntr, 1, spread=0.05)
    e, g = S.pair_potential(xtr, eq)
    db = D.molecular_database.from_arrays(S.atomic_numbers(nat), xtr, energy=e, energy_gradients=g)
    main = M.kreg(hyperparameters={"sigma": 1.0, "lambda": 1e-8})
    main.train(molecular_database=db, property_to_learn="energy", xyz_derivative_property_to_learn="energy_gradients",
               save_model=False)
    aux = M.kreg(hyperparameters={"sigma": 1.0, "lambda": 1e-8})
    aux.train(molecular_database=db, property_to_learn="energy", save_model=False)
    pm, pa = main.kreg_api._ensure_model(), aux.kreg_api._ensure_model()

    pair = al.UncertaintyPair(pm, pa)
    xval = torch.from_numpy(S.geometries(eq, 200, 9, spread=0.05)).cuda()
    out = pair.predict(xval)
    e_m, _ = pm.predict(xval)
    e_a, _ = pa.predict(xval, True, False)
    assert torch.equal(out["energy"], e_m) and torch.equal(out["uq"], (e_m - e_a).abs())
    thr = pair.calibrate(xval, "m+3mad")
    uq = out["uq"].cpu().numpy()
    assert abs(thr - (np.median(uq) + 3 * 1.4826 * np.median(np.abs(uq - np.median(uq))))) <= 1e-15
    assert al.threshold_metric(uq, "max") == uq.max() and al.threshold_metric(uq[:1], "m+2.5mad") == 0.0
    with pytest.raises(ValueError):
        al.threshold_metric(uq, "mean")
    assert pair.predict(xval)["uncertain"].dtype == torch.bool

    b = 96
    x0 = torch.from_numpy(S.geometries(eq, b, 5, spread=0.02)).cuda()
    v0 = torch.from_numpy(np.random.default_rng(6).standard_normal((b, nat, 3)) * 0.06).cuda()   # hot: leaves the data
    masses = torch.from_numpy(np.where(S.atomic_numbers(nat) == 1, 1.0, 12.0)).cuda()
    runs = {}
    for use_graph in (True, False):
        sim = md.BatchedNVE(pm, x0, v0, masses, dt=0.01, use_graph=use_graph, aux_model=pa, uq_threshold=thr)
        mid = sim.run(150)
        x_mid, stop_mid = mid["xyz"].clone(), mid["stop_step"].clone()
        fin = sim.run(150)
        runs[use_graph] = (fin["xyz"].clone(), fin["stop_step"].clone(), fin["uq"].clone())
        frozen = stop_mid >= 0
        assert frozen.any() and not frozen.all(), "the test needs both stopped and running trajectories"
        # frozen trajectories did not move during the second leg and keep their stop step
        assert torch.equal(fin["xyz"][frozen], x_mid[frozen]) and torch.equal(fin["stop_step"][frozen], stop_mid[frozen])
        assert (fin["uq"][fin["uncertain"]] > thr).all() and (fin["uq"][~fin["uncertain"]] <= thr).all()
        assert int(fin["stop_step"].max()) <= 300
    assert all(torch.equal(a, c) for a, c in zip(runs[True], runs[False]))


def test_batched_md_reproduces_the_reference_md_driver():
    """tests/golden/md/md_nve_nat9.npz is a 100-step NVE trajectory produced by the REFERENCE's own driver, mlatom.md
    (md.py:158-205), on a KREG model trained through this package's API (tools/make_golden_md.py).  BatchedNVE -- fused
    velocity-Verlet kernels + kreg_predict replayed as one CUDA graph -- must reproduce it: positions to 1e-9 Angstrom,
    energies to 1e-9 Hartree, for a single trajectory and for every member of a batch of identical ones."""
    from mlatom_b200 import engine, md

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "md", "md_nve_nat9.npz"))
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    c_acc, c_kin = md.mlatom_unit_factors()
    assert abs(c_acc - float(z["c_acc"])) <= 1e-15 * c_acc and abs(c_kin - float(z["c_kin"])) <= 1e-15 * c_kin
    pm = engine.PackedModel.from_training_set(t(z["xyz_train"]), engine.xeq(t(z["req"])), t(z["alpha"]), float(z["sigma"]),
                                              float(z["prior"]), int(z["ntrval"]), int(z["ntrgr"]))
    nsteps = int(z["nsteps"])
    for batch, use_graph in ((1, True), (1, False), (64, True)):
        x0 = t(np.repeat(z["x0"][None], batch, 0))
        v0 = t(np.repeat(z["v0"][None], batch, 0))
        sim = md.BatchedNVE(pm, x0, v0, t(z["masses"]), dt=float(z["dt"]), use_graph=use_graph, units="mlatom")
        assert abs(sim.epot[0].item() - z["epot"][0]) <= 1e-9 and abs(sim.ekin[0].item() - z["ekin"][0]) <= 1e-12
        out = sim.run(nsteps, record_every=1, record_xyz=True)
        traj = out["trajectory"].cpu().numpy()           # (nsteps, B, Nat, 3)
        en = out["energies"].cpu().numpy()               # (nsteps, 2, B)
        for k in range(batch if batch == 1 else 3):
            assert np.abs(traj[:, k] - z["xyz"][1:]).max() <= 1e-9, (batch, use_graph, np.abs(traj[:, k] - z["xyz"][1:]).max())
            assert np.abs(en[:, 0, k] - z["epot"][1:]).max() <= 1e-9 and np.abs(en[:, 1, k] - z["ekin"][1:]).max() <= 1e-9
        assert np.abs(out["velocities"][0].cpu().numpy() - z["vel"][-1]).max() <= 1e-9
        if batch > 1:
            assert (out["xyz"] - out["xyz"][0]).abs().max().item() <= 1e-12  # members of the batch stay together
