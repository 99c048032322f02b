"""CPU: host-side logic of the drop-in API (Ytrain ordering, prior handling, .npz layout, property
write-back, grid visiting order) with the device engine replaced by oracle-backed fakes -- and, when the
reference checkout is present, driven by REAL mlatom.data objects to prove duck-type compatibility."""
import os
import sys
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from mlatom_b200 import data as D, synthetic as S  # noqa: E402
from mlatom_b200 import kreg_api as KA, models as M  # noqa: E402


from oracle_engine import FakePacked, install_oracle_engine  # noqa: E402,F401


@pytest.fixture
def fake_engine(monkeypatch, oracle):
    return install_oracle_engine(monkeypatch.setattr, oracle)


def make_db(mod, nat, n, seed, with_labels=True):
    eq = S.equilibrium(nat)
    z = S.atomic_numbers(nat)
    xyz = S.geometries(eq, n, seed)
    e, g = S.pair_potential(xyz, eq)
    db = mod.molecular_database()
    for i in range(n):
        mol = mod.molecule()
        for a in range(nat):
            mol.atoms.append(mod.atom(atomic_number=int(z[a]), xyz_coordinates=xyz[i, a]))
        if with_labels:
            mol.energy = float(e[i])
            mol.add_xyz_vectorial_property(g[i], "energy_gradients")
        db.molecules.append(mol)
    return db, xyz, e, g


def reference_mlatom():
    from conftest import reference_mlatom_data

    return reference_mlatom_data()


@pytest.mark.parametrize("use_grads", [False, True])
def test_train_predict_host_logic(fake_engine, oracle, tmp_path, use_grads):
    nat, ntr, nte = 5, 14, 6
    db, xyz, e, g = make_db(D, nat, ntr, 1)
    test_db, xte, _, _ = make_db(D, nat, nte, 2, with_labels=False)
    model = M.kreg(model_file=str(tmp_path / "m"), hyperparameters={"sigma": 0.9, "lambda": 1e-6})
    model.train(molecular_database=db, property_to_learn="energy",
                xyz_derivative_property_to_learn="energy_gradients" if use_grads else None, prior="mean")
    assert model.model_file.endswith(".npz") and os.path.exists(model.model_file)
    api = model.kreg_api
    # equilibrium molecule = lowest energy (models.py:447-457); Ytrain order (kreg_api.py:51-72)
    ieq = int(np.argmin(e))
    assert np.array_equal(api.Req, xyz[ieq])
    expect_y = np.concatenate([e - e.mean()] + ([g.reshape(-1)] if use_grads else []))
    assert np.allclose(api.Ytrain, expect_y, rtol=0, atol=0)
    assert api.NtrVal == ntr and api.NtrGrXYZ == (3 * nat * ntr if use_grads else 0)
    # model-file layout
    z = np.load(model.model_file)
    assert set(z.files) == {"sigma", "Natoms", "Xsize", "Ntrain", "NtrVal", "NtrGrXYZ", "lambdav", "lambdagradxyz", "X",
                            "XYZ", "ac2dArray", "Req", "Xeq", "alpha", "prior"}
    assert z["alpha"].shape == (1, api.Ksize) and z["X"].shape == (ntr, nat * (nat - 1) // 2)
    assert z["XYZ"].shape == (ntr, nat, 3) and z["ac2dArray"].dtype == np.int32
    assert np.array_equal(z["ac2dArray"], oracle.ac2d(nat))
    # prediction is written into the molecules; gradients (not forces) per atom
    model.predict(molecular_database=test_db, calculate_energy=True, calculate_energy_gradients=True)
    xeq = oracle.xeq(xyz[ieq])
    X, Xte = oracle.descriptors(xyz, xeq), oracle.descriptors(xte, xeq)
    a = z["alpha"].reshape(-1)
    v = oracle.precompute_v(xyz, X, a[ntr:]) if use_grads else None
    e_ref, g_ref = oracle.predict_factored(X, a[:ntr], v, 0.9, float(z["prior"]), xte, Xte)
    assert np.allclose(test_db.get_properties("energy"), e_ref, rtol=1e-13)
    assert np.allclose(test_db.get_xyz_vectorial_properties("energy_gradients"), g_ref, atol=1e-12)
    # a fresh object auto-loads the file (models.py:370-375) and predicts a single molecule
    again = M.kreg(model_file=model.model_file)
    mol = test_db.molecules[0].copy()
    again.predict(molecule=mol, calculate_energy=True, calculate_energy_gradients=True)
    assert abs(mol.energy - e_ref[0]) <= 1e-12 * abs(e_ref[0])
    assert np.allclose(mol.get_energy_gradients(), g_ref[0], atol=1e-12)
    # custom property names
    again.predict(molecular_database=test_db, property_to_predict="e_ml", xyz_derivative_property_to_predict="g_ml")
    assert np.allclose(test_db.get_properties("e_ml"), e_ref, rtol=1e-13)


def test_real_mlatom_objects_are_accepted(fake_engine, tmp_path):
    ref_data = reference_mlatom()
    if ref_data is None:
        pytest.skip("reference checkout not available")
    nat, ntr, nte = 5, 10, 4
    db_ref, xyz, e, g = make_db(ref_data, nat, ntr, 1)
    db_own, _, _, _ = make_db(D, nat, ntr, 1)
    t_ref, xte, _, _ = make_db(ref_data, nat, nte, 2, with_labels=False)
    t_own, _, _, _ = make_db(D, nat, nte, 2, with_labels=False)
    out = []
    for db, tdb, name in ((db_ref, t_ref, "ref"), (db_own, t_own, "own")):
        model = M.kreg(model_file=str(tmp_path / name), hyperparameters={"sigma": 1.1, "lambda": 1e-5})
        model.train(molecular_database=db, property_to_learn="energy", xyz_derivative_property_to_learn="energy_gradients")
        model.predict(molecular_database=tdb, calculate_energy=True, calculate_energy_gradients=True)
        out.append((tdb.get_properties("energy"), tdb.get_xyz_vectorial_properties("energy_gradients")))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    # per-atom write-back as the reference does it (kreg_api.py:126-131)
    assert t_ref.molecules[0].atoms[0].energy_gradients.shape == (3,)


def test_hyperparameter_objects():
    m = M.kreg()
    assert m.hyperparameters["lambda"].value == 2 ** -35 and m.hyperparameters["sigma"].value == 1.0
    assert m.hyperparameters["sigma"].minval == 2 ** -5 and m.hyperparameters["sigma"].maxval == 2 ** 9
    assert m.hyperparameters["lambda"].optimization_space == "log"
    m.hyperparameters["sigma"] = 3
    assert m.hyperparameters["sigma"].value == 3.0 and isinstance(m.hyperparameters["sigma"].value, float)
    assert m.hyperparameters.sigma == 3.0
    assert M.kreg().hyperparameters["sigma"].value == 1.0  # class default untouched
    m2 = M.kreg(hyperparameters={"sigma": 2.0, "lambdaGradXYZ": 1e-3})
    assert m2.hyperparameters["lambdaGradXYZ"].value == 1e-3
    with pytest.raises(ValueError):
        M.kreg(ml_program="PyTorch")


def test_ml_program_mlatomf_uses_unf_model_files(fake_engine, tmp_path):
    """models.py:376-378, 478-496, 542-546: with ml_program='MLatomF' the model file is MLatomF's .unf (MLmodel.f90
    layout); training, reloading and predicting give what the .npz flavour gives."""
    db, xyz, e, g = make_db(D, 5, 12, 1)
    test_db, xte, _, _ = make_db(D, 5, 5, 2, with_labels=False)
    out = {}
    for prog in ("KREG_API", "MLatomF"):
        path = str(tmp_path / ("m_" + prog + (".unf" if prog == "MLatomF" else "")))
        model = M.kreg(model_file=path, ml_program=prog, hyperparameters={"sigma": 0.9, "lambda": 1e-6})
        model.train(molecular_database=db, property_to_learn="energy", xyz_derivative_property_to_learn="energy_gradients",
                    prior="mean")
        assert os.path.exists(model.model_file) and model.model_file.endswith(".unf" if prog == "MLatomF" else ".npz")
        again = M.kreg(model_file=model.model_file, ml_program=prog)
        tdb = test_db if prog == "KREG_API" else make_db(D, 5, 5, 2, with_labels=False)[0]
        again.predict(molecular_database=tdb, calculate_energy=True, calculate_energy_gradients=True)
        out[prog] = (tdb.get_properties("energy"), tdb.get_xyz_vectorial_properties("energy_gradients"))
    assert np.allclose(out["KREG_API"][0], out["MLatomF"][0], rtol=1e-13, atol=0)
    assert np.allclose(out["KREG_API"][1], out["MLatomF"][1], rtol=1e-12, atol=1e-13)
    auto = M.kreg(ml_program="MLatomF", hyperparameters={"sigma": 0.9, "lambda": 1e-6})
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        auto.train(molecular_database=db, property_to_learn="energy")
        assert auto.model_file.startswith("kreg_") and auto.model_file.endswith(".unf") and os.path.exists(auto.model_file)
    finally:
        os.chdir(cwd)


def test_optimize_grid_visiting_order_and_ties():
    seen = []

    def f(p):
        seen.append(tuple(p))
        return 1.0 if p[1] != 2 else 0.5  # tie between (1,2) and (3,2)

    params, val = M.optimize_grid(f, [[1, 3], [5, 2, 7]])
    assert seen == [(1, 5), (1, 2), (1, 7), (3, 5), (3, 2), (3, 7)]
    assert params == [1, 2] and val == 0.5


def test_grid_search_fused_equals_pointwise(fake_engine, tmp_path):
    nat = 5
    sub, _, _, _ = make_db(D, nat, 12, 1)
    val, _, _, _ = make_db(D, nat, 7, 3)
    res = []
    for fused in (True, False):
        model = M.kreg(model_file=str(tmp_path / f"g{fused}"))
        model.hyperparameters["lambda"].minval, model.hyperparameters["lambda"].maxval = 1e-8, 1e-2
        model.hyperparameters["sigma"].minval, model.hyperparameters["sigma"].maxval = 0.3, 8.0
        model.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="grid",
                                       optimization_algorithm_kwargs={"grid_size": 3},
                                       training_kwargs={"property_to_learn": "energy", "prior": "mean"},
                                       prediction_kwargs={"property_to_predict": "estimated_energy"},
                                       subtraining_molecular_database=sub, validation_molecular_database=val, fused=fused)
        res.append((model.hyperparameters["lambda"].value, model.hyperparameters["sigma"].value, model.validation_loss))
    assert res[0][:2] == res[1][:2]
    assert abs(res[0][2] - res[1][2]) <= 1e-9 * res[1][2]


@pytest.mark.parametrize("kernel", ["Matern", "exponential", "Laplacian"])
def test_grid_search_on_the_other_kernels_fused_equals_pointwise(fake_engine, tmp_path, kernel):
    """Values-trained models on MLatomF's other kernel functions: the fused grid (K(sigma) once per sigma through
    kernel_block_ex, lambda on the diagonal of a copy) visits the points and returns the winner and the loss of the
    pointwise route, and the log2 grid runs on them too."""
    nat = 5
    sub, _, _, _ = make_db(D, nat, 12, 1)
    val, _, _, _ = make_db(D, nat, 7, 3)
    kw = dict(training_kwargs={"property_to_learn": "energy", "prior": "mean"},
              prediction_kwargs={"property_to_predict": "estimated_energy"},
              subtraining_molecular_database=sub, validation_molecular_database=val)
    res = []
    for fused in (True, False):
        model = M.kreg(model_file=str(tmp_path / f"g{fused}"), ml_program="MLatomF")
        model.additional_mlatomf_args = [f"kernel={kernel}", "nn=1"]
        model.hyperparameters["lambda"].minval, model.hyperparameters["lambda"].maxval = 1e-8, 1e-2
        model.hyperparameters["sigma"].minval, model.hyperparameters["sigma"].maxval = 0.3, 8.0
        model.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="grid",
                                       optimization_algorithm_kwargs={"grid_size": 3}, fused=fused, **kw)
        assert type(model.kreg_api).__name__ == "RadialKREG" and model.kreg_api.nn == 1
        res.append((model.hyperparameters["lambda"].value, model.hyperparameters["sigma"].value, model.validation_loss))
        if fused:
            assert len(model.grid_losses) == 9
            assert abs(min(model.grid_losses.values()) - model.validation_loss) <= 1e-9 * model.validation_loss
    assert res[0][:2] == res[1][:2]
    assert abs(res[0][2] - res[1][2]) <= 1e-9 * res[1][2]
    model = M.kreg(model_file=str(tmp_path / "lg"), ml_program="MLatomF")
    model.additional_mlatomf_args = [f"kernel={kernel}", "nn=1"]
    model.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="log2grid",
                                   optimization_algorithm_kwargs={"depth": 2, "points": {"sigma": 5, "lambda": 3}}, **kw)
    best = min(t[2] for t in model.grid_trace)
    assert len(model.grid_trace) == 2 * 5 * 2 * 3 and abs(model.validation_loss - best) <= 1e-9 * best
    with pytest.raises(NotImplementedError):  # gradient-trained models on these kernels are not built: fused or not
        model.optimize_hyperparameters(
            hyperparameters=["sigma"], optimization_algorithm="grid", optimization_algorithm_kwargs={"grid_size": 2},
            training_kwargs={"property_to_learn": "energy", "xyz_derivative_property_to_learn": "energy_gradients"},
            prediction_kwargs={"property_to_predict": "estimated_energy"},
            subtraining_molecular_database=sub, validation_molecular_database=val)


@pytest.mark.parametrize("kernel", ["Matern", "Laplacian"])
def test_other_kernel_models_through_the_class(fake_engine, oracle, tmp_path, kernel):
    """kreg(ml_program='MLatomF') + additional_mlatomf_args=['kernel=...'] (mlatom/models.py:494-495): train on values,
    predict energies and gradients into the molecules, reload from the .npz file (KREG_API keys + kernel, nn)."""
    nat = 5
    db, xyz, e, _ = make_db(D, nat, 16, 1)
    model = M.kreg(model_file=str(tmp_path / "m"), ml_program="MLatomF", hyperparameters={"sigma": 2.0, "lambda": 1e-6})
    model.additional_mlatomf_args = [f"kernel={kernel}", "nn=1"]
    model.train(molecular_database=db, property_to_learn="energy", prior="mean")
    assert type(model.kreg_api).__name__ == "RadialKREG" and model.model_file.endswith(".npz") and os.path.exists(model.model_file)
    tdb, xte, _, _ = make_db(D, nat, 6, 2, with_labels=False)
    model.predict(molecular_database=tdb, calculate_energy=True, calculate_energy_gradients=True)
    xeq = oracle.xeq(np.asarray(model.kreg_api.Req))  # no equilibrium molecule given: the lowest-energy training geometry
    a = np.asarray(model.kreg_api.alpha).reshape(-1)
    e_ref, g_ref = oracle.predict_generic(kernel.lower(), oracle.descriptors(xyz, xeq), a, 2.0, float(e.mean()), xte,
                                          oracle.descriptors(xte, xeq), nn=1)
    assert np.abs(tdb.get_properties("energy") - e_ref).max() <= 1e-12 * np.abs(e_ref).max()
    assert np.abs(tdb.get_xyz_vectorial_properties("energy_gradients") - g_ref).max() <= 1e-12 * max(1.0, np.abs(g_ref).max())
    again = M.kreg(model_file=model.model_file, ml_program="MLatomF")
    assert type(again.kreg_api).__name__ == "RadialKREG" and again.kreg_api.kernel == kernel and again.kreg_api.nn == 1
    e2, g2 = again.predict_xyz(xte)
    assert np.array_equal(e2, tdb.get_properties("energy"))


def test_log2_grid_search_follows_mlatomf_recursion(fake_engine, tmp_path):
    """optLog2Grid / optLambdaLog2Grid (A_KRR.f90:1049-1376): nested log2 grids, zoom +-1 step around the best,
    lgOptDepth levels, first minimum wins -- restated here with one train+predict per point."""
    nat = 5
    sub, _, _, _ = make_db(D, nat, 12, 1)
    val, _, _, _ = make_db(D, nat, 7, 3)
    tk = {"property_to_learn": "energy", "prior": "mean"}
    pk = {"property_to_predict": "estimated_energy"}
    lam_rng, sig_rng, nl, ns, depth = (2.0 ** -30, 2.0 ** -6), (0.5, 8.0), 3, 4, 2

    model = M.kreg(model_file=str(tmp_path / "lg"))
    model.hyperparameters["lambda"].minval, model.hyperparameters["lambda"].maxval = lam_rng
    model.hyperparameters["sigma"].minval, model.hyperparameters["sigma"].maxval = sig_rng
    model.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="log2grid",
                                   optimization_algorithm_kwargs={"depth": depth, "points": {"lambda": nl, "sigma": ns}},
                                   training_kwargs=tk, prediction_kwargs=pk,
                                   subtraining_molecular_database=sub, validation_molecular_database=val)
    got = (model.hyperparameters["lambda"].value, model.hyperparameters["sigma"].value, model.validation_loss)
    assert len(model.grid_trace) > 0

    ref = M.kreg(model_file=str(tmp_path / "lgref"))

    def loss(lam, sigma):
        ref.hyperparameters["lambda"].value, ref.hyperparameters["sigma"].value = lam, sigma
        return ref.calculate_validation_loss(training_kwargs=dict(tk), prediction_kwargs=dict(pk),
                                             subtraining_molecular_database=sub, validation_molecular_database=val)

    def zoom(lo, hi, n, f):
        best = None
        for _ in range(depth):
            step = 0.0 if n == 1 else (hi - lo) / (n - 1)
            lvl = None
            for i in range(n):
                x = lo + step * i
                r = f(x)
                if lvl is None or r[0] < lvl[1][0]:
                    lvl = (x, r)
            best = lvl
            lo, hi = best[0] - step, best[0] + step
        return best

    def at_sigma(lg_sigma):
        x, r = zoom(np.log2(lam_rng[0]), np.log2(lam_rng[1]), nl, lambda lg_lam: (loss(2.0 ** lg_lam, 2.0 ** lg_sigma),))
        return (r[0], 2.0 ** x)

    xs, (err, lam) = zoom(np.log2(sig_rng[0]), np.log2(sig_rng[1]), ns, at_sigma)
    assert got[0] == lam and got[1] == 2.0 ** xs
    assert abs(got[2] - err) <= 1e-9 * err


def test_get_descriptor_like_the_reference(fake_engine):
    """models.py:402-437: RE descriptor stored on every molecule; the reference distance matrix is cached."""
    db, xyz, e, _ = make_db(D, 5, 6, 1)
    model = M.kreg()
    model.get_descriptor(molecular_database=db)
    eq = xyz[int(np.argmin(e))]
    pairs = np.triu_indices(5, 1)
    dist = lambda m: np.linalg.norm(m[:, None] - m[None], axis=-1)[pairs]
    for mol, m in zip(db.molecules, xyz):
        assert np.allclose(mol.re, dist(eq) / dist(m), rtol=1e-14) and mol.descriptor is mol.re
    assert np.allclose(model.reference_distance_matrix, np.linalg.norm(eq[:, None] - eq[None], axis=-1))
    one = D.molecule.from_arrays(S.atomic_numbers(5), xyz[2])
    model.get_descriptor(molecule=one, descriptor_name="x")     # cached reference geometry, no energies needed
    assert np.array_equal(one.x, db.molecules[2].re)
    with pytest.raises(ValueError):
        model.get_descriptor()
    ref_data = reference_mlatom()
    if ref_data is not None:  # the reference class itself on the reference's data objects
        import mlatom.models as RM

        rdb, _, _, _ = make_db(ref_data, 5, 6, 1)
        rmodel = RM.kreg.__new__(RM.kreg)
        RM.kreg.get_descriptor(rmodel, molecular_database=rdb)
        for a, b in zip(rdb.molecules, db.molecules):
            assert np.allclose(a.re, b.re, rtol=1e-14, atol=0)


def test_al_threshold_metric_matches_reference_formula():
    """al_utils.py:1436-1455 + stats.py:122-128 ('max' and 'm+<n>mad' with MAD = 1.4826 median|v - median v|)."""
    from mlatom_b200 import al

    v = np.abs(np.random.default_rng(3).standard_normal(101))
    med = np.median(v)
    mad = 1.4826 * np.median(np.abs(v - med))
    assert al.threshold_metric(v, "max") == v.max()
    assert abs(al.threshold_metric(v, "m+3mad") - (med + 3 * mad)) <= 1e-15
    assert abs(al.threshold_metric(list(v), "M+2.5MAD") - (med + 2.5 * mad)) <= 1e-15
    assert al.threshold_metric(v[:1], "m+3mad") == 0.0
    with pytest.raises(ValueError):
        al.threshold_metric(v, "m+mad")
    if reference_mlatom() is not None:  # (stubs h5py, puts /root/reference on sys.path)
        from mlatom import stats as ref_stats

        assert abs(al.median_absolute_deviation(v) - ref_stats.calc_median_absolute_deviation(v)) <= 1e-16


def test_reference_drivers_run_on_this_model(fake_engine, tmp_path):
    """Drop-in at the driver level: the REFERENCE's own geometry optimiser and MD driver (mlatom/simulations.py,
    md.py:158-205) are handed this package's kreg model and reference data objects; they call model.predict(molecule=...)
    per step and read mol.energy / atom.energy_gradients back (build container only: needs /root/reference)."""
    ref_data = reference_mlatom()
    if ref_data is None:
        pytest.skip("reference checkout not available")
    import mlatom as ml

    nat = 5
    db, xyz, e, g = make_db(ref_data, nat, 60, 1)
    model = M.kreg(model_file=str(tmp_path / "drv"), hyperparameters={"sigma": 1.0, "lambda": 1e-8})
    model.train(molecular_database=db, property_to_learn="energy", xyz_derivative_property_to_learn="energy_gradients")

    start = db.molecules[3].copy()
    opt = ml.optimize_geometry(model=model, initial_molecule=start, program="scipy")
    final = opt.optimized_molecule
    assert final.energy <= start.energy + 1e-12 if hasattr(start, "energy") else True
    gmax = np.abs(final.get_energy_gradients()).max()
    assert gmax < 5e-3, gmax                                    # converged on the model surface

    init = db.molecules[5].copy()
    for atom in init.atoms:
        atom.xyz_velocities = np.zeros(3)
    dyn = ml.md(model=model, molecule_with_initial_conditions=init, ensemble="NVE", time_step=0.2,
                maximum_propagation_time=4.0)
    steps = dyn.molecular_trajectory.steps
    assert len(steps) >= 20
    etot = np.array([s.molecule.total_energy for s in steps])
    assert np.abs(etot - etot[0]).max() < 1e-4 * max(1.0, abs(etot[0]))   # velocity Verlet conserves E on a smooth surface

    # batch driver (md_parallel.py:155-200): one predict(molecular_database=...) per step for all trajectories
    batch = ref_data.molecular_database()
    for k in (1, 2, 7, 9):
        mol = db.molecules[k].copy()
        for atom in mol.atoms:
            atom.xyz_velocities = np.zeros(3)
        batch.molecules.append(mol)
    par = ml.md_parallel(model=model, molecular_database=batch, ensemble="NVE", time_step=0.2, maximum_propagation_time=1.0)
    assert len(par.molecular_trajectory.steps) >= 5 and len(par.molecular_trajectory.steps[0]) == 4

    # Hessian consumer: predict(calculate_hessian=True) writes mol.hessian (3Nat x 3Nat) like kreg_api.py:130-131
    hmol = final.copy()
    model.predict(molecule=hmol, calculate_energy=True, calculate_energy_gradients=True, calculate_hessian=True)
    assert np.asarray(hmol.hessian).shape == (3 * nat, 3 * nat)
    assert np.abs(hmol.hessian - hmol.hessian.T).max() < 1e-10 * max(1.0, np.abs(hmol.hessian).max())


def test_active_learning_trainer_call_sequence(fake_engine, tmp_path):
    """The exact calls of al_utils.py:1086-1126 (main and auxiliary KREG trainers): widen sigma's range, grid-optimise
    lambda and sigma (default 9 x 9 grid, prior='mean', the auxiliary one with its odd 'estimated'+name property), then
    save through model.kreg_api; a second construction on the same file finds the model and skips training."""
    mod = reference_mlatom() or D
    sub, _, _, _ = make_db(mod, 5, 14, 1)
    val, _, _, _ = make_db(mod, 5, 8, 3)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        main = M.kreg(model_file="main.npz", ml_program="KREG_API")
        main.hyperparameters["sigma"].minval = 2 ** -5
        main.optimize_hyperparameters(subtraining_molecular_database=sub, validation_molecular_database=val,
                                      optimization_algorithm="grid", hyperparameters=["lambda", "sigma"],
                                      training_kwargs={"property_to_learn": "energy",
                                                       "xyz_derivative_property_to_learn": "energy_gradients", "prior": "mean"},
                                      prediction_kwargs={"property_to_predict": "estimated_energy",
                                                         "xyz_derivative_property_to_predict": "estimated_energy_gradients"},
                                      validation_loss_function=None)
        assert 2 ** -35 <= main.hyperparameters["lambda"].value <= 1.0 and 2 ** -5 <= main.hyperparameters["sigma"].value <= 2 ** 9
        main.kreg_api.save_model("main.npz")
        aux = M.kreg(model_file="aux.npz", ml_program="KREG_API")
        aux.hyperparameters["sigma"].minval = 2 ** -5
        aux.optimize_hyperparameters(subtraining_molecular_database=sub, validation_molecular_database=val,
                                     optimization_algorithm="grid", hyperparameters=["lambda", "sigma"],
                                     training_kwargs={"property_to_learn": "energy", "prior": "mean"},
                                     prediction_kwargs={"property_to_predict": "estimated" + "energy"})
        aux.kreg_api.save_model("aux.npz")
        assert os.path.exists("main.npz") and os.path.exists("aux.npz")
        # al_utils.py:1088-1090: the file exists -> the model is loaded, not retrained
        again = M.kreg(model_file="main.npz", ml_program="KREG_API")
        assert again.kreg_api.NtrGrXYZ == 3 * 5 * 14 and float(again.kreg_api.sigma) == main.hyperparameters["sigma"].value
        # al_utils.py:1409-1424: uncertainty = |E_main - E_aux| written into the molecules
        pool, _, _, _ = make_db(mod, 5, 6, 4, with_labels=False)
        again.predict(molecular_database=pool, property_to_predict="energy", xyz_derivative_property_to_predict="energy_gradients")
        M.kreg(model_file="aux.npz").predict(molecular_database=pool, property_to_predict="aux_energy",
                                             xyz_derivative_property_to_predict="aux_energy_gradients")
        uq = [abs(m.energy - m.aux_energy) for m in pool.molecules]
        assert len(uq) == 6 and all(np.isfinite(uq))
    finally:
        os.chdir(cwd)


def test_cross_validation_loss_matches_the_reference_class(fake_engine, oracle, tmp_path):
    """k-fold cross-validation (model_cls.py:742-770 and the CV branch of calculate_validation_loss :529-575): this
    package's kreg against the REFERENCE's own kreg class driving the same backend, on the reference's data objects."""
    ref_data = reference_mlatom()
    if ref_data is None:
        pytest.skip("reference checkout not available")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import make_golden_grid as G

    saved = sys.modules.get("mlatom.kreg_api")
    try:
        ref_models, _ = G.reference_kreg_class(oracle)  # binds mlatom.kreg_api.KREG_API to this package's class
        splits_a = [make_db(ref_data, 5, 9, 20 + k)[0] for k in range(3)]
        splits_b = [make_db(ref_data, 5, 9, 20 + k)[0] for k in range(3)]
        tk = {"property_to_learn": "energy", "xyz_derivative_property_to_learn": "energy_gradients"}
        pk = {"property_to_predict": "estimated_energy", "xyz_derivative_property_to_predict": "estimated_energy_gradients"}
        cwd = os.getcwd()
        os.chdir(tmp_path)
        try:
            ref = ref_models.kreg(model_file="cv_ref", hyperparameters={"sigma": 1.3, "lambda": 1e-7})
            ref.equilibrium_molecule = splits_a[0].molecules[0]
            want, want_splits = ref.calculate_validation_loss(training_kwargs=dict(tk), prediction_kwargs=dict(pk),
                                                              cv_splits_molecular_databases=splits_a,
                                                              calculate_CV_split_errors=True)
        finally:
            os.chdir(cwd)
        ours = M.kreg(model_file=str(tmp_path / "cv_ours"), hyperparameters={"sigma": 1.3, "lambda": 1e-7})
        ours.equilibrium_molecule = splits_b[0].molecules[0]
        got, got_splits = ours.calculate_validation_loss(training_kwargs=dict(tk), prediction_kwargs=dict(pk),
                                                         cv_splits_molecular_databases=splits_b,
                                                         calculate_CV_split_errors=True)
        assert abs(got - want) <= 1e-12 * want
        assert np.allclose(got_splits, want_splits, rtol=1e-12, atol=0)
        # and through the optimiser: a 3 x 3 grid scored by cross-validation picks the same point
        for m_, cls_splits in ((ref, splits_a), (ours, splits_b)):
            m_.hyperparameters["lambda"].minval, m_.hyperparameters["lambda"].maxval = 1e-9, 1e-3
            m_.hyperparameters["sigma"].minval, m_.hyperparameters["sigma"].maxval = 0.5, 8.0
        os.chdir(tmp_path)
        try:
            ref.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="grid",
                                         optimization_algorithm_kwargs={"grid_size": 3}, training_kwargs=dict(tk),
                                         prediction_kwargs=dict(pk), cv_splits_molecular_databases=splits_a)
        finally:
            os.chdir(cwd)
        ours.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="grid",
                                      optimization_algorithm_kwargs={"grid_size": 3}, training_kwargs=dict(tk),
                                      prediction_kwargs=dict(pk), cv_splits_molecular_databases=splits_b)
        assert ours.hyperparameters["lambda"].value == ref.hyperparameters["lambda"].value
        assert ours.hyperparameters["sigma"].value == ref.hyperparameters["sigma"].value
        assert abs(ours.validation_loss - ref.validation_loss) <= 1e-12 * ref.validation_loss
    finally:
        if saved is None:
            sys.modules.pop("mlatom.kreg_api", None)
        else:
            sys.modules["mlatom.kreg_api"] = saved
