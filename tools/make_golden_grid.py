#!/usr/bin/env python3
"""Golden fixture for the hyper-parameter grid search.

Runs the REFERENCE's own optimiser -- mlatom.models.kreg.optimize_hyperparameters (mlatom/model_cls.py:609-727) with
its recursive optimize_grid (:819-854) -- over the reference's default ranges (lambda in [2^-35, 1], sigma in
[2^-5, 2^9], 9 x 9 log grid, mlatom/models.py:354-362).  Underneath, `mlatom.kreg_api` is replaced by this package's
KREG_API class (the binding INTEGRATION.md describes) whose device engine is the CPU oracle (tests/oracle_engine.py:
oracle/kreg_oracle.c arithmetic, LAPACK dpotrf/dpotrs with the dsytrf/dsytrs fallback of mathUtils.f90:24-26).  Every
(lambda, sigma) the reference visits and the loss it gets are recorded, in visiting order, together with the
condition number of K + lambda at that point (so a GPU test knows how many digits the loss can carry).

Only runs in the build container (needs /root/reference); writes tests/golden/grid_*.json.
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def reference_kreg_class(oracle):
    """mlatom.models.kreg of the reference checkout with `mlatom.kreg_api` bound to this package's KREG_API."""
    os.environ["MLATOM_NO_VERSION_CHECK"] = "1"
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    from oracle_engine import install_oracle_engine

    from mlatom_b200 import kreg_api as ours

    install_oracle_engine(setattr, oracle)
    shim = types.ModuleType("mlatom.kreg_api")
    shim.KREG_API = ours.KREG_API
    sys.modules["mlatom.kreg_api"] = shim
    import mlatom  # noqa: F401
    import mlatom.models as ref_models

    return ref_models, mlatom.data


def make_db(mod, nat, n, seed):
    from mlatom_b200 import synthetic as S

    eq = S.equilibrium(nat)
    z = S.atomic_numbers(nat)
    xyz = S.geometries(eq, n, seed)
    e, g = S.pair_potential(xyz, eq)
    db = mod.molecular_database()
    for i in range(n):
        mol = mod.molecule()
        for a in range(nat):
            mol.atoms.append(mod.atom(atomic_number=int(z[a]), xyz_coordinates=xyz[i, a]))
        mol.energy = float(e[i])
        mol.add_xyz_vectorial_property(g[i], "energy_gradients")
        db.molecules.append(mol)
    return db, xyz


def run_reference_grid(ref_models, ref_data, oracle, nat, nsub, nval, use_grads, tmpdir, grid_size=9):
    sub, xsub = make_db(ref_data, nat, nsub, 1)
    val, _ = make_db(ref_data, nat, nval, 3)
    visited = []

    class Recording(ref_models.kreg):
        def calculate_validation_loss(self, **kw):
            loss = super().calculate_validation_loss(**kw)
            visited.append((float(self.hyperparameters["lambda"].value), float(self.hyperparameters["sigma"].value),
                            float(loss)))
            return loss

    cwd = os.getcwd()
    os.chdir(tmpdir)  # the reference writes model files relative to the working directory
    try:
        model = Recording(model_file="grid_model")
        tk = {"property_to_learn": "energy"}
        pk = {"property_to_predict": "estimated_energy"}
        if use_grads:
            tk["xyz_derivative_property_to_learn"] = "energy_gradients"
            pk["xyz_derivative_property_to_predict"] = "estimated_energy_gradients"
        model.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="grid",
                                       optimization_algorithm_kwargs={"grid_size": grid_size}, training_kwargs=tk,
                                       prediction_kwargs=pk, subtraining_molecular_database=sub,
                                       validation_molecular_database=val)
    finally:
        os.chdir(cwd)
    # conditioning of K + lambda at every visited point (symmetric eigenvalues of the oracle's matrix)
    from mlatom_b200 import synthetic as S

    eq_idx = int(np.argmin([m.energy for m in sub.molecules]))
    xeq = oracle.xeq(xsub[eq_idx])
    X = oracle.descriptors(xsub, xeq)
    nv, ng = nsub, (3 * nat * nsub if use_grads else 0)
    conds, kcache = [], {}
    for lam, sigma, _ in visited:
        if sigma not in kcache:
            kcache[sigma] = np.linalg.eigvalsh(oracle.kernel_matrix(xsub, X, sigma, nv, ng, factored=True))
        w = kcache[sigma] + lam
        conds.append(float(np.abs(w).max() / max(np.abs(w).min(), 1e-300)) if w.min() > 0 else float("inf"))
    return {
        "what": "reference optimize_hyperparameters('grid') over its default ranges; backend = mlatom_b200.KREG_API on the CPU oracle",
        "natoms": nat, "nsub": nsub, "nval": nval, "use_gradients": bool(use_grads), "grid_size": grid_size,
        "train_seed": 1, "validation_seed": 3,
        "visited": [{"lambda": l, "sigma": s, "loss": v, "cond": c} for (l, s, v), c in zip(visited, conds)],
        "winner": {"lambda": float(model.hyperparameters["lambda"].value), "sigma": float(model.hyperparameters["sigma"].value),
                   "validation_loss": float(model.validation_loss)},
    }


def main():
    import tempfile

    from oracle import kreg_oracle as O

    O.load()
    ref_models, ref_data = reference_kreg_class(O)
    out_dir = os.path.join(ROOT, "tests", "golden")
    with tempfile.TemporaryDirectory() as tmp:
        for name, nat, nsub, nval, grads in (("grid_values_nat9", 9, 60, 40, False), ("grid_valgrad_nat9", 9, 24, 16, True)):
            rec = run_reference_grid(ref_models, ref_data, O, nat, nsub, nval, grads, tmp)
            with open(os.path.join(out_dir, name + ".json"), "w") as f:
                json.dump(rec, f, indent=1)
            print(name, "visited", len(rec["visited"]), "winner", rec["winner"])


if __name__ == "__main__":
    main()
