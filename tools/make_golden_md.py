#!/usr/bin/env python3
"""Golden NVE trajectory from the REFERENCE's own MD driver.

mlatom.md (mlatom/md.py:158-205, velocity Verlet, NVE) is run in the build container on a 9-atom KREG model trained
through this package's API on the CPU oracle (tests/oracle_engine.py); the model (training geometries, alpha, sigma,
prior), the initial conditions, the masses the reference used and every step's positions, velocities and energies go to
tests/golden/md/md_nve_nat9.npz.  The GPU test rebuilds the model from the fixture's alpha and requires
mlatom_b200.md.BatchedNVE (fused velocity-Verlet kernels + kreg_predict in a CUDA graph) to reproduce the positions.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    os.environ["MLATOM_NO_VERSION_CHECK"] = "1"
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, "/root/reference")
    import mlatom as ml
    from oracle_engine import install_oracle_engine

    from mlatom_b200 import models as M, synthetic as S
    from oracle import kreg_oracle as O

    O.load()
    install_oracle_engine(setattr, O)
    nat, ntr, nsteps, dt = 9, 60, 100, 0.25
    eq = S.equilibrium(nat)
    z = S.atomic_numbers(nat)
    xyz = S.geometries(eq, ntr, 1, spread=0.08)
    e, g = S.pair_potential(xyz, eq)
    db = ml.data.molecular_database()
    for i in range(ntr):
        mol = ml.data.molecule()
        for a in range(nat):
            mol.atoms.append(ml.data.atom(atomic_number=int(z[a]), xyz_coordinates=xyz[i, a]))
        mol.energy = float(e[i])
        mol.add_xyz_vectorial_property(g[i], "energy_gradients")
        db.molecules.append(mol)
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        model = M.kreg(model_file=os.path.join(tmp, "md_model"), hyperparameters={"sigma": 1.0, "lambda": 1e-8})
        model.train(molecular_database=db, property_to_learn="energy", xyz_derivative_property_to_learn="energy_gradients")
        api = model.kreg_api
        x0 = S.geometries(eq, 1, 77, spread=0.02)[0]
        v0 = np.random.default_rng(78).standard_normal((nat, 3)) * 0.003  # Angstrom / fs
        init = ml.data.molecule()  # a fresh molecule: no stale labels, so md.py:166-169 predicts the forces at t = 0
        for a in range(nat):
            init.atoms.append(ml.data.atom(atomic_number=int(z[a]), xyz_coordinates=x0[a].copy()))
            init.atoms[a].xyz_velocities = v0[a].copy()
        dyn = ml.md(model=model, molecule_with_initial_conditions=init, ensemble="NVE", time_step=dt,
                    maximum_propagation_time=nsteps * dt)
        steps = dyn.molecular_trajectory.steps
        traj = np.array([s.molecule.xyz_coordinates for s in steps])
        vel = np.array([s.molecule.get_xyz_vectorial_properties("xyz_velocities") for s in steps])
        epot = np.array([s.molecule.energy for s in steps])
        ekin = np.array([s.molecule.kinetic_energy for s in steps])
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", "md", "md_nve_nat9.npz"),
            natoms=nat, dt=dt, nsteps=len(steps) - 1, masses=np.asarray(init.get_nuclear_masses(), dtype=np.float64),
            xyz_train=np.asarray(api.XYZ), req=np.asarray(api.Req), alpha=np.asarray(api.alpha).reshape(-1),
            sigma=float(api.sigma), prior=float(api.prior), ntrval=int(api.NtrVal), ntrgr=int(api.NtrGrXYZ),
            x0=x0, v0=v0, xyz=traj, vel=vel, epot=epot, ekin=ekin,
            c_acc=1.0 / ml.constants.ram2au * ml.constants.Bohr2Angstrom ** 2 * ml.constants.fs2au ** 2,
            c_kin=ml.constants.ram2au * (ml.constants.au2fs / ml.constants.Bohr2Angstrom) ** 2)
        print("steps", len(steps), "E_tot drift", np.abs((epot + ekin) - (epot + ekin)[0]).max(),
              "max displacement", np.abs(traj - traj[0]).max())


if __name__ == "__main__":
    main()
