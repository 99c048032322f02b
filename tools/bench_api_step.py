#!/usr/bin/env python3
"""What an MD / geometry-optimisation driver pays per step: models.kreg.predict(molecule=...) on ONE molecule object
(mlatom/md.py:169,198), energies + gradients written back into the molecule, 9 atoms, Ntrain = 1000 and 10000."""
import cProfile, json, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mlatom_b200 import data as D, models as M, synthetic as S

nat = 9
eq = S.equilibrium(nat)
out = {}
for ntr in (1000, 10000):
    xyz = S.geometries(eq, ntr, 1)
    e, g = S.pair_potential(xyz, eq)
    db = D.molecular_database.from_arrays(S.atomic_numbers(nat), xyz, energy=e)
    model = M.kreg(hyperparameters={"sigma": 1.0, "lambda": 1e-6})
    model.train(molecular_database=db, property_to_learn="energy", save_model=False)
    mol = D.molecule.from_arrays(S.atomic_numbers(nat), S.geometries(eq, 1, 2)[0])
    for _ in range(50):
        model.predict(molecule=mol, calculate_energy=True, calculate_energy_gradients=True)
    t0 = time.perf_counter()
    n = 2000
    for _ in range(n):
        model.predict(molecule=mol, calculate_energy=True, calculate_energy_gradients=True)
    out[ntr] = round((time.perf_counter() - t0) / n * 1e6, 1)
print(json.dumps({"us_per_predict_call_on_one_molecule": out}))
if "--profile" in sys.argv:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(2000):
        model.predict(molecule=mol, calculate_energy=True, calculate_energy_gradients=True)
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
