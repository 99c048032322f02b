#!/usr/bin/env python3
"""(lambda, sigma) grid search for a values-trained model on MLatomF's Matern kernel through
models.kreg.optimize_hyperparameters: the fused route (K(sigma) once per sigma, re-factorised per lambda, all resident)
beside the pointwise route (train + predict per grid point, as the reference does)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mlatom_b200 import data as D, models, synthetic as S

nat, ntr, nval, grid = 9, 4000, 1000, 9
eq = S.equilibrium(nat)
z = S.atomic_numbers(nat)
def db(n, seed):
    xyz = S.geometries(eq, n, seed)
    e, _ = S.pair_potential(xyz, eq)
    return D.molecular_database.from_arrays(z, xyz, energy=e)
sub, val = db(ntr, 1), db(nval, 3)
out = {"workload": f"grid search {grid}x{grid} (lambda, sigma), Matern n=2, natoms={nat}, ntrain={ntr} values, nval={nval}"}
for fused in (True, False, True):
    m = models.kreg(model_file=f"/tmp/grid_radial_{fused}", ml_program="MLatomF", equilibrium_molecule=D.molecule.from_arrays(z, eq))
    m.additional_mlatomf_args = ["kernel=Matern", "nn=2"]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="grid",
                               optimization_algorithm_kwargs={"grid_size": grid}, fused=fused,
                               training_kwargs={"property_to_learn": "energy", "prior": "mean"},
                               prediction_kwargs={"property_to_predict": "estimated_energy"},
                               subtraining_molecular_database=sub, validation_molecular_database=val)
    torch.cuda.synchronize()
    out["fused_s" if fused else "pointwise_s"] = round(time.perf_counter() - t0, 3)  # the second fused run is the warm one
    out["winner_fused" if fused else "winner_pointwise"] = [m.hyperparameters["lambda"].value, m.hyperparameters["sigma"].value,
                                                            m.validation_loss]
print(json.dumps(out))
