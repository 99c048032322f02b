"""Hyper-parameter grid search against the REFERENCE's own optimiser.

tests/golden/grid_*.json hold every (lambda, sigma) point and loss that mlatom.models.kreg.optimize_hyperparameters
(model_cls.py:609-727, optimize_grid :819-854) visits on its default ranges (9 x 9 log grid over lambda in [2^-35, 1],
sigma in [2^-5, 2^9]) with this package's KREG_API bound underneath on the CPU oracle (tools/make_golden_grid.py).

  CPU : (1) with the reference checkout present, the reference optimiser is run again, live, and must reproduce the
            fixture; (2) this package's fused grid (K assembled once per sigma, re-factorised per lambda) on the same
            oracle arithmetic must give the same 81-entry table and the same winner.
  GPU : the fused grid on the CUDA path must reproduce the fixture: rel 1e-8 where K + lambda is well conditioned,
        10 cond eps elsewhere (the loss inherits alpha's conditioning), same winner.
"""
import json
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))

from mlatom_b200 import data as D, synthetic as S  # noqa: E402
from mlatom_b200 import models as M  # noqa: E402

NAMES = ["grid_values_nat9", "grid_valgrad_nat9"]


def load(name):
    with open(os.path.join(HERE, "golden", name + ".json")) as f:
        return json.load(f)


def make_db(nat, n, seed):
    eq = S.equilibrium(nat)
    xyz = S.geometries(eq, n, seed)
    e, g = S.pair_potential(xyz, eq)
    return D.molecular_database.from_arrays(S.atomic_numbers(nat), xyz, energy=e, energy_gradients=g)


def fused_table(gold, tmp_path):
    sub, val = make_db(gold["natoms"], gold["nsub"], gold["train_seed"]), make_db(gold["natoms"], gold["nval"], gold["validation_seed"])
    model = M.kreg(model_file=str(tmp_path / "fused_grid"))
    tk = {"property_to_learn": "energy"}
    pk = {"property_to_predict": "estimated_energy"}
    if gold["use_gradients"]:
        tk["xyz_derivative_property_to_learn"] = "energy_gradients"
        pk["xyz_derivative_property_to_predict"] = "estimated_energy_gradients"
    model.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="grid",
                                   optimization_algorithm_kwargs={"grid_size": gold["grid_size"]}, training_kwargs=tk,
                                   prediction_kwargs=pk, subtraining_molecular_database=sub,
                                   validation_molecular_database=val, fused=True)
    return model


def compare(model, gold, tol_of_cond):
    grid = gold["visited"][:-1]  # the last visit is the retrain at the winner
    assert len(grid) == gold["grid_size"] ** 2 == len(model.grid_losses)
    worst = 0.0
    for g in grid:  # the fused path visits sigma-major (one assembly per sigma); the table is keyed by (lambda, sigma)
        lam, sigma = min(model.grid_losses, key=lambda k: abs(np.log(k[0] / g["lambda"])) + abs(np.log(k[1] / g["sigma"])))
        assert abs(lam - g["lambda"]) <= 1e-14 * g["lambda"] and abs(sigma - g["sigma"]) <= 1e-14 * g["sigma"]
        rel = abs(model.grid_losses[(lam, sigma)] - g["loss"]) / g["loss"]
        assert rel <= tol_of_cond(g["cond"]), (lam, sigma, rel, g["cond"])
        worst = max(worst, rel / tol_of_cond(g["cond"]))
    w = gold["winner"]
    assert abs(model.hyperparameters["lambda"].value - w["lambda"]) <= 1e-12 * w["lambda"]
    assert abs(model.hyperparameters["sigma"].value - w["sigma"]) <= 1e-12 * w["sigma"]
    return worst


@pytest.mark.parametrize("name", NAMES)
def test_reference_optimiser_reproduces_fixture(name, oracle, tmp_path):
    """Build container only: the reference's optimize_hyperparameters, live, on this package's KREG_API."""
    if not os.path.isdir("/root/reference/mlatom"):
        pytest.skip("reference checkout not present")
    import make_golden_grid as G

    saved = {k: sys.modules.get(k) for k in ("mlatom.kreg_api",)}
    import mlatom_b200.engine as eng

    keep = {n: getattr(eng, n) for n in ("xeq", "descriptors", "hessian", "GraphPredictor", "build_kernel_matrix", "solve",
                                         "solve_indefinite")}
    keep_from = eng.PackedModel.__dict__["from_training_set"]
    from mlatom_b200 import kreg_api as KA

    keep_dev, keep_sync = KA.KREG_API._dev, torch.cuda.synchronize
    try:
        ref_models, ref_data = G.reference_kreg_class(oracle)
        gold = load(name)
        rec = G.run_reference_grid(ref_models, ref_data, oracle, gold["natoms"], gold["nsub"], gold["nval"],
                                   gold["use_gradients"], str(tmp_path), gold["grid_size"])
    finally:
        for n, v in keep.items():
            setattr(eng, n, v)
        eng.PackedModel.from_training_set = keep_from
        KA.KREG_API._dev, torch.cuda.synchronize = keep_dev, keep_sync
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert len(rec["visited"]) == len(gold["visited"]) == gold["grid_size"] ** 2 + 1
    for a, b in zip(rec["visited"], gold["visited"]):
        assert a["lambda"] == b["lambda"] and a["sigma"] == b["sigma"]
        assert abs(a["loss"] - b["loss"]) <= 1e-12 * b["loss"] * max(1.0, b["cond"] * 1e-10)
    assert rec["winner"]["lambda"] == gold["winner"]["lambda"] and rec["winner"]["sigma"] == gold["winner"]["sigma"]


@pytest.mark.parametrize("name", NAMES)
def test_fused_grid_equals_reference_optimiser_on_the_oracle(name, oracle, monkeypatch, tmp_path):
    """Same arithmetic (CPU oracle) under both optimisers: the fused restructuring must not change a single loss."""
    from oracle_engine import install_oracle_engine

    install_oracle_engine(monkeypatch.setattr, oracle)
    gold = load(name)
    model = fused_table(gold, tmp_path)
    compare(model, gold, lambda cond: 1e-12 * max(1.0, cond * 1e-10))
    assert abs(model.validation_loss - gold["winner"]["validation_loss"]) <= 1e-10 * gold["winner"]["validation_loss"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_fused_grid_on_device_equals_reference_optimiser(name, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    gold = load(name)
    model = fused_table(gold, tmp_path)
    eps = np.finfo(float).eps
    worst = compare(model, gold, lambda cond: max(1e-8, 10.0 * cond * eps))
    w = gold["winner"]
    assert abs(model.validation_loss - w["validation_loss"]) <= 1e-8 * w["validation_loss"]
    print(f"{name}: worst deviation / tolerance = {worst:.3g}")
