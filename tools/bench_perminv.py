"""Throughput of the permutationally invariant kernel path (mlatom_b200/perminv.py) on one GPU: K assembly and E + dE/dxyz
prediction for ethanol with the CH2 / CH3 hydrogens permuted (12 permutations).  Prints one JSON line."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mlatom_b200 import engine, perminv  # noqa: E402
from mlatom_b200 import synthetic as S  # noqa: E402


def timed(fn, reps=3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for it in range(reps + 1):
        torch.cuda.synchronize()
        ev[0].record()
        out = fn()
        ev[1].record()
        torch.cuda.synchronize()
        if it:
            ms.append(ev[0].elapsed_time(ev[1]))
    return float(np.mean(ms)), out


def main():
    nat, ntr, nte, sigma = 9, 4000, 200_000, 1.0
    dev = torch.device("cuda")
    eq = S.equilibrium(nat)
    perms = perminv.Permutations(perminv.atom_maps(nat, perminv.parse_perm_inv_nuclei("4-5.6-7-8")), dev)
    xeq = engine.xeq(torch.from_numpy(eq).to(dev))
    xtr = torch.from_numpy(S.geometries(eq, ntr, 1)).to(dev)
    xte = torch.from_numpy(S.geometries(eq, nte, 2)).to(dev)
    k_ms, k = timed(lambda: perminv.kernel_block(xtr, xtr, xeq, perms, sigma, symmetric=True, lam=1e-8))
    e_tr, _ = S.pair_potential(xtr.cpu().numpy(), eq)
    alpha = engine.solve(k, torch.from_numpy(e_tr - e_tr.mean()).to(dev))
    model = perminv.PermInvModel(xtr, xeq, alpha, sigma, float(e_tr.mean()), perms)
    p_ms, (e, g) = timed(lambda: model.predict(xte))
    d = nat * (nat - 1) // 2
    pairs_k = float(ntr) * ntr * perms.nperm
    pairs_p = float(nte) * ntr * perms.nperm
    print(json.dumps({
        "workload": "ethanol, permInvNuclei=4-5.6-7-8 (12 permutations)", "ntrain": ntr, "ntest": nte, "nperm": perms.nperm,
        "kernel_matrix_ms": k_ms, "kernel_matrix_gaussian_pairs_per_s": pairs_k / k_ms * 1e3,
        "kernel_matrix_algorithmic_tflops": pairs_k * (3 * d + 3) / k_ms * 1e-9,
        "predict_ms": p_ms, "predict_geometries_per_s": nte / p_ms * 1e3,
        "predict_algorithmic_tflops": pairs_p * (5 * d + 4) / p_ms * 1e-9,
        "finite": bool(torch.isfinite(e).all() and torch.isfinite(g).all()),
    }))


if __name__ == "__main__":
    main()
