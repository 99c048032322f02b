#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE's own compiled KREG library.

Run in the build container only (needs oracle/_ref, i.e. /root/reference's
mlatom/fortran/np2/KREG.so copied by `make -C oracle ref`):

    python tools/make_golden.py

Each fixture holds seeded synthetic geometries (mlatom_b200.synthetic), the hyper-parameters,
and what the reference computed from them through kreg_mp_calc_re_descriptors_ /
kreg_mp_calc_kernel_matrix_ / kreg_mp_train_ / kreg_mp_predict_ (KREG.f90:257-271, 293-346,
495-532, 534-596): X, K, alpha, E, dE/dxyz.  alpha comes from OpenBLAS dpotrf/dpotrs behind the
MKL shim (the reference links MKL), so tests treat alpha as an input, not as a pinned output.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mlatom_b200 import synthetic as S  # noqa: E402
from oracle import ref_kreg as R  # noqa: E402

CASES = [
    # name, natoms, ntrain, ntest, use_values, use_grads, sigma, lambda
    ("values_nat5", 5, 24, 8, True, False, 0.8, 1e-6),
    ("values_nat9", 9, 40, 10, True, False, 1.0, 1e-6),
    ("values_nat21", 21, 16, 6, True, False, 2.0, 1e-6),
    ("valgrad_nat5", 5, 8, 6, True, True, 0.9, 1e-6),
    ("valgrad_nat9", 9, 6, 5, True, True, 1.1, 1e-4),
    ("gradonly_nat5", 5, 7, 5, False, True, 1.2, 1e-5),
]


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    R.set_threads(1)
    for name, nat, ntr, nte, use_v, use_g, sigma, lam in CASES:
        eq = S.equilibrium(nat)
        xyz_tr = S.geometries(eq, ntr, seed=1)
        xyz_te = S.geometries(eq, nte, seed=2)
        e_tr, g_tr = S.pair_potential(xyz_tr, eq)
        xeq = R.get_xeq(eq)
        x_tr = R.calc_re_descriptors(xyz_tr, xeq)
        x_te = R.calc_re_descriptors(xyz_te, xeq)
        ntrval = ntr if use_v else 0
        ntrgr = 3 * nat * ntr if use_g else 0
        prior = float(np.mean(e_tr)) if use_v else 0.0
        parts = []
        if use_v:
            parts.append(e_tr - prior)
        if use_g:
            parts.append(g_tr.reshape(-1))
        ytrain = np.concatenate(parts)
        k = R.calc_kernel_matrix(xyz_tr, x_tr, sigma, ntrval, ntrgr)
        _, alpha = R.train(xyz_tr, x_tr, ytrain, sigma, lam, lam, ntrval, ntrgr)
        e_te, g_te, h_te = R.predict(xyz_tr, x_tr, xyz_te, x_te, alpha, sigma, ntrval, ntrgr, calc_hess=True)
        np.savez_compressed(
            os.path.join(outdir, name + ".npz"),
            natoms=nat, ntrain=ntr, ntest=nte, ntrval=ntrval, ntrgr=ntrgr, sigma=sigma, lambdav=lam,
            lambdagradxyz=lam, prior=prior, eq=eq, xyz_train=xyz_tr, xyz_test=xyz_te, y_train=e_tr,
            ygrad_train=g_tr, ytrain=ytrain, ac2d=R.get_ac2darray(nat), xeq=xeq, X_train=x_tr, X_test=x_te, K=k,
            alpha=alpha, E_test=e_te, G_test=g_te, H_test=h_te,
        )
        print(f"{name}: M={ntrval + ntrgr} K{k.shape} |E|max={np.abs(e_te).max():.3e} |G|max={np.abs(g_te).max():.3e}")


if __name__ == "__main__":
    main()
