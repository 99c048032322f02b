"""CPU: MLatomF .unf model files -- byte layout against the writer in MLmodel.f90:819-967, and round trips."""
import struct

import numpy as np
import pytest

from mlatom_b200 import unf


def test_values_only_layout_and_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    ntr, nat = 7, 4
    D = nat * (nat - 1) // 2
    X, alpha = rng.random((ntr, D)), rng.standard_normal(ntr)
    req, z = rng.random((nat, 3)), np.array([6, 1, 1, 8])
    p = str(tmp_path / "m.unf")
    unf.write_unf(p, X=X, alpha=alpha, sigma=1.5, lambdav=1e-6, prior=-3.25, prior_label="mean", xyz_eq=req, z_eq=z)
    b = open(p, "rb").read()
    # MLmodel.f90:845-847 header, :858-862 descriptor tag, :880-886 X block
    assert b[:21] == b"MLmod1.0   REunsorted"
    assert struct.unpack("<ii", b[21:29]) == (ntr, D)
    assert np.array_equal(np.frombuffer(b[29:29 + 8 * ntr * D], "<f8").reshape(ntr, D), X)
    o = 29 + 8 * ntr * D
    # :890-893 'KRm' + prior string(256) + Yprior ; :903 kernel tag ; :905-909 lambda, sigma
    assert b[o:o + 3] == b"KRm" and b[o + 3:o + 259].rstrip() == b"mean"
    assert struct.unpack("<d", b[o + 259:o + 267])[0] == -3.25
    assert b[o + 267:o + 277] == b"Gaussian  "
    assert struct.unpack("<dd", b[o + 277:o + 293]) == (1e-6, 1.5)
    # :935-945 Ntrain + alpha ; :949-959 equilibrium tail
    assert struct.unpack("<i", b[o + 293:o + 297])[0] == ntr
    assert len(b) == o + 297 + 8 * ntr + 4 + 24 * nat + 4 * nat
    m = unf.read_unf(p)
    assert np.array_equal(m["X"], X) and np.array_equal(m["alpha"], alpha) and np.array_equal(m["Req"], req)
    assert m["sigma"] == 1.5 and m["lambdav"] == 1e-6 and m["prior"] == -3.25 and m["prior_label"] == "mean"
    assert m["NtrVal"] == ntr and m["NtrGrXYZ"] == 0 and list(m["Zeq"]) == list(z) and m["Natoms"] == nat


def test_gradient_model_roundtrip_and_no_prior(tmp_path):
    rng = np.random.default_rng(1)
    ntr, nat = 5, 3
    D = 3
    X, xyz = rng.random((ntr, D)), rng.random((ntr, nat, 3))
    alpha = rng.standard_normal(ntr + 9 * ntr)
    p = str(tmp_path / "g.unf")
    unf.write_unf(p, X=X, alpha=alpha, sigma=0.7, lambdav=1e-5, xyz=xyz, atomic_numbers=[8, 1, 1], ntrval=ntr,
                  ntrgr=9 * ntr, lambdagradxyz=2e-5, xyz_eq=xyz[0], z_eq=[8, 1, 1])
    b = open(p, "rb").read()
    o = 29 + 8 * ntr * D
    assert b[o:o + 3] == b"KRR" and b[o + 3:o + 13] == b"Gaussian  "      # no prior: :894-895
    assert struct.unpack("<i", b[o + 29:o + 33])[0] == -9                     # :917 gradient marker
    assert struct.unpack("<5i", b[o + 41:o + 61]) == (ntr, 10 * ntr, ntr, 0, 9 * ntr)  # :919
    m = unf.read_unf(p)
    assert np.array_equal(m["XYZ"], xyz) and np.array_equal(m["alpha"], alpha) and m["lambdagradxyz"] == 2e-5
    assert m["NtrVal"] == ntr and m["NtrGrXYZ"] == 9 * ntr and list(m["Z"]) == [8, 1, 1] and m["prior"] == 0.0


def test_foreign_models_are_refused(tmp_path):
    p = str(tmp_path / "x.unf")
    open(p, "wb").write(b"MLmod1.0   CMsorted  " + b"\0" * 64)
    with pytest.raises(ValueError):
        unf.read_unf(p)
    open(p, "wb").write(b"garbage")
    with pytest.raises(ValueError):
        unf.read_unf(p)
