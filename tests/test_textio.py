"""CPU: MLatom / MLatomF text data files (mlatom_b200/textio.py) -- formats of dataset.f90:826-887, 1671-1766 and
data.py:1568-1990, 2120-2130; cross-checked against the reference's own reader/writer when it is importable."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlatom_b200 import data as D, synthetic as S, textio as T  # noqa: E402


def make(nat=9, n=5, seed=4):
    eq = S.equilibrium(nat)
    xyz = S.geometries(eq, n, seed)
    e, g = S.pair_potential(xyz, eq)
    return S.atomic_numbers(nat), xyz, e, g


def test_formats_are_the_references(tmp_path):
    z, xyz, e, g = make()
    T.write_y(tmp_path / "y.dat", e)
    T.write_xyz_vectors(tmp_path / "g.dat", g)
    T.write_xyz_vectors(tmp_path / "gf.dat", g, fortran_style=True)
    T.write_xyz(tmp_path / "x.xyz", z, xyz)
    ylines = open(tmp_path / "y.dat").read().splitlines()
    assert len(ylines) == 5 and all(len(ln) == 25 for ln in ylines) and ylines[0] == "%25.13f" % e[0]
    glines = open(tmp_path / "g.dat").read().splitlines()
    assert glines[0] == "9" and glines[1] == "" and glines[2] == " %25.13f %25.13f %25.13f" % tuple(g[0, 0])
    assert len(glines) == 5 * 11
    flines = open(tmp_path / "gf.dat").read().splitlines()
    assert flines[2] == "%25.13f%25.13f%25.13f" % tuple(g[0, 0]) and len(flines[2]) == 75   # '(3F25.13)'
    xlines = open(tmp_path / "x.xyz").read().splitlines()
    assert xlines[0] == "9" and xlines[2].split()[0] == "C" and xlines[2].startswith("C  ")


def test_round_trip_and_database_methods(tmp_path):
    z, xyz, e, g = make()
    T.write_xyz(tmp_path / "x.xyz", z, xyz)
    T.write_y(tmp_path / "y.dat", e)
    T.write_xyz_vectors(tmp_path / "g.dat", g)
    z2, xyz2 = T.read_xyz(tmp_path / "x.xyz")
    assert np.array_equal(z2, np.broadcast_to(z, z2.shape))
    tol = 0.5000001e-13  # 13 decimals
    assert np.abs(xyz2 - xyz).max() <= tol and np.abs(T.read_y(tmp_path / "y.dat") - e).max() <= tol
    for style in (False, True):
        T.write_xyz_vectors(tmp_path / "g2.dat", g, fortran_style=style)
        assert np.abs(T.read_xyz_vectors(tmp_path / "g2.dat") - g).max() <= tol
    db = D.molecular_database.from_xyz_file(str(tmp_path / "x.xyz"))
    db.add_scalar_properties_from_file(str(tmp_path / "y.dat"), "energy")
    db.add_xyz_vectorial_properties_from_file(str(tmp_path / "g.dat"), "energy_gradients")
    assert len(db) == 5 and np.abs(db.get_properties("energy") - e).max() <= tol
    assert np.abs(db.get_xyz_vectorial_properties("energy_gradients") - g).max() <= tol
    db.write_file_with_xyz_coordinates(str(tmp_path / "x2.xyz"))
    db.write_file_with_properties(str(tmp_path / "y2.dat"), "energy")
    db.write_file_energy_gradients(str(tmp_path / "g3.dat"))
    assert open(tmp_path / "x2.xyz").read() == open(tmp_path / "x.xyz").read()
    assert open(tmp_path / "y2.dat").read() == open(tmp_path / "y.dat").read()
    assert open(tmp_path / "g3.dat").read() == open(tmp_path / "g.dat").read()


def test_atomic_numbers_instead_of_symbols_and_errors(tmp_path):
    (tmp_path / "n.xyz").write_text("2\n\n1 0.0 0.0 0.0\n8 0.0 0.0 0.96\n2\ncomment\nh 0 0 0\nO 0 0 1\n")
    z, xyz = T.read_xyz(tmp_path / "n.xyz")
    assert z.tolist() == [[1, 8], [1, 8]] and xyz[1, 1, 2] == 1.0
    (tmp_path / "bad.xyz").write_text("2\n\nH 0 0 0\nO 0 0 1\n3\n\nH 0 0 0\nH 0 0 1\nO 0 1 0\n")
    with pytest.raises(ValueError):
        T.read_xyz(tmp_path / "bad.xyz")
    (tmp_path / "short.xyz").write_text("3\n\nH 0 0 0\n")
    with pytest.raises(ValueError):
        T.read_xyz(tmp_path / "short.xyz")
    (tmp_path / "el.xyz").write_text("1\n\nQq 0 0 0\n")
    with pytest.raises(ValueError):
        T.read_xyz(tmp_path / "el.xyz")
    assert T.read_xyz_vectors(os.devnull).shape == (0, 0, 3)


def test_files_interchange_with_the_reference(tmp_path):
    from conftest import reference_mlatom_data

    RD = reference_mlatom_data()
    if RD is None:
        pytest.skip("reference tree only exists in the build container")
    z, xyz, e, g = make(nat=5, n=4)
    T.write_xyz(tmp_path / "x.xyz", z, xyz)
    T.write_y(tmp_path / "y.dat", e)
    T.write_xyz_vectors(tmp_path / "g.dat", g)
    rdb = RD.molecular_database.from_xyz_file(str(tmp_path / "x.xyz"))
    rdb.add_scalar_properties_from_file(str(tmp_path / "y.dat"), "energy")
    rdb.add_xyz_vectorial_properties_from_file(str(tmp_path / "g.dat"), "energy_gradients")
    # the reference writes; we read
    rdb.write_file_with_xyz_coordinates(str(tmp_path / "rx.xyz"))
    rdb.write_file_with_properties(str(tmp_path / "ry.dat"), "energy")
    rdb.write_file_with_xyz_derivative_properties(str(tmp_path / "rg.dat"), "energy_gradients")
    assert open(tmp_path / "rx.xyz").read() == open(tmp_path / "x.xyz").read()
    assert open(tmp_path / "ry.dat").read() == open(tmp_path / "y.dat").read()
    assert open(tmp_path / "rg.dat").read() == open(tmp_path / "g.dat").read()
    z2, xyz2 = T.read_xyz(tmp_path / "rx.xyz")
    assert np.array_equal(xyz2, np.asarray(rdb.xyz_coordinates)) and z2[0].tolist() == list(rdb[0].atomic_numbers)
    assert np.array_equal(T.read_y(tmp_path / "ry.dat"), rdb.get_properties("energy"))
    assert np.array_equal(T.read_xyz_vectors(tmp_path / "rg.dat"), rdb.get_xyz_vectorial_properties("energy_gradients"))
