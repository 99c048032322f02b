"""Text data files of MLatom / MLatomF, array-first.

The reference reads and writes its data sets as text: geometries in XYZ format (``dataset.f90:826-887`` readXYZ;
``data.py:1568-1578, 1956-1972``), one reference value per line (``dataset.f90:1671-1690`` writeYest, ``'(F25.13)'``;
``data.py:1869-1882, 1983-1990``), and XYZ-vectorial properties as XYZ-like blocks -- number of atoms, empty line,
three ``F25.13`` columns per atom (``dataset.f90:1716-1766`` writeYgradXYZest; ``data.py:1944-1954, 2120-2130``).

These functions go straight between such files and the fp64 arrays the device path consumes -- (N,), (N, Nat, 3) --
without building one Python object per atom; ``mlatom_b200.data.molecular_database`` wraps them with the reference's
method names.  All molecules of a file must have the same number of atoms (KREG's descriptor requires it).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

SYMBOLS = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr "
    "Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt "
    "Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl Mc Lv "
    "Ts Og"
).split()
_Z = {s.casefold(): z for z, s in enumerate(SYMBOLS)}


def atomic_number(token: str) -> int:
    """Element symbol (any case) or atomic number, as readXYZ accepts either (dataset.f90:864-876)."""
    try:
        return int(token)
    except ValueError:
        try:
            return _Z[token.casefold()]
        except KeyError:
            raise ValueError(f"unknown element {token!r}") from None


def parse_xyz(text: str) -> Tuple[np.ndarray, np.ndarray, list]:
    """XYZ text -> (Z (N, Nat) int, xyz (N, Nat, 3) float64, comments)."""
    lines = text.splitlines()
    pos, zs, comments, rows = 0, [], [], []
    nat_all = None
    while pos < len(lines):
        if not lines[pos].strip():
            pos += 1
            continue
        nat = int(lines[pos].split()[0])
        if nat_all is None:
            nat_all = nat
        elif nat != nat_all:
            raise ValueError(f"molecule {len(zs) + 1} has {nat} atoms, the first one has {nat_all}")
        if pos + 2 + nat > len(lines):
            raise ValueError("Error while reading coordinates: file ends inside a molecule")
        comments.append(lines[pos + 1].rstrip("\n"))
        block = [ln.split() for ln in lines[pos + 2: pos + 2 + nat]]
        if any(len(b) < 4 for b in block):
            raise ValueError(f"Error while reading coordinates of molecule {len(zs) + 1}")
        zs.append([atomic_number(b[0]) for b in block])
        rows.extend(b[1:4] for b in block)
        pos += 2 + nat
    if nat_all is None:
        return np.zeros((0, 0), dtype=np.int64), np.zeros((0, 0, 3)), []
    xyz = np.array(rows, dtype=np.float64).reshape(len(zs), nat_all, 3)
    return np.array(zs, dtype=np.int64), xyz, comments


def read_xyz(filename: str) -> Tuple[np.ndarray, np.ndarray]:
    with open(filename, "r") as f:
        z, xyz, _ = parse_xyz(f.read())
    return z, xyz


def write_xyz(filename: str, atomic_numbers, xyz, comments=None) -> None:
    """data.py:1956-1972: '%d', comment line, '%-3s %25.13f %25.13f %25.13f' per atom."""
    xyz = np.asarray(xyz, dtype=np.float64)
    z = np.broadcast_to(np.asarray(atomic_numbers, dtype=np.int64), xyz.shape[:2])
    with open(filename, "w") as f:
        for i in range(xyz.shape[0]):
            f.write("%d\n%s\n" % (xyz.shape[1], comments[i] if comments else ""))
            f.write("".join("%-3s %25.13f %25.13f %25.13f\n" % (SYMBOLS[z[i, a]], *xyz[i, a]) for a in range(xyz.shape[1])))


def read_y(filename: str) -> np.ndarray:
    """One value per line (data.py:1869-1882)."""
    with open(filename, "r") as f:
        return np.array([float(line) for line in f if line.strip()], dtype=np.float64)


def write_y(filename: str, y) -> None:
    """'(F25.13)' per line (dataset.f90:1686; data.py:1990)."""
    with open(filename, "w") as f:
        f.write("".join("%25.13f\n" % v for v in np.asarray(y, dtype=np.float64).reshape(-1)))


def parse_xyz_vectors(text: str) -> np.ndarray:
    """Blocks of: number of atoms, one (ignored) line, Nat lines of three numbers -> (N, Nat, 3)."""
    lines = text.splitlines()
    pos, rows, nat_all, n = 0, [], None, 0
    while pos < len(lines):
        if not lines[pos].strip():
            pos += 1
            continue
        nat = int(lines[pos].split()[0])
        if nat_all is None:
            nat_all = nat
        elif nat != nat_all:
            raise ValueError(f"block {n + 1} has {nat} atoms, the first one has {nat_all}")
        block = [ln.split() for ln in lines[pos + 2: pos + 2 + nat]]
        if len(block) != nat or any(len(b) < 3 for b in block):
            raise ValueError(f"bad vector block {n + 1}")
        rows.extend(b[-3:] for b in block)
        pos += 2 + nat
        n += 1
    if nat_all is None:
        return np.zeros((0, 0, 3))
    return np.array(rows, dtype=np.float64).reshape(n, nat_all, 3)


def read_xyz_vectors(filename: str) -> np.ndarray:
    with open(filename, "r") as f:
        return parse_xyz_vectors(f.read())


def write_xyz_vectors(filename: str, vectors, fortran_style: bool = False) -> None:
    """Python flavour ' %25.13f %25.13f %25.13f' (data.py:2120-2130) or MLatomF's '(3F25.13)' (dataset.f90:1741)."""
    v = np.asarray(vectors, dtype=np.float64)
    fmt = "%25.13f%25.13f%25.13f\n" if fortran_style else " %25.13f %25.13f %25.13f\n"
    with open(filename, "w") as f:
        for i in range(v.shape[0]):
            f.write("%d\n\n" % v.shape[1])
            f.write("".join(fmt % tuple(v[i, a]) for a in range(v.shape[1])))
