"""MLatomF ``.unf`` model files (Fortran ``access='stream'``, little-endian) for KREG models, so that models
trained here interoperate with ``ml_program='MLatomF'`` files and vice versa.

Layout written by ``writeMLmodel`` (mlatom/MLatomF_src/MLmodel.f90:819-967) and parsed by ``readMLmodel``
(:484-815), restricted to what KREG uses (RE descriptor 'unsorted', Gaussian kernel):

    'MLmod'(5) '1.0   '(6) 'REunsorted'(10)  Ntrain:i4  XvecSize:i4  X[Ntrain][XvecSize]:f8 (point-major)
    'KRR'(3)                                   -- or 'KRm'(3) prior:char(256) Yprior:f8 when a prior is used
    'Gaussian  '(10)  lambda:f8  sigma:f8
    values only :  Ntrain:i4  alpha[Ntrain]:f8
    with grads  :  -9:i4  lambdaGradXYZ:f8  (Ntrain, NtrTot, NtrVal, 0, NtrGrXYZ):5*i4  (NgradXYZmax, NAtomsMax):2*i4
                   per point: Natoms:i4 NgradXYZ:i4 then per atom xyz:3*f8 Z:i4 ;  alpha[NtrTot]:f8
    optional tail: NAtomsMax:i4  XYZeq[NAtomsMax][3]:f8  Zeq[NAtomsMax]:i4

alpha is ordered [values ; gradients (point, atom, xyz)] like KREG.f90 (A_KRR.f90:357-424).
"""
from __future__ import annotations

import struct
from typing import Optional

import numpy as np

_MAGIC, _VERSION, _DESCR, _KERNEL = b"MLmod", b"1.0   ", b"REunsorted", b"Gaussian  "


def write_unf(path: str, X: np.ndarray, alpha: np.ndarray, sigma: float, lambdav: float, prior: float = 0.0,
              prior_label: Optional[str] = None, xyz: Optional[np.ndarray] = None,
              atomic_numbers: Optional[np.ndarray] = None, ntrval: Optional[int] = None, ntrgr: int = 0,
              lambdagradxyz: Optional[float] = None, xyz_eq: Optional[np.ndarray] = None,
              z_eq: Optional[np.ndarray] = None) -> None:
    X = np.ascontiguousarray(X, dtype="<f8")
    alpha = np.ascontiguousarray(alpha, dtype="<f8").reshape(-1)
    ntrain, xsize = X.shape
    ntrval = ntrain if ntrval is None else int(ntrval)
    with open(path, "wb") as f:
        f.write(_MAGIC + _VERSION + _DESCR)
        f.write(struct.pack("<ii", ntrain, xsize))
        f.write(X.tobytes())
        if prior_label is None:
            prior_label = "0" if prior == 0.0 else repr(float(prior))
        if prior_label != "0":
            f.write(b"KRm" + prior_label.encode("ascii")[:256].ljust(256) + struct.pack("<d", float(prior)))
        else:
            f.write(b"KRR")
        f.write(_KERNEL + struct.pack("<dd", float(lambdav), float(sigma)))
        if ntrgr:
            if xyz is None:
                raise ValueError("gradient-trained models store the training geometries: xyz is required")
            xyz = np.ascontiguousarray(xyz, dtype="<f8")
            nat = xyz.shape[1]
            z = np.zeros(nat, dtype="<i4") if atomic_numbers is None else np.asarray(atomic_numbers, dtype="<i4")
            f.write(struct.pack("<i", -9) + struct.pack("<d", float(lambdav if lambdagradxyz is None else lambdagradxyz)))
            f.write(struct.pack("<5i", ntrain, ntrval + ntrgr, ntrval, 0, ntrgr))
            f.write(struct.pack("<2i", 3 * nat, nat))
            for p in range(ntrain):
                f.write(struct.pack("<2i", nat, 3 * nat))
                for a in range(nat):
                    f.write(xyz[p, a].tobytes() + struct.pack("<i", int(z[a])))
            assert alpha.size == ntrval + ntrgr
            f.write(alpha.tobytes())
        else:
            assert alpha.size == ntrain
            f.write(struct.pack("<i", ntrain) + alpha.tobytes())
        if xyz_eq is not None:
            xyz_eq = np.ascontiguousarray(xyz_eq, dtype="<f8")
            nat = xyz_eq.shape[0]
            ze = np.zeros(nat, dtype="<i4") if z_eq is None else np.asarray(z_eq, dtype="<i4")
            f.write(struct.pack("<i", nat) + xyz_eq.tobytes() + ze.tobytes())


def read_unf(path: str) -> dict:
    """Parse a KREG-compatible .unf model (RE 'unsorted' descriptor, Gaussian kernel); anything else raises."""
    buf = open(path, "rb").read()
    pos = 0

    def take(n):
        nonlocal pos
        out = buf[pos:pos + n]
        if len(out) != n:
            raise ValueError("truncated .unf file")
        pos += n
        return out

    def i4(n=1):
        v = struct.unpack("<%di" % n, take(4 * n))
        return v[0] if n == 1 else v

    def f8(n=1):
        v = np.frombuffer(take(8 * n), dtype="<f8").copy()
        return float(v[0]) if n == 1 else v

    if take(5) != _MAGIC or take(6).strip() != b"1.0":
        raise ValueError("not an MLatomF model file (format 1.0)")
    descr = take(10)
    if descr != _DESCR:
        raise ValueError(f"unsupported molecular descriptor {descr!r}: only 'REunsorted' is on the KREG path")
    ntrain, xsize = i4(2)
    out = {"Ntrain": ntrain, "Xsize": xsize, "X": f8(ntrain * xsize).reshape(ntrain, xsize)}
    algo = take(3)
    out["prior"], out["prior_label"] = 0.0, "0"
    if algo == b"KRm":
        out["prior_label"] = take(256).decode("ascii").strip()
        out["prior"] = f8()
    elif algo != b"KRR":
        raise ValueError(f"unsupported algorithm tag {algo!r}")
    kernel = take(10)
    if kernel != _KERNEL:
        raise ValueError(f"unsupported kernel {kernel!r}: only the Gaussian kernel is on the KREG path")
    out["lambdav"], out["sigma"] = f8(), f8()
    tag = i4()
    if tag == -9:
        out["lambdagradxyz"] = f8()
        n2, ntot, nval, _zero, ngr = i4(5)
        _ngmax, natmax = i4(2)
        xyz = np.zeros((ntrain, natmax, 3))
        z = np.zeros((ntrain, natmax), dtype=np.int32)
        for p in range(ntrain):
            nat_p, _ng = i4(2)
            for a in range(natmax):
                xyz[p, a] = f8(3)
                z[p, a] = i4()
        out.update({"NtrVal": nval, "NtrGrXYZ": ngr, "XYZ": xyz, "Z": z[0].copy(), "Natoms": natmax, "alpha": f8(ntot)})
    elif tag == ntrain:
        out.update({"NtrVal": ntrain, "NtrGrXYZ": 0, "lambdagradxyz": out["lambdav"], "alpha": f8(ntrain)})
    else:
        raise ValueError(f"unsupported .unf variant (tag {tag})")
    if pos < len(buf):
        nat = i4()
        out["Req"] = f8(3 * nat).reshape(nat, 3)
        out["Zeq"] = np.asarray(i4(nat) if nat > 1 else [i4()], dtype=np.int32)
        out.setdefault("Natoms", nat)
    return out
