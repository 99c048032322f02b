"""``KREG_API`` -- the backend adapter ``mlatom.models.kreg`` talks to, re-implemented on the B200
kernels (reference: mlatom/kreg_api.py:11-177, an f2py adapter around mlatom/fortran/KREG.f90).

Same method names, argument meaning, attribute names and ``.npz`` model layout
(kreg_api.py:134-156), so ``models.kreg`` and its callers (MD, geometry
optimisation, active learning: ``model.kreg_api.save_model(path)``, al_utils.py:1105) work
unchanged.  Differences that are deliberate:
  * nothing runs on the CPU: K assembly, the Cholesky solve and prediction are CUDA kernels /
    cuSOLVER on the current device; numpy arrays are only the host-side copies kept for the model file;
  * ``self.K`` is not kept after training (it is factorised in place; the reference holds K and a
    copy, 2 x Ksize^2 doubles, KREG.f90:613-616);
  * the Hessian buffer the reference always allocates (kreg_api.py:116, 5.8 GB at 1M geometries) does
    not exist; Hessians are computed only when ``hessian_to_predict`` is given (kreg_hessian);
  * failures raise Python exceptions instead of Fortran ``stop`` killing the interpreter
    (mlatom/fortran/stopper.f90:13).
"""
from __future__ import annotations

import warnings
from typing import Optional

import numpy as np
import torch

from . import data as _data
from . import engine
from ._lib import FILL_LOWER, KregError


def _as_database(molecular_database=None, molecule=None):
    """model.predict's input normalisation (mlatom/model_cls.py:78-107)."""
    if molecular_database is not None:
        return molecular_database
    if molecule is not None:
        return _data.molecular_database([molecule])
    raise ValueError("Either molecule or molecular_database should be provided in input")


def ac2d_array(natoms: int) -> np.ndarray:
    """get_ac2dArray (KREG.f90:34-52): 1-based pair index table, int32 (Nat,Nat), Fortran order."""
    out = np.zeros((natoms, natoms), dtype=np.int32, order="F")
    for i in range(1, natoms):
        for j in range(i + 1, natoms + 1):
            out[i - 1, j - 1] = out[j - 1, i - 1] = (2 * natoms - i) * (i - 1) // 2 + j - i
    return out


class KREG_API:
    def __init__(self, device: Optional[torch.device] = None):
        self.device = device
        # hyper-parameters
        self.sigma = None
        self.lambdav = None
        self.lambdagradxyz = None
        # sizes
        self.Natoms = None
        self.Xsize = None
        self.Ntrain = None
        self.NtrVal = None
        self.NtrGrXYZ = None
        self.Ksize = None
        self.ac2dArray = None
        self.Req = None   # equilibrium CARTESIAN coordinates (Nat,3) -- name kept from the reference
        self.Xeq = None   # equilibrium pair distances (Xsize,)
        # training
        self.prior = 0.0
        self.XYZ = None
        self.X = None
        self.K = None
        self.alpha = None
        self.Yref = None
        self.YgradXYZref = None
        self.Ytrain = None
        self.nthreads = None  # accepted and ignored: there are no host threads on the path
        self._model = None    # engine.PackedModel, device resident
        self._graphs = {}     # (n, want_E, want_G) -> engine.GraphPredictor for small fixed-shape batches
        self.graph_batch_limit = 64

    # -- helpers ----------------------------------------------------------------------------
    def _dev(self):
        if self.device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("mlatom_b200 needs a CUDA (sm_100) device: there is no CPU fallback")
            self.device = torch.device("cuda", torch.cuda.current_device())
        return self.device

    def _t(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(self._dev())

    # -- training ---------------------------------------------------------------------------
    def train(self, sigma, lambdav, lambdagradxyz, molecular_database, equilibrium_molecule, property_to_learn=None,
              xyz_derivative_property_to_learn=None, prior=None, shutup=True):
        """Same contract as kreg_api.py:39-88."""
        self.clean_up()
        xyz = np.asarray(molecular_database.xyz_coordinates, dtype=np.float64)
        y = molecular_database.get_properties(property_to_learn) if property_to_learn is not None else None
        ygrad = (molecular_database.get_xyz_vectorial_properties(xyz_derivative_property_to_learn)
                 if xyz_derivative_property_to_learn is not None else None)
        return self.train_arrays(sigma, lambdav, lambdagradxyz, xyz, np.asarray(equilibrium_molecule.xyz_coordinates),
                                 y=y, ygrad=ygrad, prior=prior, shutup=shutup)

    def train_arrays(self, sigma, lambdav, lambdagradxyz, xyz, req, y=None, ygrad=None, prior=None, shutup=True):
        """Array entry point: xyz (N,Nat,3), req (Nat,3), y (N,), ygrad (N,Nat,3); numpy or cuda tensors."""
        if y is None and ygrad is None:
            raise ValueError("nothing to learn: give property_to_learn and/or xyz_derivative_property_to_learn")
        self._model, self._graphs = None, {}  # a retrain through this entry point must not keep the old model's graphs alive
        self.sigma, self.lambdav, self.lambdagradxyz = float(sigma), float(lambdav), float(lambdagradxyz)
        d_xyz = xyz if torch.is_tensor(xyz) else self._t(xyz)
        d_xyz = d_xyz.to(self._dev(), torch.float64).contiguous()
        self.Ntrain, self.Natoms = int(d_xyz.shape[0]), int(d_xyz.shape[1])
        self.Xsize = self.Natoms * (self.Natoms - 1) // 2
        parts = []
        if y is not None:
            self.NtrVal = self.Ntrain
            yv = y.detach().cpu().numpy() if torch.is_tensor(y) else np.asarray(y, dtype=np.float64)
            self.Yref = yv
            if prior is None:
                self.prior = 0.0
            elif isinstance(prior, str):
                if prior.casefold() == "mean":
                    self.prior = float(np.mean(yv))
                else:
                    raise ValueError(f"Unsupported prior: {prior}")
            else:
                self.prior = float(prior)
            parts.append(yv - self.prior)
        else:
            self.NtrVal = 0
        if ygrad is not None:
            self.NtrGrXYZ = 3 * self.Natoms * self.Ntrain
            gv = ygrad.detach().cpu().numpy() if torch.is_tensor(ygrad) else np.asarray(ygrad, dtype=np.float64)
            self.YgradXYZref = gv
            parts.append(gv.reshape(self.NtrGrXYZ))  # (point, atom, xyz), xyz fastest: KREG.f90:273-291
        else:
            self.NtrGrXYZ = 0
        self.Ytrain = np.concatenate(parts)
        self.Ksize = self.NtrVal + self.NtrGrXYZ
        self.ac2dArray = ac2d_array(self.Natoms)
        self.Req = np.asarray(req.detach().cpu().numpy() if torch.is_tensor(req) else req, dtype=np.float64)
        d_xeq = engine.xeq(self._t(self.Req))
        d_y = self._t(self.Ytrain)

        use_v, use_g = self.NtrVal > 0, self.NtrGrXYZ > 0
        k = engine.build_kernel_matrix(d_xyz, d_xeq, self.sigma, self.lambdav, self.lambdagradxyz, use_values=use_v,
                                       use_gradients=use_g, fill=FILL_LOWER)
        try:
            d_alpha = engine.solve(k, d_y)
        except KregError as err:
            if err.code <= 0:
                raise
            # The reference falls back to Bunch-Kaufman (mathUtils.f90:24-26 -> solveSysLinEqsBK :106-186) when dpotrf
            # fails.  Same here, behind the C ABI (kreg_solve_indefinite: cuSOLVER sytrf/sytrs), on a RE-ASSEMBLED
            # matrix written into the same buffer: potrf has overwritten part of the first one (the reference hands BK
            # that clobbered array; not imitated).
            if not shutup:
                print(" Cholesky decomposition failed. Attempting Bunch-Kaufman decomposition")
            warnings.warn(f"K + lambda is not positive definite (leading minor {err.code}); "
                          "falling back to a Bunch-Kaufman solve -- consider a larger lambda")
            engine.build_kernel_matrix(d_xyz, d_xeq, self.sigma, self.lambdav, self.lambdagradxyz, use_values=use_v,
                                       use_gradients=use_g, fill=FILL_LOWER, out=k)
            d_alpha = engine.solve_indefinite(k, d_y)
        del k
        self._model = engine.PackedModel.from_training_set(d_xyz, d_xeq, d_alpha, self.sigma, self.prior, self.NtrVal,
                                                           self.NtrGrXYZ)
        # host copies for the model file (kreg_api.py:134-150)
        self.XYZ = d_xyz.cpu().numpy()
        self.X = engine.descriptors(d_xyz, d_xeq).cpu().numpy()
        self.Xeq = d_xeq.cpu().numpy()
        self.alpha = d_alpha.cpu().numpy().reshape(1, self.Ksize)
        return self

    # -- prediction -------------------------------------------------------------------------
    def _ensure_model(self):
        if self._model is None:
            if self.alpha is None or (self.XYZ is None and self.X is None):
                raise RuntimeError("KREG_API model not found: train or load a model first")
            req = np.asarray(self.Req, dtype=np.float64)
            if self.XYZ is None:  # values-only model read from an MLatomF .unf file: descriptors, no geometries
                self._model = engine.PackedModel.from_descriptors(
                    self._t(self.X), engine.xeq(self._t(req)), self._t(np.asarray(self.alpha).reshape(-1)),
                    float(self.sigma), float(self.prior), int(self.Natoms))
                return self._model
            self._model = engine.PackedModel.from_training_set(
                self._t(self.XYZ), engine.xeq(self._t(req)), self._t(np.asarray(self.alpha).reshape(-1)),
                float(self.sigma), float(self.prior), int(self.NtrVal), int(self.NtrGrXYZ))
        return self._model

    def predict_arrays(self, xyz, calculate_energy=True, calculate_energy_gradients=True):
        """Tensor fast path: xyz (N,Nat,3) as a cuda fp64 tensor (results stay on the device) or a numpy
        array (streamed through pinned buffers; numpy results)."""
        model = self._ensure_model()
        if torch.is_tensor(xyz) and xyz.is_cuda:
            return model.predict(xyz.to(torch.float64).contiguous(), calculate_energy, calculate_energy_gradients)
        xyz = np.asarray(xyz, dtype=np.float64)
        n = xyz.shape[0]
        if 0 < n <= self.graph_batch_limit:
            # MD / geometry optimisation: the same small batch shape every step -> replay a captured CUDA graph
            key = (n, bool(calculate_energy), bool(calculate_energy_gradients))
            gp = self._graphs.get(key)
            if gp is None or gp.model is not model:
                gp = self._graphs[key] = engine.GraphPredictor(model, n, calculate_energy, calculate_energy_gradients)
            e, g = gp(xyz)
            return (None if e is None else e.copy()), (None if g is None else g.copy())
        e, g = model.predict_host(xyz, calculate_energy, calculate_energy_gradients)
        torch.cuda.synchronize(model.device)
        return (None if e is None else e.numpy()), (None if g is None else g.numpy())

    def hessian_arrays(self, xyz) -> np.ndarray:
        """Hessian as the reference defines it (values part of alpha, no d2X/dM2 term; KREG.f90:180-221, 393-420)."""
        self._ensure_model()
        req = np.asarray(self.Req, dtype=np.float64)
        nv = int(self.NtrVal)
        x_tr = self._t(np.asarray(self.X)[:nv] if nv else np.zeros((0, int(self.Xsize))))
        a_v = self._t(np.asarray(self.alpha).reshape(-1)[:nv]) if nv else None
        h = engine.hessian(x_tr, a_v, engine.xeq(self._t(req)), float(self.sigma), self._t(xyz))
        return h.cpu().numpy()

    def predict(self, molecular_database=None, molecule=None, property_to_predict=None,
                xyz_derivative_property_to_predict=None, hessian_to_predict=None):
        """Same contract as kreg_api.py:92-131: results are written INTO the molecules; the XYZ property
        is the gradient dE/dxyz (not the force)."""
        mol_db = _as_database(molecular_database, molecule)
        want_e = property_to_predict is not None
        want_g = xyz_derivative_property_to_predict is not None
        if not (want_e or want_g or hessian_to_predict is not None):
            return
        xyz = np.asarray(mol_db.xyz_coordinates, dtype=np.float64)
        if hessian_to_predict is not None:
            # one (3Nat, 3Nat) array per molecule, stored like the reference does (kreg_api.py:130-131)
            mol_db.add_scalar_properties(self.hessian_arrays(xyz), property_name=hessian_to_predict)
        if not (want_e or want_g):
            return
        e, g = self.predict_arrays(xyz, want_e, want_g)
        if want_e:
            mol_db.add_scalar_properties(e, property_name=property_to_predict)
        if want_g:
            mol_db.add_xyz_vectorial_properties(g, xyz_vectorial_property=xyz_derivative_property_to_predict)

    # -- model file ---------------------------------------------------------------------------
    _KEYS = ["sigma", "Natoms", "Xsize", "Ntrain", "NtrVal", "NtrGrXYZ", "lambdav", "lambdagradxyz", "X", "XYZ",
             "ac2dArray", "Req", "Xeq", "alpha", "prior"]

    def save_model(self, filename):
        """np.savez with exactly the reference's keys (kreg_api.py:134-150)."""
        model_dict = {k: self.__dict__[k] for k in self._KEYS if self.__dict__.get(k) is not None}
        np.savez(filename, **model_dict)

    def load_model(self, filename):
        """kreg_api.py:152-156; files written by the reference load here and vice versa."""
        with np.load(filename) as model:
            # model files of the other model kinds carry extra keys; read as a plain KREG model they would silently
            # predict with the wrong kernel
            if type(self) is KREG_API and ("pmap" in model.files or "kernel" in model.files):
                kind = "perminv.PermInvKREG" if "pmap" in model.files else "kernels.RadialKREG"
                raise ValueError(f"{filename} is a {kind} model file: load it with that class (mlatom_b200.models.kreg does)")
            for key in model.files:
                self.__dict__[key] = model[key]
        self._model = None
        self._graphs = {}

    def save_unf(self, filename, atomic_numbers=None):
        """Write the model as an MLatomF ``.unf`` file (MLmodel.f90:819-967), readable by ``MLatomF useMLmodel``."""
        from . import unf

        prior = float(self.prior)
        unf.write_unf(filename, X=self.X, alpha=np.asarray(self.alpha).reshape(-1), sigma=float(self.sigma),
                      lambdav=float(self.lambdav), prior=prior, xyz=self.XYZ, atomic_numbers=atomic_numbers,
                      ntrval=int(self.NtrVal), ntrgr=int(self.NtrGrXYZ), lambdagradxyz=float(self.lambdagradxyz),
                      xyz_eq=self.Req, z_eq=atomic_numbers)

    def load_unf(self, filename):
        """Read an MLatomF ``.unf`` KREG model (RE 'unsorted' descriptor, Gaussian kernel)."""
        from . import unf

        m = unf.read_unf(filename)
        if "Req" not in m:
            raise ValueError("the .unf file carries no equilibrium geometry: the RE descriptor cannot be rebuilt")
        self.clean_up()
        self.sigma, self.lambdav, self.lambdagradxyz = m["sigma"], m["lambdav"], m["lambdagradxyz"]
        self.Natoms, self.Xsize, self.Ntrain = int(m["Natoms"]), int(m["Xsize"]), int(m["Ntrain"])
        self.NtrVal, self.NtrGrXYZ = int(m["NtrVal"]), int(m["NtrGrXYZ"])
        self.Ksize = self.NtrVal + self.NtrGrXYZ
        self.ac2dArray = ac2d_array(self.Natoms)
        self.Req, self.X, self.XYZ = m["Req"], m["X"], m.get("XYZ")
        self.alpha, self.prior = m["alpha"].reshape(1, -1), m["prior"]
        return self

    def clean_up(self):
        dev = self.device
        self.__init__(dev)
