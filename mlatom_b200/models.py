"""``kreg`` -- drop-in for ``mlatom.models.kreg`` (reference: mlatom/models.py:343-552) running on the
B200 KREG kernels.

The constructor, ``train`` and ``predict`` signatures, the hyper-parameter objects, the ``.npz`` model
file and the convention that predictions are written into the molecule objects (gradients, not forces)
are those of the reference, so ``ml.md``, ``ml.optimize_geometry`` and the active-learning drivers can
use this class in place of ``ml.models.kreg``.  ``ml_program`` keeps its meaning as far as a user can see it:
'KREG_API' (default; also 'B200') stores models as ``.npz``, 'MLatomF' as MLatomF's unformatted ``.unf`` file (the
Fortran program itself is not part of this package; its files are read and written, its arithmetic is the device path)."""
from __future__ import annotations

import json
import os
import uuid
from typing import Any, Dict, Optional, Union

import numpy as np
import torch

from . import data as _data
from . import engine
from ._lib import FILL_LOWER, KregError
from .hyperparameters import hyperparameter, hyperparameters
from .kreg_api import KREG_API, _as_database


def rmse(a, b) -> float:
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.sqrt(np.mean((a - b) ** 2)))


def optimize_grid(func, grid):
    """Minimum of ``func`` over the Cartesian product of ``grid`` visited in lexicographic order with the
    first hyper-parameter outermost; ties keep the earliest point (same visiting order and strict '<'
    as the reference's recursive optimize_grid, mlatom/model_cls.py:819-854)."""
    best, best_params = None, None

    def rec(prefix, rest):
        nonlocal best, best_params
        if not rest:
            val = func(list(prefix))
            if best is None or val < best:
                best, best_params = val, list(prefix)
            return
        for p in rest[0]:
            rec(prefix + [p], rest[1:])

    rec([], [list(g) for g in grid])
    return best_params, best


class kreg:
    """KREG model object (Gaussian kernel on the RE descriptor)."""

    hyperparameters = hyperparameters({
        "lambda": hyperparameter(value=2 ** -35, minval=2 ** -35, maxval=1.0, optimization_space="log", name="lambda"),
        "sigma": hyperparameter(value=1.0, minval=2 ** -5, maxval=2 ** 9, optimization_space="log", name="sigma"),
    })
    nthreads = 0

    def __init__(self, model_file: Optional[str] = None, ml_program: str = "KREG_API", equilibrium_molecule=None,
                 prior: float = 0, nthreads: Optional[int] = None,
                 hyperparameters: Union[Dict[str, Any], "hyperparameters"] = {}, device=None):
        self.model_file = model_file
        self.equilibrium_molecule = equilibrium_molecule
        self.ml_program = ml_program
        self.device = device
        self.prior = prior
        if ml_program.casefold() not in ("kreg_api", "b200", "mlatomf"):
            raise ValueError(f"ml_program={ml_program!r} is not available in mlatom_b200 (use 'KREG_API' or 'MLatomF')")
        # 'MLatomF' (models.py:376-378, 478-496, 542-546) selects MLatomF's model-file format, the unformatted .unf of
        # MLmodel.f90:819-967, instead of the .npz; the arithmetic is the same device path either way (the two
        # reference programs implement the same equations)
        self._unf = ml_program.casefold() == "mlatomf"
        if self.model_file is not None:
            if not self._unf and self.model_file[-4:] != ".npz":
                self.model_file += ".npz"
            if os.path.exists(self.model_file):
                self.kreg_api = KREG_API(device)
                if self._unf and self.model_file[-4:] != ".npz":
                    self.kreg_api.load_unf(self.model_file)
                else:
                    with np.load(self.model_file) as header:
                        permuted, other = "pmap" in header.files, "kernel" in header.files
                    if permuted:  # a permutationally invariant model (perminv.PermInvKREG.save_model)
                        from .perminv import PermInvKREG

                        self.kreg_api = PermInvKREG(device=device)
                    elif other:   # a model on another kernel function (kernels.RadialKREG.save_model)
                        from .kernels import RadialKREG

                        self.kreg_api = RadialKREG(device=device)
                    self.kreg_api.load_model(self.model_file)
        self.hyperparameters = type(self).hyperparameters.copy()
        self.hyperparameters.update(hyperparameters)
        self.nthreads = nthreads

    # -- small compat surface (mlatom/model_cls.py:20-38, 392-427) --------------------------------
    def set_num_threads(self, nthreads=0):
        if nthreads:
            self.nthreads = nthreads

    def config_multiprocessing(self):
        """Nothing to prepare before the model is used from parallel drivers (model_cls.py:29-33)."""

    def parse_args(self, args):
        """Command-line parsing belongs to MLatom's CLI layer, which is outside this package (model_cls.py:35-37)."""

    def _predict_geomopt(self, return_string=False, dump_trajectory_interval=None, filename=None, format="json",
                         print_properties=None, molecule=None, **kwargs):
        """The hook MLatom's geometry optimisers call for every step (model_cls.py:39-76, simulations.py:382):
        predict, warn about missing gradients, and -- when a dump interval is given -- append the step to the
        optimisation trajectory file using the trajectory classes of the data module the molecule comes from."""
        self.predict(molecule=molecule, **kwargs)
        if np.any(np.isnan(molecule.energy_gradients)):
            print(" * Warning* No gradients were calculated, check the logs for any reasons of this critical failure.")
            if "error_message" in molecule.__dict__:
                print(" Error message retrieved from the calculations:\n" + "-" * 10)
                print(molecule.error_message)
                print("-" * 10)
        if dump_trajectory_interval is None:
            return None
        import sys

        dmod = sys.modules[type(molecule).__module__]
        if not hasattr(dmod, "molecular_trajectory"):
            raise NotImplementedError("trajectory dumps need MLatom's data module (mlatom.data molecules)")
        traj = dmod.molecular_trajectory()
        traj.load(filename=filename, format=format)
        nsteps = len(traj.steps)
        text = None
        if print_properties == "all" or isinstance(print_properties, list):
            rule = " %s " % ("-" * 78)
            text = "\n".join([rule, f" Iteration {nsteps + 1}", rule + "\n",
                              molecule.info(properties=print_properties, return_string=True)]) + "\n"
            if not return_string:
                print(text)
        traj.steps.append(dmod.molecular_trajectory_step(step=nsteps, molecule=molecule))
        traj.dump(filename=filename, format=format)
        moldb = dmod.molecular_database()
        moldb.molecules = [step.molecule for step in traj.steps]
        moldb.write_file_with_xyz_coordinates(os.path.splitext(os.path.basename(filename))[0] + ".xyz")
        return text if return_string else None

    def reset(self):
        if self.model_file and os.path.exists(self.model_file):
            os.remove(self.model_file)

    def generate_model_dict(self):
        return {"type": "ml_model", "ml_model_type": "mlatom_b200.models.kreg",
                "kwargs": {"model_file": os.path.abspath(self.model_file) if self.model_file else None,
                           "ml_program": self.ml_program}, "nthreads": self.nthreads}

    def dump(self, filename=None, format="json"):
        model_dict = self.generate_model_dict()
        if format == "dict":
            return model_dict
        with open(filename, "w") as f:
            json.dump(model_dict, f, indent=4)

    # -- equilibrium geometry (mlatom/models.py:439-457) ---------------------------------------------
    def get_descriptor(self, molecule=None, molecular_database=None, descriptor_name="re", equilibrium_molecule=None):
        """RE descriptor of every molecule, stored as ``mol.<descriptor_name>`` and ``mol.descriptor`` like the
        reference does (mlatom/models.py:402-437) -- computed by ``kreg_descriptors`` on the device, one batch."""
        if molecular_database is not None:
            db = molecular_database
        elif molecule is not None:
            db = _data.molecular_database([molecule])
        else:
            raise ValueError("Either molecule or molecular_database should be provided in input")
        if getattr(self, "reference_distance_matrix", None) is None:
            if equilibrium_molecule is None:
                equilibrium_molecule = self.get_equilibrium_molecule(molecular_database=db)
            self.reference_distance_matrix = equilibrium_molecule.get_internuclear_distance_matrix()
            self._reference_xyz = np.asarray(equilibrium_molecule.xyz_coordinates, dtype=np.float64)
        api = self.__dict__.get("kreg_api") or KREG_API(self.device)
        dev = api._dev()
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        x = engine.descriptors(t(db.xyz_coordinates), engine.xeq(t(self._reference_xyz))).cpu().numpy()
        for mol, row in zip(db.molecules, x):
            setattr(mol, descriptor_name, row)
            mol.descriptor = row

    def get_equilibrium_molecule(self, molecular_database=None):
        first = molecular_database.molecules[0].__dict__
        for name in ("energy", "y"):
            if name in first:
                values = molecular_database.get_properties(property_name=name)
                return molecular_database.molecules[int(np.argmin(values))]
        raise ValueError("equilibrium molecule is not provided and no energies are found in molecular database")

    def get_equilibrium_distances(self, molecular_database=None):
        if self.equilibrium_molecule is None:
            self.equilibrium_molecule = self.get_equilibrium_molecule(molecular_database=molecular_database)
        x = np.asarray(self.equilibrium_molecule.xyz_coordinates)
        dm = np.linalg.norm(x[:, None] - x[None], axis=-1)
        return dm[np.triu_indices(len(x), 1)]

    # -- train / predict (mlatom/models.py:459-552) ------------------------------------------------------
    def _lambdas_and_prior(self, prior):
        lam = self.hyperparameters["lambda"].value
        lam_g = self.hyperparameters["lambdaGradXYZ"].value if "lambdaGradXYZ" in self.hyperparameters.keys() else lam
        if "prior" in self.hyperparameters.keys():
            prior = self.hyperparameters["prior"].value
        return lam, lam_g, prior

    def _other_kernel(self):
        """(kernel, nn) of an MLatomF ``kernel=`` request other than Gaussian in ``additional_mlatomf_args`` (else None)."""
        kernel, nn = None, 2  # MLatomF's default order of the Matern kernel (optionsModule.f90:105)
        for a in [str(a) for a in self.__dict__.get("additional_mlatomf_args", None) or []]:
            key, _, val = a.partition("=")
            if key.casefold() == "kernel":
                kernel = val
            elif key.casefold() == "nn":
                nn = int(val)
        if kernel is None or kernel.casefold() == "gaussian":
            return None
        if kernel.casefold() not in ("exponential", "matern", "laplacian"):
            raise NotImplementedError(f"kernel={kernel} is not built (Gaussian, exponential, Matern, Laplacian are)")
        return kernel, nn

    def _perm_inv_groups(self):
        """Atom groups of MLatomF's ``permInvKernel`` request in ``additional_mlatomf_args`` (None: not requested).
        Arguments that would select a model this package does not build raise instead of being dropped silently."""
        args = [str(a) for a in self.__dict__.get("additional_mlatomf_args", None) or []]
        for a in args:
            key, _, val = a.casefold().partition("=")
            if (key == "moldescriptor" and val != "re") or (key == "moldescrtype" and val not in ("unsorted", "permuted")) \
                    or key in ("selectperm", "periodkernel", "decaykernel", "usepermind"):
                raise NotImplementedError(f"MLatomF argument {a!r} selects a model outside this package (RE descriptor, "
                                          "Gaussian / exponential / Matern / Laplacian kernels, permInvKernel with permInvNuclei)")
        if not any(a.casefold() == "perminvkernel" for a in args):
            if any(a.casefold() == "moldescrtype=permuted" for a in args):
                raise NotImplementedError("molDescrType=permuted without permInvKernel (the plain kernel on the concatenated "
                                          "permuted descriptor) is not built")
            return None
        from .perminv import parse_perm_inv_nuclei

        for a in args:
            if a.casefold().startswith("perminvgroups="):
                raise NotImplementedError("permInvGroups (permuting whole groups of atoms) is not built; use permInvNuclei")
        for a in args:
            if a.casefold().startswith("perminvnuclei="):
                return parse_perm_inv_nuclei(a.split("=", 1)[1])
        raise ValueError("permInvKernel needs permInvNuclei=<atoms separated by '-', groups by '.'>")

    def train(self, molecular_database=None, property_to_learn=None, xyz_derivative_property_to_learn=None,
              save_model=True, invert_matrix=False, matrix_decomposition=None, prior=None,
              hyperparameters: Union[Dict[str, Any], "hyperparameters"] = {}):
        self.hyperparameters.update(hyperparameters)
        if save_model:
            if self.model_file is None:
                self.model_file = f"kreg_{uuid.uuid4()}" + (".unf" if self._unf else ".npz")
            elif not self._unf and self.model_file[-4:] != ".npz":
                self.model_file += ".npz"
        groups = self._perm_inv_groups()
        other = self._other_kernel()
        if other is not None:
            # MLatomF's other kernel functions (additional_mlatomf_args=['kernel=Matern', 'nn=2'], models.py:494-495)
            from .kernels import RadialKREG

            if groups is not None:
                raise NotImplementedError("permInvKernel is built for the Gaussian kernel only")
            self.kreg_api = RadialKREG(kernel=other[0], nn=other[1], device=self.device)
            if save_model and self.model_file[-4:] != ".npz":
                self.model_file += ".npz"
            groups = ()  # not a .unf model either
        elif groups is not None:
            # MLatomF's permutationally invariant kernel (additional_mlatomf_args=['permInvKernel', 'permInvNuclei=..',
            # 'molDescrType=permuted'], models.py:494-495): same device path with the expanded training set, .npz model file
            from .perminv import PermInvKREG

            self.kreg_api = PermInvKREG(groups=groups, device=self.device)
            if save_model and self.model_file[-4:] != ".npz":
                self.model_file += ".npz"
        else:
            self.kreg_api = KREG_API(self.device)
        self.kreg_api.nthreads = self.nthreads
        if self.equilibrium_molecule is None:
            self.equilibrium_molecule = self.get_equilibrium_molecule(molecular_database=molecular_database)
        lam, lam_g, prior = self._lambdas_and_prior(prior)
        self.kreg_api.train(self.hyperparameters["sigma"].value, lam, lam_g, molecular_database,
                            self.equilibrium_molecule, property_to_learn=property_to_learn,
                            xyz_derivative_property_to_learn=xyz_derivative_property_to_learn, prior=prior)
        if save_model:
            if self._unf and groups is None:
                self.kreg_api.save_unf(self.model_file, atomic_numbers=np.asarray(self.equilibrium_molecule.atomic_numbers))
            else:
                self.kreg_api.save_model(self.model_file)

    def predict(self, molecular_database=None, molecule=None, calculate_energy=False, calculate_energy_gradients=False,
                calculate_hessian=False, property_to_predict=None, xyz_derivative_property_to_predict=None,
                hessian_to_predict=None):
        mol_db = _as_database(molecular_database, molecule)
        # ml_model.predict's naming rule (mlatom/model_cls.py:366-390)
        if calculate_energy:
            property_to_predict = "energy"
        if calculate_energy_gradients:
            xyz_derivative_property_to_predict = "energy_gradients"
        if calculate_hessian:
            hessian_to_predict = "hessian"
        if "kreg_api" not in self.__dict__:
            raise RuntimeError("KREG_API model not found")
        self.kreg_api.nthreads = self.nthreads
        self.kreg_api.predict(molecular_database=mol_db, property_to_predict=property_to_predict,
                              xyz_derivative_property_to_predict=xyz_derivative_property_to_predict,
                              hessian_to_predict=hessian_to_predict)

    def predict_xyz(self, xyz, calculate_energy=True, calculate_energy_gradients=True):
        """Tensor entry point for large batches: xyz (N,Nat,3) cuda fp64 tensor or numpy -> (E, dE/dxyz).
        (Python molecule objects cannot carry 10^6-10^7 geometries)"""
        return self.kreg_api.predict_arrays(xyz, calculate_energy, calculate_energy_gradients)

    # -- validation / hyper-parameter search (mlatom/model_cls.py:450-739) -------------------------------------
    def holdout_validation(self, subtraining_molecular_database=None, validation_molecular_database=None,
                           training_kwargs=None, prediction_kwargs=None):
        self.train(molecular_database=subtraining_molecular_database, **(training_kwargs or {}))
        self.predict(molecular_database=validation_molecular_database, **(prediction_kwargs or {}))

    @staticmethod
    def _names(training_kwargs, prediction_kwargs):
        tk, pk = dict(training_kwargs or {}), dict(prediction_kwargs or {})
        p_learn, g_learn = tk.get("property_to_learn"), tk.get("xyz_derivative_property_to_learn")
        if p_learn is None and g_learn is None:
            p_learn = tk["property_to_learn"] = "y"
        p_pred, g_pred = pk.get("property_to_predict"), pk.get("xyz_derivative_property_to_predict")
        if p_pred is None and g_pred is None:
            if p_learn is not None:
                p_pred = pk["property_to_predict"] = f"estimated_{p_learn}"
            if g_learn is not None:
                g_pred = pk["xyz_derivative_property_to_predict"] = f"estimated_{g_learn}"
        return tk, pk, p_learn, g_learn, p_pred, g_pred

    def cross_validation(self, cv_splits_molecular_databases=None, training_kwargs=None, prediction_kwargs=None):
        """k-fold cross-validation (mlatom/model_cls.py:742-770): for every split, train on the union of the others and
        predict the split; predictions are left in the splits' molecules."""
        tk, pk = dict(training_kwargs or {}), dict(prediction_kwargs or {})
        splits = list(cv_splits_molecular_databases)
        for ii, held_out in enumerate(splits):
            sub = type(held_out)()
            for jj, other in enumerate(splits):
                if ii != jj:
                    sub.molecules += other.molecules
            self.reset()
            self.train(molecular_database=sub, **tk)
            self.predict(molecular_database=held_out, **pk)

    def calculate_validation_loss(self, training_kwargs=None, prediction_kwargs=None,
                                  cv_splits_molecular_databases=None, calculate_CV_split_errors=False,
                                  subtraining_molecular_database=None, validation_molecular_database=None,
                                  validation_loss_function=None, validation_loss_function_kwargs={}, debug=False,
                                  **unsupported):
        """Validation loss for the current hyper-parameters: RMSE, geometric mean over scalar and XYZ-vectorial
        properties (mlatom/model_cls.py:450-607) -- hold-out (sub-training / validation databases) or k-fold
        cross-validation over ``cv_splits_molecular_databases`` (then over all splits' predictions together, and with
        ``calculate_CV_split_errors`` also per split, returned as ``(error, [split errors])``)."""
        for k, v in unsupported.items():  # the ml_database variants (cv_splits_ml_databases, subtraining_ml_database, ...)
            if v is not None and v is not False:
                raise NotImplementedError(f"{k} is not supported by mlatom_b200.kreg (use the molecular_database arguments)")
        tk, pk, p_learn, g_learn, p_pred, g_pred = self._names(training_kwargs, prediction_kwargs)

        def geom_rmse(dbs):
            total = 1.0
            if p_learn is not None:
                total *= rmse(np.concatenate([np.asarray(d.get_properties(p_pred)).reshape(-1) for d in dbs]),
                              np.concatenate([np.asarray(d.get_properties(p_learn)).reshape(-1) for d in dbs]))
            if g_learn is not None:
                total *= rmse(np.concatenate([np.asarray(d.get_xyz_vectorial_properties(g_pred)).reshape(-1) for d in dbs]),
                              np.concatenate([np.asarray(d.get_xyz_vectorial_properties(g_learn)).reshape(-1) for d in dbs]))
            if p_learn is not None and g_learn is not None:
                total = float(np.sqrt(total))
            return total

        split_errors = None
        if cv_splits_molecular_databases is None:
            self.holdout_validation(subtraining_molecular_database, validation_molecular_database, tk, pk)
            scored = [validation_molecular_database]
        else:
            self.cross_validation(cv_splits_molecular_databases, tk, pk)
            scored = list(cv_splits_molecular_databases)
        error = geom_rmse(scored) if validation_loss_function is None else validation_loss_function(**validation_loss_function_kwargs)
        self.reset()
        if cv_splits_molecular_databases is not None and calculate_CV_split_errors:
            split_errors = [geom_rmse([d]) if validation_loss_function is None
                            else validation_loss_function(**validation_loss_function_kwargs) for d in scored]
        if debug:
            for name in self.hyperparameters.keys():
                print(f"  Hyperparameter {name} = {self.hyperparameters[name].value}")
            print(f"    Validation loss: {error}")
        return (error, split_errors) if split_errors is not None else error

    def _grid(self, names, grid_size):
        slices = []
        for key in names:
            hp = self.hyperparameters[key]
            if hp.optimization_space == "linear":
                slices.append(list(np.linspace(hp.minval, hp.maxval, num=grid_size)))
            elif hp.optimization_space == "log":
                slices.append(list(np.logspace(np.log(hp.minval), np.log(hp.maxval), num=grid_size, base=np.exp(1))))
            else:
                raise NotImplementedError(hp.optimization_space)
        return slices

    def optimize_hyperparameters(self, hyperparameters=None, training_kwargs=None, prediction_kwargs=None,
                                 subtraining_molecular_database=None, validation_molecular_database=None,
                                 optimization_algorithm=None, optimization_algorithm_kwargs={}, maximum_evaluations=10000,
                                 validation_loss_function=None, validation_loss_function_kwargs={}, debug=False,
                                 fused=True, cv_splits_molecular_databases=None, **unsupported):
        """Minimise the hold-out validation loss over the named hyper-parameters, then retrain with the
        winners (mlatom/model_cls.py:609-727).  'grid' over 'lambda'/'sigma' takes the fused device path
        (K assembled once per sigma, re-factorised per lambda, validation kernels stay in HBM); every other
        case evaluates train+predict per point exactly like the reference."""
        names = list(hyperparameters)
        algo = optimization_algorithm.casefold()
        vkw = dict(training_kwargs=training_kwargs, prediction_kwargs=prediction_kwargs,
                   subtraining_molecular_database=subtraining_molecular_database,
                   validation_molecular_database=validation_molecular_database,
                   validation_loss_function=validation_loss_function,
                   validation_loss_function_kwargs=validation_loss_function_kwargs, debug=debug,
                   cv_splits_molecular_databases=cv_splits_molecular_databases, **unsupported)

        def loss_at(values):
            for name, v in zip(names, values):
                self.hyperparameters[name].value = v
            return self.calculate_validation_loss(**vkw)

        saved_name = self.model_file
        import tempfile

        with tempfile.TemporaryDirectory() as tmp:
            self.model_file = os.path.join(tmp, os.path.basename(saved_name or "kreg_tmp.npz"))
            if algo in ("grid", "brute"):
                slices = self._grid(names, optimization_algorithm_kwargs.get("grid_size", 9))
                # the fused path assembles the Gaussian KREG kernel or (values-trained) one of MLatomF's other kernel
                # functions; models on the permutationally invariant kernel take the pointwise route (train + predict per
                # grid point)
                plain_kernel = self._perm_inv_groups() is None
                can_fuse = (fused and plain_kernel and set(names) <= {"lambda", "sigma"} and validation_loss_function is None
                            and cv_splits_molecular_databases is None and not any(v for v in unsupported.values()))
                if can_fuse:
                    table = self._fused_grid_losses(names, slices, training_kwargs, prediction_kwargs,
                                                    subtraining_molecular_database, validation_molecular_database)
                    self.grid_losses = table
                    params, _ = optimize_grid(lambda p: table[tuple(p)], slices)
                else:
                    params, _ = optimize_grid(loss_at, slices)
            elif algo in ("log2grid", "lggrid"):
                # MLatomF's own optimiser (sigma=opt lambda=opt with lgOptDepth levels); not offered by the reference's
                # Python class, available here because the fused path makes it cheap
                if not set(names) <= {"lambda", "sigma"}:
                    raise NotImplementedError("log2grid optimises 'sigma' and/or 'lambda'")
                if self._perm_inv_groups() is not None:
                    raise NotImplementedError("log2grid runs on the fused path, which the permutationally invariant kernel "
                                              "does not have; use 'grid'")
                best, _, self.grid_trace = self._log2_grid_search(names, optimization_algorithm_kwargs, training_kwargs,
                                                                  prediction_kwargs, subtraining_molecular_database,
                                                                  validation_molecular_database)
                params = [best[n] for n in names]
            elif algo == "tpe":
                # Tree-structured Parzen estimator through hyperopt, exactly the reference's call (model_cls.py:688-727);
                # hyperopt is an optional dependency there too (a plain ImportError if it is not installed)
                import hyperopt

                def space_of(key):
                    hp = self.hyperparameters[key]
                    if hp.optimization_space == "log":
                        return hyperopt.hp.loguniform(key, np.log(hp.minval), np.log(hp.maxval))
                    if hp.optimization_space == "linear":
                        return hyperopt.hp.uniform(key, hp.minval, hp.maxval)
                    raise NotImplementedError(hp.optimization_space)

                res = hyperopt.fmin(fn=lambda d: loss_at([d[k] for k in names]), space={k: space_of(k) for k in names},
                                    algo=hyperopt.tpe.suggest, max_evals=maximum_evaluations, show_progressbar=False)
                params = [res[k] for k in names]
            else:
                import scipy.optimize

                x0 = np.array([self.hyperparameters[k].value for k in names])
                bounds = np.array([[self.hyperparameters[k].minval, self.hyperparameters[k].maxval] for k in names])
                res = scipy.optimize.minimize(loss_at, x0, method=optimization_algorithm, bounds=bounds,
                                              options={"maxiter": maximum_evaluations})
                params = list(res.x)
            for name, v in zip(names, params):
                self.hyperparameters[name].value = v
            self.model_file = saved_name
        self.validation_loss = loss_at([self.hyperparameters[k].value for k in names])

    def _fused_evaluator(self, training_kwargs, prediction_kwargs, sub_db, val_db):
        """Everything of the hold-out loss that does not depend on (sigma, lambda), resident on the device."""
        tk, _, p_learn, g_learn, _, _ = self._names(training_kwargs, prediction_kwargs)
        api = KREG_API(self.device)
        dev = api._dev()
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        if self.equilibrium_molecule is None:
            self.equilibrium_molecule = self.get_equilibrium_molecule(molecular_database=sub_db)
        _, _, prior = self._lambdas_and_prior(tk.get("prior"))
        fixed_lam_g = self.hyperparameters["lambdaGradXYZ"].value if "lambdaGradXYZ" in self.hyperparameters.keys() else None
        other = self._other_kernel()
        if other is not None and g_learn is not None:
            raise NotImplementedError("models on the non-Gaussian kernels are trained on values only here")
        return _FusedHoldout(t, self.equilibrium_molecule.xyz_coordinates, sub_db, val_db, p_learn, g_learn, prior,
                             fixed_lam_g, kernel=other)

    def _fused_grid_losses(self, names, slices, training_kwargs, prediction_kwargs, sub_db, val_db):
        """All (lambda, sigma) grid losses with K(sigma) built once per sigma."""
        ev = self._fused_evaluator(training_kwargs, prediction_kwargs, sub_db, val_db)
        current = {k: self.hyperparameters[k].value for k in ("lambda", "sigma")}
        sig_list = slices[names.index("sigma")] if "sigma" in names else [current["sigma"]]
        lam_list = slices[names.index("lambda")] if "lambda" in names else [current["lambda"]]
        # grid points are independent: with torch.distributed initialised every rank takes a round-robin share of the
        # (sigma, lambda) points (grouped by sigma so K(sigma) is still assembled once per sigma it needs) and the
        # losses are all-gathered; single process: all points
        from . import dist as kdist

        points = [(si, li) for si in range(len(sig_list)) for li in range(len(lam_list))]
        mine = set(kdist.grid_shard(points))
        local = {}
        for si, sigma in enumerate(sig_list):
            todo = [li for li in range(len(lam_list)) if points.index((si, li)) in mine]
            if not todo:
                continue
            ev.set_sigma(sigma)
            for li in todo:
                local[points.index((si, li))] = ev.loss(lam_list[li])
        losses = kdist.allgather_losses(local, len(points))
        table = {}
        for idx, (si, li) in enumerate(points):
            key = tuple(lam_list[li] if n == "lambda" else sig_list[si] for n in names)
            table[key] = losses[idx]
        return table

    def _log2_grid_search(self, names, kwargs, training_kwargs, prediction_kwargs, sub_db, val_db):
        """MLatomF's recursive logarithmic grid (optLog2Grid / optLambdaLog2Grid, A_KRR.f90:1049-1376): sigma on a
        log2 grid; for every sigma, lambda on its own log2 grid re-using K(sigma); after each level the range shrinks
        to one grid step either side of the best point; ``depth`` levels (lgOptDepth, default 3); ties keep the first
        minimum.  Returns ({name: value}, loss, trace)."""
        depth = int(kwargs.get("depth", 3))
        pts = kwargs.get("points", {})
        default_pts = {"sigma": 11, "lambda": 6}  # the ranges MLatomF spans by default at unit log2 steps (-35..-6, 2..9)
        npts = {k: int(pts.get(k, kwargs.get("grid_size", default_pts[k]))) for k in ("sigma", "lambda")}
        lg = lambda k: (float(np.log2(self.hyperparameters[k].minval)), float(np.log2(self.hyperparameters[k].maxval)))
        ev = self._fused_evaluator(training_kwargs, prediction_kwargs, sub_db, val_db)
        trace = []

        def level_points(lo, hi, n):
            step = 0.0 if n == 1 else (hi - lo) / (n - 1)
            return step, [lo + step * i for i in range(n)]

        def best_lambda(sigma):
            if "lambda" not in names:
                lam = self.hyperparameters["lambda"].value
                return lam, ev.loss(lam)
            lo, hi = lg("lambda")  # reset for every sigma (getLambda, A_KRR.f90:963-981)
            best = None
            for _ in range(depth):
                step, grid = level_points(lo, hi, npts["lambda"])
                lvl = None
                for x in grid:
                    err = ev.loss(2.0 ** x)
                    trace.append((sigma, 2.0 ** x, err))
                    if lvl is None or err < lvl[1]:
                        lvl = (x, err)
                best = lvl
                lo, hi = best[0] - step, best[0] + step
            return 2.0 ** best[0], best[1]

        if "sigma" not in names:
            ev.set_sigma(self.hyperparameters["sigma"].value)
            lam, err = best_lambda(self.hyperparameters["sigma"].value)
            return {"lambda": lam}, err, trace
        lo, hi = lg("sigma")
        best = None
        for _ in range(depth):
            step, grid = level_points(lo, hi, npts["sigma"])
            lvl = None
            for x in grid:
                sigma = 2.0 ** x
                ev.set_sigma(sigma)
                lam, err = best_lambda(sigma)
                if lvl is None or err < lvl[2]:
                    lvl = (x, lam, err)
            best = lvl
            lo, hi = best[0] - step, best[0] + step
        out = {"sigma": 2.0 ** best[0]}
        if "lambda" in names:
            out["lambda"] = best[1]
        return out, best[2], trace


class _FusedHoldout:
    """Hold-out validation loss as a function of lambda at a fixed sigma, everything resident in HBM: K(sigma) is
    assembled once by ``set_sigma``; ``loss(lambda)`` shifts the diagonal of a copy, factorises it, predicts the
    validation set and returns the reference's loss (RMSE, geometric mean of scalar and vectorial parts,
    model_cls.py:577-585).  A matrix that is not positive definite at some lambda is solved by the Bunch-Kaufman
    fallback, as the reference does (mathUtils.f90:24-26)."""

    def __init__(self, t, eq_xyz, sub_db, val_db, p_learn, g_learn, prior, fixed_lam_g, kernel=None):
        self.p_learn, self.g_learn, self.fixed_lam_g = p_learn, g_learn, fixed_lam_g
        self.kernel = kernel  # None: the Gaussian KREG model; (name, nn): values-trained model on that kernel function
        self.xeq = engine.xeq(t(eq_xyz))
        self.xyz, self.xyz_val = t(sub_db.xyz_coordinates), t(val_db.xyz_coordinates)
        ntr, nat = int(self.xyz.shape[0]), int(self.xyz.shape[1])
        parts, self.prior_value = [], 0.0
        if p_learn is not None:
            y = np.asarray(sub_db.get_properties(p_learn), dtype=np.float64)
            self.prior_value = float(np.mean(y)) if isinstance(prior, str) and prior.casefold() == "mean" else float(prior or 0.0)
            parts.append(y - self.prior_value)
        if g_learn is not None:
            parts.append(np.asarray(sub_db.get_xyz_vectorial_properties(g_learn), dtype=np.float64).reshape(-1))
        self.ytrain = t(np.concatenate(parts))
        self.nv, self.ng = (ntr if p_learn is not None else 0), (3 * nat * ntr if g_learn is not None else 0)
        self.y_val = np.asarray(val_db.get_properties(p_learn)) if p_learn is not None else None
        self.g_val = np.asarray(val_db.get_xyz_vectorial_properties(g_learn)) if g_learn is not None else None
        m = self.nv + self.ng
        dev = self.xyz.device
        if kernel is not None:
            self.k0 = None  # kernel_block_ex returns K(sigma)
            self.work = torch.empty((m, m), dtype=torch.float64, device=dev)
        elif self.xyz.is_cuda:
            self.k0 = engine.empty_kernel_matrix(m, self.nv if self.ng else 0, dev)
            self.work = engine.empty_kernel_matrix(m, self.nv if self.ng else 0, dev)
        else:
            self.k0 = torch.empty((m, m), dtype=torch.float64, device=dev)
            self.work = torch.empty_like(self.k0)
        self.diag = torch.arange(m, device=dev)
        self.sigma = None
        self.fallbacks = 0  # grid points solved by the symmetric-indefinite route
        ws_bytes = engine.build_workspace_bytes(nat, ntr, self.ng) if self.xyz.is_cuda and kernel is None else 0
        self.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None  # one scratch for every sigma

    def set_sigma(self, sigma):
        self.sigma = float(sigma)
        if self.kernel is not None:
            self.k0 = None  # release the previous sigma's matrix before the next one is allocated
            self.k0 = engine.kernel_block_ex(self.kernel[0], self.xyz, self.xyz, self.xeq, self.sigma, nn=self.kernel[1],
                                             symmetric=True, lam=0.0)
            return
        engine.build_kernel_matrix(self.xyz, self.xeq, self.sigma, 0.0, 0.0, use_values=self.nv > 0,
                                   use_gradients=self.ng > 0, fill=FILL_LOWER, out=self.k0, workspace=self.ws)

    def loss(self, lam):
        nv, diag, work = self.nv, self.diag, self.work
        lam_g = lam if self.fixed_lam_g is None else self.fixed_lam_g
        work.copy_(self.k0)
        work[diag[:nv], diag[:nv]] += lam
        work[diag[nv:], diag[nv:]] += lam_g
        try:
            alpha = engine.solve(work, self.ytrain)
        except KregError as err:
            if err.code <= 0:
                raise
            # not positive definite at this (lambda, sigma) -- the regime the default lambda = 2^-35 visits: the same
            # Bunch-Kaufman fallback as KREG_API.train_arrays (mathUtils.f90:24-26), on a fresh copy of K(sigma)
            work.copy_(self.k0)
            work[diag[:nv], diag[:nv]] += lam
            work[diag[nv:], diag[nv:]] += lam_g
            try:
                alpha = engine.solve_indefinite(work, self.ytrain)
            except KregError as err2:
                if err2.code <= 0:
                    raise
                return float("inf")  # exactly singular pivot: no solution to score
            self.fallbacks += 1
        if self.kernel is not None:
            from .kernels import RadialModel

            model = RadialModel(self.kernel[0], self.kernel[1], self.xyz, self.xeq, alpha, self.sigma, self.prior_value)
        else:
            model = engine.PackedModel.from_training_set(self.xyz, self.xeq, alpha, self.sigma, self.prior_value, nv, self.ng)
        e, g = model.predict(self.xyz_val, self.p_learn is not None, self.g_learn is not None)
        total = 1.0
        if self.p_learn is not None:
            total *= rmse(e.cpu().numpy(), self.y_val)
        if self.g_learn is not None:
            total *= rmse(g.cpu().numpy(), self.g_val)
        if self.p_learn is not None and self.g_learn is not None:
            total = float(np.sqrt(total))
        return total
