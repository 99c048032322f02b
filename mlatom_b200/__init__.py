"""mlatom_b200 -- the KREG hot path of MLatom (RE descriptor + Gaussian kernel ridge regression:
kernel-matrix assembly, Cholesky solve, fused energy + gradient prediction) on NVIDIA B200 (sm_100a).

    from mlatom_b200 import models, data
    model = models.kreg(model_file='ethanol.npz', hyperparameters={'sigma': 1.0, 'lambda': 1e-6})
    model.train(molecular_database=db, property_to_learn='energy', xyz_derivative_property_to_learn='energy_gradients')
    model.predict(molecular_database=test_db, calculate_energy=True, calculate_energy_gradients=True)

Native code: mlatom_b200/libkreg_b200.so (C ABI in include/kreg_b200.h).  No CPU fallback exists.
"""
__version__ = "0.1.0"

from . import data, hyperparameters, synthetic  # noqa: F401  (pure Python, no device needed)


def __getattr__(name):
    # modules that need torch / the native library are imported on first use
    if name in ("models", "kreg_api", "engine", "dist", "_lib"):
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
