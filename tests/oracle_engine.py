"""TEST INFRASTRUCTURE: replaces the device engine of mlatom_b200 by oracle-backed fakes (numpy / SciPy on the CPU) so
that the host-side logic of the drop-in API can be exercised without a GPU -- by the CPU tests and by
tools/make_golden_grid.py.  Never imported by the product package."""
import numpy as np
import torch

from mlatom_b200 import kreg_api as KA
from mlatom_b200._lib import KregError


class FakePacked:
    def __init__(self, oracle, xyz_tr, xeq, alpha, sigma, prior, nv, ng):
        self.o, self.sigma, self.prior, self.device = oracle, sigma, prior, torch.device("cpu")
        self.xeq = xeq.numpy()
        self.xyz_tr = xyz_tr.numpy()
        self.X = oracle.descriptors(self.xyz_tr, self.xeq)
        a = alpha.numpy().reshape(-1)
        self.av = a[:nv] if nv else None
        self.v = oracle.precompute_v(self.xyz_tr, self.X, a[nv:]) if ng else None

    def predict(self, xyz, energy=True, gradient=True, **_):
        x = xyz.numpy()
        e, g = self.o.predict_factored(self.X, self.av, self.v, self.sigma, self.prior, x, self.o.descriptors(x, self.xeq))
        return (torch.from_numpy(e) if energy else None), (torch.from_numpy(g) if gradient else None)

    def predict_host(self, xyz, energy=True, gradient=True, **_):
        return self.predict(torch.from_numpy(np.ascontiguousarray(xyz)), energy, gradient)


def install_oracle_engine(setattr_, oracle):
    """setattr_(obj, name, value): pytest's monkeypatch.setattr, or plain setattr for a script."""
    eng = KA.engine
    setattr_(KA.KREG_API, "_dev", lambda self: torch.device("cpu"))
    setattr_(torch.cuda, "synchronize", lambda *a, **k: None)
    setattr_(eng, "xeq", lambda req: torch.from_numpy(oracle.xeq(req.numpy())))
    setattr_(eng, "descriptors", lambda xyz, xeq: torch.from_numpy(oracle.descriptors(xyz.numpy(), xeq.numpy())))

    def build(xyz, xeq, sigma, lv, lg, use_values=True, use_gradients=False, fill=0, rows=None, out=None, workspace=None):
        x = xyz.numpy()
        n, nat, _ = x.shape
        nv, ng = (n if use_values else 0), (3 * nat * n if use_gradients else 0)
        k = oracle.kernel_matrix(x, oracle.descriptors(x, xeq.numpy()), sigma, nv, ng, factored=True)
        k[np.diag_indices_from(k)] += np.r_[np.full(nv, lv), np.full(ng, lg)]
        k = torch.from_numpy(k)
        if out is not None:
            out.copy_(k)
            return out
        return k

    def _sym(k):
        a = np.tril(k.numpy())
        return a + np.tril(a, -1).T

    def solve(k, y):
        """dpotrf/dpotrs like mathUtils.f90:65,79; a failed factorisation surfaces as the ABI's positive info code."""
        from scipy.linalg import LinAlgError, cho_factor, cho_solve

        try:
            return torch.from_numpy(cho_solve(cho_factor(_sym(k)), y.numpy()))
        except LinAlgError as err:
            raise KregError("kreg_solve", 1, str(err))

    def solve_indefinite(k, y, method=0):
        """dsytrf/dsytrs (Bunch-Kaufman) like solveSysLinEqsBK, mathUtils.f90:106-186."""
        import warnings

        from scipy.linalg import LinAlgWarning, solve as sl_solve

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", LinAlgWarning)
            return torch.from_numpy(sl_solve(_sym(k), y.numpy(), assume_a="sym"))

    class FakeGraph:
        def __init__(self, model, n, energy=True, gradient=True):
            self.model, self.energy, self.gradient = model, energy, gradient

        def __call__(self, xyz):
            e, g = self.model.predict_host(xyz, self.energy, self.gradient)
            return (None if e is None else e.numpy()), (None if g is None else g.numpy())

    def hessian(x_train, alpha_val, xeq, sigma, xyz):
        xte = xyz.numpy()
        nv = 0 if alpha_val is None else alpha_val.numel()
        a = np.zeros(0) if alpha_val is None else alpha_val.numpy()
        return torch.from_numpy(oracle.hessian_kregf90(x_train.numpy(), a, sigma, nv, xte,
                                                       oracle.descriptors(xte, xeq.numpy())))

    def kernel_block_ex(kind, xyz_a, xyz_b, xeq, sigma, nn=2, symmetric=False, lam=0.0):
        xa, xb, q = xyz_a.numpy(), xyz_b.numpy(), xeq.numpy()
        k = oracle.kernel_block_generic(kind, oracle.descriptors(xa, q), oracle.descriptors(xb, q), sigma, nn=nn)
        if symmetric:
            k[np.diag_indices_from(k)] = 1.0 + lam
        return torch.from_numpy(k)

    class FakeRadialModel:
        def __init__(self, kind, nn, xyz_tr, xeq, alpha, sigma, prior):
            self.kind, self.nn, self.sigma, self.prior, self.device = kind.lower(), int(nn), sigma, prior, torch.device("cpu")
            self.xeq = xeq.numpy()
            self.X = oracle.descriptors(xyz_tr.numpy(), self.xeq)
            self.alpha = alpha.numpy().reshape(-1)

        def predict(self, xyz, energy=True, gradient=True):
            x = np.ascontiguousarray(xyz.numpy() if torch.is_tensor(xyz) else xyz)
            e, g = oracle.predict_generic(self.kind, self.X, self.alpha, self.sigma, self.prior, x,
                                          oracle.descriptors(x, self.xeq), nn=self.nn, calc_grad=gradient)
            return (torch.from_numpy(e) if energy else None), (torch.from_numpy(g) if gradient else None)

    from mlatom_b200 import kernels as KN

    setattr_(KN, "RadialModel", FakeRadialModel)
    setattr_(KN, "_dev", lambda t, device=None: torch.as_tensor(np.ascontiguousarray(
        t.detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float64)))
    setattr_(eng, "kernel_block_ex", kernel_block_ex)
    setattr_(eng, "hessian", hessian)
    setattr_(eng, "GraphPredictor", FakeGraph)
    setattr_(eng, "build_kernel_matrix", build)
    setattr_(eng, "solve", solve)
    setattr_(eng, "solve_indefinite", solve_indefinite)
    setattr_(eng.PackedModel, "from_training_set",
             classmethod(lambda cls, xyz, xeq, alpha, sigma, prior, nv, ng, device=None:
                         FakePacked(oracle, xyz, xeq, alpha, sigma, prior, nv, ng)))
    return eng
