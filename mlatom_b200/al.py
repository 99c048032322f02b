"""Active-learning consumers of the KREG path on device-resident tensors.

The reference's physics-informed active learning (mlatom/al_utils.py) keeps a main model (energies + gradients) and an
auxiliary one (energies only); a geometry is *uncertain* when ``|E_main - E_aux|`` exceeds a threshold
(``al_utils.py:1409-1424``), the threshold is ``median + n * MAD`` of the validation deviations (``:1436-1455``), and
the sampler propagates a batch of trajectories until each one hits an uncertain geometry (``:246-322``).  The reference
does all of it through per-molecule Python objects; here the two predictions and the comparison stay in HBM.
"""
from __future__ import annotations

import re
from typing import Dict, Optional

import numpy as np
import torch

from .engine import PackedModel


def median_absolute_deviation(values) -> float:
    """stats.calc_median_absolute_deviation (mlatom/stats.py:122-128): 1.4826 * median(|v - median(v)|)."""
    v = np.asarray(values, dtype=np.float64)
    return float(1.4826 * np.median(np.abs(v - np.median(v))))


def threshold_metric(absdevs, metric: str = "m+3mad") -> float:
    """al_utils.py:1436-1455: 'max', or 'm+<n>mad' = median + n * MAD (0.0 for fewer than two values)."""
    v = np.asarray(absdevs.detach().cpu().numpy() if torch.is_tensor(absdevs) else absdevs, dtype=np.float64)
    if metric.casefold() == "max":
        return float(np.max(v))
    m = re.fullmatch(r"m\+(\d+(?:\.\d+)?)mad", metric.casefold())
    if not m:
        raise ValueError(f"Unknown threshold metric: {metric!r}")
    if len(v) < 2:
        return 0.0
    return float(np.median(v) + float(m.group(1)) * median_absolute_deviation(v))


class UncertaintyPair:
    """Main + auxiliary model evaluated together (al_utils.py:1409-1424).

    ``predict(xyz)`` returns device tensors: E and dE/dxyz of the main model, the auxiliary energy (energies-only pass:
    the second GEMM is skipped), ``uq = |E - E_aux|`` and, when a threshold is set, the ``uncertain`` mask."""

    def __init__(self, main: PackedModel, aux: PackedModel, uq_threshold: Optional[float] = None):
        self.main, self.aux, self.uq_threshold = main, aux, uq_threshold

    def predict(self, xyz: torch.Tensor, out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        out = out if out is not None else {}
        e, g = self.main.predict(xyz, True, True, out_e=out.get("energy"), out_g=out.get("energy_gradients"))
        ea, _ = self.aux.predict(xyz, True, False, out_e=out.get("aux_energy"))
        if "uq" in out:
            torch.sub(e, ea, out=out["uq"]).abs_()
        else:
            out["uq"] = (e - ea).abs_()
        out.update(energy=e, energy_gradients=g, aux_energy=ea)
        if self.uq_threshold is not None:
            if "uncertain" in out:
                torch.gt(out["uq"], self.uq_threshold, out=out["uncertain"])
            else:
                out["uncertain"] = out["uq"] > self.uq_threshold
        return out

    def calibrate(self, xyz_validation: torch.Tensor, metric: str = "m+3mad") -> float:
        """New threshold from the deviations on a validation set (al_utils.py:1394-1398)."""
        self.uq_threshold = threshold_metric(self.predict(xyz_validation)["uq"], metric)
        return self.uq_threshold
