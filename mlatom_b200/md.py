"""Batched molecular dynamics on device-resident tensors.

The reference propagates trajectories in numpy and calls ``model.predict`` once per step on Python molecule
objects (mlatom/md.py:158-205 for one trajectory, md_parallel.py:155-200 for a batch).  Here the whole
velocity-Verlet step stays on the GPU as THREE kernels -- ``kreg_md_first_half`` (positions + first half kick),
``kreg_predict`` (one fused E + dE/dxyz pass for all trajectories), ``kreg_md_second_half`` (second half kick +
kinetic energies) -- captured once in a CUDA graph, so a step costs one graph replay and no host traffic.

Units: with ``units='mlatom'`` they are the reference's -- coordinates in Angstrom, time in fs, energies in Hartree,
masses in relative atomic mass units -- and the conversion factors of md.py:171 and data.py:862 are applied the way
the reference applies them, so a trajectory reproduces ``mlatom.md``'s (tests/golden/md/md_nve_nat9.npz).  ``units=None``
(default) applies no factors: ``force / mass`` must then be length/time^2 for the given ``dt``.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ._lib import call
from .engine import F64, PackedModel, _on, _ptr, _stream

# mlatom/constants.py:10,28 and au2fs (:30): the three numbers md.py:171 and data.py:862 combine
BOHR2ANGSTROM = 0.52917721092
RAM2AU = 1822.888515
AU2FS = 2.4188843265857e-2


def mlatom_unit_factors():
    """(c_acc, c_kin): acceleration = force / mass * c_acc in Angstrom/fs^2 for forces in Hartree/Angstrom and masses in
    relative atomic mass (md.py:171: / ram2au * Bohr2Angstrom**2 * fs2au**2); E_kin = sum m v^2 / 2 * c_kin in Hartree
    (data.py:862: * ram2au * (au2fs / Bohr2Angstrom)**2)."""
    fs2au = 1.0 / AU2FS
    return 1.0 / RAM2AU * (BOHR2ANGSTROM ** 2) * fs2au ** 2, RAM2AU * (AU2FS / BOHR2ANGSTROM) ** 2


class BatchedNVE:
    """Velocity Verlet for B independent trajectories of the same molecule (md.py:163-205).

    With ``aux_model`` and ``uq_threshold`` it is the active-learning sampler of al_utils.py:246-322: every step also
    evaluates the auxiliary model (energies only), and a trajectory whose ``|E - E_aux|`` exceeds the threshold is
    frozen at that geometry (the reference's stop function on ``molecule.uncertain``); ``stop_step`` records when."""

    def __init__(self, model: PackedModel, xyz: torch.Tensor, velocities: torch.Tensor, masses: torch.Tensor,
                 dt: float, use_graph: bool = True, aux_model: Optional[PackedModel] = None,
                 uq_threshold: Optional[float] = None, units: Optional[str] = None):
        assert xyz.is_cuda and xyz.dtype == F64 and xyz.dim() == 3
        self.model, self.dt = model, float(dt)
        self.aux, self.uq_threshold = aux_model, uq_threshold
        dev = xyz.device
        b, nat, _ = xyz.shape
        self.ntraj, self.natoms = int(b), int(nat)
        c_acc, c_kin = mlatom_unit_factors() if units == "mlatom" else (1.0, 1.0)
        m = masses.to(dev, F64).reshape(-1)
        assert m.numel() == nat
        # per Cartesian component, with the unit factors folded in the way the reference multiplies them
        self.acc_scale = (1.0 / m * c_acc).repeat_interleave(3).contiguous()
        self.kin_scale = (0.5 * m * c_kin).repeat_interleave(3).contiguous()
        self.active = None
        if aux_model is not None:
            assert uq_threshold is not None, "an auxiliary model needs a threshold"
            self.eaux = torch.empty(b, dtype=F64, device=dev)
            self.uq = torch.zeros(b, dtype=F64, device=dev)
            self.active = torch.ones(b, dtype=F64, device=dev)   # 1.0 while the trajectory runs
            self.stop_step = torch.full((b,), -1, dtype=torch.int64, device=dev)
            self.step_no = torch.zeros((), dtype=torch.int64, device=dev)
        self.xyz = xyz.clone().contiguous()
        self.vel = velocities.to(dev, F64).clone().contiguous()
        self.epot = torch.empty(b, dtype=F64, device=dev)
        self.grad = torch.empty((b, nat, 3), dtype=F64, device=dev)
        self.ekin = torch.empty(b, dtype=F64, device=dev)
        # scratch owned by this integrator: the captured graph keeps the raw pointers (never the model's shared scratch)
        self.ws = model.workspace(b)
        self.ws_aux = aux_model.workspace(b) if aux_model is not None else None
        model.predict(self.xyz, True, True, out_e=self.epot, out_g=self.grad, ws=self.ws)  # forces at t = 0
        self._kinetic()
        if self.aux is not None:
            self._check_uncertain()  # an initial condition can already be uncertain (stop_step = 0)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        if use_graph:
            with torch.cuda.device(dev):
                state = [self.xyz, self.vel, self.epot, self.grad, self.ekin]
                if self.aux is not None:
                    state += [self.eaux, self.uq, self.active, self.stop_step, self.step_no]
                saved = tuple(t.clone() for t in state)
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize(dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=side):  # a stream of this device (see engine.GraphPredictor)
                    self._step()
                for dst, src in zip(state, saved):
                    dst.copy_(src)  # warm-up and capture advanced the state: restore t = 0

    def _second_half(self, dt: float):
        with _on(self.xyz):
            call("kreg_md_second_half", self.natoms, self.ntraj, dt, _ptr(self.acc_scale), _ptr(self.kin_scale),
                 _ptr(self.grad), _ptr(self.vel), _ptr(self.ekin), _ptr(self.active), _stream(self.xyz.device))

    def _kinetic(self):
        self._second_half(0.0)  # dt = 0: no kick, kinetic energies only

    def _check_uncertain(self):
        """uq = |E - E_aux| (al_utils.py:1418-1424); running trajectories above the threshold stop here."""
        self.aux.predict(self.xyz, True, False, out_e=self.eaux, ws=self.ws_aux)
        with _on(self.xyz):
            call("kreg_md_uncertainty", self.ntraj, _ptr(self.epot), _ptr(self.eaux), float(self.uq_threshold),
                 _ptr(self.step_no), _ptr(self.uq), _ptr(self.active), _ptr(self.stop_step), _stream(self.xyz.device))

    def _step(self):
        if self.aux is not None:
            self.step_no.add_(1)
        with _on(self.xyz):
            call("kreg_md_first_half", self.natoms, self.ntraj, self.dt, _ptr(self.acc_scale), _ptr(self.grad),
                 _ptr(self.xyz), _ptr(self.vel), _ptr(self.active), _stream(self.xyz.device))
        self.model.predict(self.xyz, True, True, out_e=self.epot, out_g=self.grad, ws=self.ws)
        self._second_half(self.dt)
        if self.aux is not None:
            self._check_uncertain()

    def run(self, nsteps: int, record_every: int = 0, record_xyz: bool = False) -> Dict[str, torch.Tensor]:
        """Advance all trajectories; returns final state and (optionally) the energies -- and positions -- every
        `record_every` steps."""
        trace, frames = [], []
        for it in range(nsteps):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._step()
            if record_every and (it + 1) % record_every == 0:
                trace.append(torch.stack((self.epot, self.ekin)).clone())
                if record_xyz:
                    frames.append(self.xyz.clone())
        out = {"xyz": self.xyz, "velocities": self.vel, "epot": self.epot, "ekin": self.ekin}
        if self.aux is not None:
            out.update(uq=self.uq, uncertain=self.stop_step >= 0, stop_step=self.stop_step)
        if trace:
            out["energies"] = torch.stack(trace)  # (nrecords, 2, B)
        if frames:
            out["trajectory"] = torch.stack(frames)  # (nrecords, B, Nat, 3)
        return out
