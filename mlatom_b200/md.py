"""Batched molecular dynamics on device-resident tensors.

The reference propagates trajectories in numpy and calls ``model.predict`` once per step on Python molecule
objects (mlatom/md.py:158-205 for one trajectory, md_parallel.py:155-200 for a batch).  Here the whole
velocity-Verlet step -- positions, one fused E + dE/dxyz prediction for all trajectories, velocities -- stays on the
GPU and is captured once in a CUDA graph, so a step costs one graph replay and no host traffic.

Units are the caller's: forces are ``-dE/dxyz`` in (energy unit)/(length unit), ``masses`` in mass units such that
``force / mass`` is length/time^2 for the given ``dt``.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .engine import F64, PackedModel


class BatchedNVE:
    """Velocity Verlet for B independent trajectories of the same molecule (md.py:163-205).

    With ``aux_model`` and ``uq_threshold`` it is the active-learning sampler of al_utils.py:246-322: every step also
    evaluates the auxiliary model (energies only), and a trajectory whose ``|E - E_aux|`` exceeds the threshold is
    frozen at that geometry (the reference's stop function on ``molecule.uncertain``); ``stop_step`` records when."""

    def __init__(self, model: PackedModel, xyz: torch.Tensor, velocities: torch.Tensor, masses: torch.Tensor,
                 dt: float, use_graph: bool = True, aux_model: Optional[PackedModel] = None,
                 uq_threshold: Optional[float] = None):
        assert xyz.is_cuda and xyz.dtype == F64 and xyz.dim() == 3
        self.model, self.dt = model, float(dt)
        self.aux, self.uq_threshold = aux_model, uq_threshold
        if aux_model is not None:
            assert uq_threshold is not None, "an auxiliary model needs a threshold"
            b0 = xyz.shape[0]
            self.eaux = torch.empty(b0, dtype=F64, device=xyz.device)
            self.uq = torch.zeros(b0, dtype=F64, device=xyz.device)
            self.active = torch.ones((b0, 1, 1), dtype=F64, device=xyz.device)   # 1.0 while the trajectory runs
            self.stop_step = torch.full((b0,), -1, dtype=torch.int64, device=xyz.device)
            self.step_no = torch.zeros((), dtype=torch.int64, device=xyz.device)
        self.xyz = xyz.clone().contiguous()
        self.vel = velocities.to(xyz.device, F64).clone().contiguous()
        self.inv_m = (1.0 / masses.to(xyz.device, F64)).reshape(1, -1, 1)
        self.half_m = (0.5 * masses.to(xyz.device, F64)).reshape(1, -1, 1)
        b, nat, _ = xyz.shape
        self.epot = torch.empty(b, dtype=F64, device=xyz.device)
        self.grad = torch.empty((b, nat, 3), dtype=F64, device=xyz.device)
        self.ekin = torch.empty(b, dtype=F64, device=xyz.device)
        # scratch owned by this integrator: the captured graph keeps the raw pointers (never the model's shared scratch)
        self.ws = model.workspace(b)
        self.ws_aux = aux_model.workspace(b) if aux_model is not None else None
        model.predict(self.xyz, True, True, out_e=self.epot, out_g=self.grad, ws=self.ws)  # forces at t = 0
        self._kinetic()
        if self.aux is not None:
            self._check_uncertain()  # an initial condition can already be uncertain (stop_step = 0)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        if use_graph:
            dev = xyz.device
            with torch.cuda.device(dev):
                state = [self.xyz, self.vel, self.epot, self.grad]
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
                with torch.cuda.graph(self.graph):
                    self._step()
                for dst, src in zip(state, saved):
                    dst.copy_(src)  # warm-up and capture advanced the state: restore t = 0
                self._kinetic()

    def _kinetic(self):
        torch.sum(self.half_m * self.vel * self.vel, dim=(1, 2), out=self.ekin)

    def _check_uncertain(self):
        """uq = |E - E_aux| (al_utils.py:1418-1424); running trajectories above the threshold stop here."""
        self.aux.predict(self.xyz, True, False, out_e=self.eaux, ws=self.ws_aux)
        torch.sub(self.epot, self.eaux, out=self.uq).abs_()
        hit = (self.uq > self.uq_threshold) & (self.active.reshape(-1) > 0)
        self.stop_step.copy_(torch.where(hit, self.step_no.expand_as(self.stop_step), self.stop_step))
        self.active.mul_((~hit).to(F64).reshape(-1, 1, 1))

    def _step(self):
        dt = self.dt
        if self.aux is None:
            self.vel.add_(self.grad * self.inv_m, alpha=-0.5 * dt)   # v(t + dt/2) = v - dt/2 * grad/m
            self.xyz.add_(self.vel, alpha=dt)                        # x(t + dt)
            self.model.predict(self.xyz, True, True, out_e=self.epot, out_g=self.grad, ws=self.ws)
            self.vel.add_(self.grad * self.inv_m, alpha=-0.5 * dt)   # v(t + dt)
            self._kinetic()
            return
        # sampler mode: frozen trajectories keep position and velocity (their updates are multiplied by 0)
        self.step_no.add_(1)
        self.vel.add_(self.grad * self.inv_m * self.active, alpha=-0.5 * dt)
        self.xyz.add_(self.vel * self.active, alpha=dt)
        self.model.predict(self.xyz, True, True, out_e=self.epot, out_g=self.grad, ws=self.ws)
        self.vel.add_(self.grad * self.inv_m * self.active, alpha=-0.5 * dt)
        self._kinetic()
        self._check_uncertain()

    def run(self, nsteps: int, record_every: int = 0) -> Dict[str, torch.Tensor]:
        """Advance all trajectories; returns final state and (optionally) the energies every `record_every` steps."""
        trace = []
        for it in range(nsteps):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._step()
            if record_every and (it + 1) % record_every == 0:
                trace.append(torch.stack((self.epot, self.ekin)).clone())
        out = {"xyz": self.xyz, "velocities": self.vel, "epot": self.epot, "ekin": self.ekin}
        if self.aux is not None:
            out.update(uq=self.uq, uncertain=self.stop_step >= 0, stop_step=self.stop_step)
        if trace:
            out["energies"] = torch.stack(trace)  # (nrecords, 2, B)
        return out
