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
    """Velocity Verlet for B independent trajectories of the same molecule (md.py:163-205)."""

    def __init__(self, model: PackedModel, xyz: torch.Tensor, velocities: torch.Tensor, masses: torch.Tensor,
                 dt: float, use_graph: bool = True):
        assert xyz.is_cuda and xyz.dtype == F64 and xyz.dim() == 3
        self.model, self.dt = model, float(dt)
        self.xyz = xyz.clone().contiguous()
        self.vel = velocities.to(xyz.device, F64).clone().contiguous()
        self.inv_m = (1.0 / masses.to(xyz.device, F64)).reshape(1, -1, 1)
        self.half_m = (0.5 * masses.to(xyz.device, F64)).reshape(1, -1, 1)
        b, nat, _ = xyz.shape
        self.epot = torch.empty(b, dtype=F64, device=xyz.device)
        self.grad = torch.empty((b, nat, 3), dtype=F64, device=xyz.device)
        self.ekin = torch.empty(b, dtype=F64, device=xyz.device)
        model.predict(self.xyz, True, True, out_e=self.epot, out_g=self.grad)  # forces at t = 0
        self._kinetic()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        if use_graph:
            dev = xyz.device
            with torch.cuda.device(dev):
                saved = (self.xyz.clone(), self.vel.clone(), self.epot.clone(), self.grad.clone())
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize(dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._step()
                for dst, src in zip((self.xyz, self.vel, self.epot, self.grad), saved):
                    dst.copy_(src)  # warm-up and capture advanced the state: restore t = 0
                self._kinetic()

    def _kinetic(self):
        torch.sum(self.half_m * self.vel * self.vel, dim=(1, 2), out=self.ekin)

    def _step(self):
        dt = self.dt
        self.vel.add_(self.grad * self.inv_m, alpha=-0.5 * dt)   # v(t + dt/2) = v - dt/2 * grad/m
        self.xyz.add_(self.vel, alpha=dt)                        # x(t + dt)
        self.model.predict(self.xyz, True, True, out_e=self.epot, out_g=self.grad)
        self.vel.add_(self.grad * self.inv_m, alpha=-0.5 * dt)   # v(t + dt)
        self._kinetic()

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
        if trace:
            out["energies"] = torch.stack(trace)  # (nrecords, 2, B)
        return out
