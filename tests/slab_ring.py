"""Several x-slabs of one grid in ONE process on ONE GPU (TEST HARNESS for the peer-memory halo path of
ceviche_b200/slab.py): the exchange blocks are plain device allocations attached by pointer instead of CUDA IPC
mappings, and the slabs are stepped in lockstep on one stream -- all H half-steps of a time step, then all D
half-steps -- so every arrival counter is already satisfied when its kernel starts.  Everything else is the
production path: the tensor-map kernels' peer stores, counters, targets, halo descriptors and source / probe
localisation."""
import ctypes as C

import numpy as np
import torch

from ceviche_b200 import _lib
from ceviche_b200.slab import SlabFDTD, partition


class LocalRing:
    def __init__(self, eps, dL, npml, P, dtype, device="cuda:0", options=()):
        eps = np.asarray(eps, dtype=np.float64)
        self.shape, self.P = eps.shape, P
        self.sims = []
        for r in range(P):
            lo, hi = partition(self.shape[0], P)[r]
            eps_local = np.concatenate([eps[(lo - 1) % self.shape[0]][None], eps[lo:hi]], 0)
            sim = SlabFDTD(self.shape, eps_local, dL, npml, dtype=dtype, device=device, path="peer", _ring=(r, P))
            for k, v in options:
                sim.set_option(k, v)
            self.sims.append(sim)
        self.blocks = [sim._peer_alloc(want_handle=False)[0] for sim in self.sims]
        for r, sim in enumerate(self.sims):
            sim._peer_attach(self.blocks[r], self.blocks[(r - 1) % P], self.blocks[(r + 1) % P])
        torch.cuda.synchronize()

    def prepare(self, sources, probes):
        for sim in self.sims:
            sim.prepare(sources, probes)

    def run(self, steps, waveforms):
        wf = torch.as_tensor(np.ascontiguousarray(waveforms, dtype=np.float64)).cuda()
        for sim in self.sims:
            sim.be.new_partials(steps)
        for n in range(steps):
            for sim in self.sims:
                sim.be.step_H(0, sim.nx, n - 1)
            for sim in self.sims:
                sim.be.step_D(0, sim.nx, n, wf[n])
        if steps:
            for sim in self.sims:
                sim.be.sample(0, steps - 1)
        for sim in self.sims:
            sim.t_index += steps
            sim._peer_check()
        return sum(sim.be.series() for sim in self.sims)

    def initialize_fields(self):
        for sim in self.sims:
            sim.initialize_fields()

    def field(self, key):
        return torch.cat([sim.be.field(key) for sim in self.sims], 0)

    def close(self):
        lib = self.sims[0].be.plan.lib
        for sim, b in zip(self.sims, self.blocks):
            lib.cev_fdtd_halo_attach(sim.be.plan.handle, None, None, None)
            lib.cev_halo_free(sim.be.device.index, b)
        self.blocks = []
