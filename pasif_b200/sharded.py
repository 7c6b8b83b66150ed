"""One process per GPU: shard the embarrassingly parallel (system x excitation) batch and all-reduce the parameter
gradients.  The reference has no multi-GPU logic; its only hook is "MPI rank -> GPU id"
(reference PASIf.py:27-30,167, src/__GpuDriver.cu:603) plus a TODO (src/__GpuDriver.cu:910).

Partitioning: the excitation axis is split into contiguous blocks when there are at least as many excitations as
ranks (every rank then holds every system, and the gradient summed over excitations needs a SUM all-reduce);
otherwise the system axis is split (coefficient blocks are not duplicated) and each rank fills its own slice of the
reduced-gradient vector, zeros elsewhere, so the same single SUM all-reduce assembles the result.  No collective
runs during time stepping: trajectories never exchange data.
"""
from __future__ import annotations

import os

import numpy as np

from .frontend import make_frontend


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def block_partition(n, world, rank):
    """Contiguous block [lo, hi) of `n` items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedPASIf:
    """Front end with the reference's call sequence (ctor -> setSystems -> setJacobian -> get*), sharded over ranks."""

    def __init__(self, excitationSet, sampleRate, numsteps_=0, dCompute_=False, dSystem_=False, dSolver_=False,
                 driver_cls=None, rank=None, world=None, axis="auto"):
        d = _dist()
        self.rank = rank if rank is not None else (d.get_rank() if d else int(os.environ.get("RANK", "0")))
        self.world = world if world is not None else (d.get_world_size() if d else int(os.environ.get("WORLD_SIZE", "1")))
        self._exc_all = [np.asarray(e, dtype=np.float64) for e in excitationSet]
        self._ctor = (sampleRate, numsteps_ if numsteps_ else len(self._exc_all[0]), dCompute_, dSystem_, dSolver_)
        if driver_cls is None:
            from .frontend import load_native_driver
            driver_cls = load_native_driver()
        self._cls = make_frontend(driver_cls)
        self.axis = axis
        self.local = None
        self.n_systems = 0
        self._sys_range = (0, 0)
        self._exc_range = (0, len(self._exc_all))

    def _choose_axis(self, n_systems):
        if self.axis != "auto":
            return self.axis
        return "systems" if n_systems >= self.world else "excitations"

    def setSystems(self, vecM, vecB, vecK, vecGamma, vecLambda, vecForcePattern, vecInitialConditions):
        self.n_systems = len(vecM)
        self._axis = self._choose_axis(self.n_systems)
        if self._axis == "systems":
            self._sys_range = block_partition(self.n_systems, self.world, self.rank)
            self._exc_range = (0, len(self._exc_all))
        else:
            self._sys_range = (0, self.n_systems)
            self._exc_range = block_partition(len(self._exc_all), self.world, self.rank)
        lo, hi = self._sys_range
        e0, e1 = self._exc_range
        if hi <= lo or e1 <= e0:
            raise ValueError("rank %d received an empty shard (%d systems, %d excitations over %d ranks)"
                             % (self.rank, self.n_systems, len(self._exc_all), self.world))
        sr, n, dc, ds, dv = self._ctor
        self.local = self._cls(self._exc_all[e0:e1], sr, n, dc, ds, dv)
        self._sizes = [m.shape[0] for m in vecM]
        sl = slice(lo, hi)
        self.local.setSystems(vecM[sl], vecB[sl], vecK[sl], vecGamma[sl], vecLambda[sl], vecForcePattern[sl],
                              vecInitialConditions[sl])

    def setJacobian(self, vecM, vecB, vecK, vecGamma, vecLambda, vecForcePattern, vecInitialConditions, vecPsi):
        if self.local is None:
            raise ValueError("The system must be set before the jacobian")
        self._adj_sizes = [m.shape[0] for m in vecM]
        sl = slice(*self._sys_range)
        self.local.setJacobian(vecM[sl], vecB[sl], vecK[sl], vecGamma[sl], vecLambda[sl], vecForcePattern[sl],
                               vecInitialConditions[sl], vecPsi[sl])

    # local shard, reference layout [e][i] over the rank's excitations / systems
    def getAmplitudes(self):
        return self.local.getAmplitudes()

    def getGradient(self):
        return self.local.getGradient()

    def getGradientReduced(self, device_tensor=None):
        """Adjoint state summed over ALL excitations for ALL systems, identical on every rank: local on-device sum over
        the rank's excitations, then ONE all-reduce(SUM) (NCCL when torch.distributed runs on it).  With
        `device_tensor` (a CUDA torch tensor of the engine's dtype and global adjoint size) the reduction result is
        written and all-reduced in place without touching the host."""
        n_adj_glob = int(sum(self._adj_sizes))
        off = int(sum(self._adj_sizes[:self._sys_range[0]]))
        n_loc = int(sum(self._adj_sizes[self._sys_range[0]:self._sys_range[1]]))
        d = _dist()
        if device_tensor is not None:
            device_tensor.zero_()
            view = device_tensor[off:off + n_loc]
            self.local._getGradientReducedDevice(view.data_ptr(), 0)
            if d is not None and self.world > 1:
                d.all_reduce(device_tensor)
            return device_tensor
        if hasattr(self.local, "_getGradientReduced"):
            loc = np.asarray(self.local._getGradientReduced(0), dtype=np.float64)
        else:                                      # driver without the additive reduced entry point
            g = np.asarray(self.local.getGradient(), dtype=np.float64)
            loc = g.reshape(-1, n_loc).sum(axis=0)
        full = np.zeros(n_adj_glob)
        full[off:off + n_loc] = loc
        if d is not None and self.world > 1:
            import torch
            t = torch.from_numpy(full)
            d.all_reduce(t)
            full = t.numpy()
        return full
