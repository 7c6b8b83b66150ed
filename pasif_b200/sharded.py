"""One process per GPU: shard the embarrassingly parallel (system x excitation) batch and all-reduce the parameter
gradients.  The reference has no multi-GPU logic; its only hook is "MPI rank -> GPU id"
(reference PASIf.py:27-30,167, src/__GpuDriver.cu:603) plus a TODO (src/__GpuDriver.cu:910).

Partitioning (`axis="auto"`): the SYSTEM axis is split into contiguous blocks when there are at least as many systems
as ranks -- coefficient blocks are then not duplicated and every rank fills its own slice of the reduced-gradient
vector, zeros elsewhere; otherwise the EXCITATION axis is split (every rank holds every system and a partial sum over
its excitations).  Either way ONE all-reduce(SUM) assembles the excitation-summed adjoint state on every rank.  No
collective runs during time stepping: trajectories never exchange data.

`presharded=True`: every rank passes only ITS OWN block of systems to setSystems / setJacobian (weak scaling: the global
batch is the concatenation over ranks and is never materialised on one host); the block offsets come from an
all-gather of the block sizes.
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


def choose_axis(n_systems, n_excitations, world):
    """'systems' when every rank gets at least one system (no coefficient duplication), else 'excitations' when every
    rank gets at least one excitation; ValueError when neither axis can feed every rank."""
    if n_systems >= world:
        return "systems"
    if n_excitations >= world:
        return "excitations"
    raise ValueError("cannot shard %d systems x %d excitations over %d ranks: both axes are shorter than the world size"
                     % (n_systems, n_excitations, world))


class ShardedPASIf:
    """Front end with the reference's call sequence (ctor -> setSystems -> setJacobian -> get*), sharded over ranks."""

    def __init__(self, excitationSet, sampleRate, numsteps_=0, dCompute_=False, dSystem_=False, dSolver_=False,
                 driver_cls=None, rank=None, world=None, axis="auto", presharded=False, options=None):
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
        self.presharded = bool(presharded)
        # every call of this (additive) front end starts from the initial conditions; the reference driver carries forward /
        # adjoint state from one _getGradient into the next, which the engine reproduces by default
        self.options = dict({"call_order": "independent"}, **(options or {}))
        self.local = None
        self.n_systems = 0
        self._sys_range = (0, 0)
        self._exc_range = (0, len(self._exc_all))
        self.last_device_ms = self.last_allreduce_ms = None

    def _gather_counts(self, n_local):
        """block sizes of all ranks (presharded mode)"""
        d = _dist()
        if d is None or self.world == 1:
            return [n_local] * self.world
        out = [None] * self.world
        d.all_gather_object(out, int(n_local))
        return [int(x) for x in out]

    def setSystems(self, vecM, vecB, vecK, vecGamma, vecLambda, vecForcePattern, vecInitialConditions):
        sr, n, dc, ds, dv = self._ctor
        if self.presharded:
            counts = self._gather_counts(len(vecM))
            self.n_systems = int(sum(counts))
            self._axis = "systems"
            lo = int(sum(counts[:self.rank]))
            self._sys_range = (lo, lo + len(vecM))
            self._exc_range = (0, len(self._exc_all))
            sl = slice(None)
            sizes = [m.shape[0] for m in vecM]
            self._size_blocks = self._gather_sizes(sizes)
        else:
            self.n_systems = len(vecM)
            self._axis = choose_axis(self.n_systems, len(self._exc_all), self.world) if self.axis == "auto" else self.axis
            if self._axis == "systems":
                self._sys_range = block_partition(self.n_systems, self.world, self.rank)
                self._exc_range = (0, len(self._exc_all))
            else:
                self._sys_range = (0, self.n_systems)
                self._exc_range = block_partition(len(self._exc_all), self.world, self.rank)
            sl = slice(*self._sys_range)
            self._size_blocks = None
        lo, hi = self._sys_range
        e0, e1 = self._exc_range
        if hi <= lo or e1 <= e0:
            raise ValueError("rank %d received an empty shard (%d systems, %d excitations over %d ranks)"
                             % (self.rank, self.n_systems, len(self._exc_all), self.world))
        self.local = self._cls(self._exc_all[e0:e1], sr, n, dc, ds, dv)
        if hasattr(self.local, "_setOption"):
            for k, v in self.options.items():
                self.local._setOption(k, v)
        self.local.setSystems(vecM[sl], vecB[sl], vecK[sl], vecGamma[sl], vecLambda[sl], vecForcePattern[sl],
                              vecInitialConditions[sl])

    def _gather_sizes(self, sizes):
        d = _dist()
        if d is None or self.world == 1:
            return [list(sizes)] * self.world
        out = [None] * self.world
        d.all_gather_object(out, [int(s) for s in sizes])
        return out

    def setJacobian(self, vecM, vecB, vecK, vecGamma, vecLambda, vecForcePattern, vecInitialConditions, vecPsi):
        if self.local is None:
            raise ValueError("The system must be set before the jacobian")
        sizes = [m.shape[0] for m in vecM]
        if self.presharded:
            blocks = self._gather_sizes(sizes)
            self._adj_off = int(sum(sum(b) for b in blocks[:self.rank]))
            self._adj_glob = int(sum(sum(b) for b in blocks))
            self._adj_loc = int(sum(sizes))
            sl = slice(None)
        else:
            lo, hi = self._sys_range
            self._adj_off = int(sum(sizes[:lo]))
            self._adj_glob = int(sum(sizes))
            self._adj_loc = int(sum(sizes[lo:hi]))
            sl = slice(lo, hi)
        self.local.setJacobian(vecM[sl], vecB[sl], vecK[sl], vecGamma[sl], vecLambda[sl], vecForcePattern[sl],
                               vecInitialConditions[sl], vecPsi[sl])

    @property
    def n_adj_global(self):
        return self._adj_glob

    @property
    def adj_slice(self):
        """[lo, hi) of this rank's systems inside the reduced-gradient vector"""
        return self._adj_off, self._adj_off + self._adj_loc

    def setExcitations(self, excitationSet, sampleRate):
        """Replace the excitation set (every rank passes the FULL set; the rank keeps its block when the excitation
        axis is split).  Same host -> device upload as the reference's `setExcitations` (PASIf.py:371-385)."""
        e0, e1 = self._exc_range
        exc = excitationSet[e0:e1] if (e0, e1) != (0, len(excitationSet)) else excitationSet
        self.local._loadExcitationsSet(exc, sampleRate)

    # local shard, reference layout [e][i] over the rank's excitations / systems
    def getAmplitudes(self):
        return self.local.getAmplitudes()

    def getGradient(self):
        return self.local.getGradient()

    def getGradientReduced(self, device_tensor=None):
        """Adjoint state summed over ALL excitations for ALL systems, identical on every rank: local on-device sum over
        the rank's excitations, then ONE all-reduce(SUM) (NCCL when torch.distributed runs on it).  With
        `device_tensor` (a CUDA torch tensor of the engine's dtype and global adjoint size) the reduction result is
        written and all-reduced in place without touching the host; `last_device_ms` / `last_allreduce_ms` then hold
        the device time of the local gradient and of the collective (CUDA events)."""
        off, n_loc = self._adj_off, self._adj_loc
        d = _dist()
        if device_tensor is not None:
            import torch
            if device_tensor.numel() != self._adj_glob:
                raise ValueError("device_tensor must hold %d values (global adjoint size)" % self._adj_glob)
            if not hasattr(self.local, "_getGradientReducedDevice"):
                raise RuntimeError("the native driver has no _getGradientReducedDevice entry point; rebuild PASIfgpu")
            device_tensor.zero_()
            # the engine writes on its own (non-blocking) stream: the zero fill must have landed before it starts
            torch.cuda.current_stream(device_tensor.device).synchronize()
            view = device_tensor[off:off + n_loc]
            self.local._getGradientReducedDevice(view.data_ptr(), 0)      # returns after the engine's stream is idle
            self.last_device_ms = self.local._query("last_device_ms")
            self.last_allreduce_ms = 0.0
            if d is not None and self.world > 1:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                d.all_reduce(device_tensor)
                ev1.record()
                ev1.synchronize()
                self.last_allreduce_ms = ev0.elapsed_time(ev1)
            return device_tensor
        return self._reduce_host(d)

    def getGradientAndReduced(self, device_tensor):
        """ONE gradient evaluation, both results: the rank's un-reduced adjoint state [e][j] on the host (what the
        reference's getGradient returns for the local shard) and the all-reduced, excitation-summed vector in
        `device_tensor` (summed on the device from the same evaluation, then all-reduced in place)."""
        import torch
        d = _dist()
        if device_tensor.numel() != self._adj_glob:
            raise ValueError("device_tensor must hold %d values (global adjoint size)" % self._adj_glob)
        local = self.local._getGradient(0)
        device_tensor.zero_()
        torch.cuda.current_stream(device_tensor.device).synchronize()
        view = device_tensor[self._adj_off:self._adj_off + self._adj_loc]
        self.local._reduceLastGradientDevice(view.data_ptr())
        if d is not None and self.world > 1:
            d.all_reduce(device_tensor)
        return local, device_tensor

    def _reduce_host(self, d):
        off, n_loc = self._adj_off, self._adj_loc
        if hasattr(self.local, "_getGradientReduced"):
            loc = np.asarray(self.local._getGradientReduced(0), dtype=np.float64)
        else:                                      # driver without the additive reduced entry point
            g = np.asarray(self.local.getGradient(), dtype=np.float64)
            loc = g.reshape(-1, n_loc).sum(axis=0)
        full = np.zeros(self._adj_glob)
        full[off:off + n_loc] = loc
        if d is not None and self.world > 1:
            import torch
            t = torch.from_numpy(full)
            if d.get_backend() == "nccl":          # NCCL reduces device tensors only: stage through this rank's GPU
                t = t.cuda()
                d.all_reduce(t)
                full = t.cpu().numpy()
            else:
                d.all_reduce(t)
                full = t.numpy()
        return full
