"""Python front end of the engine: same public surface as the reference's ``PASIf.py``
(class ``coo_tensor`` ``PASIf.py:32-146``; class ``PASIf(__GpuDriver)`` ``PASIf.py:148-632``) so callers of the
reference can switch by changing one import.  Method names, argument order, validation and the exceptions raised
follow the reference; what changes is *how* the block-diagonal global system is assembled: vectorised NumPy
instead of the reference's per-system ``np.concatenate`` / per-nnz Python loops (``PASIf.py:86-115,475-543``),
and the caller's input lists are never mutated (the reference overwrites ``vecForcePattern[0]`` /
``vecInitialConditions[0]`` and, for Jacobians, ``vecGamma[0]`` -- SURVEY.md quirk G4).

The hand-off to the native driver is element-for-element what the reference sends (checked against the unmodified
reference front end in ``tests/test_frontend.py``):

    K' = -M^-1 (B + K)  as sorted CSR         -> _setFwdK / _setBwdK              (PASIf.py:250-255, 328-335)
    Gamma', Lambda', Psi' scaled by -1/M[out]  -> _set{Fwd,Bwd}{Gamma,Lambda,Psi}  (PASIf.py:607-633)
    F' = F / M  (forward only)                 -> _setFwdForcePattern             (PASIf.py:618)
"""
from __future__ import annotations

import copy as cp
import os

import numpy as np
from scipy.sparse import coo_matrix, csc_matrix, dia_matrix


class coo_tensor:
    """Sparse tensor in COO format, indices stored per non-zero as [dim0, dim1, ..., dimN-1]
    (reference ``PASIf.py:32-146``; same attributes and helper methods)."""

    def __init__(self, dimensions_=None):
        self.dimensions = list(dimensions_) if dimensions_ is not None else [2, 2]
        self.val = []
        self.indices = []

    def __str__(self):
        nd = len(self.dimensions)
        lines = ["val: " + " ".join(str(v) for v in self.val)]
        for d in range(nd):
            lines.append(f"dim{d + 1}: " + " ".join(str(self.indices[k * nd + d]) for k in range(len(self.val))))
        return "\n" + "\n".join(lines) + "\n"

    def concatenateTensor(self, tensor_):
        if len(self.dimensions) != len(tensor_.dimensions):
            raise Exception("The number of dimensions of the two tensors are not the same")
        nd = len(self.dimensions)
        off = np.tile(np.asarray(self.dimensions, dtype=np.int64), len(tensor_.val))
        self.val = list(self.val) + list(tensor_.val)
        self.indices = list(self.indices) + [int(i) for i in (np.asarray(tensor_.indices, dtype=np.int64) + off[:len(tensor_.indices)])]
        for i in range(nd):
            self.dimensions[i] += tensor_.dimensions[i]

    def multiplyByDiagMatrix(self, diagMatrix_):
        nd = len(self.dimensions)
        if len(self.val):
            first = np.asarray(self.indices, dtype=np.int64)[::nd]
            self.val = list(np.asarray(self.val, dtype=np.float64) * np.asarray(diagMatrix_)[first])

    def getIndicesDim(self, dim_):
        nd = len(self.dimensions)
        return [self.indices[dim_ + i * nd] for i in range(len(self.val))]

    def isSquare(self):
        return all(d == self.dimensions[0] for d in self.dimensions)

    def extendDimmensions(self, extension_):
        self.dimensions = [d + extension_ for d in self.dimensions]

    def offsetDimmension(self, targetDim_, offset_):
        nd = len(self.dimensions)
        for i in range(targetDim_, len(self.indices), nd):
            self.indices[i] += offset_


def _local_rank():
    """GPU id the way the reference picks it: MPI rank (PASIf.py:27-30,167); torchrun's LOCAL_RANK when mpi4py is absent."""
    try:
        from mpi4py import MPI  # type: ignore
        return int(MPI.COMM_WORLD.Get_rank())
    except Exception:
        return int(os.environ.get("LOCAL_RANK", "0"))


def _concat_tensors(tensors, order):
    """Block-diagonal concatenation of COO tensors: every index of mode m of system s is offset by the sum of
    dimension m over the previous systems (reference ``coo_tensor.concatenateTensor``, PASIf.py:86-100)."""
    dims = np.array([t.dimensions for t in tensors], dtype=np.int64).reshape(len(tensors), order)
    offs = np.concatenate([np.zeros((1, order), np.int64), np.cumsum(dims, axis=0)[:-1]])
    vals, idxs = [], []
    for t, off in zip(tensors, offs):
        nnz = len(t.val)
        if nnz:
            vals.append(np.asarray(t.val, dtype=np.float64))
            idxs.append(np.asarray(t.indices, dtype=np.int64).reshape(nnz, order) + off[None, :])
    out = coo_tensor([int(d) for d in dims.sum(axis=0)])
    out._val = np.concatenate(vals) if vals else np.zeros(0)
    out._idx = np.concatenate(idxs) if idxs else np.zeros((0, order), np.int64)
    return out


def _concat_matrices(mats, sizes):
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    val = np.concatenate([np.asarray(m.data, dtype=np.float64).reshape(-1) for m in mats])
    row = np.concatenate([np.asarray(m.row, dtype=np.int64) + o for m, o in zip(mats, offs)])
    col = np.concatenate([np.asarray(m.col, dtype=np.int64) + o for m, o in zip(mats, offs)])
    n = int(np.sum(sizes))
    return coo_matrix((val, (row, col)), shape=(n, n))


def make_frontend(driver_cls):
    """Class factory: ``PASIf`` bound to a native driver class exposing the reference's ``__GpuDriver`` interface."""

    class PASIf(driver_cls):
        def __init__(self, excitationSet, sampleRate, numsteps_=0, dCompute_=False, dSystem_=False, dSolver_=False):
            self.numsteps = len(excitationSet[0]) if numsteps_ == 0 else numsteps_
            super().__init__(excitationSet, sampleRate, self.numsteps, dCompute_, dSystem_, dSolver_, _local_rank())
            self.sampleRate = sampleRate
            self.globalSystemSize = 0
            self.globalAdjointSize = 0
            self.interpolSize = 0
            self.saveSteps = 1
            self.n_excitations = len(excitationSet)
            self.systemSet = False
            self.numberOfSystems = 0
            self.system_M = self.system_B = self.system_K = None
            self.system_Gamma = self.system_Lambda = None
            self.system_forcePattern = self.system_initialConditions = None
            self.jacobianSet = False
            self.jacobian_M = self.jacobian_B = self.jacobian_K = None
            self.jacobian_Gamma = self.jacobian_Lambda = self.jacobian_Psi = None
            self.jacobian_forcePattern = self.jacobian_initialConditions = None

        # ------------------------------------------------------------------ excitations (PASIf.py:196-200)
        def setExcitations(self, excitationSet, sampleRate):
            self.sampleRate = sampleRate
            self.n_excitations = len(excitationSet)
            self._loadExcitationsSet(excitationSet, self.sampleRate)

        # ------------------------------------------------------------------ forward system (PASIf.py:202-277)
        def setSystems(self, vecM, vecB, vecK, vecGamma, vecLambda, vecForcePattern, vecInitialConditions):
            self.systemSet = False
            assert len(vecM) > 0
            assert len(vecM) == len(vecB) == len(vecK) == len(vecGamma) == len(vecLambda) == len(vecForcePattern) \
                == len(vecInitialConditions), "The number of Systems in the input vectors must be the same."
            assert (type(vecM[0]) == dia_matrix and type(vecB[0]) == coo_matrix and type(vecK[0]) == coo_matrix
                    and isinstance(vecGamma[0], coo_tensor) and isinstance(vecLambda[0], coo_tensor)
                    and type(vecForcePattern[0]) == np.ndarray and type(vecInitialConditions[0]) == np.ndarray), \
                "Inputs vectors types error."
            self.globalSystemSize = 0
            self._checkInputs(vecM, vecB, vecK, vecGamma, vecLambda, vecForcePattern, vecInitialConditions)

            self.numberOfSystems = len(vecM)
            sizes = np.array([m.shape[0] for m in vecM], dtype=np.int64)
            self.globalSystemSize = int(sizes.sum())
            N = self.globalSystemSize
            diagM = np.concatenate([np.asarray(m.data, dtype=np.float64).reshape(-1) for m in vecM])
            self.system_M = dia_matrix((diagM, [0]), shape=(N, N))
            self.system_B = _concat_matrices(vecB, sizes)
            self.system_K = _concat_matrices(vecK, sizes)
            self.system_Gamma = _concat_tensors(vecGamma, 3)
            self.system_Lambda = _concat_tensors(vecLambda, 4)
            self.system_forcePattern = np.concatenate([np.asarray(f, dtype=np.float64) for f in vecForcePattern])
            self.system_initialConditions = np.concatenate([np.asarray(f, dtype=np.float64) for f in vecInitialConditions])

            # -M^-1 pre-scaling (PASIf.py:607-618)
            self.system_M, minv = _neg_inverse_of_diagonal(diagM)
            self.system_B = coo_matrix(self.system_M.dot(self.system_B))
            self.system_K = coo_matrix(self.system_M.dot(self.system_K))
            _scale_tensor(self.system_Gamma, minv)
            _scale_tensor(self.system_Lambda, minv)
            self.system_forcePattern = self.system_forcePattern * (-1. * minv)

            csr_K = (self.system_B + self.system_K).tocsr()
            csr_K.sort_indices()
            self._setFwdK(self.system_K.shape, csr_K.data, csr_K.indices, csr_K.indptr)
            self._setFwdGamma(self.system_Gamma.dimensions, self.system_Gamma.val, self.system_Gamma.indices)
            self._setFwdLambda(self.system_Lambda.dimensions, self.system_Lambda.val, self.system_Lambda.indices)
            self._setFwdPsi([1] * 5, [], [])
            self._setFwdForcePattern(self.system_forcePattern)
            self._setFwdInitialConditions(self.system_initialConditions)
            self.systemSet = True
            self._allocateSystemOnDevice()

        # ------------------------------------------------------------------ adjoint system (PASIf.py:279-369)
        def setJacobian(self, vecM, vecB, vecK, vecGamma, vecLambda, vecForcePattern, vecInitialConditions, vecPsi):
            self.jacobianSet = False
            self.globalAdjointSize = 0
            if self.systemSet == False:  # noqa: E712  (reference wording)
                raise ValueError("The system must be set before the jacobian")
            assert len(vecM) > 0
            assert len(vecM) == self.numberOfSystems, "The number of jacobian must be the same as the number of systems."
            assert len(vecM) == len(vecB) == len(vecK) == len(vecGamma) == len(vecLambda) == len(vecForcePattern) \
                == len(vecInitialConditions), "The number of Systems in the input vectors must be the same."
            assert (type(vecM[0]) == dia_matrix and type(vecB[0]) == coo_matrix and type(vecK[0]) == coo_matrix
                    and isinstance(vecGamma[0], coo_tensor) and isinstance(vecLambda[0], coo_tensor)
                    and type(vecForcePattern[0]) == np.ndarray and type(vecInitialConditions[0]) == np.ndarray
                    and isinstance(vecPsi[0], coo_tensor)), "Inputs vectors types error."
            self._checkInputs(vecM, vecB, vecK, vecGamma, vecLambda, vecForcePattern, vecInitialConditions, vecPsi)

            sizes = np.array([m.shape[0] for m in vecM], dtype=np.int64)
            self.globalAdjointSize = int(sizes.sum())
            A = self.globalAdjointSize
            diagM = np.concatenate([np.asarray(m.data, dtype=np.float64).reshape(-1) for m in vecM])
            self.jacobian_M = dia_matrix((diagM, [0]), shape=(A, A))
            self.jacobian_B = _concat_matrices(vecB, sizes)
            self.jacobian_K = _concat_matrices(vecK, sizes)
            self.jacobian_Gamma = _concat_tensors(vecGamma, 3)
            self.jacobian_Lambda = _concat_tensors(vecLambda, 4)
            self.jacobian_Psi = _concat_tensors(vecPsi, 5)
            self.jacobian_forcePattern = np.concatenate([np.asarray(f, dtype=np.float64) for f in vecForcePattern])
            self.jacobian_initialConditions = np.concatenate([np.asarray(f, dtype=np.float64) for f in vecInitialConditions])

            # -M^-1 pre-scaling; the adjoint force pattern is NOT scaled (PASIf.py:620-633)
            self.jacobian_M, minv = _neg_inverse_of_diagonal(diagM)
            self.jacobian_B = coo_matrix(self.jacobian_M.dot(self.jacobian_B))
            self.jacobian_K = coo_matrix(self.jacobian_M.dot(self.jacobian_K))
            _scale_tensor(self.jacobian_Gamma, minv)
            _scale_tensor(self.jacobian_Lambda, minv)
            _scale_tensor(self.jacobian_Psi, minv)

            jac_csr_K = (self.jacobian_B + self.jacobian_K).tocsr()
            jac_csr_K.sort_indices()
            self._setBwdK(self.jacobian_K.shape, jac_csr_K.data, jac_csr_K.indices, jac_csr_K.indptr)
            self._setBwdGamma(self.jacobian_Gamma.dimensions, self.jacobian_Gamma.val, self.jacobian_Gamma.indices)
            self._setBwdLambda(self.jacobian_Lambda.dimensions, self.jacobian_Lambda.val, self.jacobian_Lambda.indices)
            self._setBwdForcePattern(self.jacobian_forcePattern)
            self._setBwdInitialConditions(self.jacobian_initialConditions)
            self._setBwdPsi(self.jacobian_Psi.dimensions, self.jacobian_Psi.val, self.jacobian_Psi.indices)
            self.jacobianSet = True
            self._allocateAdjointOnDevice()

        # ------------------------------------------------------------------ options (PASIf.py:371-396)
        def setInterpolationMatrix(self, interpolationMatrix_):
            for row in interpolationMatrix_:
                if len(row) % 2 != 0:
                    raise ValueError("The windows size must be even.")
                elif len(row) != len(interpolationMatrix_[0]):
                    raise ValueError("The windows size must be the same for all the rows.")
            self.interpolSize = len(interpolationMatrix_)
            self.interpolationMatrix = np.array(interpolationMatrix_).flatten()
            self._setInterpolationMatrix(self.interpolationMatrix, len(interpolationMatrix_[0]))

        def setLinearInterpolation(self, windowSize):
            """Additive: up-sample the excitations by `windowSize` / 2 RK4 steps per sample with the LINEAR 2 x W table,
            generated on the device -- the same result as ``setInterpolationMatrix`` with the table
            ``[[j / W], [1 - j / W]]`` (what the reference's callers build on the host, PASIf.py:371-385)."""
            if windowSize <= 0 or windowSize & (windowSize - 1):
                raise ValueError("The windows size must be a power of two.")
            self.interpolSize = 2
            self._setInterpolationLinear(windowSize)

        def setModulationBuffer(self, modulationBuffer_):
            if len(modulationBuffer_) == 0:
                raise ValueError("The modulation buffer must be non-empty.")
            elif len(modulationBuffer_) > 32000:
                raise ValueError("The modulation buffer don't fit in the GPU cst memory.")
            self.modulationBuffer = modulationBuffer_
            self._setModulationBuffer(self.modulationBuffer)

        # ------------------------------------------------------------------ solvers (PASIf.py:398-429)
        def getAmplitudes(self):
            if self.systemSet == False:  # noqa: E712
                raise ValueError("The system must be set before computing the amplitudes.")
            return self._getAmplitudes()

        def getTrajectory(self, saveSteps=1):
            if self.systemSet == False:  # noqa: E712
                raise ValueError("The system must be set before computing the trajectory.")
            self.saveSteps = saveSteps
            trajectory = np.asarray(self._getTrajectory(saveSteps), dtype=np.float64)
            numOfSavedSteps = int(self.numsteps / self.saveSteps)
            width = self.globalSystemSize * self.n_excitations
            out = np.zeros((width + 1, numOfSavedSteps))
            out[0] = np.arange(numOfSavedSteps) * self.saveSteps / self.sampleRate
            out[1:] = trajectory[:numOfSavedSteps * width].reshape(numOfSavedSteps, width).T
            return out

        def getGradient(self, save=0):
            if self.jacobianSet == False:  # noqa: E712
                raise ValueError("The jacobian must be set before computing the gradient.")
            return self._getGradient(save)

        def getMatrices(self):
            if not self.systemSet:
                raise ValueError("System is not set yet")
            return (self.system_B + self.system_K), self.system_Gamma, self.system_Lambda, \
                self.system_forcePattern, self.system_initialConditions

        # ------------------------------------------------------------------ validation (PASIf.py:439-472)
        def _checkInputs(self, vecM, vecB, vecK, vecGamma, vecLambda, vecF, vecIC, vecPsi=None):
            for i in range(len(vecM)):
                ndofs = vecM[i].shape[0]
                assert vecM[i].shape[0] == vecM[i].shape[1], "The matrix M must be squared."
                assert vecB[i].shape[0] == vecB[i].shape[1], "The matrix B must be squared."
                assert vecK[i].shape[0] == vecK[i].shape[1], "The matrix K must be squared."
                if vecPsi is None:
                    assert vecGamma[i].dimensions[0] == vecGamma[i].dimensions[1] == vecGamma[i].dimensions[2], \
                        "The System Gamma tensor must be squared."
                    assert vecLambda[i].dimensions[0] == vecLambda[i].dimensions[1] == vecLambda[i].dimensions[2] \
                        == vecLambda[i].dimensions[3], "The System Lambda tensor must be squared."
                assert vecB[i].shape[0] == ndofs, "The matrix B must have the same size as the matrix M."
                assert vecK[i].shape[0] == ndofs, "The matrix K must have the same size as the matrix M."
                assert vecGamma[i].dimensions[0] == ndofs, "The Gamma tensor must have the same size as the matrix M."
                assert vecLambda[i].dimensions[0] == ndofs, "The Lambda tensor must have the same size as the matrix M."
                assert vecF[i].shape[0] == ndofs, "The force pattern vector must have the same size as the matrix M."
                assert vecIC[i].shape[0] == ndofs, "The initial conditions vector must have the same size as the matrix M."
                if vecPsi is not None:
                    assert vecPsi[i].dimensions[0] == ndofs, "The Psi tensor must have the same size as the matrix M."
                    assert vecPsi[i].dimensions[0] == vecPsi[i].dimensions[1], \
                        "The two highest dim of the Psi tensor should be of the size of the adjoint system."
                    assert vecPsi[i].dimensions[2] == vecPsi[i].dimensions[3] == vecPsi[i].dimensions[4], \
                        "The tree smallest dim of the Psi tensor should be of the size of the forward system."

    PASIf.__name__ = "PASIf"
    return PASIf


def _neg_inverse_of_diagonal(diag):
    """-M^-1 of the (diagonal) global mass matrix as a sparse matrix + its diagonal (PASIf.py:609,627 call
    ``scipy.sparse.linalg.inv``, a SuperLU solve against the identity column by column: O(N) Python-level solves,
    87 s for the 4096-system config).  For a diagonal matrix that solve is ``1.0 / d`` per entry, which is what is
    computed here (element-identical: tests/test_frontend.py compares the hand-off with the reference front end).
    A zero mass entry is singular for the reference as well (SuperLU raises) and is rejected here."""
    diag = np.asarray(diag, dtype=np.float64).reshape(-1)
    if not np.all(diag != 0.0):
        raise RuntimeError("Factor is exactly singular")
    minv = -1. * (1.0 / diag)
    n = diag.size
    return csc_matrix((minv, np.arange(n), np.arange(n + 1)), shape=(n, n)), minv


def _scale_tensor(t, minv):
    """val_k *= (-1/M)[out_k]  (``coo_tensor.multiplyByDiagMatrix``, PASIf.py:102-115) and publish list views."""
    if len(t._val):
        t._val = t._val * minv[t._idx[:, 0]]
    t.val = t._val.tolist()
    t.indices = t._idx.reshape(-1).tolist()


def load_native_driver():
    """Import the in-tree pybind11 module ``PASIfgpu`` (built by ``__graft_entry__.build()``).  There is no fallback."""
    import importlib
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    try:
        mod = importlib.import_module("PASIfgpu")
    except ImportError as e:  # loud: no CPU path exists
        raise ImportError("pasif_b200: native module PASIfgpu is not built (run `python -c 'import __graft_entry__ as g; "
                          "g.build()'`): " + str(e)) from e
    return getattr(mod, "__GpuDriver")


_PASIf = None


def __getattr__(name):
    global _PASIf
    if name == "PASIf":
        if _PASIf is None:
            _PASIf = make_frontend(load_native_driver())
        return _PASIf
    raise AttributeError(name)
