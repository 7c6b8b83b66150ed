"""pasif_b200 -- B200-native batched RK4 forward / adjoint-gradient engine behind the PASIf front end.

    from pasif_b200 import PASIf, coo_tensor          # same public API as the reference's PASIf.py

Layout (only what the hot path needs):
    csrc/rk4_kernels.cuh   register-resident RK4 forward / adjoint kernels, templated on the sparsity pattern
    csrc/pattern.hpp       hand-off analysis: components -> pattern classes -> compile-time traits
    csrc/engine.cu         host engine, CUDA-graph-free launch schedule, NVRTC specialisation, C ABI
    csrc/PASIfgpu.cpp      pybind11 module `PASIfgpu.__GpuDriver` (the reference's boundary) over the C ABI
    capi.py                ctypes binding of the C ABI (NumPy host buffers)
    frontend.py            PASIf / coo_tensor front end (vectorised assembly, reference-identical hand-off)
    sharded.py             one-process-per-GPU sharding of the (system x excitation) batch + gradient all-reduce
    workloads.py           synthetic systems / excitations of the BASELINE configs
The CUDA extension is mandatory: there is no CPU path in this package (the CPU oracle lives in /oracle, tests only).
"""
from .frontend import coo_tensor, make_frontend  # noqa: F401


def __getattr__(name):
    if name == "PASIf":
        from . import frontend
        return frontend.PASIf
    raise AttributeError(name)
