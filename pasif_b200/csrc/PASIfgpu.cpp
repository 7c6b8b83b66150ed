// PASIfgpu.cpp -- pybind11 module `PASIfgpu`, class `__GpuDriver`: same constructor and the same 21 methods as the
// reference's binding (reference src/PASIfgpu.cpp:22-105), implemented as a thin shim over the C ABI in
// include/pasif_b200.h.  The unmodified reference front end (`class PASIf(__GpuDriver)`, PASIf.py:148) runs on it.
//
// Differences a caller can observe, all deliberate:
//   * errors raise RuntimeError / ValueError instead of printing and calling std::exit(0) (src/helpers.cuh:43-51);
//   * the compute type (`reel`, a compile-time macro in the reference, src/helpers.cuh:29) is chosen at run time:
//     extra keyword `dtype_` ("f64" | "f32") or environment PASIF_DTYPE, default f64;
//   * the GIL is released while the device works;
//   * engine options (adjoint schedule semantics etc.) through `_setOption(key, value)` or PASIF_COMPAT.
// NOT a difference: repeated solver calls carry state like the reference driver's (a second _getGradient without setJacobian
// continues from the previous forward / adjoint state, src/__GpuDriver.cu:522-523,743; getAmplitudes / getTrajectory reset the
// forward state afterwards, :447,482) -- pinned by tests/golden/ref_calls_*.npz; `_setOption("call_order", "independent")` opts out.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pasif_b200.h"

namespace py = pybind11;
typedef unsigned int uint;
using darr = py::array_t<double, py::array::c_style | py::array::forcecast>;
using uarr = py::array_t<uint32_t, py::array::c_style | py::array::forcecast>;

class GpuDriver {
 public:
  GpuDriver(py::object excitationSet_, uint sampleRate_, uint numsteps_, bool dCompute_, bool dSystem_, bool dSolver_,
            uint GPUId_, std::string dtype_)
      : dCompute(dCompute_), dSystem(dSystem_), dSolver(dSolver_) {
    if (dtype_.empty()) {
      const char* env = std::getenv("PASIF_DTYPE");
      dtype_ = env ? env : "f64";
    }
    int dt;
    if (dtype_ == "f64" || dtype_ == "double" || dtype_ == "float64") dt = PASIF_F64;
    else if (dtype_ == "f32" || dtype_ == "float" || dtype_ == "float32") dt = PASIF_F32;
    else throw std::invalid_argument("dtype_ must be 'f64' or 'f32'");
    darr exc = excitation_matrix(excitationSet_);
    if (pasif_create(&h, exc.data(), (uint32_t)exc.shape(0), (uint32_t)exc.shape(1), sampleRate_, numsteps_, dt, (int)GPUId_) != 0)
      throw std::runtime_error(std::string("PASIfgpu: ") + pasif_last_error(nullptr));
    if (const char* c = std::getenv("PASIF_COMPAT")) check(pasif_set_option(h, "compat", c));
  }
  ~GpuDriver() { pasif_destroy(h); }
  GpuDriver(const GpuDriver&) = delete;
  GpuDriver& operator=(const GpuDriver&) = delete;

  void setFwdK(std::array<uint, 2> n_, darr values_, uarr indices_, uarr indptr_) {
    check_csr(n_, values_, indices_, indptr_);
    check(pasif_set_fwd_k(h, n_[0], n_[1], values_.data(), indices_.data(), indptr_.data(), (uint32_t)values_.size()));
  }
  void setFwdGamma(std::array<uint, 3> n_, darr v, uarr i) { check_t(3, v, i); check(pasif_set_fwd_gamma(h, n_.data(), v.data(), i.data(), (uint32_t)v.size())); }
  void setFwdLambda(std::array<uint, 4> n_, darr v, uarr i) { check_t(4, v, i); check(pasif_set_fwd_lambda(h, n_.data(), v.data(), i.data(), (uint32_t)v.size())); }
  void setFwdPsi(std::array<uint, 5> n_, darr v, uarr i) { check_t(5, v, i); check(pasif_set_fwd_psi(h, n_.data(), v.data(), i.data(), (uint32_t)v.size())); }
  void setFwdForcePattern(darr f) { check(pasif_set_fwd_force(h, f.data(), (uint32_t)f.size())); }
  void setFwdInitialConditions(darr ic) { check(pasif_set_fwd_ic(h, ic.data(), (uint32_t)ic.size())); }
  void allocateSystemOnDevice() { py::gil_scoped_release r; check(pasif_alloc_system(h)); }

  void setBwdK(std::array<uint, 2> n_, darr values_, uarr indices_, uarr indptr_) {
    check_csr(n_, values_, indices_, indptr_);
    check(pasif_set_bwd_k(h, n_[0], n_[1], values_.data(), indices_.data(), indptr_.data(), (uint32_t)values_.size()));
  }
  void setBwdGamma(std::array<uint, 3> n_, darr v, uarr i) { check_t(3, v, i); check(pasif_set_bwd_gamma(h, n_.data(), v.data(), i.data(), (uint32_t)v.size())); }
  void setBwdLambda(std::array<uint, 4> n_, darr v, uarr i) { check_t(4, v, i); check(pasif_set_bwd_lambda(h, n_.data(), v.data(), i.data(), (uint32_t)v.size())); }
  void setBwdPsi(std::array<uint, 5> n_, darr v, uarr i) { check_t(5, v, i); check(pasif_set_bwd_psi(h, n_.data(), v.data(), i.data(), (uint32_t)v.size())); }
  void setBwdForcePattern(darr f) { check(pasif_set_bwd_force(h, f.data(), (uint32_t)f.size())); }
  void setBwdInitialConditions(darr ic) { check(pasif_set_bwd_ic(h, ic.data(), (uint32_t)ic.size())); }
  void allocateAdjointOnDevice() { py::gil_scoped_release r; check(pasif_alloc_adjoint(h)); }

  int loadExcitationsSet(py::object excitationSet_, uint sampleRate_) {
    darr exc = excitation_matrix(excitationSet_);
    check(pasif_load_excitations(h, exc.data(), (uint32_t)exc.shape(0), (uint32_t)exc.shape(1), sampleRate_));
    return 0;
  }
  void setInterpolationMatrix(darr m, uint window) { check(pasif_set_interp(h, m.data(), (uint32_t)m.size(), window)); }
  void setInterpolationLinear(uint window) { check(pasif_set_interp_linear(h, window)); }
  void setModulationBuffer(darr m) { check(pasif_set_modulation(h, m.data(), (uint32_t)m.size())); }
  void setOption(const std::string& k, const std::string& v) { check(pasif_set_option(h, k.c_str(), v.c_str())); }
  double query(const std::string& k) {
    double v = 0;
    if (pasif_query(h, k.c_str(), &v) != 0) throw std::invalid_argument("unknown query key " + k);
    return v;
  }

  std::vector<double> getAmplitudes() {
    std::vector<double> out((size_t)(query("n_fwd") * query("n_excitations")));
    { py::gil_scoped_release r; check(pasif_get_amplitudes(h, out.data(), out.size())); }
    return out;
  }
  std::vector<double> getTrajectory(uint saveSteps_) {
    if (saveSteps_ == 0) throw std::invalid_argument("saveSteps_ must be >= 1");
    std::vector<double> out(pasif_trajectory_size(h, saveSteps_));
    { py::gil_scoped_release r; check(pasif_get_trajectory(h, saveSteps_, out.data(), out.size())); }
    return out;
  }
  std::vector<double> getGradient(uint save) {
    std::vector<double> out((size_t)(query("n_adj") * query("n_excitations")));
    { py::gil_scoped_release r; check(pasif_get_gradient(h, save, out.data(), out.size())); }
    return out;
  }
  std::vector<double> getGradientReduced(uint save) {
    std::vector<double> out((size_t)query("n_adj"));
    { py::gil_scoped_release r; check(pasif_get_gradient_reduced(h, save, out.data(), out.size())); }
    return out;
  }
  // excitation-summed adjoint state written to DEVICE memory of the engine's GPU (`device_ptr`: n_adj values of the
  // engine's compute type, e.g. torch_tensor.data_ptr()); the write is complete when the call returns
  void getGradientReducedDevice(uintptr_t device_ptr, uint save) {
    py::gil_scoped_release r;
    check(pasif_get_gradient_reduced_device(h, save, reinterpret_cast<void*>(device_ptr), (uint64_t)query("n_adj")));
  }
  void reduceLastGradientDevice(uintptr_t device_ptr) {
    py::gil_scoped_release r;
    check(pasif_reduce_last_gradient_device(h, reinterpret_cast<void*>(device_ptr), (uint64_t)query("n_adj")));
  }

 private:
  pasif_engine* h = nullptr;
  bool dCompute, dSystem, dSolver;

  void check(int rc) {
    if (rc != 0) throw std::runtime_error(std::string("PASIfgpu: ") + pasif_last_error(h));
  }
  static void check_csr(const std::array<uint, 2>& n, const darr& v, const uarr& idx, const uarr& ptr) {
    if (v.size() != idx.size()) throw std::invalid_argument("CSR values/indices length mismatch");
    if ((size_t)ptr.size() != (size_t)n[0] + 1) throw std::invalid_argument("CSR indptr must have n_rows + 1 entries");
  }
  static void check_t(int order, const darr& v, const uarr& idx) {
    if ((size_t)idx.size() != (size_t)order * (size_t)v.size())
      throw std::invalid_argument("COO tensor needs `order` indices per value");
  }
  // list of equally long 1-D arrays (or a 2-D array) -> contiguous [E][L] doubles; ragged input is an error
  // (the reference prints "[Error] ... not of the same size" and carries on, src/__GpuDriver.cu:358-363)
  static darr excitation_matrix(const py::object& o) {
    darr a;
    try {
      a = darr::ensure(o);
    } catch (...) {
      a = darr();
    }
    if (!a || a.ndim() != 2) {
      PyErr_Clear();
      throw std::invalid_argument("Excitations vectors are not of the same size (need a rectangular [E][L] set)");
    }
    if (a.shape(0) == 0 || a.shape(1) == 0) throw std::invalid_argument("empty excitation set");
    return a;
  }
};

PYBIND11_MODULE(PASIfgpu, m) {
  m.doc() = "Parallel Accelerated Solver Interface -- B200-native engine (pasif_b200)";
  py::class_<GpuDriver>(m, "__GpuDriver")
      .def(py::init<py::object, uint, uint, bool, bool, bool, uint, std::string>(), py::arg("excitationSet_"),
           py::arg("sampleRate_"), py::arg("numsteps_"), py::arg("dCompute_") = false, py::arg("dSystem_") = false,
           py::arg("dSolver_") = false, py::arg("GPUId_") = 0, py::arg("dtype_") = "")
      // forward system interface
      .def("_setFwdK", &GpuDriver::setFwdK, py::arg("n_"), py::arg("values_"), py::arg("indices_"), py::arg("indptr_"))
      .def("_setFwdGamma", &GpuDriver::setFwdGamma, py::arg("n_"), py::arg("values_"), py::arg("indices_"))
      .def("_setFwdLambda", &GpuDriver::setFwdLambda, py::arg("n_"), py::arg("values_"), py::arg("indices_"))
      .def("_setFwdPsi", &GpuDriver::setFwdPsi, py::arg("n_"), py::arg("values_"), py::arg("indices_"))
      .def("_setFwdForcePattern", &GpuDriver::setFwdForcePattern, py::arg("forcePattern_"))
      .def("_setFwdInitialConditions", &GpuDriver::setFwdInitialConditions, py::arg("initialConditions_"))
      .def("_allocateSystemOnDevice", &GpuDriver::allocateSystemOnDevice)
      // backward system interface
      .def("_setBwdK", &GpuDriver::setBwdK, py::arg("n_"), py::arg("values_"), py::arg("indices_"), py::arg("indptr_"))
      .def("_setBwdGamma", &GpuDriver::setBwdGamma, py::arg("n_"), py::arg("values_"), py::arg("indices_"))
      .def("_setBwdLambda", &GpuDriver::setBwdLambda, py::arg("n_"), py::arg("values_"), py::arg("indices_"))
      .def("_setBwdPsi", &GpuDriver::setBwdPsi, py::arg("n_"), py::arg("values_"), py::arg("indices_"))
      .def("_setBwdForcePattern", &GpuDriver::setBwdForcePattern, py::arg("forcePattern_"))
      .def("_setBwdInitialConditions", &GpuDriver::setBwdInitialConditions, py::arg("initialConditions_"))
      .def("_allocateAdjointOnDevice", &GpuDriver::allocateAdjointOnDevice)
      // compute options interface
      .def("_loadExcitationsSet", &GpuDriver::loadExcitationsSet, py::arg("excitationSet_"), py::arg("sampleRate_"))
      .def("_setInterpolationMatrix", &GpuDriver::setInterpolationMatrix, py::arg("interpolationMatrix_"),
           py::arg("interpolationWindowSize_"))
      .def("_setInterpolationLinear", &GpuDriver::setInterpolationLinear, py::arg("interpolationWindowSize_"))
      .def("_setModulationBuffer", &GpuDriver::setModulationBuffer, py::arg("modulationBuffer_"))
      // solvers interface
      .def("_getAmplitudes", &GpuDriver::getAmplitudes)
      .def("_getTrajectory", &GpuDriver::getTrajectory, py::arg("saveSteps_") = 1)
      .def("_getGradient", &GpuDriver::getGradient, py::arg("save"))
      // additive
      .def("_getGradientReduced", &GpuDriver::getGradientReduced, py::arg("save") = 0)
      .def("_getGradientReducedDevice", &GpuDriver::getGradientReducedDevice, py::arg("device_ptr"), py::arg("save") = 0)
      .def("_reduceLastGradientDevice", &GpuDriver::reduceLastGradientDevice, py::arg("device_ptr"))
      .def("_setOption", &GpuDriver::setOption, py::arg("key"), py::arg("value"))
      .def("_query", &GpuDriver::query, py::arg("key"));
}
