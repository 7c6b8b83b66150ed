// pattern.hpp -- host-side structural analysis of the hand-off (no CUDA in this file).
//
// The reference keeps the block-diagonal GLOBAL system it receives from PASIf.py (CSR K', COO Gamma'/Lambda'/Psi',
// dense F', IC; reference src/helpers.cu:24-776) and replicates it over the excitations
// (src/__GpuDriver.cu:896-938, 1196-1214).  This engine instead
//   1. splits the global system into its independent COMPONENTS (connected components of the coupling graph --
//      the user's "systems", or finer),
//   2. groups components with the same local sparsity pattern into CLASSES (one compiled kernel per class, the
//      coefficient values of every instance in an instance-minor table),
//   3. emits the compile-time traits (`struct` with constexpr index tables) the kernels in rk4_kernels.cuh are
//      instantiated with.
// For the gradient the same is done on the JOINT graph of forward + adjoint variables: adjoint tensors reference
// forward variables in their trailing modes (dims [N_adj, N_adj, N_fwd, ...], src/__GpuDriver.cu:832-833).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace pasif_host {

struct Tensor {
  int order = 3;
  std::vector<uint32_t> dims;
  std::vector<double> val;
  std::vector<uint32_t> idx;  // flat, `order` per nnz
  size_t nnz() const { return val.size(); }
};

struct SystemData {
  uint32_t n = 0;
  bool have_k = false, have_f = false, have_ic = false;
  std::vector<double> kval;
  std::vector<uint32_t> kidx, kptr;
  Tensor G, L, P;
  std::vector<double> F, IC;
  SystemData() { G.order = 3; L.order = 4; P.order = 5; }
};

// component-local sparsity pattern of one differentiated system
struct Pattern {
  int N = 0;
  std::vector<int> nv;  // adjoint: needed forward variables (forward-local indices); tensor p-indices refer to this list
  std::vector<int> k_row, k_col, g_o, g_a, g_b, l_o, l_a, l_b, l_c, p_o, p_a, p_b, p_c, p_d, f_row;
  // lane-block classes (lane_split.hpp): the component is integrated by `lanes` threads; x-operand indices >= N and
  // p-operand indices >= nv.size() address HALO entries: entry h = published slot pub_slot[hx_pub[h]] (== hx_slot[h]) of
  // lane (lane + hx_delta[h]) mod lanes; likewise hp_* / ppub_slot over the set-point (nv) list of the adjoint side
  // K terms whose coefficient is exactly 1 for every instance (x' = v of a second-order system in first-order form): the
  // kernels add the operand instead of multiplying by a register (empty: none)
  std::vector<int> k_unit;
  int lanes = 1;
  std::vector<int> hx_slot, hx_delta, hx_pub, pub_slot, hp_slot, hp_delta, hp_pub, ppub_slot;
  int nc() const { return (int)(k_row.size() + g_o.size() + l_o.size() + p_o.size() + f_row.size()); }

  std::string signature() const {
    std::ostringstream s;
    auto put = [&](const char* tag, const std::vector<int>& v) {
      s << tag;
      for (int x : v) s << x << ',';
      s << ';';
    };
    s << "N" << N << ';';
    put("nv", nv); put("kr", k_row); put("kc", k_col);
    put("go", g_o); put("ga", g_a); put("gb", g_b);
    put("lo", l_o); put("la", l_a); put("lb", l_b); put("lc", l_c);
    put("po", p_o); put("pa", p_a); put("pb", p_b); put("pc", p_c); put("pd", p_d);
    put("f", f_row);
    bool any_unit = false;
    for (int u : k_unit) any_unit = any_unit || u != 0;
    if (any_unit) put("ku", k_unit);
    if (lanes > 1) {
      s << "G" << lanes << ';';
      put("hxs", hx_slot); put("hxd", hx_delta); put("hps", hp_slot); put("hpd", hp_delta);
    }
    return s.str();
  }
};

struct ClassData {
  bool joint = false;
  int lanes = 1;     // > 1: lane-block class (lane_split.hpp): n_inst counts VIRTUAL instances (instance * lanes + lane)
  bool lane_pair = false;   // lane-block class sized for the fp32 pair kernels (two trajectories per thread: state slots count twice)
  Pattern fwd;  // forward pattern (joint: forward part of the joint component)
  Pattern adj;  // joint only
  uint64_t hash = 0;
  std::string name;  // F_<hash> or J_<hash>
  int n_inst = 0;
  // instance-minor tables
  std::vector<double> coef, ic;        // [NC][n_inst], [N][n_inst]
  std::vector<int> varmap;             // [N][n_inst] global forward index (-1: padding variable)
  std::vector<double> coef_adj, ic_adj;
  std::vector<int> varmap_adj;
};

inline uint64_t fnv1a(const std::string& s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}

struct UnionFind {
  std::vector<uint32_t> p;
  explicit UnionFind(size_t n) : p(n) { std::iota(p.begin(), p.end(), 0u); }
  uint32_t find(uint32_t x) {
    while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
    return x;
  }
  void unite(uint32_t a, uint32_t b) {
    a = find(a); b = find(b);
    if (a != b) { if (a < b) p[b] = a; else p[a] = b; }
  }
};

inline void check_system(const SystemData& s, const char* what, uint32_t n_setpoint) {
  auto fail = [&](const std::string& m) { throw std::runtime_error(std::string(what) + ": " + m); };
  if (!s.have_k) fail("K was not set");
  if (!s.have_ic) fail("initial conditions were not set");
  if (s.kptr.size() != (size_t)s.n + 1) fail("CSR indptr length != n_rows + 1");
  if (s.IC.size() != s.n) fail("initial conditions length != system size");
  if (s.have_f && s.F.size() != s.n) fail("force pattern length != system size");
  if (s.kptr.back() != s.kval.size()) fail("CSR indptr[n] != nnz");
  for (uint32_t c : s.kidx) if (c >= s.n) fail("CSR column index out of range");
  const Tensor* ts[3] = {&s.G, &s.L, &s.P};
  for (const Tensor* t : ts) {
    if (t->idx.size() != t->nnz() * (size_t)t->order) fail("tensor index array length != order * nnz");
    for (size_t k = 0; k < t->nnz(); ++k) {
      const uint32_t* ix = &t->idx[k * t->order];
      if (ix[0] >= s.n || ix[1] >= s.n) fail("tensor out/in1 index out of range");
      for (int m = 2; m < t->order; ++m) if (ix[m] >= n_setpoint) fail("tensor setpoint index out of range");
    }
  }
}

namespace detail {

struct CompBuild {
  std::vector<uint32_t> fvars, avars;  // global indices, ascending
  Pattern fwd, adj;
  std::vector<double> cf, ca;          // coefficient values in table order, built per kind then concatenated
  std::vector<double> kf, gf, lf, pf, ff, ka, ga, la, pa;
};

// fills the forward-type pattern of every component: `comp_of[v]` component id of variable v, `local[v]` local index
inline void fill_differentiated(const SystemData& s, const std::vector<int>& comp_of, const std::vector<int>& local,
                                std::vector<CompBuild>& comps, bool adjoint, const std::vector<int>* fcomp_local,
                                double eps, bool as_float) {
  // K rows
  for (uint32_t r = 0; r < s.n; ++r) {
    int c = comp_of[r];
    if (c < 0) continue;
    Pattern& p = adjoint ? comps[c].adj : comps[c].fwd;
    auto& kv = adjoint ? comps[c].ka : comps[c].kf;
    for (uint32_t k = s.kptr[r]; k < s.kptr[r + 1]; ++k) {
      p.k_row.push_back(local[r]);
      p.k_col.push_back(local[s.kidx[k]]);
      kv.push_back(s.kval[k]);
    }
  }
  auto setpoint_local = [&](uint32_t g) { return fcomp_local ? (*fcomp_local)[g] : local[g]; };
  for (size_t k = 0; k < s.G.nnz(); ++k) {
    const uint32_t* ix = &s.G.idx[3 * k];
    int c = comp_of[ix[0]];
    if (c < 0) continue;
    Pattern& p = adjoint ? comps[c].adj : comps[c].fwd;
    p.g_o.push_back(local[ix[0]]); p.g_a.push_back(local[ix[1]]); p.g_b.push_back(setpoint_local(ix[2]));
    (adjoint ? comps[c].ga : comps[c].gf).push_back(s.G.val[k]);
  }
  for (size_t k = 0; k < s.L.nnz(); ++k) {
    const uint32_t* ix = &s.L.idx[4 * k];
    int c = comp_of[ix[0]];
    if (c < 0) continue;
    Pattern& p = adjoint ? comps[c].adj : comps[c].fwd;
    p.l_o.push_back(local[ix[0]]); p.l_a.push_back(local[ix[1]]);
    p.l_b.push_back(setpoint_local(ix[2])); p.l_c.push_back(setpoint_local(ix[3]));
    (adjoint ? comps[c].la : comps[c].lf).push_back(s.L.val[k]);
  }
  for (size_t k = 0; k < s.P.nnz(); ++k) {
    const uint32_t* ix = &s.P.idx[5 * k];
    int c = comp_of[ix[0]];
    if (c < 0) continue;
    Pattern& p = adjoint ? comps[c].adj : comps[c].fwd;
    p.p_o.push_back(local[ix[0]]); p.p_a.push_back(local[ix[1]]);
    p.p_b.push_back(setpoint_local(ix[2])); p.p_c.push_back(setpoint_local(ix[3])); p.p_d.push_back(setpoint_local(ix[4]));
    (adjoint ? comps[c].pa : comps[c].pf).push_back(s.P.val[k]);
  }
  if (!adjoint && s.have_f) {
    // COOVector ctor: entries with |F_i| <= eps(reel) are dropped (reference src/helpers.cu:674-681)
    for (uint32_t v = 0; v < s.n; ++v) {
      int c = comp_of[v];
      if (c < 0) continue;
      double f = as_float ? (double)(float)s.F[v] : s.F[v];
      if (std::fabs(f) > eps) { comps[c].fwd.f_row.push_back(local[v]); comps[c].ff.push_back(s.F[v]); }
    }
  }
}

inline void unite_system(UnionFind& uf, const SystemData& s, uint32_t base, uint32_t setpoint_base) {
  for (uint32_t r = 0; r < s.n; ++r)
    for (uint32_t k = s.kptr[r]; k < s.kptr[r + 1]; ++k) uf.unite(base + r, base + s.kidx[k]);
  const Tensor* ts[3] = {&s.G, &s.L, &s.P};
  for (const Tensor* t : ts)
    for (size_t k = 0; k < t->nnz(); ++k) {
      const uint32_t* ix = &t->idx[k * t->order];
      uf.unite(base + ix[0], base + ix[1]);
      for (int m = 2; m < t->order; ++m) uf.unite(base + ix[0], setpoint_base + ix[m]);
    }
}

inline void transpose_into(std::vector<double>& dst, const std::vector<std::vector<double>>& rows, int width) {
  const int n = (int)rows.size();
  dst.assign((size_t)width * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < width; ++i) dst[(size_t)i * n + j] = rows[j][i];
}
inline void transpose_into(std::vector<int>& dst, const std::vector<std::vector<int>>& rows, int width) {
  const int n = (int)rows.size();
  dst.assign((size_t)width * n, 0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < width; ++i) dst[(size_t)i * n + j] = rows[j][i];
}

struct ClassAccum {
  ClassData cd;
  std::vector<std::vector<double>> coef, ic, coef_adj, ic_adj;
  std::vector<std::vector<int>> vm, vm_adj;
};

inline std::vector<ClassData> finish(std::vector<ClassAccum>& acc) {
  std::vector<ClassData> out;
  for (auto& a : acc) {
    ClassData& c = a.cd;
    c.n_inst = (int)a.coef.size();
    transpose_into(c.coef, a.coef, c.fwd.nc());
    transpose_into(c.ic, a.ic, c.fwd.N);
    transpose_into(c.varmap, a.vm, c.fwd.N);
    if (c.joint) {
      transpose_into(c.coef_adj, a.coef_adj, c.adj.nc());
      transpose_into(c.ic_adj, a.ic_adj, c.adj.N);
      transpose_into(c.varmap_adj, a.vm_adj, c.adj.N);
    }
    out.push_back(std::move(c));
  }
  return out;
}

inline std::string hex64(uint64_t h) {
  char b[32];
  std::snprintf(b, sizeof b, "%016llx", (unsigned long long)h);
  return b;
}

}  // namespace detail

// ---- forward-only analysis (getAmplitudes / getTrajectory)
inline std::vector<ClassData> analyze_forward(const SystemData& f, double eps, bool as_float) {
  using namespace detail;
  check_system(f, "forward system", f.n);
  UnionFind uf(f.n);
  unite_system(uf, f, 0, 0);
  std::vector<int> comp_of(f.n, -1), local(f.n, 0);
  std::vector<CompBuild> comps;
  {
    std::unordered_map<uint32_t, int> root2comp;
    for (uint32_t v = 0; v < f.n; ++v) {
      uint32_t r = uf.find(v);
      auto it = root2comp.find(r);
      int c;
      if (it == root2comp.end()) { c = (int)comps.size(); root2comp[r] = c; comps.emplace_back(); }
      else c = it->second;
      comp_of[v] = c;
      local[v] = (int)comps[c].fvars.size();
      comps[c].fvars.push_back(v);
    }
  }
  for (auto& c : comps) c.fwd.N = (int)c.fvars.size();
  fill_differentiated(f, comp_of, local, comps, false, nullptr, eps, as_float);

  std::vector<ClassAccum> acc;
  std::unordered_map<std::string, int> sig2class;
  for (auto& c : comps) {
    std::string sig = "F|" + c.fwd.signature();
    auto it = sig2class.find(sig);
    int ci;
    if (it == sig2class.end()) {
      ci = (int)acc.size();
      sig2class[sig] = ci;
      acc.emplace_back();
      acc[ci].cd.fwd = c.fwd;
      acc[ci].cd.hash = fnv1a(sig);
      acc[ci].cd.name = "F_" + hex64(acc[ci].cd.hash);
    } else ci = it->second;
    std::vector<double> co;
    for (auto* v : {&c.kf, &c.gf, &c.lf, &c.pf, &c.ff}) co.insert(co.end(), v->begin(), v->end());
    acc[ci].coef.push_back(std::move(co));
    std::vector<double> ic;
    std::vector<int> vm;
    for (uint32_t g : c.fvars) { ic.push_back(f.IC[g]); vm.push_back((int)g); }
    acc[ci].ic.push_back(std::move(ic));
    acc[ci].vm.push_back(std::move(vm));
  }
  return finish(acc);
}

// ---- joint analysis (getGradient)
inline std::vector<ClassData> analyze_joint(const SystemData& f, const SystemData& b, double eps, bool as_float) {
  using namespace detail;
  check_system(f, "forward system", f.n);
  check_system(b, "adjoint system", f.n);
  const uint32_t NF = f.n, NA = b.n;
  UnionFind uf((size_t)NF + NA);
  unite_system(uf, f, 0, 0);
  unite_system(uf, b, NF, 0);
  // components that contain at least one adjoint variable, numbered by their first adjoint variable
  std::unordered_map<uint32_t, int> root2comp;
  std::vector<CompBuild> comps;
  std::vector<int> acomp(NA, -1), alocal(NA, 0), fcomp(NF, -1), flocal(NF, 0);
  for (uint32_t v = 0; v < NA; ++v) {
    uint32_t r = uf.find(NF + v);
    auto it = root2comp.find(r);
    int c;
    if (it == root2comp.end()) { c = (int)comps.size(); root2comp[r] = c; comps.emplace_back(); }
    else c = it->second;
    acomp[v] = c;
    alocal[v] = (int)comps[c].avars.size();
    comps[c].avars.push_back(v);
  }
  for (uint32_t v = 0; v < NF; ++v) {
    auto it = root2comp.find(uf.find(v));
    if (it == root2comp.end()) continue;   // forward component no adjoint variable depends on
    int c = it->second;
    fcomp[v] = c;
    flocal[v] = (int)comps[c].fvars.size();
    comps[c].fvars.push_back(v);
  }
  for (auto& c : comps) { c.fwd.N = (int)c.fvars.size(); c.adj.N = (int)c.avars.size(); }
  fill_differentiated(f, fcomp, flocal, comps, false, nullptr, eps, as_float);
  fill_differentiated(b, acomp, alocal, comps, true, &flocal, eps, as_float);
  // needed forward variables: remap the adjoint tensors' setpoint indices onto the sorted list of referenced locals
  for (auto& c : comps) {
    std::map<int, int> used;
    for (auto* v : {&c.adj.g_b, &c.adj.l_b, &c.adj.l_c, &c.adj.p_b, &c.adj.p_c, &c.adj.p_d})
      for (int x : *v) used[x] = 0;
    int k = 0;
    for (auto& kv : used) { kv.second = k++; c.adj.nv.push_back(kv.first); }
    for (auto* v : {&c.adj.g_b, &c.adj.l_b, &c.adj.l_c, &c.adj.p_b, &c.adj.p_c, &c.adj.p_d})
      for (int& x : *v) x = used[x];
  }

  std::vector<ClassAccum> acc;
  std::unordered_map<std::string, int> sig2class;
  for (auto& c : comps) {
    std::string sig = "J|" + c.fwd.signature() + "|A|" + c.adj.signature();
    auto it = sig2class.find(sig);
    int ci;
    if (it == sig2class.end()) {
      ci = (int)acc.size();
      sig2class[sig] = ci;
      acc.emplace_back();
      acc[ci].cd.joint = true;
      acc[ci].cd.fwd = c.fwd;
      acc[ci].cd.adj = c.adj;
      if (acc[ci].cd.fwd.N == 0) acc[ci].cd.fwd.N = 1;   // adjoint component without forward variables: padding variable
      acc[ci].cd.hash = fnv1a(sig);
      acc[ci].cd.name = "J_" + hex64(acc[ci].cd.hash);
    } else ci = it->second;
    std::vector<double> co, ca;
    for (auto* v : {&c.kf, &c.gf, &c.lf, &c.pf, &c.ff}) co.insert(co.end(), v->begin(), v->end());
    for (auto* v : {&c.ka, &c.ga, &c.la, &c.pa}) ca.insert(ca.end(), v->begin(), v->end());
    acc[ci].coef.push_back(std::move(co));
    acc[ci].coef_adj.push_back(std::move(ca));
    std::vector<double> ic, ica;
    std::vector<int> vm, vma;
    for (uint32_t g : c.fvars) { ic.push_back(f.IC[g]); vm.push_back((int)g); }
    if (c.fvars.empty()) { ic.push_back(0.0); vm.push_back(-1); }
    for (uint32_t g : c.avars) { ica.push_back(b.IC[g]); vma.push_back((int)g); }
    acc[ci].ic.push_back(std::move(ic));
    acc[ci].vm.push_back(std::move(vm));
    acc[ci].ic_adj.push_back(std::move(ica));
    acc[ci].vm_adj.push_back(std::move(vma));
  }
  return finish(acc);
}

// ---- unit coefficients: flags the K terms whose value is exactly 1 for every instance of the class (lane classes: for
// every lane that owns the row; a lane where both the row and the operand are padding slots may take it as 1 as well,
// padding variables start at 0 and then stay 0).  The flags are part of the class signature (a different kernel).
inline void mark_unit_side(Pattern& p, const std::vector<double>& coef, const std::vector<int>& varmap, int n_inst, int lanes) {
  const int nk = (int)p.k_row.size();
  p.k_unit.assign((size_t)nk, 0);
  bool any = false;
  for (int k = 0; k < nk; ++k) {
    const int row = p.k_row[(size_t)k], col = p.k_col[(size_t)k];
    if (col >= p.N) continue;                                   // halo operand
    bool ok = n_inst > 0;
    for (int j = 0; j < n_inst && ok; ++j) {
      const double v = coef[(size_t)k * n_inst + j];
      if (v == 1.0) continue;
      const bool pad = lanes > 1 && v == 0.0 && varmap[(size_t)row * n_inst + j] < 0 && varmap[(size_t)col * n_inst + j] < 0;
      if (!pad) ok = false;
    }
    if (ok) { p.k_unit[(size_t)k] = 1; any = true; }
  }
  if (!any) p.k_unit.clear();
}
inline void mark_unit_terms(ClassData& c) {
  mark_unit_side(c.fwd, c.coef, c.varmap, c.n_inst, c.lanes);
  if (c.joint) mark_unit_side(c.adj, c.coef_adj, c.varmap_adj, c.n_inst, c.lanes);
  if (c.fwd.k_unit.empty() && c.adj.k_unit.empty()) return;
  const std::string sig = c.joint ? std::string(c.lanes > 1 ? "JL|" : "J|") + c.fwd.signature() + "|A|" + c.adj.signature()
                                  : std::string(c.lanes > 1 ? "FL|" : "F|") + c.fwd.signature();
  c.hash = fnv1a(sig);
  c.name = std::string(c.joint ? "J_" : "F_") + detail::hex64(c.hash);
}

// ---- traits emission (consumed by rk4_kernels.cuh; identical text for the AOT include and the NVRTC source)
inline void emit_table(std::ostringstream& s, const char* name, const std::vector<int>& v) {
  s << "  PASIF_TAB(" << name;
  if (v.empty()) s << ", 0";
  for (int x : v) s << ", " << x;
  s << ")\n";
}

inline std::string emit_pattern_struct(const std::string& name, const Pattern& p, const char* indent = "") {
  std::ostringstream s;
  const int nk = (int)p.k_row.size(), ng = (int)p.g_o.size(), nl = (int)p.l_o.size(), np = (int)p.p_o.size(),
            nf = (int)p.f_row.size();
  s << "struct " << name << " {\n";
  s << "  static constexpr int N = " << p.N << ", NV = " << p.nv.size() << ", NNZK = " << nk << ", NNZG = " << ng
    << ", NNZL = " << nl << ", NNZP = " << np << ", NF = " << nf << ";\n";
  s << "  static constexpr int G = " << p.lanes << ", H = " << p.hx_slot.size() << ", NPUB = " << p.pub_slot.size() << ", HP = "
    << p.hp_slot.size() << ", NPPUB = " << p.ppub_slot.size() << ";\n";
  s << "  static constexpr int OFF_K = 0, OFF_G = " << nk << ", OFF_L = " << nk + ng << ", OFF_P = " << nk + ng + nl
    << ", OFF_F = " << nk + ng + nl + np << ", NC = " << nk + ng + nl + np + nf << ";\n";
  emit_table(s, "nv", p.nv);
  emit_table(s, "k_row", p.k_row); emit_table(s, "k_col", p.k_col);
  {
    std::vector<int> unit((size_t)nk, 0), first((size_t)nk, 0);
    for (int k = 0; k < nk; ++k) {
      unit[(size_t)k] = k < (int)p.k_unit.size() ? p.k_unit[(size_t)k] : 0;
      first[(size_t)k] = 1;
      for (int j = 0; j < k; ++j) if (p.k_row[(size_t)j] == p.k_row[(size_t)k]) first[(size_t)k] = 0;
    }
    emit_table(s, "k_unit", unit); emit_table(s, "k_first", first);
  }
  emit_table(s, "g_o", p.g_o); emit_table(s, "g_a", p.g_a); emit_table(s, "g_b", p.g_b);
  emit_table(s, "l_o", p.l_o); emit_table(s, "l_a", p.l_a); emit_table(s, "l_b", p.l_b); emit_table(s, "l_c", p.l_c);
  emit_table(s, "p_o", p.p_o); emit_table(s, "p_a", p.p_a); emit_table(s, "p_b", p.p_b); emit_table(s, "p_c", p.p_c);
  emit_table(s, "p_d", p.p_d);
  emit_table(s, "f_row", p.f_row);
  emit_table(s, "pub_slot", p.pub_slot); emit_table(s, "hx_pub", p.hx_pub); emit_table(s, "hx_delta", p.hx_delta);
  emit_table(s, "ppub_slot", p.ppub_slot); emit_table(s, "hp_pub", p.hp_pub); emit_table(s, "hp_delta", p.hp_delta);
  s << "};\n";
  (void)indent;
  return s.str();
}

inline std::string emit_class_traits(const ClassData& c) {
  std::ostringstream s;
  if (!c.joint) {
    s << emit_pattern_struct(c.name, c.fwd);
  } else {
    s << emit_pattern_struct(c.name + "_Fwd", c.fwd);
    s << emit_pattern_struct(c.name + "_Adj", c.adj);
    s << "struct " << c.name << " { typedef " << c.name << "_Fwd Fwd; typedef " << c.name << "_Adj Adj; };\n";
  }
  return s.str();
}

}  // namespace pasif_host
