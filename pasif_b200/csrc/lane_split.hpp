// lane_split.hpp -- host-side partitioner of the LANE-BLOCK path (no CUDA in this file).
//
// A component with more coupled state variables than one thread can keep in registers (BASELINE config 4: a 256-DOF
// chain, 514 states) is integrated by a GROUP of G threads (G = 8 .. 1024 lanes; above 128 the group is a whole CTA): every lane owns R = ceil(N / G)-ish
// state variables in registers and evaluates their rows of the derivative; the few operands owned by other lanes
// ("halo") are exchanged once per RK4 stage.  For the kernels this is just another pattern class: the big component of
// n_inst instances becomes n_inst * G VIRTUAL instances (virtual instance = instance * G + lane) of ONE lane-local
// pattern with R rows whose operand indices run over the extended vector [own slots | halo entries]; lanes that lack a
// term of the common pattern carry a zero coefficient for it (fma(0, x, y) == y: results are unchanged).  Checkpoints,
// chunk trajectories, the adjoint schedule and the user scatter are the register path's own, on n_traj * G threads.
//
// What is decided here:
//   1. UNITS: heavy-edge matching on the coupling graph pairs the variables that exchange the most terms (x_d with v_d
//      for a second-order oscillator written in first-order form), so that they land on the same lane;
//   2. ORDER: units are laid out along a line, either in natural (index) order or in Cuthill-McKee order of the unit
//      graph, whichever gives the smaller common pattern; lane l owns units [l U, (l + 1) U);
//   3. ADJOINT ROWS follow the forward unit they reference most (set-point operands first, then adjoint couplings);
//   4. the COMMON lane pattern: the ordered union over the lanes of their local term lists (kept consistent with the
//      hand-off order inside a row), halo tables, per-lane coefficient / initial-condition / variable-map tables;
//   5. a cost model (registers per thread, share of real terms) that accepts the split or leaves the component to the
//      shared-memory tile kernels / the dense-operator path.
// The result is verified on the host against the original pattern (one random derivative evaluation per side) before
// it is handed to the kernels.
//
// Replaces, for such components, the reference's N x E global vectors and ~17 launches per step
// (reference src/__GpuDriver.cu:693-713, 785-838; src/kernels.cu:21-63).
#pragma once

#include <algorithm>
#include <cstdint>
#include <map>
#include <set>

#include "pattern.hpp"

namespace pasif_host {

namespace lane_detail {

// (lane, slot) of every variable of one side (forward state or adjoint state) of the component
struct Assign {
  std::vector<int> lane, slot;
  int R = 0;
};

inline int enc(int delta, int idx) { return delta * 65536 + idx; }

struct TermRef {
  int lane = 0;
  std::vector<int> sig;   // kind-specific: [out slot, encoded operand refs..., occurrence]
  int orig = 0;           // index inside its kind in the source pattern
};

// ordered union of the lanes' term sequences: a common supersequence built greedily, so that every lane's own order
// (the hand-off order inside a row) is kept wherever the lanes agree
inline std::vector<std::vector<int>> merge_sequences(const std::vector<std::vector<std::vector<int>>>& seqs) {
  std::vector<std::vector<int>> uni;
  for (const auto& seq : seqs) {
    size_t cursor = 0;
    for (const auto& s : seq) {
      size_t found = uni.size();
      for (size_t i = 0; i < uni.size(); ++i)
        if (uni[i] == s) { found = i; break; }
      if (found == uni.size()) { uni.insert(uni.begin() + (long)cursor, s); ++cursor; }
      else if (found >= cursor) cursor = found + 1;
    }
  }
  return uni;
}

struct SideResult {
  Pattern pat;
  std::vector<int> term_union;   // per source coefficient index (K, G, L, P, F concatenated): row of the new table
  std::vector<int> term_lane;    // per source coefficient index: owning lane
  long long real_terms = 0;
};

// Builds the lane-local pattern of one side.
//   src      source pattern (adjoint: tensor p-indices refer to src.nv)
//   xa       assignment of the differentiated variables (rows / x operands)
//   fa       adjoint only: assignment of the forward variables (set-point operands); nullptr for the forward side
inline SideResult build_side(const Pattern& src, const Assign& xa, const Assign* fa, int G) {
  SideResult r;
  const bool adjoint = fa != nullptr;
  // set-point list of the lane pattern: forward slots referenced by any lane
  std::map<int, int> nvidx;   // forward slot -> position in the new nv list
  if (adjoint) {
    std::set<int> s;
    for (int f : src.nv) s.insert(fa->slot[f]);
    int k = 0;
    for (int v : s) { nvidx[v] = k++; r.pat.nv.push_back(v); }
  }
  auto xref = [&](int out_lane, int var) { return enc(((xa.lane[var] - out_lane) % G + G) % G, xa.slot[var]); };
  auto pref = [&](int out_lane, int p) {
    if (!adjoint) return xref(out_lane, p);
    const int f = src.nv[p];
    return enc(((fa->lane[f] - out_lane) % G + G) % G, nvidx[fa->slot[f]]);
  };
  const int nk = (int)src.k_row.size(), ng = (int)src.g_o.size(), nl = (int)src.l_o.size(), np = (int)src.p_o.size(),
            nf = (int)src.f_row.size();
  std::vector<TermRef> terms[5];
  for (int k = 0; k < nk; ++k) {
    TermRef t; t.lane = xa.lane[src.k_row[k]]; t.orig = k;
    t.sig = {xa.slot[src.k_row[k]], xref(t.lane, src.k_col[k])};
    terms[0].push_back(t);
  }
  for (int k = 0; k < ng; ++k) {
    TermRef t; t.lane = xa.lane[src.g_o[k]]; t.orig = k;
    t.sig = {xa.slot[src.g_o[k]], xref(t.lane, src.g_a[k]), pref(t.lane, src.g_b[k])};
    terms[1].push_back(t);
  }
  for (int k = 0; k < nl; ++k) {
    TermRef t; t.lane = xa.lane[src.l_o[k]]; t.orig = k;
    t.sig = {xa.slot[src.l_o[k]], xref(t.lane, src.l_a[k]), pref(t.lane, src.l_b[k]), pref(t.lane, src.l_c[k])};
    terms[2].push_back(t);
  }
  for (int k = 0; k < np; ++k) {
    TermRef t; t.lane = xa.lane[src.p_o[k]]; t.orig = k;
    t.sig = {xa.slot[src.p_o[k]], xref(t.lane, src.p_a[k]), pref(t.lane, src.p_b[k]), pref(t.lane, src.p_c[k]),
             pref(t.lane, src.p_d[k])};
    terms[3].push_back(t);
  }
  for (int k = 0; k < nf; ++k) {
    TermRef t; t.lane = xa.lane[src.f_row[k]]; t.orig = k;
    t.sig = {xa.slot[src.f_row[k]]};
    terms[4].push_back(t);
  }
  // per kind: per-lane sequences sorted by (out slot, hand-off order); duplicates get an occurrence counter
  std::vector<std::vector<int>> uni[5];
  for (int kind = 0; kind < 5; ++kind) {
    std::vector<std::vector<TermRef*>> by_lane((size_t)G);
    for (auto& t : terms[kind]) by_lane[(size_t)t.lane].push_back(&t);
    std::vector<std::vector<std::vector<int>>> seqs((size_t)G);
    for (int l = 0; l < G; ++l) {
      auto& v = by_lane[(size_t)l];
      std::stable_sort(v.begin(), v.end(), [](const TermRef* a, const TermRef* b) { return a->sig[0] < b->sig[0]; });
      std::map<std::vector<int>, int> occ;
      for (TermRef* t : v) {
        const int o = occ[t->sig]++;
        t->sig.push_back(o);
        seqs[(size_t)l].push_back(t->sig);
      }
    }
    uni[kind] = merge_sequences(seqs);
    // rows stay grouped: stable sort of the union by out slot keeps the merged order inside a row
    std::stable_sort(uni[kind].begin(), uni[kind].end(), [](const std::vector<int>& a, const std::vector<int>& b) { return a[0] < b[0]; });
  }
  // halo registries in order of first use
  std::map<int, int> hx, hp;
  auto xext = [&](int code) {
    const int d = code / 65536, s = code % 65536;
    if (d == 0) return s;
    auto it = hx.find(code);
    if (it == hx.end()) { it = hx.emplace(code, (int)hx.size()).first; r.pat.hx_slot.push_back(s); r.pat.hx_delta.push_back(d); }
    return xa.R + it->second;
  };
  auto pext = [&](int code) {
    if (!adjoint) return xext(code);
    const int d = code / 65536, s = code % 65536;
    if (d == 0) return s;
    auto it = hp.find(code);
    if (it == hp.end()) { it = hp.emplace(code, (int)hp.size()).first; r.pat.hp_slot.push_back(s); r.pat.hp_delta.push_back(d); }
    return (int)r.pat.nv.size() + it->second;
  };
  Pattern& p = r.pat;
  p.N = xa.R;
  p.lanes = G;
  for (auto& s : uni[0]) { p.k_row.push_back(s[0]); p.k_col.push_back(xext(s[1])); }
  for (auto& s : uni[1]) { p.g_o.push_back(s[0]); p.g_a.push_back(xext(s[1])); p.g_b.push_back(pext(s[2])); }
  for (auto& s : uni[2]) { p.l_o.push_back(s[0]); p.l_a.push_back(xext(s[1])); p.l_b.push_back(pext(s[2])); p.l_c.push_back(pext(s[3])); }
  for (auto& s : uni[3]) {
    p.p_o.push_back(s[0]); p.p_a.push_back(xext(s[1])); p.p_b.push_back(pext(s[2])); p.p_c.push_back(pext(s[3]));
    p.p_d.push_back(pext(s[4]));
  }
  for (auto& s : uni[4]) p.f_row.push_back(s[0]);
  // published slots of the halo exchange
  {
    std::map<int, int> pub;
    for (int s : p.hx_slot) pub.emplace(s, 0);
    int k = 0;
    for (auto& kv : pub) { kv.second = k++; p.pub_slot.push_back(kv.first); }
    for (int s : p.hx_slot) p.hx_pub.push_back(pub[s]);
    std::map<int, int> ppub;
    for (int s : p.hp_slot) ppub.emplace(s, 0);
    k = 0;
    for (auto& kv : ppub) { kv.second = k++; p.ppub_slot.push_back(kv.first); }
    for (int s : p.hp_slot) p.hp_pub.push_back(ppub[s]);
  }
  // source coefficient index -> (row of the new table, lane)
  const int off[5] = {0, nk, nk + ng, nk + ng + nl, nk + ng + nl + np};
  int uoff[5];
  uoff[0] = 0;
  for (int kind = 1; kind < 5; ++kind) uoff[kind] = uoff[kind - 1] + (int)uni[kind - 1].size();
  r.term_union.assign((size_t)src.nc(), -1);
  r.term_lane.assign((size_t)src.nc(), 0);
  for (int kind = 0; kind < 5; ++kind) {
    std::map<std::vector<int>, int> pos;
    for (size_t i = 0; i < uni[kind].size(); ++i) pos[uni[kind][i]] = (int)i;
    for (auto& t : terms[kind]) {
      r.term_union[(size_t)(off[kind] + t.orig)] = uoff[kind] + pos.at(t.sig);
      r.term_lane[(size_t)(off[kind] + t.orig)] = t.lane;
    }
  }
  r.real_terms = src.nc();
  return r;
}

// derivative of the source pattern (instance 0) at x (and set point p) -- reference evaluation for the self-check
inline void eval_source(const Pattern& s, const std::vector<double>& coef, int n_inst, const std::vector<double>& x,
                        const std::vector<double>& p, double fe, std::vector<double>& y) {
  y.assign((size_t)s.N, 0.0);
  size_t c = 0;
  auto cv = [&](size_t k) { return coef[k * (size_t)n_inst]; };
  for (size_t k = 0; k < s.k_row.size(); ++k, ++c) y[s.k_row[k]] = std::fma(cv(c), x[s.k_col[k]], y[s.k_row[k]]);
  for (size_t k = 0; k < s.g_o.size(); ++k, ++c) y[s.g_o[k]] = std::fma(cv(c) * x[s.g_a[k]], p[s.g_b[k]], y[s.g_o[k]]);
  for (size_t k = 0; k < s.l_o.size(); ++k, ++c) y[s.l_o[k]] = std::fma(cv(c) * x[s.l_a[k]] * p[s.l_b[k]], p[s.l_c[k]], y[s.l_o[k]]);
  for (size_t k = 0; k < s.p_o.size(); ++k, ++c)
    y[s.p_o[k]] = std::fma(cv(c) * x[s.p_a[k]] * p[s.p_b[k]] * p[s.p_c[k]], p[s.p_d[k]], y[s.p_o[k]]);
  for (size_t k = 0; k < s.f_row.size(); ++k, ++c) y[s.f_row[k]] = std::fma(cv(c), fe, y[s.f_row[k]]);
}

// the same derivative through the lane pattern: xl[lane][slot], pl[lane][nv entry] (forward: pl == xl)
inline void eval_lanes(const Pattern& q, const std::vector<double>& coef, int n_vinst, int G, const std::vector<std::vector<double>>& xl,
                       const std::vector<std::vector<double>>& pl, bool adjoint, double fe, std::vector<std::vector<double>>& yl) {
  yl.assign((size_t)G, std::vector<double>((size_t)q.N, 0.0));
  for (int l = 0; l < G; ++l) {
    std::vector<double> xe = xl[(size_t)l];
    for (size_t h = 0; h < q.hx_slot.size(); ++h) xe.push_back(xl[(size_t)((l + q.hx_delta[h]) % G)][(size_t)q.pub_slot[(size_t)q.hx_pub[h]]]);
    std::vector<double> pe;
    if (adjoint) {
      pe = pl[(size_t)l];
      for (size_t h = 0; h < q.hp_slot.size(); ++h) pe.push_back(pl[(size_t)((l + q.hp_delta[h]) % G)][(size_t)q.ppub_slot[(size_t)q.hp_pub[h]]]);
    } else pe = xe;
    std::vector<double>& y = yl[(size_t)l];
    size_t c = 0;
    auto cv = [&](size_t k) { return coef[k * (size_t)n_vinst + (size_t)l]; };   // instance 0: virtual instance = lane
    for (size_t k = 0; k < q.k_row.size(); ++k, ++c) y[q.k_row[k]] = std::fma(cv(c), xe[q.k_col[k]], y[q.k_row[k]]);
    for (size_t k = 0; k < q.g_o.size(); ++k, ++c) y[q.g_o[k]] = std::fma(cv(c) * xe[q.g_a[k]], pe[q.g_b[k]], y[q.g_o[k]]);
    for (size_t k = 0; k < q.l_o.size(); ++k, ++c) y[q.l_o[k]] = std::fma(cv(c) * xe[q.l_a[k]] * pe[q.l_b[k]], pe[q.l_c[k]], y[q.l_o[k]]);
    for (size_t k = 0; k < q.p_o.size(); ++k, ++c)
      y[q.p_o[k]] = std::fma(cv(c) * xe[q.p_a[k]] * pe[q.p_b[k]] * pe[q.p_c[k]], pe[q.p_d[k]], y[q.p_o[k]]);
    for (size_t k = 0; k < q.f_row.size(); ++k, ++c) y[q.f_row[k]] = std::fma(cv(c), fe, y[q.f_row[k]]);
  }
}

inline double pseudo(uint64_t& s) {   // deterministic values in (-1, 1) for the self-check
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return ((double)((s >> 11) & 0xFFFFFFFFull) / 4294967296.0) * 2.0 - 1.0;
}

// ---- coupling graph helpers
typedef std::vector<std::map<int, int>> Graph;   // adjacency with term-count weights

inline void add_edge(Graph& g, int a, int b) {
  if (a == b) return;
  ++g[(size_t)a][b];
  ++g[(size_t)b][a];
}

inline Graph forward_graph(const Pattern& f) {
  Graph g((size_t)f.N);
  for (size_t k = 0; k < f.k_row.size(); ++k) add_edge(g, f.k_row[k], f.k_col[k]);
  for (size_t k = 0; k < f.g_o.size(); ++k) { add_edge(g, f.g_o[k], f.g_a[k]); add_edge(g, f.g_o[k], f.g_b[k]); }
  for (size_t k = 0; k < f.l_o.size(); ++k) { add_edge(g, f.l_o[k], f.l_a[k]); add_edge(g, f.l_o[k], f.l_b[k]); add_edge(g, f.l_o[k], f.l_c[k]); }
  for (size_t k = 0; k < f.p_o.size(); ++k) {
    add_edge(g, f.p_o[k], f.p_a[k]); add_edge(g, f.p_o[k], f.p_b[k]); add_edge(g, f.p_o[k], f.p_c[k]); add_edge(g, f.p_o[k], f.p_d[k]);
  }
  return g;
}

// heavy-edge matching in index order: unit_of[v], units[u] = members ascending
inline void match_units(const Graph& g, std::vector<int>& unit_of, std::vector<std::vector<int>>& units) {
  const int n = (int)g.size();
  unit_of.assign((size_t)n, -1);
  units.clear();
  for (int v = 0; v < n; ++v) {
    if (unit_of[(size_t)v] >= 0) continue;
    int best = -1, bw = 0;
    for (auto& kv : g[(size_t)v])
      if (unit_of[(size_t)kv.first] < 0 && kv.first != v && kv.second > bw) { best = kv.first; bw = kv.second; }
    const int u = (int)units.size();
    units.emplace_back();
    units[(size_t)u].push_back(v);
    unit_of[(size_t)v] = u;
    if (best >= 0) { units[(size_t)u].push_back(best); unit_of[(size_t)best] = u; }
    std::sort(units[(size_t)u].begin(), units[(size_t)u].end());
  }
}

// Observer units -- units no row outside them reads (probe variables such as p1' = x_0^2, p2' = p1) -- do not belong on the
// line of units: each is absorbed as extra members into the unit it reads most from, so that it costs a few extra slots
// instead of a whole position on every lane.
inline void absorb_observers(const Pattern& f, const Graph& g, std::vector<int>& unit_of, std::vector<std::vector<int>>& units) {
  for (bool changed = true; changed;) {
    changed = false;
    const int nu = (int)units.size();
    std::vector<char> feeds((size_t)nu, 0);
    auto use = [&](int out, int operand) { if (unit_of[(size_t)out] != unit_of[(size_t)operand]) feeds[(size_t)unit_of[(size_t)operand]] = 1; };
    for (size_t k = 0; k < f.k_row.size(); ++k) use(f.k_row[k], f.k_col[k]);
    for (size_t k = 0; k < f.g_o.size(); ++k) { use(f.g_o[k], f.g_a[k]); use(f.g_o[k], f.g_b[k]); }
    for (size_t k = 0; k < f.l_o.size(); ++k) { use(f.l_o[k], f.l_a[k]); use(f.l_o[k], f.l_b[k]); use(f.l_o[k], f.l_c[k]); }
    for (size_t k = 0; k < f.p_o.size(); ++k) { use(f.p_o[k], f.p_a[k]); use(f.p_o[k], f.p_b[k]); use(f.p_o[k], f.p_c[k]); use(f.p_o[k], f.p_d[k]); }
    for (int u = 0; u < nu && !changed; ++u) {
      if (feeds[(size_t)u] || units[(size_t)u].empty()) continue;
      std::map<int, int> w;
      for (int v : units[(size_t)u])
        for (auto& kv : g[(size_t)v])
          if (unit_of[(size_t)kv.first] != u) w[unit_of[(size_t)kv.first]] += kv.second;
      int best = -1, bw = 0;
      for (auto& kv : w) if (kv.second > bw) { best = kv.first; bw = kv.second; }
      if (best < 0) continue;
      for (int v : units[(size_t)u]) { units[(size_t)best].push_back(v); unit_of[(size_t)v] = best; }
      units[(size_t)u].clear();
      changed = true;
    }
    if (changed) {   // compact
      std::vector<std::vector<int>> nu2;
      for (auto& m : units) if (!m.empty()) nu2.push_back(m);
      units.swap(nu2);
      for (size_t u = 0; u < units.size(); ++u) for (int v : units[u]) unit_of[(size_t)v] = (int)u;
    }
  }
}

// Cuthill-McKee order of the unit graph from a pseudo-peripheral unit
inline std::vector<int> cm_order(const Graph& g, const std::vector<int>& unit_of, int n_units) {
  std::vector<std::set<int>> ug((size_t)n_units);
  for (size_t v = 0; v < g.size(); ++v)
    for (auto& kv : g[v]) {
      const int a = unit_of[v], b = unit_of[(size_t)kv.first];
      if (a != b) { ug[(size_t)a].insert(b); ug[(size_t)b].insert(a); }
    }
  auto bfs = [&](int start, std::vector<int>& order, std::vector<int>& seen) {
    size_t head = order.size();
    order.push_back(start);
    seen[(size_t)start] = 1;
    int last = start;
    while (head < order.size()) {
      const int u = order[head++];
      last = u;
      std::vector<int> nb;
      for (int w : ug[(size_t)u]) if (!seen[(size_t)w]) nb.push_back(w);
      std::sort(nb.begin(), nb.end(), [&](int a, int b) {
        if (ug[(size_t)a].size() != ug[(size_t)b].size()) return ug[(size_t)a].size() < ug[(size_t)b].size();
        return a < b;
      });
      for (int w : nb) { seen[(size_t)w] = 1; order.push_back(w); }
    }
    return last;
  };
  std::vector<int> order, seen((size_t)n_units, 0);
  for (int s = 0; s < n_units; ++s) {
    if (seen[(size_t)s]) continue;
    // two sweeps to reach a pseudo-peripheral start inside this connected part
    std::vector<int> tmp, tseen = seen;
    int far = bfs(s, tmp, tseen);
    tmp.clear(); tseen = seen;
    far = bfs(far, tmp, tseen);
    bfs(far, order, seen);
  }
  return order;
}

}  // namespace lane_detail

struct LaneSplitReport {
  bool ok = false;
  int G = 0, R = 0, RA = 0, nv = 0, nc_fwd = 0, nc_adj = 0, halo_fwd = 0, halo_adj = 0, halo_p = 0, regs_est = 0;
  double efficiency = 0;
  std::string why;
};

// One attempt with a given group width and unit order.  Returns false when the split is not viable.
// pair: the class will run on the fp32 pair kernels (rk4_kernels.cuh, W = 2) -- every state / stage / halo slot is a 64-bit
// register pair, the coefficients stay scalars
inline bool lane_split_try(const ClassData& c, int G, bool cm, int tbytes, int reg_budget, double min_eff, bool pair, ClassData& out,
                           LaneSplitReport& rep) {
  using namespace lane_detail;
  const Pattern& F = c.fwd;
  const int n_inst = c.n_inst;
  Graph g = forward_graph(F);
  std::vector<int> unit_of;
  std::vector<std::vector<int>> units;
  match_units(g, unit_of, units);
  absorb_observers(F, g, unit_of, units);
  const int n_units = (int)units.size();
  std::vector<int> order;
  if (cm) order = cm_order(g, unit_of, n_units);
  else { order.resize((size_t)n_units); std::iota(order.begin(), order.end(), 0); }   // units are created in index order
  const int U = (n_units + G - 1) / G;
  std::vector<int> unit_lane((size_t)n_units), unit_pos((size_t)n_units);
  for (int i = 0; i < n_units; ++i) { unit_lane[(size_t)order[(size_t)i]] = i / U; unit_pos[(size_t)order[(size_t)i]] = i % U; }

  // forward slots: key (position, member) -> slot, union over the lanes
  Assign fa;
  fa.lane.assign((size_t)F.N, 0); fa.slot.assign((size_t)F.N, 0);
  {
    std::set<std::pair<int, int>> keys;
    for (int u = 0; u < n_units; ++u)
      for (size_t m = 0; m < units[(size_t)u].size(); ++m) keys.insert({unit_pos[(size_t)u], (int)m});
    std::map<std::pair<int, int>, int> slot;
    for (auto& k : keys) { const int s = (int)slot.size(); slot[k] = s; }
    fa.R = (int)slot.size();
    for (int u = 0; u < n_units; ++u)
      for (size_t m = 0; m < units[(size_t)u].size(); ++m) {
        const int v = units[(size_t)u][m];
        fa.lane[(size_t)v] = unit_lane[(size_t)u];
        fa.slot[(size_t)v] = slot[{unit_pos[(size_t)u], (int)m}];
      }
  }
  SideResult sf = build_side(F, fa, nullptr, G);

  Assign aa;
  SideResult sa;
  if (c.joint) {
    const Pattern& A = c.adj;
    // adjoint row -> forward unit.  Phase 1: set-point operands; phase 2 (waves): couplings to assigned adjoint rows.
    std::vector<int> aunit((size_t)A.N, -1);
    std::vector<std::map<int, int>> aff((size_t)A.N);
    auto fu = [&](int pidx) { return unit_of[(size_t)A.nv[(size_t)pidx]]; };
    for (size_t k = 0; k < A.g_o.size(); ++k) ++aff[(size_t)A.g_o[k]][fu(A.g_b[k])];
    for (size_t k = 0; k < A.l_o.size(); ++k) { ++aff[(size_t)A.l_o[k]][fu(A.l_b[k])]; ++aff[(size_t)A.l_o[k]][fu(A.l_c[k])]; }
    for (size_t k = 0; k < A.p_o.size(); ++k) {
      ++aff[(size_t)A.p_o[k]][fu(A.p_b[k])]; ++aff[(size_t)A.p_o[k]][fu(A.p_c[k])]; ++aff[(size_t)A.p_o[k]][fu(A.p_d[k])];
    }
    // forward variable with the same GLOBAL index as the adjoint row (lambda_i <-> q_i in the usual layout): tie-break
    std::map<int, int> f_by_global;
    for (int i = 0; i < F.N; ++i) f_by_global[c.varmap[(size_t)i * n_inst]] = i;
    auto pick = [&](int row, const std::map<int, int>& w) {
      int pref = -1, bw = -1;
      auto it = f_by_global.find(c.varmap_adj[(size_t)row * n_inst]);
      if (it != f_by_global.end()) pref = unit_of[(size_t)it->second];
      for (auto& kv : w) bw = std::max(bw, kv.second);
      if (pref >= 0 && w.count(pref) && w.at(pref) == bw) return pref;
      for (auto& kv : w)
        if (kv.second == bw) return kv.first;          // lowest unit index among the heaviest
      return -1;
    };
    for (int r2 = 0; r2 < A.N; ++r2)
      if (!aff[(size_t)r2].empty()) aunit[(size_t)r2] = pick(r2, aff[(size_t)r2]);
    Graph ag((size_t)A.N);
    for (size_t k = 0; k < A.k_row.size(); ++k) add_edge(ag, A.k_row[k], A.k_col[k]);
    for (size_t k = 0; k < A.g_o.size(); ++k) add_edge(ag, A.g_o[k], A.g_a[k]);
    for (size_t k = 0; k < A.l_o.size(); ++k) add_edge(ag, A.l_o[k], A.l_a[k]);
    for (size_t k = 0; k < A.p_o.size(); ++k) add_edge(ag, A.p_o[k], A.p_a[k]);
    for (bool changed = true; changed;) {
      changed = false;
      std::vector<int> next = aunit;
      for (int r2 = 0; r2 < A.N; ++r2) {
        if (aunit[(size_t)r2] >= 0) continue;
        std::map<int, int> w;
        for (auto& kv : ag[(size_t)r2])
          if (aunit[(size_t)kv.first] >= 0) w[aunit[(size_t)kv.first]] += kv.second;
        if (!w.empty()) { next[(size_t)r2] = pick(r2, w); changed = true; }
      }
      aunit.swap(next);
    }
    for (int r2 = 0; r2 < A.N; ++r2)
      if (aunit[(size_t)r2] < 0) aunit[(size_t)r2] = order[0];
    // adjoint slots: key (position, member among the unit's adjoint rows in index order)
    std::vector<std::vector<int>> rows_of((size_t)n_units);
    for (int r2 = 0; r2 < A.N; ++r2) rows_of[(size_t)aunit[(size_t)r2]].push_back(r2);
    std::set<std::pair<int, int>> keys;
    for (int u = 0; u < n_units; ++u)
      for (size_t m = 0; m < rows_of[(size_t)u].size(); ++m) keys.insert({unit_pos[(size_t)u], (int)m});
    std::map<std::pair<int, int>, int> slot;
    for (auto& k : keys) { const int s = (int)slot.size(); slot[k] = s; }
    aa.R = std::max<int>((int)slot.size(), 1);
    aa.lane.assign((size_t)A.N, 0); aa.slot.assign((size_t)A.N, 0);
    for (int u = 0; u < n_units; ++u)
      for (size_t m = 0; m < rows_of[(size_t)u].size(); ++m) {
        const int v = rows_of[(size_t)u][m];
        aa.lane[(size_t)v] = unit_lane[(size_t)u];
        aa.slot[(size_t)v] = slot[{unit_pos[(size_t)u], (int)m}];
      }
    sa = build_side(A, aa, &fa, G);
  }

  // ---- cost model
  rep.G = G; rep.R = fa.R; rep.RA = c.joint ? aa.R : 0;
  rep.nc_fwd = sf.pat.nc(); rep.nc_adj = c.joint ? sa.pat.nc() : 0;
  rep.halo_fwd = (int)sf.pat.hx_slot.size();
  rep.halo_adj = c.joint ? (int)sa.pat.hx_slot.size() : 0;
  rep.halo_p = c.joint ? (int)sa.pat.hp_slot.size() : 0;
  rep.nv = c.joint ? (int)sa.pat.nv.size() : 0;
  const int words = tbytes / 4, sw = pair ? 2 * words : words;      // registers per coefficient / per state slot
  const int regs_f = rep.nc_fwd * words + (4 * fa.R + rep.halo_fwd) * sw + 24;
  const int regs_a = c.joint ? rep.nc_adj * words + (4 * aa.R + rep.halo_adj + 3 * ((int)sa.pat.nv.size() + rep.halo_p)) * sw + 24 : 0;
  rep.regs_est = std::max(regs_f, regs_a);
  const double real = (double)F.nc() + (c.joint ? (double)c.adj.nc() : 0.0);
  rep.efficiency = real / ((double)G * (double)(rep.nc_fwd + rep.nc_adj));
  // groups wider than the default CTA (128 threads) get a CTA of their own width: the register file then caps a thread at
  // 65536 / G registers
  if (G > 128) reg_budget = std::min(reg_budget, 65536 / G - 8);
  if (rep.regs_est > reg_budget) { rep.why = "register estimate " + std::to_string(rep.regs_est) + " > " + std::to_string(reg_budget); return false; }
  if (rep.efficiency < min_eff) { rep.why = "only " + std::to_string((int)(rep.efficiency * 100)) + " % of the common pattern's terms are real"; return false; }
  if (fa.R > 48 || (c.joint && aa.R > 48)) { rep.why = "more than 48 slots per lane"; return false; }
  {
    // static shared memory of the halo exchange: 2 buffers x published slots x CTA width, per exchange kind
    const size_t cta = (size_t)std::max(G, 128);
    const size_t ebytes = (size_t)tbytes * (pair ? 2 : 1);            // a published slot is a pair in the pair kernels
    const size_t smem_f = 2 * sf.pat.pub_slot.size() * cta * ebytes;
    const size_t smem_a = c.joint ? 2 * (sa.pat.pub_slot.size() + sa.pat.ppub_slot.size()) * cta * ebytes : 0;
    if (std::max(smem_f, smem_a) > 40 * 1024) { rep.why = "halo exchange would need more than 40 KB of shared memory"; return false; }
  }
  // a lane that reads more operands from other lanes than it owns is a dense coupling: the exchange would dominate, and
  // the dense-operator path (tensor-core GEMM per stage) is the one built for it
  if (min_eff > 0 && (rep.halo_fwd > std::max(24, 2 * fa.R) || (c.joint && rep.halo_adj + rep.halo_p > std::max(24, 2 * aa.R)))) {
    rep.why = "too many halo operands per lane (dense coupling)";
    return false;
  }

  // ---- tables of the virtual instances
  out = ClassData();
  out.joint = c.joint;
  out.lanes = G;
  out.fwd = sf.pat;
  out.n_inst = n_inst * G;
  const int nvi = out.n_inst;
  auto fill_coef = [&](const SideResult& s, const std::vector<double>& src, std::vector<double>& dst) {
    dst.assign((size_t)s.pat.nc() * nvi, 0.0);
    for (size_t t = 0; t < s.term_union.size(); ++t)
      for (int j = 0; j < n_inst; ++j)
        dst[(size_t)s.term_union[t] * nvi + (size_t)j * G + s.term_lane[t]] = src[t * (size_t)n_inst + j];
  };
  auto fill_vars = [&](const Assign& a, int nvar, const std::vector<double>& ic, const std::vector<int>& vm, std::vector<double>& ic2,
                       std::vector<int>& vm2) {
    ic2.assign((size_t)a.R * nvi, 0.0);
    vm2.assign((size_t)a.R * nvi, -1);
    for (int v = 0; v < nvar; ++v)
      for (int j = 0; j < n_inst; ++j) {
        const size_t o = (size_t)a.slot[(size_t)v] * nvi + (size_t)j * G + a.lane[(size_t)v];
        ic2[o] = ic[(size_t)v * n_inst + j];
        vm2[o] = vm[(size_t)v * n_inst + j];
      }
  };
  fill_coef(sf, c.coef, out.coef);
  fill_vars(fa, F.N, c.ic, c.varmap, out.ic, out.varmap);
  if (c.joint) {
    out.adj = sa.pat;
    fill_coef(sa, c.coef_adj, out.coef_adj);
    fill_vars(aa, c.adj.N, c.ic_adj, c.varmap_adj, out.ic_adj, out.varmap_adj);
  }
  const std::string sig = std::string(c.joint ? "JL|" : "FL|") + out.fwd.signature() + (c.joint ? "|A|" + out.adj.signature() : std::string());
  out.hash = fnv1a(sig);
  out.name = std::string(c.joint ? "J_" : "F_") + detail::hex64(out.hash);

  // ---- self-check against the source pattern (instance 0): one derivative evaluation per side
  {
    uint64_t seed = 0x9E3779B97F4A7C15ull ^ out.hash;
    std::vector<double> x((size_t)F.N), y;
    for (double& v : x) v = pseudo(seed);
    const double fe = pseudo(seed);
    eval_source(F, c.coef, n_inst, x, x, fe, y);
    std::vector<std::vector<double>> xl((size_t)G, std::vector<double>((size_t)fa.R, 0.0)), yl;
    for (int v = 0; v < F.N; ++v) xl[(size_t)fa.lane[(size_t)v]][(size_t)fa.slot[(size_t)v]] = x[(size_t)v];
    eval_lanes(out.fwd, out.coef, nvi, G, xl, xl, false, fe, yl);
    double scale = 1e-300, err = 0;
    for (int v = 0; v < F.N; ++v) {
      scale = std::max(scale, std::fabs(y[(size_t)v]));
      err = std::max(err, std::fabs(y[(size_t)v] - yl[(size_t)fa.lane[(size_t)v]][(size_t)fa.slot[(size_t)v]]));
    }
    if (!(err <= 1e-12 * scale)) throw std::runtime_error("internal: lane split of the forward pattern failed its self-check");
    if (c.joint) {
      const Pattern& A = c.adj;
      std::vector<double> lam((size_t)A.N), p(A.nv.size());
      for (double& v : lam) v = pseudo(seed);
      for (size_t i = 0; i < p.size(); ++i) p[i] = x[(size_t)A.nv[i]];
      eval_source(A, c.coef_adj, n_inst, lam, p, 0.0, y);
      std::vector<std::vector<double>> ll((size_t)G, std::vector<double>((size_t)aa.R, 0.0)),
          pl((size_t)G, std::vector<double>(out.adj.nv.size(), 0.0));
      for (int v = 0; v < A.N; ++v) ll[(size_t)aa.lane[(size_t)v]][(size_t)aa.slot[(size_t)v]] = lam[(size_t)v];
      for (int l = 0; l < G; ++l)
        for (size_t i = 0; i < out.adj.nv.size(); ++i) pl[(size_t)l][i] = xl[(size_t)l][(size_t)out.adj.nv[i]];
      eval_lanes(out.adj, out.coef_adj, nvi, G, ll, pl, true, 0.0, yl);
      scale = 1e-300; err = 0;
      for (int v = 0; v < A.N; ++v) {
        scale = std::max(scale, std::fabs(y[(size_t)v]));
        err = std::max(err, std::fabs(y[(size_t)v] - yl[(size_t)aa.lane[(size_t)v]][(size_t)aa.slot[(size_t)v]]));
      }
      if (!(err <= 1e-12 * scale)) throw std::runtime_error("internal: lane split of the adjoint pattern failed its self-check");
    }
  }
  rep.ok = true;
  return true;
}

// Tries group widths from narrow to wide and both unit orders; keeps the first width that fits the register budget,
// with the order that gives the smaller common pattern.
inline bool lane_split(const ClassData& c, int tbytes, int max_lanes, int reg_budget, double min_eff, ClassData& out, LaneSplitReport& rep,
                       bool pair = false, int min_lanes = 8) {
  const long long terms = (long long)c.fwd.nc() + c.adj.nc();
  const long long rows = (long long)c.fwd.N + c.adj.N;
  if (terms > 64 * rows) { rep.why = "operator too dense for a lane split (" + std::to_string(terms / std::max<long long>(rows, 1)) + " terms per row)"; return false; }
  for (int G = std::max(2, min_lanes); G <= max_lanes; G *= 2) {
    ClassData best;
    LaneSplitReport brep;
    bool have = false;
    for (int cm = 0; cm < 2; ++cm) {
      ClassData cand;
      LaneSplitReport r;
      if (!lane_split_try(c, G, cm != 0, tbytes, reg_budget, min_eff, pair, cand, r)) { if (!have) brep = r; continue; }
      if (!have || r.nc_fwd + r.nc_adj < brep.nc_fwd + brep.nc_adj) { best = std::move(cand); brep = r; have = true; }
    }
    if (have) { out = std::move(best); out.lane_pair = pair; rep = brep; return true; }
    rep = brep;
  }
  return false;
}

}  // namespace pasif_host
