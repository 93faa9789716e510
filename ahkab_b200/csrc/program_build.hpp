// program_build.hpp — host-side topology compiler: ahk_circuit -> stamp program.
//
// No reference counterpart: ahkab re-walks the Python element list on every
// assembly (dc_analysis.py:949-1069, transient.py:472-561, ac.py:323-443,
// dc_analysis.py:843-881).  Here the walk happens once per topology and yields
//   * the merged, (row,col)-sorted list of matrix positions any stamp touches,
//   * per linear element the positions it adds to / sets (reference order and
//     assign-vs-accumulate semantics kept),
//   * per position / per row the CSR lists of device outputs to gather,
//   * locked node pairs, source tables, the OP-guess operator.
// Plain C++ (no CUDA) so that the host-emulation test harness can use it too.
#pragma once
#include <algorithm>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "program.h"

namespace ahk {

struct HostProg {
  Prog hdr;   // scalar fields valid; pointers filled by view()/upload
  std::vector<int> row_ptr, nl_ptr, tx_ptr, lock, gslot, slot_vary, pstage, gpslot;
  int nst = 0;       // doubles of per-device state
  std::vector<unsigned short> pcol, prowi, pcd;
  std::vector<unsigned char> pdiag;
  std::vector<int> term_ptr, term_slot;
  std::vector<signed char> term_kind;
  std::vector<short> nl_src, tx_src;
  std::vector<signed char> nl_sgn, tx_sgn;
  std::vector<LinOp> lin;
  std::vector<SrcOp> src;
  std::vector<DevOp> dev;
  std::vector<double> guessG, gscale, gconst, nominal;
  std::vector<int> ic_kind, ic_a, ic_b, ic_slot;   // unused on device (host x0 helpers)
  int n_nodes = 0, n_vde = 0;
  int ncon = 0;      // doubles of per-instance device constants
  bool nonlinear = false;

  // Prog whose pointers reference this object's host vectors
  Prog view() const {
    Prog p = hdr;
    p.row_ptr = row_ptr.data(); p.pcol = pcol.data(); p.pcd = pcd.data(); p.prowi = prowi.data(); p.pdiag = pdiag.data();
    p.term_ptr = term_ptr.data(); p.term_slot = term_slot.data(); p.term_kind = term_kind.data();
    p.nl_ptr = nl_ptr.data(); p.nl_src = nl_src.data(); p.nl_sgn = nl_sgn.data();
    p.tx_ptr = tx_ptr.data(); p.tx_src = tx_src.data(); p.tx_sgn = tx_sgn.data();
    p.lin = lin.data(); p.src = src.data(); p.dev = dev.data(); p.lock = lock.data();
    p.guessG = guessG.data(); p.gslot = gslot.data(); p.gscale = gscale.data();
    p.gconst = gconst.data(); p.nominal = nominal.data(); p.slot_vary = slot_vary.data();
    p.pstage = pstage.data(); p.gpslot = gpslot.data();
    return p;
  }
};

inline int dout_count(int type) {
  switch (type) {
    case AHK_D: return 2;
    case AHK_EKV: return 4;
    case AHK_MOSQ: return 9;
    case AHK_SW: return 3;
  }
  return 0;
}

inline void assign_constant_blocks(HostProg& H);

inline void build_program(const ahk_circuit* c, HostProg& H) {
  if (!c || c->abi_version != AHK_ABI_VERSION) throw std::runtime_error("ahk_circuit: bad abi_version");
  const int nn = c->n_nodes, nv1 = nn - 1, n = nv1 + c->n_vde;
  if (n <= 0) throw std::runtime_error("empty circuit");
  H.n_nodes = nn; H.n_vde = c->n_vde;
  // reduced index of an unreduced row/col u (u = node, or nn + vde); -1 = ground
  auto red = [&](int u) { return u - 1; };
  typedef std::pair<int, int> RC;
  std::map<RC, int> posmap;
  auto touch = [&](int ur, int uc) { if (ur > 0 && uc > 0) posmap[RC(red(ur), red(uc))] = 0; };
  // pass 1: collect every position any stamp can touch
  for (int i = 1; i < nn; i++) touch(i, i);   // Gmin / cmin diagonals
  for (int k = 0; k < c->n_elems; k++) {
    const ahk_elem& e = c->elems[k];
    const int idx = nn + e.vde;
    switch (e.type) {
      case AHK_R: case AHK_C:
        touch(e.n1, e.n1); touch(e.n1, e.n2); touch(e.n2, e.n1); touch(e.n2, e.n2); break;
      case AHK_G:
        touch(e.n1, e.n3); touch(e.n1, e.n4); touch(e.n2, e.n3); touch(e.n2, e.n4); break;
      case AHK_V: case AHK_L: case AHK_E: case AHK_H:
        touch(e.n1, idx); touch(e.n2, idx); touch(idx, e.n1); touch(idx, e.n2);
        if (e.type == AHK_E) { touch(idx, e.n3); touch(idx, e.n4); }
        if (e.type == AHK_H) touch(idx, nn + e.ref);
        if (e.type == AHK_L) touch(idx, idx);
        break;
      case AHK_F: touch(e.n1, nn + e.ref); touch(e.n2, nn + e.ref); break;
      case AHK_K: touch(nn + e.ref, nn + e.ref2); touch(nn + e.ref2, nn + e.ref); break;
      case AHK_D:
        touch(e.n1, e.n1); touch(e.n1, e.n2); touch(e.n2, e.n1); touch(e.n2, e.n2); break;
      case AHK_EKV: {   // out (d,s); drive ports (d,b),(g,b),(s,b)
        int o[2] = {e.n1, e.n3}, dp[4] = {e.n1, e.n2, e.n3, e.n4};
        for (int a = 0; a < 2; a++) for (int b = 0; b < 4; b++) touch(o[a], dp[b]);
      } break;
      case AHK_MOSQ: {  // rows d and s of the 4x4 stamp over (d,g,s,b)
        int o[2] = {e.n1, e.n3}, dp[4] = {e.n1, e.n2, e.n3, e.n4};
        for (int a = 0; a < 2; a++) for (int b = 0; b < 4; b++) touch(o[a], dp[b]);
      } break;
      case AHK_SW: {
        int o[2] = {e.n1, e.n2}, dp[4] = {e.n1, e.n2, e.n3, e.n4};
        for (int a = 0; a < 2; a++) for (int b = 0; b < 4; b++) touch(o[a], dp[b]);
      } break;
      case AHK_I: break;
      default: throw std::runtime_error("unknown element type");
    }
  }
  int np = 0;
  for (auto& kv : posmap) kv.second = np++;   // std::map iterates sorted by (row, col)
  auto P = [&](int ur, int uc) -> int {
    if (ur <= 0 || uc <= 0) return -1;
    return posmap.at(RC(red(ur), red(uc)));
  };
  H.row_ptr.assign(n + 1, 0);
  H.pcol.resize(np); H.prowi.resize(np); H.pdiag.resize(np);
  for (auto& kv : posmap) {
    int r = kv.first.first, cc = kv.first.second, k = kv.second;
    H.row_ptr[r + 1]++;
    H.pcol[k] = (unsigned short)cc;
    H.prowi[k] = (unsigned short)r;
    H.pdiag[k] = (r == cc && r < nv1) ? 1 : 0;
  }
  for (int r = 0; r < n; r++) H.row_ptr[r + 1] += H.row_ptr[r];
  if (n >= 32768) throw std::runtime_error("circuit too large (n >= 32768)");
  H.pcd.resize(np);
  for (int k = 0; k < np; k++) H.pcd[k] = (unsigned short)(H.pcol[k] | (H.pdiag[k] ? 0x8000 : 0));

  // pass 2: linear ops in the reference's loop order (dc_analysis.py:989-1062)
  auto mk = [&](int kind, const ahk_elem& e) { LinOp o; std::memset(&o, 0, sizeof o); o.kind = kind; o.pbase = e.pbase; o.aux0 = e.aux0; o.aux1 = e.aux1; for (int i = 0; i < 6; i++) o.pos[i] = -1; return o; };
  for (int k = 0; k < c->n_elems; k++) {
    const ahk_elem& e = c->elems[k];
    if (e.type == AHK_R) { LinOp o = mk(K_R, e); o.pos[0] = P(e.n1, e.n1); o.pos[1] = P(e.n1, e.n2); o.pos[2] = P(e.n2, e.n1); o.pos[3] = P(e.n2, e.n2); H.lin.push_back(o); }
    if (e.type == AHK_G) { LinOp o = mk(K_G, e); o.pos[0] = P(e.n1, e.n3); o.pos[1] = P(e.n1, e.n4); o.pos[2] = P(e.n2, e.n3); o.pos[3] = P(e.n2, e.n4); H.lin.push_back(o); }
  }
  for (int k = 0; k < c->n_elems; k++) {
    const ahk_elem& e = c->elems[k];
    if (e.vde < 0) continue;
    const int idx = nn + e.vde;
    LinOp o = mk(e.type == AHK_E ? K_E : (e.type == AHK_H ? K_H : K_VDE), e);
    o.pos[0] = P(e.n1, idx); o.pos[1] = P(e.n2, idx); o.pos[2] = P(idx, e.n1); o.pos[3] = P(idx, e.n2);
    if (e.type == AHK_E) { o.pos[4] = P(idx, e.n3); o.pos[5] = P(idx, e.n4); }
    if (e.type == AHK_H) o.pos[4] = P(idx, nn + e.ref);
    H.lin.push_back(o);
  }
  for (int k = 0; k < c->n_elems; k++) {
    const ahk_elem& e = c->elems[k];
    if (e.type == AHK_F) { LinOp o = mk(K_F, e); o.pos[0] = P(e.n1, nn + e.ref); o.pos[1] = P(e.n2, nn + e.ref); H.lin.push_back(o); }
  }
  // D ops (transient.py:527-550)
  for (int k = 0; k < c->n_elems; k++) {
    const ahk_elem& e = c->elems[k];
    if (e.type == AHK_C) { LinOp o = mk(K_C, e); o.pos[0] = P(e.n1, e.n1); o.pos[1] = P(e.n1, e.n2); o.pos[2] = P(e.n2, e.n2); o.pos[3] = P(e.n2, e.n1); H.lin.push_back(o); }
    if (e.type == AHK_L) { LinOp o = mk(K_L, e); o.pos[0] = P(nn + e.vde, nn + e.vde); H.lin.push_back(o); }
  }
  for (int k = 0; k < c->n_elems; k++) {
    const ahk_elem& e = c->elems[k];
    if (e.type == AHK_K) { LinOp o = mk(K_K, e); o.pos[0] = P(nn + e.ref, nn + e.ref2); o.pos[1] = P(nn + e.ref2, nn + e.ref); H.lin.push_back(o); }
  }
  // gather form of the MNA stamps: replay the ops with their add / assign semantics
  {
    std::vector<std::vector<std::pair<int, signed char> > > terms(np);
    auto add = [&](int q, int slot, int kind) { if (q >= 0) terms[q].push_back(std::make_pair(slot, (signed char)kind)); };
    auto set = [&](int q, int slot, int kind) { if (q >= 0) { terms[q].clear(); terms[q].push_back(std::make_pair(slot, (signed char)kind)); } };
    for (auto& o : H.lin) {
      switch (o.kind) {
        case K_R: add(o.pos[0], o.pbase, 1); add(o.pos[1], o.pbase, -1); add(o.pos[2], o.pbase, -1); add(o.pos[3], o.pbase, 1); break;
        case K_G: add(o.pos[0], o.pbase, 2); add(o.pos[1], o.pbase, -2); add(o.pos[2], o.pbase, -2); add(o.pos[3], o.pbase, 2); break;
        case K_VDE: case K_E: case K_H:
          set(o.pos[0], 0, 3); set(o.pos[1], 0, -3); set(o.pos[2], 0, 3); set(o.pos[3], 0, -3);
          if (o.kind == K_E) { set(o.pos[4], o.pbase, -2); set(o.pos[5], o.pbase, 2); }
          if (o.kind == K_H) set(o.pos[4], o.pbase, 2);
          break;
        case K_F: add(o.pos[0], o.pbase, 2); add(o.pos[1], o.pbase, -2); break;
      }
    }
    H.term_ptr.assign(np + 1, 0);
    for (int q = 0; q < np; q++) {
      H.term_ptr[q + 1] = H.term_ptr[q] + (int)terms[q].size();
      for (auto& t : terms[q]) { H.term_slot.push_back(t.first); H.term_kind.push_back(t.second); }
    }
    if (H.term_slot.empty()) { H.term_slot.push_back(0); H.term_kind.push_back(0); }
  }
  // persistent parameter slots: the values still needed after the per-instance set-up
  // (sources, diode / mosq / switch devices, OP-guess terms); everything else (R, C, L,
  // EKV model + geometry) is read once from HBM while the instance is loaded
  auto persist = [&](int slot0, int cnt) {
    int at = (int)H.pstage.size();
    for (int i = 0; i < cnt; i++) H.pstage.push_back(slot0 + i);
    return at;
  };
  // sources: I first, then V (the reference's two loops)
  int ntd = 0, nac = 0;
  for (int pass = 0; pass < 2; pass++)
    for (int k = 0; k < c->n_elems; k++) {
      const ahk_elem& e = c->elems[k];
      if ((pass == 0 && e.type != AHK_I) || (pass == 1 && e.type != AHK_V)) continue;
      SrcOp s; std::memset(&s, 0, sizeof s);
      s.is_v = e.type == AHK_V; s.elem = k; s.pbase = e.pbase; s.tf = e.tf;
      s.has_ac = (e.flags & AHK_FLAG_HAS_AC) ? 1 : 0;
      if (s.is_v) { s.row0 = nv1 + e.vde; s.row1 = -1; }
      else { s.row0 = e.n1 - 1; s.row1 = e.n2 - 1; }
      s.ntf = !e.tf ? 0 : (e.tf == AHK_TF_PWL ? 4 + 2 * (int)c->nominal[e.pbase + 3] : 8);
      if (e.tf == AHK_TF_PWL && (int)c->nominal[e.pbase + 3] < 2) throw std::runtime_error("pwl needs at least 2 points");
      // persistent block of a source: its dc value, then its waveform block; |ac| and
      // arg(ac) are read from HBM by the AC kernel (once per instance)
      s.ppb = persist(e.pbase, 1);
      if (s.ntf) persist(e.pbase + 3, s.ntf);
      if (s.tf) ntd++;
      if (s.has_ac) nac++;
      H.src.push_back(s);
    }
  // devices + gather lists
  std::vector<std::vector<std::pair<short, signed char> > > nl(np), tx(n);
  int ndout = 0;
  auto addJ = [&](int ur, int uc, int srcidx, int sgn) { int q = P(ur, uc); if (q >= 0) nl[q].push_back(std::make_pair((short)srcidx, (signed char)sgn)); };
  auto addT = [&](int ur, int srcidx, int sgn) { if (ur > 0) tx[ur - 1].push_back(std::make_pair((short)srcidx, (signed char)sgn)); };
  for (int k = 0; k < c->n_elems; k++) {
    const ahk_elem& e = c->elems[k];
    int cnt = dout_count(e.type);
    if (!cnt) continue;
    H.nonlinear = true;
    DevOp d; std::memset(&d, 0, sizeof d);
    d.type = e.type; d.pbase = e.pbase; d.mbase = e.mbase; d.flags = e.flags;
    d.dout = ndout; d.active = 1; d.con = -1; d.ppb = d.pmb = d.st = -1;   // con: assign_constant_blocks
    if (e.type == AHK_MOSQ) { d.ppb = persist(e.pbase, 6); d.pmb = persist(e.mbase, 7); }
    if (e.type == AHK_SW) { d.ppb = persist(e.pbase, 1); d.pmb = persist(e.mbase, 4); }
    if (e.type == AHK_D || e.type == AHK_SW) { d.st = H.nst; H.nst += 2; }
    const int b = ndout;
    switch (e.type) {
      case AHK_D:   // stamp path, diode.py:150-198
        d.node[0] = e.n1 - 1; d.node[1] = e.n2 - 1; d.node[2] = d.node[3] = -1;
        addJ(e.n1, e.n1, b + 1, +1); addJ(e.n1, e.n2, b + 1, -1);
        addJ(e.n2, e.n1, b + 1, -1); addJ(e.n2, e.n2, b + 1, +1);
        addT(e.n1, b, +1); addT(e.n2, b, -1);
        H.lock.push_back(e.n1); H.lock.push_back(e.n2);
        break;
      case AHK_EKV: {  // generic path, dc_analysis.py:863-881; ports (d,b),(g,b),(s,b)
        d.node[0] = e.n1 - 1; d.node[1] = e.n2 - 1; d.node[2] = e.n3 - 1; d.node[3] = e.n4 - 1;
        int o1 = e.n1, o2 = e.n3, dp[3] = {e.n1, e.n2, e.n3};
        if (!(o1 || o2)) d.active = 0;
        else {
          addT(o1, b, +1); addT(o2, b, -1);
          for (int t = 0; t < 3; t++) {
            addJ(o1, dp[t], b + 1 + t, +1); addJ(o1, e.n4, b + 1 + t, -1);
            addJ(o2, dp[t], b + 1 + t, -1); addJ(o2, e.n4, b + 1 + t, +1);
          }
        }
        H.lock.push_back(e.n1); H.lock.push_back(e.n4);
        H.lock.push_back(e.n2); H.lock.push_back(e.n4);
        H.lock.push_back(e.n3); H.lock.push_back(e.n4);
      } break;
      case AHK_MOSQ: { // stamp path, mosq.py:340-403; ports (d,s),(g,s),(b,s)
        d.node[0] = e.n1 - 1; d.node[1] = e.n2 - 1; d.node[2] = e.n3 - 1; d.node[3] = e.n4 - 1;
        int nd4[4] = {e.n1, e.n2, e.n3, e.n4};
        for (int cc = 0; cc < 4; cc++) { addJ(e.n1, nd4[cc], b + 1 + cc, +1); addJ(e.n3, nd4[cc], b + 5 + cc, +1); }
        addT(e.n1, b, +1); addT(e.n3, b, -1);
        H.lock.push_back(e.n1); H.lock.push_back(e.n3);
        H.lock.push_back(e.n2); H.lock.push_back(e.n3);
        H.lock.push_back(e.n4); H.lock.push_back(e.n3);
      } break;
      case AHK_SW: {   // generic path; ports (n1,n2),(sn1,sn2)
        d.node[0] = e.n1 - 1; d.node[1] = e.n2 - 1; d.node[2] = e.n3 - 1; d.node[3] = e.n4 - 1;
        if (!(e.n1 || e.n2)) d.active = 0;
        else {
          addT(e.n1, b, +1); addT(e.n2, b, -1);
          int dp[2][2] = {{e.n1, e.n2}, {e.n3, e.n4}};
          for (int t = 0; t < 2; t++) {
            addJ(e.n1, dp[t][0], b + 1 + t, +1); addJ(e.n1, dp[t][1], b + 1 + t, -1);
            addJ(e.n2, dp[t][0], b + 1 + t, -1); addJ(e.n2, dp[t][1], b + 1 + t, +1);
          }
        }
        H.lock.push_back(e.n1); H.lock.push_back(e.n2);
        H.lock.push_back(e.n3); H.lock.push_back(e.n4);
      } break;
    }
    ndout += cnt;
    H.dev.push_back(d);
  }
  H.nl_ptr.assign(np + 1, 0);
  for (int q = 0; q < np; q++) {
    H.nl_ptr[q + 1] = H.nl_ptr[q] + (int)nl[q].size();
    for (auto& pr : nl[q]) { H.nl_src.push_back(pr.first); H.nl_sgn.push_back(pr.second); }
  }
  H.tx_ptr.assign(n + 1, 0);
  for (int r = 0; r < n; r++) {
    H.tx_ptr[r + 1] = H.tx_ptr[r] + (int)tx[r].size();
    for (auto& pr : tx[r]) { H.tx_src.push_back(pr.first); H.tx_sgn.push_back(pr.second); }
  }
  // guess operator
  const int ng = c->n_guess > 0 ? c->n_guess : 0;
  H.guessG.assign(c->guess_G, c->guess_G + (size_t)n * ng);
  H.gslot.assign(c->guess_slot, c->guess_slot + ng);
  H.gscale.assign(c->guess_scale, c->guess_scale + ng);
  H.gconst.assign(c->guess_const, c->guess_const + ng);
  H.gpslot.assign(ng > 0 ? ng : 1, -1);
  for (int j = 0; j < ng; j++) if (H.gslot[j] >= 0) H.gpslot[j] = persist(H.gslot[j], 1);
  if (H.pstage.empty()) H.pstage.push_back(0);
  H.nominal.assign(c->nominal, c->nominal + c->n_params);
  H.slot_vary.assign(c->n_params > 0 ? c->n_params : 1, -1);
  assign_constant_blocks(H);
  // keep vectors non-empty so data() is a valid pointer everywhere
  if (H.nl_src.empty()) { H.nl_src.push_back(0); H.nl_sgn.push_back(0); }
  if (H.tx_src.empty()) { H.tx_src.push_back(0); H.tx_sgn.push_back(0); }
  Prog& h = H.hdr;
  std::memset(&h, 0, sizeof h);
  h.n = n; h.nv1 = nv1; h.npos = np; h.nlin = (int)H.lin.size(); h.nsrc = (int)H.src.size();
  h.ntd = ntd; h.nac = nac; h.ndev = (int)H.dev.size(); h.ndout = ndout;
  h.nlock = (int)H.lock.size() / 2; h.nguess = ng; h.nparams = c->n_params;
  h.npersist = 0;
  for (auto& so : H.src) h.npersist = std::max(h.npersist, so.ppb + 1 + so.ntf);
  for (auto& dv : H.dev) {
    if (dv.pmb >= 0) h.npersist = std::max(h.npersist, dv.pmb + (dv.type == AHK_MOSQ ? 7 : 4));
  }
  for (int j = 0; j < ng; j++) h.npersist = std::max(h.npersist, H.gpslot[j] + 1);
}

// Per-instance device-constant blocks (EKV 16 doubles, diode 10).  Devices that share a
// model and have identical, un-varied own parameters (e.g. the three NMOS of a ring
// oscillator) share ONE block: fewer shared-memory bytes per instance.  Depends on which
// slots the batch varies, so it is recomputed whenever they change.
inline void assign_constant_blocks(HostProg& H) {
  H.ncon = 0;
  std::map<std::vector<double>, int> seen;
  for (auto& d : H.dev) {
    if (d.type != AHK_EKV && d.type != AHK_D) { d.con = -1; continue; }
    const int own = d.type == AHK_EKV ? 4 : 1, size = d.type == AHK_EKV ? 16 : 10;
    bool varied = false;
    std::vector<double> key;
    key.push_back(d.type); key.push_back(d.mbase); key.push_back(d.flags);
    for (int i = 0; i < own; i++) {
      if (H.slot_vary[d.pbase + i] >= 0) varied = true;
      key.push_back(H.nominal[d.pbase + i]);
    }
    auto it = varied ? seen.end() : seen.find(key);
    if (it != seen.end()) { d.con = it->second; continue; }
    d.con = H.ncon;
    H.ncon += size;
    if (!varied) seen[key] = d.con;
  }
}

inline void set_vary_slots(HostProg& H, int n_vary, const int* slots) {
  std::fill(H.slot_vary.begin(), H.slot_vary.end(), -1);
  for (int k = 0; k < n_vary; k++) {
    if (slots[k] < 0 || slots[k] >= H.hdr.nparams) throw std::runtime_error("vary slot out of range");
    H.slot_vary[slots[k]] = k;
  }
  assign_constant_blocks(H);
}

enum { MODE_OP = 0, MODE_TRAN = 1 };

// Shared-memory layout of one instance for G lanes per instance.
// NC > 0: register variant with NC matrix columns (staging rows of NC - 1 doubles).
// Every array is allocated only when the mode / variant / circuit needs it: the
// bytes per instance bound how many instances are resident per SM.
// jit: layout of a circuit-specialised kernel (jit_gen.hpp) — it assembles straight into
// registers and needs no staging rows.
inline Layout make_layout(const HostProg& H, int mode, int method, int use_step_control, int G,
                          int NC = 0, int R = 1, bool jit = false, bool big = false) {
  const Prog& h = H.hdr;
  Layout L; std::memset(&L, 0, sizeof L);
  const int n = h.n;
  L.RS = NC > 0 ? NC - 1 : (big ? ((n + 3) & ~3) : ((n + 1) | 1));   // big: rows start on 32-byte sectors
  L.ARS = L.RS;
  int off = 0;
  auto take = [&](int cnt) { int o = off; off += cnt > 0 ? cnt : 0; return o; };
  // register variants with several rows per lane stage through one scratch row per lane
  // big: run-time n beyond AHK_SMALL_NMAX — the n x RS matrix lives in the global workspace
  L.big = big ? 1 : 0;
  L.A = take(big ? 0 : (jit && NC > 0 ? 0 : (NC > 0 && R > 1 ? (G < n ? G : n) : n) * L.RS));
  L.brhs = take(big ? n : 0);
  L.nzm = take(big ? (n * ((n + 31) / 32) * 4 + 7) / 8 : 0);
  // (dense stamp patterns eliminate in dense panels from the first step: no column bitmaps,
  // their shared memory is worth more as occupancy — mesh n = 202: 233 ms without, 279 ms with)
  const bool cz = big && (long long)h.npos * 4 <= (long long)n * n;
  L.czm = cz ? take((n * ((n + 31) / 32) * 4 + 7) / 8) : -1;

  L.x = take(n); L.dx = take(n); L.T = take(n); L.res = take(n);
  // (many matrix positions — dense patterns — do not fit next to the vectors: global workspace)
  L.vglob = big && h.npos > AHK_VGLOB_NPOS ? 1 : 0;
  int voff = 0;
  auto takev = [&](int cnt) { if (!L.vglob) return take(cnt); int o2 = voff; voff += cnt > 0 ? cnt : 0; return o2; };
  L.Mv = takev(h.npos);
  // the matrix values are formed while assembling: Mv (+ C1 Dv in a transient) plus Gmin
  // on the node diagonals; Bv is kept as an alias of Mv
  L.Bv = L.Mv;
  L.Nc = take(n);
  L.Nd = h.ntd > 0 ? take(n) : L.Nc;      // N + Tt(t); no waveforms: N itself
  L.dout = take(h.ndout); L.dst = take(H.nst);
  // OP: xs = residual of pass 1, x1 = solution of pass 1.  Transient: xs = x(tstart), used
  // only before the first accepted step, when the history ring still holds it in slot 0
  // (aliased below); x1 is not used.
  if (mode != MODE_TRAN) { L.xs = take(n); L.x1 = take(n); }
  // pivot bookkeeping of the generic solves: prow / stepof as bytes (n <= 64), or as 16-bit
  // entries plus the row list of the current elimination step (large n)
  L.bytes = take(NC > 0 ? 0 : (big ? (6 * n + 4 * ((n + 31) / 32) + 7) / 8 + 2 : 16));   // + the used-row bitmap
  L.Pv = take(h.npersist); L.dcon = take(H.ncon);
  L.pvb = take(NC > 0 && G > 1 ? 2 * NC : 0);
  L.xin = take(n);
  L.buflen = 1;
  if (mode == MODE_TRAN) {
    int order, max_x, max_dx, pmax_x;
    if (method == AHK_IMPLICIT_EULER) { order = 1; max_x = 0; max_dx = -1; pmax_x = 1; }
    else if (method == AHK_TRAP) { order = 2; max_x = 0; max_dx = 0; pmax_x = 2; }
    else { order = method - 10; max_x = order; max_dx = -1; pmax_x = order; }
    (void)order;
    // transient.py:251-260 (fixed step with max_dx None raises TypeError on
    // python 3 — SURVEY quirk 13 — so it is sized like the step-control case)
    int bl = max_x > max_dx ? max_x : max_dx;
    if (use_step_control || max_dx < 0) { if (pmax_x > bl) bl = pmax_x; }
    L.buflen = bl + 1;
    L.Dv = takev(h.npos); L.Ntr = take(n); L.xacc = take(n);
    L.hx = take(L.buflen * n); L.hdx = take(n); L.ht = take(L.buflen);
    L.xs = L.hx; L.x1 = L.hx;
    L.C0 = take(n);
    L.pred = use_step_control ? take(n) : L.C0;   // the predictor exists only with LTE control
  } else {
    L.Dv = L.Mv; L.Ntr = L.Nd; L.xacc = L.xs;
    L.hx = L.hdx = L.ht = L.C0 = L.pred = 0;
  }
  // bank rule: consecutive groups of a half-warp must land on distinct 8-byte banks
  // (one lane per instance: any odd stride spreads 16 lanes over the 16 banks)
  if (G == 1) { if (off % 2 == 0) off++; }
  else {
    int want = (G * L.RS) % 16;
    if (want % 2 == 0) want = (want + 1) % 16;
    while (off % 16 != want) off++;
  }
  L.stride = off;
  L.vlen = (voff + 3) & ~3;      // (keeps every instance's workspace 32-byte aligned: double2 tiles of the dense panels)
  // multipliers of a dense panel (lu_panel_dense): in the instance's global workspace, behind
  // the matrix and the value block — sparse circuits never touch it and pay no shared memory
  L.lpan = big ? n * L.RS + L.vlen : 0;
  return L;
}

}  // namespace ahk
