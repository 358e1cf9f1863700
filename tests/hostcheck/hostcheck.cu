// hostcheck.cu - TEST FIXTURE.  Runs the per-index bodies of the CUDA tree build
// (csrc/build_logic.h, csrc/tree_logic.h) and the per-lane walk arithmetic in plain
// sequential host loops, so that `pytest -m "not gpu"` can check the construction logic
// against the unmodified reference on a machine without a GPU.  It is compiled only by
// tests/hostcheck/build.py, lives under tests/, and is never loaded by the product package.
#include <vector>
#include <algorithm>
#include <numeric>
#include <cuda_runtime.h>
#include "../../sidm-nbody_b200/csrc/build_logic.h"

using namespace b200;

namespace {
struct Host {
  int n = 0, m = 0, maxlev = 0;
  std::vector<float4> posm, geom, leaf_posm;
  std::vector<uint64_t> hi, lo, shi, slo;
  std::vector<int> sidx, nodestart, nstart, nend, nparent, npstart, ndp, narrive, nminidx, nlstart, leaf_orig, orig_leaf, krank, lrank, leaf_parent;
  std::vector<signed char> clev;
  std::vector<unsigned char> nlevel, nnp, nnchild;
  std::vector<NodeRec> nodes;
  std::vector<Moments> nmom;
  RootBox root;
  RootBox roots[8];                       // several particle types: one root cell per type
  std::vector<unsigned char> stype;       // type of every sorted particle (empty: one type)
  int flags[16];
  BuildView v;                           // the views of the last build (hc_refit goes on with them)
  std::vector<float> next;
} H;
}

// types == nullptr: one tree.  Otherwise one tree per particle type as csrc/tree_build.cu lays them out: a root
// cell per type, keys along the particle's own tree, key order, then a stable sort by type.
static int hc_build_impl(int n, const float *pos, const float *mass, const int *types, int want_lorder) {
  H.n = n;
  H.posm.resize(n);
  double mn[8][3], mx[8][3];
  for (int t = 0; t < 8; t++) for (int k = 0; k < 3; k++) { mn[t][k] = 1e300; mx[t][k] = -1e300; }
  for (int i = 0; i < n; i++) {
    H.posm[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], mass[i]);
    const int t = types ? (types[i] & 7) : 0;
    for (int k = 0; k < 3; k++) { mn[t][k] = std::min(mn[t][k], (double)pos[3 * i + k]); mx[t][k] = std::max(mx[t][k], (double)pos[3 * i + k]); }
  }
  for (int t = 0; t < 8; t++) if (mn[t][0] <= mx[t][0]) H.roots[t] = make_root(mn[t], mx[t]);
  H.root = H.roots[0];
  H.hi.resize(n); H.lo.resize(n);
  for (int i = 0; i < n; i++) make_key(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], H.roots[types ? (types[i] & 7) : 0], H.hi[i], H.lo[i]);
  H.sidx.resize(n); std::iota(H.sidx.begin(), H.sidx.end(), 0);
  std::stable_sort(H.sidx.begin(), H.sidx.end(), [&](int a, int b) { return key_less(H.hi[a], H.lo[a], H.hi[b], H.lo[b]); });
  H.stype.clear();
  if (types) {
    std::stable_sort(H.sidx.begin(), H.sidx.end(), [&](int a, int b) { return (types[a] & 7) < (types[b] & 7); });
    H.stype.resize(n);
    for (int j = 0; j < n; j++) H.stype[j] = (unsigned char)(types[H.sidx[j]] & 7);
  }
  H.shi.resize(n); H.slo.resize(n);
  for (int j = 0; j < n; j++) { H.shi[j] = H.hi[H.sidx[j]]; H.slo[j] = H.lo[H.sidx[j]]; }
  const int cap = 4 * n + 64;
  H.clev.assign(n + 1, 0); H.nodestart.assign(n + 2, 0);
  H.nodes.assign(cap, NodeRec()); H.geom.resize(cap); H.nstart.assign(cap, 0); H.nend.assign(cap, 0); H.nparent.assign(cap, -1);
  H.npstart.assign(cap + 1, 0); H.nlevel.assign(cap, 0); H.nnp.assign(cap, 0); H.nnchild.assign(cap, 0); H.ndp.assign(8 * (size_t)cap, -1);
  H.narrive.assign(cap, 0); H.nminidx.assign(cap, 0); H.nlstart.assign(cap, 0); H.nmom.resize(cap);
  H.leaf_posm.resize(n); H.leaf_orig.assign(n, 0); H.orig_leaf.assign(n, 0); H.krank.assign(n, 0); H.lrank.assign(n, 0); H.leaf_parent.assign(n, 0);
  for (int k = 0; k < 16; k++) H.flags[k] = 0;
  BuildView &v = H.v; v = BuildView();
  v.n = n; v.maxnodes = cap; v.posm = H.posm.data(); v.shi = H.shi.data(); v.slo = H.slo.data(); v.sidx = H.sidx.data();
  v.clev = H.clev.data(); v.nodestart = H.nodestart.data(); v.root = types ? H.roots : &H.root;
  v.stype = types ? H.stype.data() : nullptr;
  v.nodes = H.nodes.data(); v.geom = H.geom.data(); v.nstart = H.nstart.data(); v.nend = H.nend.data(); v.nparent = H.nparent.data();
  v.npstart = H.npstart.data(); v.nlevel = H.nlevel.data(); v.nnp = H.nnp.data(); v.nnchild = H.nnchild.data(); v.ndp = H.ndp.data();
  v.narrive = H.narrive.data(); v.nminidx = H.nminidx.data(); v.nlstart = H.nlstart.data(); v.nmom = H.nmom.data();
  v.leaf_posm = H.leaf_posm.data(); v.leaf_orig = H.leaf_orig.data(); v.orig_leaf = H.orig_leaf.data(); v.krank = H.krank.data();
  v.lrank = H.lrank.data(); v.leaf_parent = H.leaf_parent.data(); v.flags = H.flags;
  // B1 + scan
  std::vector<int> cnt(n + 1, 0);
  int maxlev = 0;
  for (int i = 0; i < n; i++) {
    int c = b1_common(v, i);
    if (c >= kMaxLevels) return 9005;
    int cp = i > 0 ? b1_common(v, i - 1) : -1;
    H.clev[i] = (signed char)c; cnt[i] = b1_count_from(cp, c, i, n);
    if (c > maxlev) maxlev = c;
  }
  int acc = 0;
  for (int i = 0; i <= n; i++) { H.nodestart[i] = acc; acc += cnt[i]; }
  H.m = H.nodestart[n]; H.maxlev = maxlev;
  if (H.m >= cap) return 1;
  for (int i = 0; i < n; i++) b2_body(v, i);
  for (int id = 0; id < H.m; id++) b3_body(v, id);
  int ps = 0;
  for (int id = 0; id <= H.m; id++) { H.npstart[id] = ps; if (id < H.m) ps += H.nnp[id]; }
  for (int id = 0; id < H.m; id++) b4_body(v, id);
  for (int lev = maxlev; lev >= 0; lev--)
    for (int id = 0; id < H.m; id++) if (H.nlevel[id] == lev) b5_body(v, id);
  if (want_lorder)
    for (int lev = 0; lev <= maxlev; lev++)
      for (int id = 0; id < H.m; id++) if (H.nlevel[id] == lev) b6_body(v, id);
  return 0;
}

extern "C" int hc_build(int n, const float *pos, const float *mass, int want_lorder) { return hc_build_impl(n, pos, mass, nullptr, want_lorder); }
extern "C" int hc_build_types(int n, const float *pos, const float *mass, const int *types, int want_lorder) {
  return hc_build_impl(n, pos, mass, types, want_lorder);
}
extern "C" int hc_num_nodes() { return H.m; }
extern "C" int hc_max_level() { return H.maxlev; }
extern "C" void hc_get_root(float *out) { out[0] = H.root.cx; out[1] = H.root.cy; out[2] = H.root.cz; out[3] = H.root.len; }

extern "C" void hc_get_tree(float *center, float *len, float *mass, float *s, float *Q, float *oc, float *bmax2, int *count,
                            int *level, int *skip, int *parent, int *minidx) {
  for (int id = 0; id < H.m; id++) {
    const NodeRec &r = H.nodes[id]; const float4 gm = H.geom[id];
    center[3 * id] = gm.x; center[3 * id + 1] = gm.y; center[3 * id + 2] = gm.z; len[id] = gm.w;
    mass[id] = r.mass; s[3 * id] = r.sx; s[3 * id + 1] = r.sy; s[3 * id + 2] = r.sz;
    float *q = Q + 7 * id; q[0] = r.q11; q[1] = r.q22; q[2] = r.q33; q[3] = r.q12; q[4] = r.q13; q[5] = r.q23; q[6] = r.p;
    oc[id] = r.oc; bmax2[id] = r.bmax2; count[id] = H.nend[id] - H.nstart[id] + 1; level[id] = H.nlevel[id];
    skip[id] = r.skip; parent[id] = H.nparent[id]; minidx[id] = H.nminidx[id];
  }
}
extern "C" void hc_get_orders(int *sidx, int *leaf_orig, int *lrank) {
  for (int i = 0; i < H.n; i++) { sidx[i] = H.sidx[i]; leaf_orig[i] = H.leaf_orig[i]; lrank[i] = H.lrank[i]; }
}

// one target's walk with the lane arithmetic of csrc/walk.cu (float interactions, float
// partial sums flushed into double)
extern "C" void hc_walk(int nt, const int *targets, const float *oldacc, int criterion, float theta, float alpha, float eps,
                        double *acc, int *cost) {
  const float h_inv = (float)(1.0 / (2.8 * (double)eps)), theta2 = theta * theta;
  for (int t = 0; t < nt; t++) {
    const float4 tp = H.posm[targets[t]];
    const float oa = oldacc[targets[t]];
    const bool bh = criterion == 0 || oa == 0.0f;
    const float oac = oa * alpha;
    double ax = 0, ay = 0, az = 0; float fx = 0, fy = 0, fz = 0; int it = 0, npart = 0, nnode = 0;
    int no = 0;
    while (no < H.m) {
      const NodeRec &n = H.nodes[no];
      const float dx = n.sx - tp.x, dy = n.sy - tp.y, dz = n.sz - tp.z;
      const float r2 = dx * dx + dy * dy + dz * dz;
      const bool open = bh ? open_bh(n.len2, r2, theta2) : open_rel(n.oc, n.bmax2, r2, oac);
      if (!open) { pn_force(dx, dy, dz, r2, n, h_inv, fx, fy, fz); nnode++; no = n.skip; }
      else {
        const int np = n.pinfo & 15, ps = n.pinfo >> 4;
        for (int k = 0; k < np; k++) { const float4 q = H.leaf_posm[ps + k]; pp_force(q.x - tp.x, q.y - tp.y, q.z - tp.z, q.w, h_inv, fx, fy, fz); npart++; }
        no = no + 1;
      }
      if (++it == 8) { ax += fx; ay += fy; az += fz; fx = fy = fz = 0; it = 0; }
    }
    ax += fx; ay += fy; az += fz;
    acc[3 * t] = ax; acc[3 * t + 1] = ay; acc[3 * t + 2] = az; cost[2 * t] = npart; cost[2 * t + 1] = nnode;
  }
}

// the same lane arithmetic over a forest (one tree per particle type, k_walk<PER, MULTI> of csrc/walk.cu): every tree in
// turn from its root to the next root, h = 2.8 max(eps of the tree's type, eps of the target's type)
extern "C" void hc_walk_types(int nt, const int *targets, const float *oldacc, const int *types, const float *eps_table, int criterion,
                              float theta, float alpha, double *acc, int *cost) {
  const float theta2 = theta * theta;
  std::vector<int> roots;
  for (int id = 0; id < H.m; id++) if (H.nparent[id] < 0) roots.push_back(id);
  roots.push_back(H.m);
  for (int t = 0; t < nt; t++) {
    const float4 tp = H.posm[targets[t]];
    const float oa = oldacc[targets[t]];
    const bool bh = criterion == 0 || oa == 0.0f;
    const float oac = oa * alpha;
    const float eps_t = eps_table[types[targets[t]] & 7];
    double ax = 0, ay = 0, az = 0; int npart = 0, nnode = 0;
    for (size_t r = 0; r + 1 < roots.size(); r++) {
      const int ttype = H.stype[H.nstart[roots[r]]];                 // type of the tree's first sorted particle
      const float h_inv = 1.0f / (2.8f * std::max(eps_table[ttype], eps_t));
      float fx = 0, fy = 0, fz = 0; int it = 0;
      int no = roots[r];
      while (no < roots[r + 1]) {
        const NodeRec &n = H.nodes[no];
        const float dx = n.sx - tp.x, dy = n.sy - tp.y, dz = n.sz - tp.z;
        const float r2 = dx * dx + dy * dy + dz * dz;
        const bool open = bh ? open_bh(n.len2, r2, theta2) : open_rel(n.oc, n.bmax2, r2, oac);
        if (!open) { pn_force(dx, dy, dz, r2, n, h_inv, fx, fy, fz); nnode++; no = n.skip; }
        else {
          const int np = n.pinfo & 15, ps = n.pinfo >> 4;
          for (int k = 0; k < np; k++) { const float4 q = H.leaf_posm[ps + k]; pp_force(q.x - tp.x, q.y - tp.y, q.z - tp.z, q.w, h_inv, fx, fy, fz); npart++; }
          no = no + 1;
        }
        if (++it == 8) { ax += fx; ay += fy; az += fz; fx = fy = fz = 0; it = 0; }
      }
      ax += fx; ay += fy; az += fz;
    }
    acc[3 * t] = ax; acc[3 * t + 1] = ay; acc[3 * t + 2] = az; cost[2 * t] = npart; cost[2 * t + 1] = nnode;
  }
}

// ---- periodic neighbour searches: the classification that lets a search use the open-boundary kernels (tree_logic.h)
extern "C" int hc_cube_clear_of_faces(const float *dom, double box, const float *cube) {
  return b200::cube_clear_of_faces(dom, box, cube[0], cube[1], cube[2], cube[3], cube[4], cube[5]) ? 1 : 0;
}
// r^2 of the wrapped and of the plain float test (forcetree.c:2195-2206 with / without ngb_periodic) for n particles and one centre
extern "C" void hc_dist2_both(int n, const float *pos, const float *c, double box, float *r2_wrapped, float *r2_plain) {
  using namespace b200;
  for (int i = 0; i < n; i++) {
    const float dx = fadd(pos[3 * i], -c[0]), dy = fadd(pos[3 * i + 1], -c[1]), dz = fadd(pos[3 * i + 2], -c[2]);
    const float wx = wrap_periodic(dx, box), wy = wrap_periodic(dy, box), wz = wrap_periodic(dz, box);
    r2_wrapped[i] = fadd(fadd(fmul(wx, wx), fmul(wy, wy)), fmul(wz, wz));
    r2_plain[i] = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
  }
}

// ---- tree reuse: the refit of csrc/tree_build.cu tree_refit_impl() on the host - same topology, leaves from the new positions,
// all moments and cell extents bottom-up (b5_body with BuildView::next set).  Returns the largest coordinate displacement.
extern "C" float hc_refit(const float *pos_new) {
  float pad = 0.f;
  for (int i = 0; i < H.n; i++) { H.posm[i].x = pos_new[3 * i]; H.posm[i].y = pos_new[3 * i + 1]; H.posm[i].z = pos_new[3 * i + 2]; }
  for (int L = 0; L < H.n; L++) {
    const float4 p = H.posm[H.leaf_orig[L]], o = H.leaf_posm[L];
    pad = std::max(pad, std::max(fabsf(p.x - o.x), std::max(fabsf(p.y - o.y), fabsf(p.z - o.z))));
    H.leaf_posm[L] = p;
  }
  H.next.assign(H.m + 1, 0.f);
  H.v.next = H.next.data();
  for (int lev = H.maxlev; lev >= 0; lev--)
    for (int id = 0; id < H.m; id++) if (H.nlevel[id] == lev) b5_body(H.v, id);
  H.v.next = nullptr;
  return pad;
}
// per node: len2 of the walk's record, the leaf range [first, end) of its subtree, the cell extent of the last refit
extern "C" void hc_get_refit(float *len2, int *leaf_first, int *leaf_end, float *extent) {
  for (int id = 0; id < H.m; id++) {
    len2[id] = H.nodes[id].len2;
    leaf_first[id] = H.npstart[id];
    const int skip = H.nodes[id].skip;
    leaf_end[id] = skip < H.m ? H.npstart[skip] : H.n;
    extent[id] = H.next.empty() ? 0.f : H.next[id];
  }
}
extern "C" void hc_get_leaves(float *leaf_pos, int *leaf_orig) {
  for (int L = 0; L < H.n; L++) { leaf_pos[3 * L] = H.leaf_posm[L].x; leaf_pos[3 * L + 1] = H.leaf_posm[L].y; leaf_pos[3 * L + 2] = H.leaf_posm[L].z; leaf_orig[L] = H.leaf_orig[L]; }
}
