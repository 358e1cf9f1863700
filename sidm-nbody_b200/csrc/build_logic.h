// build_logic.h - per-index bodies of the octree construction kernels (host+device).
// tree_build.cu launches one CUDA thread per index; tests/hostcheck runs the same bodies in a
// sequential loop to check the construction against the reference without a GPU.
//
// Construction (replaces the sequential insertion of forcetree.c:166-399):
//   sorted keys -> clev[i] = levels shared by sorted particles i and i+1
//   a node of level L starts at sorted particle a  <=>  clev[a-1] < L <= clev[a]
//   => nodes starting at a : levels clev[a-1]+1 .. clev[a]; exclusive scan -> node ids,
//      which come out in depth-first pre-order (children after parents, siblings by octant).
#pragma once
#include "tree_logic.h"

namespace b200 {

struct BuildView {
  int n;                      // particles in the tree
  int maxnodes;
  const float4 *posm;         // original order: PosPred.xyz, mass
  const uint64_t *shi, *slo;  // sorted keys
  const int *sidx;            // sorted position -> original index
  signed char *clev;          // [n]
  int *nodestart;             // [n+1]
  const RootBox *root;        // one root cell; with several particle types: root[type]
  const unsigned char *stype = nullptr;   // several types (one tree per type, forcetree.c:90-158): type of every sorted particle
  // per node
  NodeRec *nodes; float4 *geom; int *nstart, *nend, *nparent, *npstart;
  unsigned char *nlevel, *nnp, *nnchild; int *ndp; int *narrive; int *nminidx; int *nlstart; Moments *nmom;
  // leaf order
  float4 *leaf_posm; int *leaf_orig, *orig_leaf; int *krank; int *lrank; int *leaf_parent;
  int *flags;                 // device status words
  // refit (tree reuse): extent of every cell's particles about the cell centre (largest coordinate distance), bottom-up;
  // null in a fresh build, where every particle lies inside its cell
  float *next = nullptr;
};

// ---- B1: shared levels with the next particle, and how many nodes start here
// levels two sorted particles share; particles of different types share none (they live in different trees)
B200_HD int common_of(const BuildView &v, int a, int b) {
  if (v.stype && v.stype[a] != v.stype[b]) return -1;
  return common_levels(v.shi[a], v.slo[a], v.shi[b], v.slo[b]);
}
B200_HD int b1_common(const BuildView &v, int i) {
  return (i + 1 < v.n) ? common_of(v, i, i + 1) : -1;
}
B200_HD int b1_count_from(int c_prev, int c_here, int i, int n) {
  int cnt = (i == 0) ? c_here + 1 : (c_here > c_prev ? c_here - c_prev : 0);
  if (n == 1) cnt = 1;     // a lone particle still gets a root node (forcetree.c:214-232)
  return cnt;
}

// ---- B2: one thread per sorted particle writes (start, level) of every node starting there
B200_HD void b2_body(const BuildView &v, int i) {
  const int c_prev = i > 0 ? v.clev[i - 1] : -1;
  const int first = v.nodestart[i], cnt = v.nodestart[i + 1] - first;
  for (int k = 0; k < cnt; k++) {
    const int id = first + k;
    if (id >= v.maxnodes) return;
    v.nstart[id] = i;
    v.nlevel[id] = (unsigned char)(c_prev + 1 + k);
  }
  v.krank[v.sidx[i]] = i;
}

// last sorted index that shares `level` leading octants with sorted particle a
B200_HD int range_end(const BuildView &v, int a, int level) {
  if (level == 0 && !v.stype) return v.n - 1;
  int lo = a, step = 1;                       // invariant: lo is inside
  while (lo + step < v.n && common_of(v, a, lo + step) >= level) { lo += step; step <<= 1; }
  int hi = lo + step; if (hi > v.n) hi = v.n;  // hi is outside (or n)
  while (hi - lo > 1) {
    const int mid = lo + ((hi - lo) >> 1);
    if (common_of(v, a, mid) >= level) lo = mid; else hi = mid;
  }
  return lo;
}

// first index in [a, e) whose octant at `level` is >= d
B200_HD int octant_lower_bound(const BuildView &v, int a, int e, int level, int d) {
  int lo = a, hi = e;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (digit_at(v.shi[mid], v.slo[mid], level) < d) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- B3: one thread per node: range, cell geometry, children
B200_HD void b3_body(const BuildView &v, int id) {
  const int a = v.nstart[id], level = v.nlevel[id];
  const int b = range_end(v, a, level);
  v.nend[id] = b;
  // cell: descend from the root along the first `level` octants of particle a
  const RootBox &rb = v.stype ? v.root[v.stype[a]] : v.root[0];
  float cx = rb.cx, cy = rb.cy, cz = rb.cz, len = rb.len;
  const uint64_t ahi = v.shi[a], alo = v.slo[a];
  for (int l = 0; l < level; l++) child_cell(digit_at(ahi, alo, l), cx, cy, cz, len);
  v.geom[id] = make_float4(cx, cy, cz, len);
  // children
  int bnd[9];
  bnd[0] = a; bnd[8] = b + 1;
  for (int d = 1; d < 8; d++) bnd[d] = (level < kMaxLevels) ? octant_lower_bound(v, a, b + 1, level, d) : b + 1;
  int np = 0, nch = 0;
  for (int d = 0; d < 8; d++) {
    const int cnt = bnd[d + 1] - bnd[d];
    if (cnt == 1) { v.ndp[8 * id + np] = bnd[d]; np++; }
    else if (cnt >= 2) {
      const int ca = bnd[d];
      const int c_prev = ca > 0 ? v.clev[ca - 1] : -1;
      const int child = v.nodestart[ca] + (level + 1) - (c_prev + 1);
      if (child < v.maxnodes) v.nparent[child] = id;
      nch++;
    }
  }
  for (int k = np; k < 8; k++) v.ndp[8 * id + k] = -1;
  v.nnp[id] = (unsigned char)np; v.nnchild[id] = (unsigned char)nch;
  v.narrive[id] = 0;
  if (level == 0) v.nparent[id] = -1;          // a root (one per particle type)
  v.nodes[id].skip = v.nodestart[b + 1];
}

// ---- B4: one thread per node: copy its direct particles into leaf order
B200_HD void b4_body(const BuildView &v, int id) {
  const int np = v.nnp[id], ps = v.npstart[id];
  int mn = 0x7fffffff;
  for (int k = 0; k < np; k++) {
    const int j = v.ndp[8 * id + k];
    const int o = v.sidx[j];
    v.leaf_posm[ps + k] = v.posm[o];
    v.leaf_orig[ps + k] = o;
    v.orig_leaf[o] = ps + k;
    v.leaf_parent[ps + k] = id;
    if (o < mn) mn = o;
  }
  v.nminidx[id] = mn;        // completed bottom-up in b5
  v.nodes[id].pinfo = (ps << 4) | np;
}

// ---- B5: moments of one node from its direct particles and its (finished) child nodes.  REFIT (tree reuse): also the extent of
// the cell's particles, see below; a compile-time switch so that the build's instantiation is exactly the code it was before
// the refit existed (the run-time form cost k_b5_levels 0.4 ms at N = 1e7)
template <bool REFIT>
B200_HD void b5_body_t(const BuildView &v, int id) {
  const float4 gm = v.geom[id];
  Moments m; moments_zero(m);
  const int np = v.nnp[id], ps = v.npstart[id];
  for (int k = 0; k < np; k++) {
    const float4 p = v.leaf_posm[ps + k];
    moments_add_particle(m, p.x, p.y, p.z, p.w, gm.x, gm.y, gm.z);
  }
  int mn = v.nminidx[id];
  const int end = v.nodes[id].skip;
  for (int c = id + 1; c < end; c = v.nodes[c].skip) {     // child nodes, in octant order
    const float4 cg = v.geom[c];
    moments_add_child(m, v.nmom[c], (double)cg.x - (double)gm.x, (double)cg.y - (double)gm.y, (double)cg.z - (double)gm.z);
    if (v.nminidx[c] < mn) mn = v.nminidx[c];
  }
  v.nmom[id] = m;
  v.nminidx[id] = mn;
  NodeRec r = v.nodes[id];
  float len = gm.w;
  if (REFIT) {
    // refitted tree: particles may have left the cell they are filed under.  The cell's size for the opening tests (len2, oc =
    // mass len^4, bmax2) grows to cover them, like the reference's ngb_update_nodes() grows `len` (forcetree.c:2486-2549): a
    // target always opens the cell it is filed under, and a spread-out cell is opened from further away.
    float ext = 0.f;
    for (int k = 0; k < np; k++) {
      const float4 p = v.leaf_posm[ps + k];
      ext = fmaxf(ext, fmaxf(fabsf(p.x - gm.x), fmaxf(fabsf(p.y - gm.y), fabsf(p.z - gm.z))));
    }
    for (int c = id + 1; c < end; c = v.nodes[c].skip) {
      const float4 cg = v.geom[c];
      ext = fmaxf(ext, v.next[c] + fmaxf(fabsf(cg.x - gm.x), fmaxf(fabsf(cg.y - gm.y), fabsf(cg.z - gm.z))));
    }
    v.next[id] = ext;
    if (2.0f * ext > len) len = 2.0f * ext;
  }
  moments_finish(m, gm.x, gm.y, gm.z, len, r);
  v.nodes[id] = r;
}
B200_HD void b5_body(const BuildView &v, int id) { if (v.next) b5_body_t<true>(v, id); else b5_body_t<false>(v, id); }

// ---- B6: position of every node / particle in the reference's next[] chain.
// forcetree.c:274-279 appends a particle at the end of the deepest existing node's chain, so
// inside every node the groups (children) are ordered by the smallest original index they
// contain; recursively that fixes the whole chain.  One thread per node, parents first.
B200_HD void b6_body(const BuildView &v, int id) {
  const int base = (v.nparent[id] < 0) ? v.nstart[id] : v.nlstart[id];   // a root starts its tree's chain
  // gather children: direct particles (size 1, minidx = own index) and child nodes
  int cmin[8], csize[8], cref[8], nc = 0;   // cref >= 0: node id, < 0: ~original particle index
  const int np = v.nnp[id];
  for (int k = 0; k < np; k++) { const int o = v.sidx[v.ndp[8 * id + k]]; cmin[nc] = o; csize[nc] = 1; cref[nc] = ~o; nc++; }
  const int end = v.nodes[id].skip;
  for (int c = id + 1; c < end; c = v.nodes[c].skip) { cmin[nc] = v.nminidx[c]; csize[nc] = v.nend[c] - v.nstart[c] + 1; cref[nc] = c; nc++; }
  for (int i = 1; i < nc; i++) {            // insertion sort by min original index
    const int m = cmin[i], s = csize[i], r = cref[i]; int j = i - 1;
    while (j >= 0 && cmin[j] > m) { cmin[j + 1] = cmin[j]; csize[j + 1] = csize[j]; cref[j + 1] = cref[j]; j--; }
    cmin[j + 1] = m; csize[j + 1] = s; cref[j + 1] = r;
  }
  int off = base;
  for (int i = 0; i < nc; i++) {
    if (cref[i] >= 0) v.nlstart[cref[i]] = off; else v.lrank[~cref[i]] = off;
    off += csize[i];
  }
}

}  // namespace b200
