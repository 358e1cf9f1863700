// sidm.cu - the SIDM scatter step on the GPU.
// Reference: sidm.c:57-627 (sidm), :630-805 (setup_nbr_sidm), :814-968
// (sidm_ensure_neighbours), init.c:431-512 (setup_smoothinglengths_sidm),
// forcetree.c:2163-2297 (ngb_treefind_variable / ngb_treesearch), :2311-2414 (ngb_treefind),
// accel.c:27-132 (compute_accelerations).
//
// The reference runs one sequential loop over a communication buffer.  Inside one bunch that
// loop only READS particle state (P[j].dVel, P[j].Vel); all writes happen afterwards in two
// ordered sweeps (sidm.c:495-537 then :559-601).  So the loop body is data-parallel once each
// buffer slot has its own random number, and the two sweeps become "own kick" followed by
// "partner kick, last slot in buffer order wins" (atomicMax on the slot index).
//
// Kernels: slot assignment (exported-first buffer order) -> pass 1: neighbour count, P_max,
// uniform, early-out (one thread per slot, range search through the gravity octree in
// pre-order with skip pointers) -> pass 2 (only slots that passed): cumulative pair
// probability and partner choice, either in tree order on the fly or, for replay parity, in
// the reference's own list order (tree order + next[] chains + swap-remove filter) ->
// resolve sweeps.  Random numbers: counter-based Philox4x32-10 keyed by (seed, call, particle)
// or the reference's own stream fed per slot (b200_replay).
#include <stdlib.h>
#include <cub/cub.cuh>
#include "ctx.cuh"

namespace b200 {

double s_a_inverse_at(double time);

// search record: the cell bounds c -/+ 0.5*len formed in double exactly as forcetree.c:2252-2276 forms them per test
struct __attribute__((aligned(16))) SearchNode { double lo[3], hi[3]; int skip, pstart, np, pend; };

// compact record for the default (tree-order) search: cell bounds rounded OUTWARD to float, so a
// float test can only err towards "overlaps" (never loses a neighbour); which cells are taken
// wholesale does not change the candidate order (always ascending leaf index) nor the set
// (every candidate still passes the exact float sphere test).  32 bytes instead of 64 halves
// the L1 traffic of the divergent per-lane loads that bound this kernel (ncu: l1tex 83 % busy).
struct __attribute__((aligned(16))) SearchNodeF { float lo[3], hi[3]; int skip; int pinfo; };   // pinfo = pstart<<5 | bucket<<4 | np
constexpr int kBucket = 8;   // subtrees with <= kBucket particles are scanned as one leaf range instead of being descended

struct SidmState {
  SearchNode *snode = nullptr;
  SearchNodeF *snodef = nullptr;
  int *last_active = nullptr; int last_nactive = 0; bool last_all = false;
  int *slot_of_sorted = nullptr;   // processing order of slots (key order)
  int *passlist = nullptr;
  int *logpos = nullptr;
  double *rr = nullptr, *rd = nullptr; size_t replay_cap = 0;   // staged replay arrays
  double *rx = nullptr; int *ro = nullptr; size_t rx_cap = 0, ro_cap = 0;   // type 4 angle trials
  float *dt = nullptr;             // per slot: 2*(time - CurrentTime) (sidm.c:196)
  unsigned char *already = nullptr;
  double *ptot = nullptr;
  int cap_nodes = 0, cap_part = 0;
  // scratch private to the SIDM chain (it may run on its own stream next to the gravity walk,
  // which owns S.x_redo / g.d_t* / S.cub_tmp)
  int *x_redo = nullptr, *x_redo2 = nullptr, *x_want = nullptr, *x_keys = nullptr, *x_keys2 = nullptr, *x_vals = nullptr, *x_shard = nullptr;
  void *cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
  // query groups of the warp-shared search (k_pass1_group): leaf range + tree node of each group
  int2 *groups = nullptr; int *gnode = nullptr, *gflag = nullptr, *gpos = nullptr, *order_leaf = nullptr; int ngroups = 0;
  int *spart = nullptr;            // particle -> slot of the current pass, -1 = not a query (warp-shared search over a large explicit list)
  int *gown = nullptr, *gownflag = nullptr; int nown = 0;   // sharded: the groups this rank searches (compact list: no idle warps in k_pass1_group)
} S;


// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ __forceinline__ uint4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) { philox_round(c0, c1, c2, c3, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ double u01(uint32_t x) { return (double)x / 4294967296.0; }   // same lattice as gsl_rng_uniform

// ------------------------------------------------------------------ range search
struct SearchCtx {
  int qcap;                        // usable entries of the warp-private cell queue (<= kQCap; tests shrink it)
  int M; const SearchNode *snode; const SearchNodeF *snodef; const float4 *leaf_posm; const int *leaf_orig;
  const int *nparent, *leaf_parent, *orig_leaf;
  double box;                      // > 0: periodic box (ngb_periodic(), forcetree.c:1999-2006)
  const float *domain;             // periodic box with one tree: float[6] DomainMin xyz, DomainMax xyz of the particles at the build (see box_interior)
  const float *pad;                // refitted tree (option "tree_reuse"): the cells are as built, the particles have moved by at most
                                   // *pad per coordinate since: every cell test is widened by it (null: a fresh build)
};
__device__ __forceinline__ float search_pad(const SearchCtx &C) { return C.pad ? __ldg(C.pad) : 0.0f; }
// Periodic box, one tree: a search cube that keeps a distance from the faces of the box larger than any particle sticks out of
// it (the reference wraps positions only at a domain decomposition, run.c:135) cannot contain a periodic image: ngb_periodic()
// returns every coordinate difference unchanged, and the plain search - started at the enclosing ancestor cell, float records -
// yields the same candidates in the same order with the same distances.  Only cubes near a face walk from the root with the
// wrapped tests of forcetree.c:2228-2276.  `domain` = DomainMin / DomainMax of the last tree build (null: always wrapped).
__device__ __forceinline__ bool box_interior(const SearchCtx &C, float lox, float loy, float loz, float hix, float hiy, float hiz) {
  if (!C.domain) return false;
  float dom[6];
  for (int k = 0; k < 6; k++) dom[k] = __ldg(C.domain + k);
  return cube_clear_of_faces(dom, C.box, lox, loy, loz, hix, hiy, hiz);      // tree_logic.h (checked on the host: tests/test_hostcheck.py)
}

// ngb_periodic(): float argument, wrapped with double Box / BoxHalf, rounded back to float (tree_logic.h)
__device__ __forceinline__ float ngb_periodic(float x, double box) { return wrap_periodic(x, box); }
__device__ __forceinline__ float dist2_per(float px, float py, float pz, float x, float y, float z, double box) {
  const float dx = ngb_periodic(fadd(px, -x), box), dy = ngb_periodic(fadd(py, -y), box), dz = ngb_periodic(fadd(pz, -z), box);
  return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

__device__ __forceinline__ float dist2_ref(float px, float py, float pz, float x, float y, float z) {
  // forcetree.c:2195-2204: float differences, float products, summed left to right, no FMA
  const float dx = fadd(px, -x), dy = fadd(py, -y), dz = fadd(pz, -z);
  return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

// Smallest ancestor cell of particle `i` that contains the search cube with a safety margin.
// The reference starts every search at the root; all it does above this ancestor is open
// the cells on the path and discard their other children (they cannot reach into the cube:
// the margin covers the <= 42 float roundings of the centre recursion), so starting here
// yields the same candidates in the same order while skipping ~15 levels of the descent.
__device__ __forceinline__ int search_start(const SearchCtx &C, int i, float x, float y, float z, float h) {
  const double lox = (double)fadd(x, -h), loy = (double)fadd(y, -h), loz = (double)fadd(z, -h);
  const double hix = (double)fadd(x, h), hiy = (double)fadd(y, h), hiz = (double)fadd(z, h);
  const double m = 1.0e-5 * (fabs((double)x) + fabs((double)y) + fabs((double)z) + (double)h) + (double)search_pad(C);
  int no = C.leaf_parent[C.orig_leaf[i]];
  if (C.box > 0 && !box_interior(C, (float)lox, (float)loy, (float)loz, (float)hix, (float)hiy, (float)hiz)) {
    while (C.nparent[no] >= 0) no = C.nparent[no];       // periodic, near a face: the root of the particle's own tree
    return no;
  }
  while (C.nparent[no] >= 0) {                 // stop at the root of the particle's tree (one tree per type)
    const SearchNode &nd = C.snode[no];
    if (nd.lo[0] + m <= lox && nd.lo[1] + m <= loy && nd.lo[2] + m <= loz && nd.hi[0] - m >= hix && nd.hi[1] - m >= hiy && nd.hi[2] - m >= hiz) break;
    no = C.nparent[no];
  }
  return no;
}

// calls f(leaf_slot, pos, r2, bulk, node) for every candidate the reference's ngb_treesearch()
// would append (forcetree.c:2224-2297), in its order, visiting the subtree of `start` in
// pre-order with skip pointers; cell tests in double on float operands like the reference.
// One query per thread (a warp-lockstep variant like the gravity walk was measured slower: the
// 32 cubes of a warp are small compared with the cells between them, so the union of their
// paths is ~10x one path).  `valid` lets padding lanes fall through.
template <class F>
__device__ __forceinline__ void range_search(const SearchCtx &C, bool valid, int start, float x, float y, float z, float h, F &&f) {
  if (!valid) return;
  const float lox = fadd(x, -h), loy = fadd(y, -h), loz = fadd(z, -h);
  const float hix = fadd(x, h), hiy = fadd(y, h), hiz = fadd(z, h);
  if (C.box > 0) {
    // periodic variant of the same walk (forcetree.c:2228-2276 under PERIODIC): everything is
    // measured relative to the search centre through ngb_periodic(); always from the root
    const double box = C.box;
    const double sminx = fadd(lox, -x), sminy = fadd(loy, -y), sminz = fadd(loz, -z);
    const double smaxx = fadd(hix, -x), smaxy = fadd(hiy, -y), smaxz = fadd(hiz, -z);
    int no = start;                                    // the root of the query's own tree
    const int stop_p = C.snode[start].skip;
    while (no < stop_p) {
      const double2 *q = reinterpret_cast<const double2 *>(C.snode + no);
      const double2 a0 = __ldg(q), a1 = __ldg(q + 1), a2 = __ldg(q + 2);
      const int4 info = __ldg(reinterpret_cast<const int4 *>(q + 3));
      const float cx = (float)(0.5 * (a0.x + a1.y)), cy = (float)(0.5 * (a0.y + a2.x)), cz = (float)(0.5 * (a1.x + a2.y));
      const double half = 0.5 * (a1.y - a0.x);
      const double px = ngb_periodic(fadd(cx, -x), box), py = ngb_periodic(fadd(cy, -y), box), pz = ngb_periodic(fadd(cz, -z), box);
      if (px + half < sminx || px - half > smaxx || py + half < sminy || py - half > smaxy || pz + half < sminz || pz - half > smaxz) { no = info.x; continue; }
      const bool inside = !(px + half > smaxx) && !(px - half < sminx) && !(py + half > smaxy) && !(py - half < sminy) && !(pz + half > smaxz) && !(pz - half < sminz);
      if (inside) {
        for (int L = info.y; L < info.w; L++) { const float4 p = __ldg(C.leaf_posm + L); f(L, p, dist2_per(p.x, p.y, p.z, x, y, z, box), true, no); }
        no = info.x;
      } else {
        for (int k = 0; k < info.z; k++) {
          const int L = info.y + k;
          const float4 p = __ldg(C.leaf_posm + L);
          const double ex = ngb_periodic(fadd(p.x, -x), box), ey = ngb_periodic(fadd(p.y, -y), box), ez = ngb_periodic(fadd(p.z, -z), box);
          if (ex < sminx || ex > smaxx || ey < sminy || ey > smaxy || ez < sminz || ez > smaxz) continue;
          f(L, p, dist2_per(p.x, p.y, p.z, x, y, z, box), false, no);
        }
        no = no + 1;
      }
    }
    return;
  }
  const double pd = (double)search_pad(C);
  const double dlx = (double)lox - pd, dly = (double)loy - pd, dlz = (double)loz - pd, dhx = (double)hix + pd, dhy = (double)hiy + pd, dhz = (double)hiz + pd;
  // Flattened walk: every trip of the loop does ONE thing for this lane - test one pending
  // particle, or test one cell - so that the lanes of a warp, which are at different places of
  // different subtrees, diverge two ways per trip instead of nesting loops of different lengths
  // (measured: 5.8 of 32 lanes active with the nested form).  Order of the callbacks is unchanged:
  // a cell's own particles (or its whole leaf range when fully inside) before its child cells.
  int no = start;
  const int stop = C.snode[start].skip;
  int pk = 0, pe = 0, pnode = 0; bool bulk = false;
  for (;;) {
    if (pk < pe) {
      const int L = pk++;
      const float4 p = __ldg(C.leaf_posm + L);
      if (bulk || !(p.x < lox || p.x > hix || p.y < loy || p.y > hiy || p.z < loz || p.z > hiz))
        f(L, p, dist2_ref(p.x, p.y, p.z, x, y, z), bulk, pnode);
      continue;
    }
    if (no >= stop) break;
    const double2 *q = reinterpret_cast<const double2 *>(C.snode + no);
    const double2 a0 = __ldg(q), a1 = __ldg(q + 1), a2 = __ldg(q + 2);      // lo.xy | lo.z hi.x | hi.yz
    const int4 info = __ldg(reinterpret_cast<const int4 *>(q + 3));         // skip, pstart, np, pend
    if (a1.y < dlx || a0.x > dhx || a2.x < dly || a0.y > dhy || a2.y < dlz || a1.x > dhz) { no = info.x; continue; }
    bulk = !(a1.y > dhx) && !(a0.x < dlx) && !(a2.x > dhy) && !(a0.y < dly) && !(a2.y > dhz) && !(a1.x < dlz);
    pnode = no; pk = info.y;
    if (bulk) { pe = info.w; no = info.x; } else { pe = info.y + info.z; no = no + 1; }
  }
}

// default search: same flattened pre-order walk on the compact float records (see SearchNodeF)
template <class F>
__device__ __forceinline__ void range_search_fast(const SearchCtx &C, bool valid, int start, float x, float y, float z, float h, F &&f) {
  if (!valid) return;
  const float lox = fadd(x, -h), loy = fadd(y, -h), loz = fadd(z, -h);
  const float hix = fadd(x, h), hiy = fadd(y, h), hiz = fadd(z, h);
  if (C.box > 0 && !box_interior(C, lox, loy, loz, hix, hiy, hiz)) { range_search(C, valid, start, x, y, z, h, f); return; }
  // cell tests: the cube widened by the refit displacement bound (0 after a fresh build); "taken whole" needs the cell inside the
  // cube even when its particles have moved by that much
  const float pd = search_pad(C);
  const float clx = lox - pd, cly = loy - pd, clz = loz - pd, chx = hix + pd, chy = hiy + pd, chz = hiz + pd;
  const float wlx = lox + pd, wly = loy + pd, wlz = loz + pd, whx = hix - pd, why = hiy - pd, whz = hiz - pd;
  int no = start;
  const int stop = C.snodef[start].skip;
  int pk = 0, pe = 0, pnode = 0; bool bulk = false;
  for (;;) {
    if (pk < pe) {
      const int L = pk++;
      const float4 p = __ldg(C.leaf_posm + L);
      if (bulk || !(p.x < lox || p.x > hix || p.y < loy || p.y > hiy || p.z < loz || p.z > hiz))
        f(L, p, dist2_ref(p.x, p.y, p.z, x, y, z), bulk, pnode);
      continue;
    }
    if (no >= stop) break;
    const float4 *q = reinterpret_cast<const float4 *>(C.snodef + no);
    const float4 a = __ldg(q), b = __ldg(q + 1);                  // lo.xyz hi.x | hi.yz skip pinfo
    const int skip = __float_as_int(b.z), pinfo = __float_as_int(b.w);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(C.snodef + skip));     // the jump target, if this cell is skipped or taken whole
    if (a.w < clx || a.x > chx || b.x < cly || a.y > chy || b.y < clz || a.z > chz) { no = skip; continue; }
    bulk = (a.x >= wlx) && (a.w <= whx) && (a.y >= wly) && (b.x <= why) && (a.z >= wlz) && (b.y <= whz);
    pnode = no; pk = pinfo >> 5;
    if (bulk) { pe = C.snodef[skip].pinfo >> 5; no = skip; } else { pe = pk + (pinfo & 15); no = (pinfo & 16) ? skip : no + 1; }
  }
}

__global__ void k_search_nodes(int m, const NodeRec *nodes, const float4 *geom, const int *npstart, const unsigned char *nnp, const int *nparent, SearchNode *out, SearchNodeF *outf) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= m) return;
  const float4 gm = geom[id]; const int skip = nodes[id].skip;
  const double half = 0.5 * (double)gm.w;
  SearchNode s;
  s.lo[0] = (double)gm.x - half; s.lo[1] = (double)gm.y - half; s.lo[2] = (double)gm.z - half;
  s.hi[0] = (double)gm.x + half; s.hi[1] = (double)gm.y + half; s.hi[2] = (double)gm.z + half;
  s.skip = skip; s.pstart = npstart[id]; s.np = nnp[id]; s.pend = npstart[skip];
  out[id] = s;
  SearchNodeF t;
  for (int k = 0; k < 3; k++) {
    float l = (float)s.lo[k], h = (float)s.hi[k];
    if ((double)l > s.lo[k]) l = nextafterf(l, -INFINITY);
    if ((double)h < s.hi[k]) h = nextafterf(h, INFINITY);
    t.lo[k] = l; t.hi[k] = h;
  }
  const int cnt = s.pend - s.pstart;
  const int bucket = (cnt <= kBucket && nparent[id] >= 0) ? 1 : 0;     // roots are never buckets
  t.skip = skip; t.pinfo = (s.pstart << 5) | (bucket << 4) | (bucket ? cnt : s.np);
  outf[id] = t;
  if (id == m - 1) { SearchNodeF e; for (int k = 0; k < 3; k++) { e.lo[k] = 0; e.hi[k] = 0; } e.skip = m; e.pinfo = npstart[m] << 5; outf[m] = e; }
}

// ------------------------------------------------------------------ slots
// sidm.c:141-161: particles within h of the domain box are flagged for export and placed first
__global__ void k_export_flag(int na, const int *active, const float4 *posm, const float4 *velh, const float *domain_all, int *flag, const int *ptype) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= na) return;
  const int i = active ? active[a] : a;
  const float *domain = ptype ? domain_all + 6 * (ptype[i] & 7) : domain_all;     // DomainMin/Max[type], forcetree.c:192-198
  const float4 p = posm[i]; const float h = velh[i].w;
  const float pp[3] = {p.x, p.y, p.z};
  int j;
  for (j = 0; j < 3; j++) {
    if (pp[j] < fadd(domain[j], h)) break;
    if (pp[j] > fadd(domain[3 + j], -h)) break;
  }
  flag[a] = (j != 3);
}
__global__ void k_assign_slots(int na, const int *active, const int *flag, const int *scan, int *slot_part, int *slot_of_active,
                               const float *curtime, const float *dvel, double time, float *dt, unsigned char *already,
                               const int *krank, int *keys, int *vals, int *flags_out, int *partner, float *dv, double *prob, double *ptot,
                               int *slot_of_part) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= na) return;
  partner[a] = -1; dv[3 * (size_t)a] = dv[3 * (size_t)a + 1] = dv[3 * (size_t)a + 2] = 0; prob[a] = 0; ptot[a] = 0;   // slot a's results
  const int nexport = scan[na];
  const int i = active ? active[a] : a;
  const int place = flag[a] ? scan[a] : nexport + (a - scan[a]);
  slot_part[place] = i;
  slot_of_active[a] = place;
  if (slot_of_part) slot_of_part[i] = place;       // large explicit list on the warp-shared search: the array was filled with -1
  dt[place] = (float)(2 * (time - (double)curtime[i]));             // sidm.c:196
  already[place] = dvel[3 * (size_t)i] != 0.0f;                    // sidm.c:189-192 (ID = 0)
  if (active) { keys[a] = krank[i]; vals[a] = place; }
  else vals[krank[i]] = place;          // every particle active: the key order is the tree build's own order
  if (a == 0) flags_out[FL_NEXPORT] = nexport;
}

// ------------------------------------------------------------------ pass 1
struct Pass1 {
  int ns; const int *order; const int *slot_part; SearchCtx C;
  const float4 *posm, *velh; const float *dt; const unsigned char *already;
  const double *replay_rand; double C_Pmax, s_a_inverse; uint32_t k0, k1;
  int *ngb; double *pmax, *rnd; int *pass; int count_only; unsigned long long *ctr;
};
__global__ void __launch_bounds__(128, 16) k_pass1(Pass1 P) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < P.ns;
  const int s = valid ? P.order[t] : 0;
  const int i = valid ? P.slot_part[s] : 0;
  const float4 p = valid ? P.posm[i] : make_float4(0, 0, 0, 0);
  const float h = valid ? P.velh[i].w : 0.0f;
  const float sr2 = fmul(h, h);
  int cnt = 0, cand = 0;
  const int start = valid ? search_start(P.C, i, p.x, p.y, p.z, h) : 0;
  range_search_fast(P.C, valid, start, p.x, p.y, p.z, h, [&](int, const float4 &, float r2, bool, int) { cand++; if (r2 < sr2) cnt++; });
  unsigned long long wc = cand;
  for (int o = 16; o > 0; o >>= 1) wc += __shfl_down_sync(0xffffffffu, wc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&P.ctr[CT_CAND], wc);
  if (!valid) return;
  P.ngb[s] = cnt;
  if (P.count_only) return;
  const double dt_h0 = (double)P.dt[s] * P.s_a_inverse;
  const double hh = 1.0 * (double)h, hinv = 1.0 / hh, hinv3 = hinv * hinv * hinv;
  const double pm = P.C_Pmax * (double)p.w * hinv3 * dt_h0;          // sidm.c:338
  double r;
  if (P.replay_rand) r = P.replay_rand[s];
  else r = u01(philox((uint32_t)i, 0u, 0u, 0u, P.k0, P.k1).x);
  P.pmax[s] = pm; P.rnd[s] = r;
  const int pass = !(pm < r) && !P.already[s];                        // sidm.c:343-346
  P.pass[t] = pass;
}

// ------------------------------------------------------------------ pass 1, warp-shared search
// When every particle is a query (all-active steps, start-up) the queries are grouped by tree
// node: a group = the particles of a subtree with <= 32 particles whose parent holds more (or the
// direct particles of a bigger cell), i.e. one compact cell, contiguous in leaf order.  One warp
// per group: lane = query.  The warp walks the tree ONCE for the union of its search cubes
// (warp-uniform pre-order walk, broadcast loads) and every lane tests every candidate of the
// overlapping leaf cells against its own sphere with the reference's float test
// (forcetree.c:2195-2206).  The neighbour SET of a query does not depend on how the candidates
// were found, so the counts equal those of the per-query walk (k_pass1) bit for bit; what changes
// is the cost: ~10x fewer instructions, coalesced candidate loads instead of divergent ones.
constexpr int kGroupCell = 512;   // query groups are cut from subtrees with at most this many particles
// A group is a run of <= 32 consecutive leaf slots inside one compact cell: a subtree with <=
// kGroupCell particles whose parent holds more, or the direct particles of a bigger cell.  On one GPU
// a cell's leaf range is cut into equal runs; sharded over several GPUs it is cut at the multiples of
// 32 of the leaf order, the blocks the leaf order is dealt out in (b200_set_shard).
__device__ __forceinline__ int2 group_range(const SearchNode &nd) {      // leaf range the node contributes
  const int cnt = nd.pend - nd.pstart;
  return cnt <= kGroupCell ? make_int2(nd.pstart, nd.pend) : make_int2(nd.pstart, nd.pstart + nd.np);
}
// aligned (sharded runs): cut at the multiples of 32 of the leaf order; else: ceil(len/32) equal runs (better filled warps)
__global__ void k_group_flag(int m, const SearchNode *sn, const int *nparent, int *flag, int aligned) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id > m) return;
  if (id == m) { flag[m] = 0; return; }
  const int cnt = sn[id].pend - sn[id].pstart;
  const int2 r = group_range(sn[id]);
  int f = 0;
  if (r.y > r.x) f = aligned ? ((r.y - 1) >> 5) - (r.x >> 5) + 1 : (r.y - r.x + 31) >> 5;
  if (cnt <= kGroupCell && nparent[id] >= 0) { const int par = nparent[id]; if (sn[par].pend - sn[par].pstart <= kGroupCell) f = 0; }
  flag[id] = f;
}
__global__ void k_group_emit(int m, const SearchNode *sn, const int *flag, const int *pos, int2 *groups, int *gnode, int aligned) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= m || !flag[id]) return;
  const int2 r = group_range(sn[id]);
  const int c = flag[id], g0 = pos[id], b0 = r.x >> 5;
  if (aligned) {
    for (int j = 0; j < c; j++) {
      const int lo = max(r.x, (b0 + j) << 5), hi = min(r.y, (b0 + j + 1) << 5);
      groups[g0 + j] = make_int2(lo, hi - lo); gnode[g0 + j] = id;
    }
  } else {
    const int len = r.y - r.x, base = len / c, rem = len % c;
    int at = r.x;
    for (int j = 0; j < c; j++) { const int nj = base + (j < rem); groups[g0 + j] = make_int2(at, nj); gnode[g0 + j] = id; at += nj; }
  }
}
// sharded runs: the global processing order of the slots = leaf order
// sharded runs: the groups of the 32-blocks b of the leaf order with b % world == rank
__global__ void k_group_own(int ng, const int2 *groups, int world, int rank, int *flag) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < ng) flag[w] = ((groups[w].x >> 5) % world) == rank;
}
__global__ void k_order_leaf(int n, const int *leaf_orig, const int *slot_of_part, int *order_leaf) {
  const int L = blockIdx.x * blockDim.x + threadIdx.x;
  if (L < n) order_leaf[L] = slot_of_part[leaf_orig[L]];
}

__device__ __forceinline__ int f2ord(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }   // monotone float -> int
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

struct Pass1G {
  int ng; const int *own; const int2 *groups; const int *gnode; SearchCtx C; const float4 *velh; const int *slot_of_part;
  const float *dt; const unsigned char *already; const double *replay_rand; double C_Pmax, s_a_inverse; uint32_t k0, k1;
  int *ngb; double *pmax, *rnd; int *pass; int *order_leaf; int count_only; unsigned long long *ctr;
  int rank, world;     // sharded: this rank handles the groups of the 32-blocks b with b % world == rank
  int masked;          // slot_of_part holds -1 for the particles that are not queries of this pass (large explicit lists, one rank)
};
constexpr int kGroupTiny = 4;
constexpr int kWarpQueryMax = 1 << 14;   // up to this many queries a pass is served one warp per query (beyond, the
                                         // extra work of 32 lanes per query costs more than the shorter tail saves)
constexpr int kQCap = 320;         // warp-private cell queue of the lane-parallel walk

// packed fp32 pairs (sm_100a add.rn.f32x2 = SASS FADD2): two candidates per instruction
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 pk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 bc2(float a) { return pk2(a, a); }
__device__ __forceinline__ void up2(f2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
// squares as two scalar __fmul_rn: ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even
// with --fmad=false, which would break the FMA-free neighbour test; scalar .rn products are left alone
__device__ __forceinline__ f2 sq2(f2 a) { float x, y; up2(a, x, y); return pk2(__fmul_rn(x, x), __fmul_rn(y, y)); }

struct Cube { float lx, ly, lz, hx, hy, hz; };
// warp-uniform pre-order walk (every lane visits every cell): fallback of the lane-parallel walk
__device__ __noinline__ void group_walk_uniform(const SearchCtx &C, const Cube &U, int A, int stop, const float4 &p, float sr2, int &cnt, unsigned &cand) {
  int no = A;
  while (no < stop) {
    const float4 *q = reinterpret_cast<const float4 *>(C.snodef + no);
    const float4 a = __ldg(q), b = __ldg(q + 1);           // lo.xyz hi.x | hi.yz skip pinfo
    const int skip = __float_as_int(b.z), pinfo = __float_as_int(b.w);
    if (a.w < U.lx || a.x > U.hx || b.x < U.ly || a.y > U.hy || b.y < U.lz || a.z > U.hz) { no = skip; continue; }
    const bool whole = (pinfo & 16) || ((a.x >= U.lx) && (a.w <= U.hx) && (a.y >= U.ly) && (b.x <= U.hy) && (a.z >= U.lz) && (b.y <= U.hz));
    const int pk = pinfo >> 5;
    int pe;
    if (whole) { pe = (pinfo & 16) ? pk + (pinfo & 15) : (C.snodef[skip].pinfo >> 5); no = skip; }
    else { pe = pk + (pinfo & 15); no = no + 1; }
    cand += (unsigned)(pe - pk);
    for (int k = pk; k < pe; k++) { const float4 c0 = __ldg(C.leaf_posm + k); cnt += dist2_ref(c0.x, c0.y, c0.z, p.x, p.y, p.z) < sr2; }
  }
}
__device__ __forceinline__ void pass1_finish(const Pass1G &P, int L, int i, const float4 &p, float h, int cnt) {
  const int s = P.slot_of_part[i];
  P.ngb[s] = cnt;
  // position of this query in the (rank's) processing order: leaf order, 32-blocks dealt round robin
  const int k = P.world > 1 ? (((L >> 5) / P.world) << 5) + (L & 31) : L;
  if (P.world == 1) P.order_leaf[L] = s;
  if (P.count_only) { P.pass[k] = 0; return; }
  const double dt_h0 = (double)P.dt[s] * P.s_a_inverse;
  const double hh = 1.0 * (double)h, hinv = 1.0 / hh, hinv3 = hinv * hinv * hinv;
  const double pm = P.C_Pmax * (double)p.w * hinv3 * dt_h0;          // sidm.c:338
  double r;
  if (P.replay_rand) r = P.replay_rand[s];
  else r = u01(philox((uint32_t)i, 0u, 0u, 0u, P.k0, P.k1).x);
  P.pmax[s] = pm; P.rnd[s] = r;
  P.pass[k] = !(pm < r) && !P.already[s];                             // sidm.c:343-346
}
__global__ void __launch_bounds__(128) k_pass1_group(Pass1G P) {
  const int wi = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (wi >= P.ng) return;                                  // warp-uniform
  const int w = P.own ? P.own[wi] : wi;                    // sharded: this rank's groups only
  const int2 gr = P.groups[w];
  bool valid = lane < gr.y;
  const int L = gr.x + (valid ? lane : 0);
  const float4 p = P.C.leaf_posm[L];
  const int i = P.C.leaf_orig[L];
  if (P.masked) {
    // a large explicit list (the repair passes of a step in which most particles left the neighbour window): the members of the
    // group that are not queries sit out; their places in the leaf-ordered pass flags are cleared here
    if (valid && P.slot_of_part[i] < 0) { valid = false; P.pass[L] = 0; P.order_leaf[L] = 0; }
    if (!__any_sync(0xffffffffu, valid)) return;
  }
  const float h = valid ? P.velh[i].w : 0.0f;
  const float sr2 = fmul(h, h);
  // a few stray particles (direct particles of a big cell, far from each other): per-query tree walks, as k_pass1.  Short runs
  // inside a small cell (the 32-aligned cuts of sharded runs leave many) stay on the warp-shared search: their cubes overlap.
  const SearchNode &gn = P.C.snode[P.gnode[w]];
  auto each_lane_alone = [&]() {
    int cnt = 0, cand = 0;
    const int start = valid ? search_start(P.C, i, p.x, p.y, p.z, h) : 0;
    range_search_fast(P.C, valid, start, p.x, p.y, p.z, h, [&](int, const float4 &, float r2, bool, int) { cand++; if (r2 < sr2) cnt++; });
    if (valid) { atomicAdd(&P.ctr[CT_CAND], (unsigned long long)cand); pass1_finish(P, L, i, p, h, cnt); }
  };
  if (gr.y <= kGroupTiny && gn.pend - gn.pstart > kGroupCell) { each_lane_alone(); return; }
  // union of the group's search cubes
  const float big = 3.0e38f;
  Cube U;
  U.lx = ord2f(__reduce_min_sync(0xffffffffu, f2ord(valid ? fadd(p.x, -h) : big)));
  U.ly = ord2f(__reduce_min_sync(0xffffffffu, f2ord(valid ? fadd(p.y, -h) : big)));
  U.lz = ord2f(__reduce_min_sync(0xffffffffu, f2ord(valid ? fadd(p.z, -h) : big)));
  U.hx = ord2f(__reduce_max_sync(0xffffffffu, f2ord(valid ? fadd(p.x, h) : -big)));
  U.hy = ord2f(__reduce_max_sync(0xffffffffu, f2ord(valid ? fadd(p.y, h) : -big)));
  U.hz = ord2f(__reduce_max_sync(0xffffffffu, f2ord(valid ? fadd(p.z, h) : -big)));
  // periodic box: a group whose union cube comes near a face is searched lane by lane (wrapped tests from the root where
  // needed, forcetree.c:2228-2276); everywhere else no periodic image can be a neighbour and the shared search applies as it is
  if (P.C.box > 0 && !box_interior(P.C, U.lx, U.ly, U.lz, U.hx, U.hy, U.hz)) { each_lane_alone(); return; }
  { const float pd = search_pad(P.C); U.lx -= pd; U.ly -= pd; U.lz -= pd; U.hx += pd; U.hy += pd; U.hz += pd; }   // refitted tree: cells as built
  // smallest ancestor cell that contains the union with a safety margin (cf. search_start)
  int A = P.gnode[w];
  {
    const double mg = 1.0e-5 * (fmax(fabs((double)U.lx), fabs((double)U.hx)) + fmax(fabs((double)U.ly), fabs((double)U.hy)) +
                                fmax(fabs((double)U.lz), fabs((double)U.hz)) + ((double)U.hx - (double)U.lx));
    while (P.C.nparent[A] >= 0) {
      const SearchNode &nd = P.C.snode[A];
      if (nd.lo[0] + mg <= (double)U.lx && nd.lo[1] + mg <= (double)U.ly && nd.lo[2] + mg <= (double)U.lz &&
          nd.hi[0] - mg >= (double)U.hx && nd.hi[1] - mg >= (double)U.hy && nd.hi[2] - mg >= (double)U.hz) break;
      A = P.C.nparent[A];
    }
  }
  const int stopA = P.C.snodef[A].skip;

  // Lane-parallel walk: a warp-private LIFO of (cell, end of its sibling chain).  Each trip every
  // lane pops one cell, tests it against the union cube, pushes the next sibling and - if the cell
  // is opened - its first child, and yields the leaf range whose particles have to be tested (the
  // cell's own particles, or the whole subtree of a cell that lies inside the cube or is a bucket).
  // The ranges of a trip are flattened over the lanes: 32 candidates are fetched at once (coalesced
  // within a range), staged in shared memory, and every lane tests all of them against its own
  // sphere two at a time (packed add.rn.f32x2 for the differences and sums, scalar products: every
  // operation individually rounded, i.e. exactly the FMA-free float test of forcetree.c:2195-2206).
  __shared__ int2 s_q[4][kQCap];
  __shared__ __align__(8) float s_cx[4][32], s_cy[4][32], s_cz[4][32];
  const int wib = threadIdx.x >> 5;
  int2 *q = s_q[wib]; float *cx = s_cx[wib], *cy = s_cy[wib], *cz = s_cz[wib];
  const unsigned lt = (1u << lane) - 1u;
  const f2 nqx = bc2(-p.x), nqy = bc2(-p.y), nqz = bc2(-p.z);
  int cnt = 0; unsigned cand = 0;
  int qn = 1;
  if (lane == 0) q[0] = make_int2(A, stopA);
  __syncwarp();
  bool overflow = false;
  while (qn > 0) {
    const int take = qn < 32 ? qn : 32, base = qn - take;
    const int2 it = lane < take ? q[base + lane] : make_int2(-1, -1);
    __syncwarp();
    qn = base;
    int pk = 0, pe = 0, sib = -1, child = -1, cstop = 0;
    if (lane < take) {
      const float4 *rq = reinterpret_cast<const float4 *>(P.C.snodef + it.x);
      const float4 a = __ldg(rq), b = __ldg(rq + 1);       // lo.xyz hi.x | hi.yz skip pinfo
      const int skip = __float_as_int(b.z), pinfo = __float_as_int(b.w);
      if (skip < it.y) sib = skip;
      if (!(a.w < U.lx || a.x > U.hx || b.x < U.ly || a.y > U.hy || b.y < U.lz || a.z > U.hz)) {
        const bool whole = (pinfo & 16) || ((a.x >= U.lx) && (a.w <= U.hx) && (a.y >= U.ly) && (b.x <= U.hy) && (a.z >= U.lz) && (b.y <= U.hz));
        pk = pinfo >> 5;
        if (whole) pe = (pinfo & 16) ? pk + (pinfo & 15) : (P.C.snodef[skip].pinfo >> 5);
        else { pe = pk + (pinfo & 15); if (it.x + 1 < skip) { child = it.x + 1; cstop = skip; } }
      }
    }
    const unsigned m1 = __ballot_sync(0xffffffffu, sib >= 0), m2 = __ballot_sync(0xffffffffu, child >= 0);
    const int n1 = __popc(m1), n2 = __popc(m2);
    if (qn + n1 + n2 > P.C.qcap) { overflow = true; break; }
    if (sib >= 0) q[qn + __popc(m1 & lt)] = make_int2(sib, it.y);
    if (child >= 0) q[qn + n1 + __popc(m2 & lt)] = make_int2(child, cstop);
    qn += n1 + n2;
    __syncwarp();
    // flatten this trip's leaf ranges over the lanes
    const int len = pe - pk;
    int inc = len;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    const int T = __shfl_sync(0xffffffffu, inc, 31), exc = inc - len;
    cand += (unsigned)T;
    for (int c0 = 0; c0 < T; c0 += 32) {
      const int idx = c0 + lane;
      int r = 0;                                            // the last lane whose range starts at or before idx
      for (int step = 16; step > 0; step >>= 1) {
        const int v = __shfl_sync(0xffffffffu, exc, (r + step) & 31);
        if (r + step < 32 && v <= idx) r += step;
      }
      const int pk_r = __shfl_sync(0xffffffffu, pk, r), exc_r = __shfl_sync(0xffffffffu, exc, r);
      float4 c = make_float4(big, big, big, 0.f);          // padding fails every sphere test (r2 = inf)
      if (idx < T) c = __ldg(P.C.leaf_posm + pk_r + (idx - exc_r));
      cx[lane] = c.x; cy[lane] = c.y; cz[lane] = c.z;
      __syncwarp();
      const int nv = (T - c0 < 32) ? T - c0 : 32, npair = (nv + 1) >> 1;
      const unsigned long long *px = reinterpret_cast<const unsigned long long *>(cx), *py = reinterpret_cast<const unsigned long long *>(cy),
                               *pz = reinterpret_cast<const unsigned long long *>(cz);
#pragma unroll 4
      for (int k = 0; k < npair; k++) {
        f2 X, Y, Z; X.v = px[k]; Y.v = py[k]; Z.v = pz[k];
        const f2 dx = add2(X, nqx), dy = add2(Y, nqy), dz = add2(Z, nqz);
        float ra, rb;
        up2(add2(add2(sq2(dx), sq2(dy)), sq2(dz)), ra, rb);
        cnt += (ra < sr2) + (rb < sr2);
      }
      __syncwarp();
    }
  }
  if (overflow) {                                           // queue full (never seen in practice): plain uniform walk
    cnt = 0; cand = 0;
    group_walk_uniform(P.C, U, A, stopA, p, sr2, cnt, cand);
  }
  const int nq = P.masked ? __popc(__ballot_sync(0xffffffffu, valid)) : gr.y;
  if (lane == 0) atomicAdd(&P.ctr[CT_CAND], (unsigned long long)cand * (unsigned)nq);
  if (valid) pass1_finish(P, L, i, p, h, cnt);
}

// ------------------------------------------------------------------ pass 1, one warp per query
// Small query sets (the repair passes of sidm_ensure_neighbours, individual-time-step active lists):
// with one thread per query the launch lasts as long as its slowest query - a cube that straddles a
// high-level cell boundary walks ~1000 cells one after the other (0.6 ms, whether the pass has 10^5
// queries or one).  Here a warp serves one query with the same lane-parallel cell walk as
// k_pass1_group and tests 32 candidates per trip, one per lane: the dependent chain shrinks to a few
// dozen trips.  Same neighbour set, same counts.
template <bool PERIODIC>                                      // the open-boundary instantiation carries no wrapped-search code (40 registers)
__global__ void __launch_bounds__(128) k_pass1_warp(Pass1 P) {
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= P.ns) return;                                   // warp-uniform
  const int s = P.order[w];
  const int i = P.slot_part[s];
  const float4 p = P.posm[i];
  const float h = P.velh[i].w;
  const float sr2 = fmul(h, h);
  Cube U;
  U.lx = fadd(p.x, -h); U.ly = fadd(p.y, -h); U.lz = fadd(p.z, -h);
  U.hx = fadd(p.x, h); U.hy = fadd(p.y, h); U.hz = fadd(p.z, h);
  { const float pd = search_pad(P.C); U.lx -= pd; U.ly -= pd; U.lz -= pd; U.hx += pd; U.hy += pd; U.hz += pd; }   // refitted tree: cells as built
  const int A = search_start(P.C, i, p.x, p.y, p.z, h);
  const int stopA = P.C.snodef[A].skip;
  __shared__ int2 s_q[4][kQCap];
  int2 *q = s_q[threadIdx.x >> 5];
  const unsigned lt = (1u << lane) - 1u;
  int cnt = 0; unsigned cand = 0;
  // periodic box, cube near a face (A is the root then): the wrapped walk of forcetree.c:2228-2276 by one lane; the few
  // queries this concerns do not pay for a lane-parallel form of the wrapped tests
  const bool wrapped = PERIODIC && P.C.box > 0 && !box_interior(P.C, U.lx, U.ly, U.lz, U.hx, U.hy, U.hz);
  if (PERIODIC) {
    if (wrapped && lane == 0)
      range_search(P.C, true, A, p.x, p.y, p.z, h, [&](int, const float4 &, float r2, bool, int) { cand++; if (r2 < sr2) cnt++; });
  }
  int qn = wrapped ? 0 : 1;
  if (lane == 0) q[0] = make_int2(A, stopA);
  __syncwarp();
  bool overflow = false;
  while (qn > 0) {
    const int take = qn < 32 ? qn : 32, base = qn - take;
    const int2 it = lane < take ? q[base + lane] : make_int2(-1, -1);
    __syncwarp();
    qn = base;
    int pk = 0, pe = 0, sib = -1, child = -1, cstop = 0;
    if (lane < take) {
      const float4 *rq = reinterpret_cast<const float4 *>(P.C.snodef + it.x);
      const float4 a = __ldg(rq), b = __ldg(rq + 1);       // lo.xyz hi.x | hi.yz skip pinfo
      const int skip = __float_as_int(b.z), pinfo = __float_as_int(b.w);
      if (skip < it.y) sib = skip;
      if (!(a.w < U.lx || a.x > U.hx || b.x < U.ly || a.y > U.hy || b.y < U.lz || a.z > U.hz)) {
        const bool whole = (pinfo & 16) || ((a.x >= U.lx) && (a.w <= U.hx) && (a.y >= U.ly) && (b.x <= U.hy) && (a.z >= U.lz) && (b.y <= U.hz));
        pk = pinfo >> 5;
        if (whole) pe = (pinfo & 16) ? pk + (pinfo & 15) : (P.C.snodef[skip].pinfo >> 5);
        else { pe = pk + (pinfo & 15); if (it.x + 1 < skip) { child = it.x + 1; cstop = skip; } }
      }
    }
    const unsigned m1 = __ballot_sync(0xffffffffu, sib >= 0), m2 = __ballot_sync(0xffffffffu, child >= 0);
    const int n1 = __popc(m1), n2 = __popc(m2);
    if (qn + n1 + n2 > P.C.qcap) { overflow = true; break; }
    if (sib >= 0) q[qn + __popc(m1 & lt)] = make_int2(sib, it.y);
    if (child >= 0) q[qn + n1 + __popc(m2 & lt)] = make_int2(child, cstop);
    qn += n1 + n2;
    __syncwarp();
    const int len = pe - pk;
    int inc = len;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    const int T = __shfl_sync(0xffffffffu, inc, 31), exc = inc - len;
    cand += (unsigned)T;
    for (int c0 = 0; c0 < T; c0 += 32) {                   // one candidate per lane
      const int idx = c0 + lane;
      int r = 0;
      for (int step = 16; step > 0; step >>= 1) {
        const int v = __shfl_sync(0xffffffffu, exc, (r + step) & 31);
        if (r + step < 32 && v <= idx) r += step;
      }
      const int pk_r = __shfl_sync(0xffffffffu, pk, r), exc_r = __shfl_sync(0xffffffffu, exc, r);
      bool in = false;
      if (idx < T) { const float4 c = __ldg(P.C.leaf_posm + pk_r + (idx - exc_r)); in = dist2_ref(c.x, c.y, c.z, p.x, p.y, p.z) < sr2; }
      cnt += __popc(__ballot_sync(0xffffffffu, in));
    }
  }
  if (overflow) { cnt = 0; cand = 0; group_walk_uniform(P.C, U, A, stopA, p, sr2, cnt, cand); }
  if (lane != 0) return;
  atomicAdd(&P.ctr[CT_CAND], (unsigned long long)cand);
  P.ngb[s] = cnt;
  if (P.count_only) return;
  const double dt_h0 = (double)P.dt[s] * P.s_a_inverse;
  const double hh = 1.0 * (double)h, hinv = 1.0 / hh, hinv3 = hinv * hinv * hinv;
  const double pm = P.C_Pmax * (double)p.w * hinv3 * dt_h0;          // sidm.c:338
  double r;
  if (P.replay_rand) r = P.replay_rand[s];
  else r = u01(philox((uint32_t)i, 0u, 0u, 0u, P.k0, P.k1).x);
  P.pmax[s] = pm; P.rnd[s] = r;
  P.pass[w] = !(pm < r) && !P.already[s];                             // sidm.c:343-346
}

// ------------------------------------------------------------------ pass 2
struct Pass2 {
  int np; const int *np_dev; unsigned long long *ctr_pass1;   // np_dev != null: the number of slots is read on the device (np = upper bound)
  const int *passlist; const int *slot_part; SearchCtx C;
  const float4 *posm, *velh; const float *dvel; const float *dt; const double *rnd;
  const double *replay_dir; double sigma, s_a_inverse; uint32_t k0, k1; int xs_type; double vc, pl_n, pl_v0;
  const double *extra; const int *extra_off;   // type 4: replayed (rand, cosO-uniform) pairs per slot
  const double *kernel;            // begrun.c:968-992 table, 1002 doubles
  int *partner; float *dv; double *prob, *ptot;
  // reference-order mode
  int ref_order; int *cand; unsigned long long *candkey; int cand_stride, cand_cap; const int *krank, *lrank, *nstart; int *flags;
};

__device__ __forceinline__ double kernel_w(const double *K, double u, double hinv3) {
  const int ii = (int)(u * 1000);
  return hinv3 * (K[ii] + (K[ii + 1] - K[ii]) * (u - ((double)ii) / 1000) * 1000);   // sidm.c:359-363
}

// probability increment of one pair, sidm.c:366-382
__device__ __forceinline__ double pair_prob(const Pass2 &P, double mj, double wk, double rv, double dt_h0) {
  switch (P.xs_type) {
    case 1: return 0.5 * mj * wk * P.sigma * dt_h0;
    case 2: { const double beta = rv / P.vc, vd = 1.0 / (1.0 + beta * beta); return 0.5 * mj * wk * rv * vd * vd * P.sigma * dt_h0; }
    case 3: return 0.5 * mj * wk * rv * pow(rv / P.pl_v0, P.pl_n) * P.sigma * dt_h0;
    default: return 0.5 * mj * wk * rv * P.sigma * dt_h0;            // types 0 and 4
  }
}

__device__ __forceinline__ void unit_vector(const Pass2 &P, int s, int i, double n[3]) {
  if (P.replay_dir) { n[0] = P.replay_dir[3 * (size_t)s]; n[1] = P.replay_dir[3 * (size_t)s + 1]; n[2] = P.replay_dir[3 * (size_t)s + 2]; return; }
  // sidm_rand.h:24-37 (Marsaglia), draws from the particle's own Philox stream
  double y1, y2, r2; uint32_t c = 1;
  do {
    const uint4 q = philox((uint32_t)i, c++, 0u, 0u, P.k0, P.k1);
    y1 = 1.0 - 2.0 * u01(q.x); y2 = 1.0 - 2.0 * u01(q.y); r2 = y1 * y1 + y2 * y2;
    if (r2 > 1.0) { y1 = 1.0 - 2.0 * u01(q.z); y2 = 1.0 - 2.0 * u01(q.w); r2 = y1 * y1 + y2 * y2; }
  } while (r2 > 1.0);
  const double sq = sqrt(1.0 - r2);
  n[0] = 2.0 * y1 * sq; n[1] = 2.0 * y2 * sq; n[2] = 1.0 - 2.0 * r2;
}

__global__ void __launch_bounds__(128) k_pass2(Pass2 P) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = P.np_dev ? *P.np_dev : P.np;
  if ((int)(blockIdx.x * blockDim.x) >= np) return;                   // launched for the upper bound: nothing here
  if (t == 0 && P.ctr_pass1) atomicAdd(P.ctr_pass1, (unsigned long long)np);
  const bool valid = t < np;
  const int s = valid ? P.passlist[t] : 0;
  const int i = valid ? P.slot_part[s] : 0;
  const float4 p = valid ? P.posm[i] : make_float4(0, 0, 0, 0);
  const float4 vi = valid ? P.velh[i] : make_float4(0, 0, 0, 0);
  const float h = vi.w;
  const float sr2 = fmul(h, h);
  const double dt_h0 = valid ? (double)P.dt[s] * P.s_a_inverse : 0.0;
  const int start = valid ? search_start(P.C, i, p.x, p.y, p.z, h) : 0;
  const double hh = (double)h, hinv = 1.0 / hh, hinv3 = hinv * hinv * hinv;
  double rnd = valid ? P.rnd[s] : 2.0;
  double prob = 0, wk = 0, ptot = 0; int partner = -1; double prv[4] = {0, 0, 0, 0}; float pmass = 0;
  bool done = false; int trial = 0; double cosO = 0;      // type 4: angle trials of this slot (sidm.c:391-398)

  auto visit = [&](int j, float r2) {       // one neighbour in list order, sidm.c:352-385
    if (P.dvel[3 * (size_t)j] != 0.0f) return;
    const double r = sqrt((double)r2);
    if (r < hh) wk = kernel_w(P.kernel, r * hinv, hinv3);
    const float4 vj = P.velh[j];
    const double rvx = (double)fadd(vi.x, -vj.x), rvy = (double)fadd(vi.y, -vj.y), rvz = (double)fadd(vi.z, -vj.z);
    const double rv = sqrt(rvx * rvx + rvy * rvy + rvz * rvz);
    const float mj = P.posm[j].w;
    const double dp = pair_prob(P, (double)mj, wk, rv, dt_h0);
    ptot += dp;
    if (done) return;
    prob += dp;
    if (prob < rnd) return;
    partner = j;                                         // SidmTarget[i] = j, kept even if the angle is rejected
    if (P.xs_type == 4) {
      // sidm.c:391-398: draw a fresh uniform (it REPLACES the slot's uniform for the rest of the scan)
      // and cos(theta); accept with probability 1/(1 + beta^2 sin^2(theta/2))^2
      double u1, u2;
      const int xo = P.extra_off ? P.extra_off[s] + trial : 0;
      if (P.extra_off && xo < P.extra_off[s + 1]) { u1 = P.extra[2 * (size_t)xo]; u2 = P.extra[2 * (size_t)xo + 1]; }
      else { const uint4 q = philox((uint32_t)i, (uint32_t)trial, 1u, 0u, P.k0, P.k1); u1 = u01(q.x); u2 = u01(q.y); }
      trial++;
      const double beta = rv / P.vc;
      rnd = u1;
      cosO = 2.0 * u2 - 1.0;
      const double sin22 = 0.5 * (1.0 - cosO), denom = 1.0 + beta * beta * sin22;
      if (rnd >= 1 / (denom * denom) || rv == 0.0) return;
    }
    done = true;
    prv[0] = rvx; prv[1] = rvy; prv[2] = rvz; prv[3] = rv; pmass = mj;
  };

  if (!P.ref_order) {
    range_search_fast(P.C, valid, start, p.x, p.y, p.z, h, [&](int L, const float4 &, float r2, bool, int) { if (r2 < sr2) visit(P.C.leaf_orig[L], r2); });
  } else {
    // gather every candidate the reference's tree search appends, with a key that reproduces
    // its order: position along the octant-ordered tree; inside fully-contained cells the
    // next[] chain rank (forcetree.c:274-279, 2270-2276)
    int nc = 0; bool over = false;
    int *cl = P.cand + t; unsigned long long *ck = P.candkey + t; const int st = P.cand_stride;
    range_search(P.C, valid, start, p.x, p.y, p.z, h, [&](int L, const float4 &, float, bool bulk, int node) {
      if (nc >= P.cand_cap) { over = true; return; }
      const int o = P.C.leaf_orig[L];
      const unsigned long long key = bulk ? (((unsigned long long)P.nstart[node] << 32) | (unsigned)P.lrank[o])
                                          : ((unsigned long long)P.krank[o] << 32);
      // insertion into the sorted candidate list (short lists: ~60 entries)
      int k = nc - 1;
      while (k >= 0 && ck[(size_t)k * st] > key) { ck[(size_t)(k + 1) * st] = ck[(size_t)k * st]; cl[(size_t)(k + 1) * st] = cl[(size_t)k * st]; k--; }
      ck[(size_t)(k + 1) * st] = key; cl[(size_t)(k + 1) * st] = o;
      nc++;
    });
    if (over) P.flags[FL_ERR_NGB] = 1;
    // swap-remove sphere filter, forcetree.c:2191-2212
    int n = nc;
    for (int a = 0; a < n; a++) {
      const int j = cl[(size_t)a * st];
      const float4 q = P.posm[j];
      const float r2 = P.C.box > 0 ? dist2_per(q.x, q.y, q.z, p.x, p.y, p.z, P.C.box) : dist2_ref(q.x, q.y, q.z, p.x, p.y, p.z);
      if (r2 >= sr2) { cl[(size_t)a * st] = cl[(size_t)(n - 1) * st]; n--; a--; }
    }
    for (int a = 0; a < n; a++) {
      const int j = cl[(size_t)a * st];
      const float4 q = P.posm[j];
      visit(j, P.C.box > 0 ? dist2_per(q.x, q.y, q.z, p.x, p.y, p.z, P.C.box) : dist2_ref(q.x, q.y, q.z, p.x, p.y, p.z));
    }
  }
  if (!valid) return;
  P.prob[s] = prob; P.ptot[s] = ptot; P.partner[s] = partner;
  float dvx = 0, dvy = 0, dvz = 0;
  if (done) {
    const double rmass = (double)(pmass / (p.w + pmass));            // float division, sidm.c:387
    double n[3];
    unit_vector(P, s, i, n);
    if (P.xs_type == 4) {                                             // sidm.c:407-423
      double np_[3] = {prv[1] * n[2] - prv[2] * n[1], prv[2] * n[0] - prv[0] * n[2], prv[0] * n[1] - prv[1] * n[0]};   // perp(), sidm.c:29-52
      const double oo = sqrt(np_[0] * np_[0] + np_[1] * np_[1] + np_[2] * np_[2]);
      np_[0] /= oo; np_[1] /= oo; np_[2] /= oo;
      const double sinO = 1.0 - cosO * cosO > 0.0 ? sqrt(1.0 - cosO * cosO) : 0.0;
      dvx = (float)(rmass * (-prv[0] + cosO * prv[0] + sinO * prv[3] * np_[0]));
      dvy = (float)(rmass * (-prv[1] + cosO * prv[1] + sinO * prv[3] * np_[1]));
      dvz = (float)(rmass * (-prv[2] + cosO * prv[2] + sinO * prv[3] * np_[2]));
    } else {
      dvx = (float)(rmass * (-prv[0] + prv[3] * n[0]));               // sidm.c:446-451
      dvy = (float)(rmass * (-prv[1] + prv[3] * n[1]));
      dvz = (float)(rmass * (-prv[2] + prv[3] * n[2]));
    }
  }
  P.dv[3 * (size_t)s] = dvx; P.dv[3 * (size_t)s + 1] = dvy; P.dv[3 * (size_t)s + 2] = dvz;
}

// ------------------------------------------------------------------ resolve sweeps
// sidm.c:495-537: own result -> particle, in/out of range decides confirmation
__global__ void k_resolve_own(int ns, const int *slot_part, const int *sngb, const float *dv, int lo, int hi, int count_only,
                              int *ngb, float *dvel, int *confirm, unsigned long long *winner, unsigned long long wbase, const int *partner, unsigned long long *ctr) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  const int i = slot_part[s];
  const int nb = sngb[s];
  ngb[i] = nb;
  if (count_only) return;
  const float d0 = dv[3 * (size_t)s];
  int conf = 0;
  if (nb < lo || nb > hi) { if (d0 != 0.0f) atomicAdd(&ctr[CT_REJECTED], 1ull); }
  else if (d0 != 0.0f) {
    dvel[3 * (size_t)i] = d0; dvel[3 * (size_t)i + 1] = dv[3 * (size_t)s + 1]; dvel[3 * (size_t)i + 2] = dv[3 * (size_t)s + 2];
    atomicAdd(&ctr[CT_SCATTERED], 1ull);
    conf = 1;
  }
  confirm[s] = conf;
  if (conf) atomicMax(&winner[partner[s]], wbase + (unsigned long long)s);
}
// sidm.c:559-601: partner gets -dv; several slots naming one partner: the last in buffer order wins
__global__ void k_resolve_partner(int ns, const int *confirm, const int *partner, const float *dv, const unsigned long long *winner, unsigned long long wbase, float *dvel,
                                  const int *logpos, b200_scatlog *log, int logcap, const int *logbase_dev, const int *slot_part,
                                  const float4 *posm, const float4 *velh, const int *pid, float time, int *kick_list, int *nkick, int kick_cap) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns || !confirm[s]) return;
  const int j = partner[s];
  if (winner[j] == wbase + (unsigned long long)s) {
    dvel[3 * (size_t)j] = -dv[3 * (size_t)s]; dvel[3 * (size_t)j + 1] = -dv[3 * (size_t)s + 1]; dvel[3 * (size_t)j + 2] = -dv[3 * (size_t)s + 2];
    const int at = atomicAdd(nkick, 1);                    // the partner need not be active: b200_download_active() sends it along
    if (at < kick_cap) kick_list[at] = j;
  }
  const int lp = *logbase_dev + logpos[s];
  if (lp < logcap) {
    const int i = slot_part[s];
    b200_scatlog e;
    e.time = time; e.id1 = pid[i]; e.id2 = pid[j]; e.Hsml1 = velh[i].w; e.Hsml2 = velh[j].w;
    const float4 pi = posm[i], pj = posm[j], vi = velh[i], vj = velh[j];
    e.x1[0] = pi.x; e.x1[1] = pi.y; e.x1[2] = pi.z; e.x2[0] = pj.x; e.x2[1] = pj.y; e.x2[2] = pj.z;
    e.v1[0] = vi.x; e.v1[1] = vi.y; e.v1[2] = vi.z; e.v2[0] = vj.x; e.v2[1] = vj.y; e.v2[2] = vj.z;
    e.dv[0] = dv[3 * (size_t)s]; e.dv[1] = dv[3 * (size_t)s + 1]; e.dv[2] = dv[3 * (size_t)s + 2];
    log[lp] = e;
  }
}

__global__ void k_bump_logbase(int *logbase, const int *nlog) { if (threadIdx.x == 0 && blockIdx.x == 0) *logbase += *nlog; }

// multi-GPU: per-slot results of this rank's share of the buffer -> all ranks
struct __attribute__((aligned(16))) SlotRec { int ngb, partner; float dv[3]; int pass, pad0, pad1; };
__global__ void k_slot_pack(int nown, const int *order, const int *sngb, const int *partner, const float *dv, const int *pass, int count_only, SlotRec *send) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nown) return;
  const int s = order[k];
  SlotRec r; r.ngb = sngb[s]; r.partner = partner[s]; r.dv[0] = dv[3 * (size_t)s]; r.dv[1] = dv[3 * (size_t)s + 1]; r.dv[2] = dv[3 * (size_t)s + 2];
  r.pass = count_only ? 0 : pass[k]; r.pad0 = r.pad1 = 0;
  send[k] = r;
}
__global__ void k_slot_unpack(int ns, int world, int per_rank, const int *sorted_slots, const SlotRec *recv, int *sngb, int *partner, float *dv, unsigned long long *ctr) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)world * per_rank) return;
  const int q = (int)(g / per_rank), k = (int)(g % per_rank);
  const long long j = ((long long)(k >> kShardShift) * world + q) * kShardBlock + (k & (kShardBlock - 1));
  if (j >= ns) return;
  const int s = sorted_slots[j];
  const SlotRec r = recv[g];
  sngb[s] = r.ngb; partner[s] = r.partner; dv[3 * (size_t)s] = r.dv[0]; dv[3 * (size_t)s + 1] = r.dv[1]; dv[3 * (size_t)s + 2] = r.dv[2];
  if (r.pass) atomicAdd(&ctr[CT_PASS1], 1ull);
}

// Compact form of the same exchange: almost every slot only has a neighbour count to report, so a rank
// sends {header | one uint16 count per slot | the few scatter proposals}: ~2.5 instead of 32 bytes per slot
// (N=1e7: 25 MB instead of 320 MB gathered by every rank).  If a count does not fit 16 bits or a rank has
// more proposals than the buffer holds, all ranks see it in the gathered headers and fall back to SlotRec.
struct CompactHdr { int nprop, npass, big, pad; };
struct __attribute__((aligned(8))) PropRec { int k, partner; float dv[3]; int pad; };
__device__ __host__ inline size_t compact_ngb_bytes(int per_rank) { return ((size_t)per_rank * 2 + 15) & ~(size_t)15; }
__global__ void k_compact_pack(int nown, const int *order, const int *sngb, const int *partner, const float *dv, const int *pass,
                               int count_only, int per_rank, int cap, char *send) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nown) return;
  CompactHdr *hdr = reinterpret_cast<CompactHdr *>(send);
  unsigned short *n16 = reinterpret_cast<unsigned short *>(send + sizeof(CompactHdr));
  PropRec *props = reinterpret_cast<PropRec *>(send + sizeof(CompactHdr) + compact_ngb_bytes(per_rank));
  const int s = order[k];
  const int nb = sngb[s];
  n16[k] = (unsigned short)(nb < 65535 ? nb : 65535);
  if (nb >= 65535) hdr->big = 1;
  if (count_only) return;
  if (pass[k]) atomicAdd(&hdr->npass, 1);
  if (partner[s] >= 0) {
    const int at = atomicAdd(&hdr->nprop, 1);
    if (at < cap) { PropRec r; r.k = k; r.partner = partner[s]; r.dv[0] = dv[3 * (size_t)s]; r.dv[1] = dv[3 * (size_t)s + 1]; r.dv[2] = dv[3 * (size_t)s + 2]; r.pad = 0; props[at] = r; }
  }
}
__global__ void k_compact_unpack_ngb(int ns, int world, int per_rank, size_t rank_bytes, const int *sorted_slots, const char *recv, int *sngb,
                                     unsigned long long *ctr) {
  const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= (long long)world * per_rank) return;
  const int q = (int)(gi / per_rank), k = (int)(gi % per_rank);
  if (k == 0) atomicAdd(&ctr[CT_PASS1], (unsigned long long)reinterpret_cast<const CompactHdr *>(recv + (size_t)q * rank_bytes)->npass);
  const long long j = ((long long)(k >> kShardShift) * world + q) * kShardBlock + (k & (kShardBlock - 1));
  if (j >= ns) return;
  const unsigned short *n16 = reinterpret_cast<const unsigned short *>(recv + (size_t)q * rank_bytes + sizeof(CompactHdr));
  sngb[sorted_slots[j]] = n16[k];
}
__global__ void k_compact_unpack_props(int ns, int world, int per_rank, int cap, size_t rank_bytes, const int *sorted_slots, const char *recv,
                                       int *partner, float *dv) {
  const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= (long long)world * cap) return;
  const int q = (int)(gi / cap), a = (int)(gi % cap);
  const char *base = recv + (size_t)q * rank_bytes;
  if (a >= reinterpret_cast<const CompactHdr *>(base)->nprop) return;
  const PropRec r = reinterpret_cast<const PropRec *>(base + sizeof(CompactHdr) + compact_ngb_bytes(per_rank))[a];
  const long long j = ((long long)(r.k >> kShardShift) * world + q) * kShardBlock + (r.k & (kShardBlock - 1));
  if (j >= ns) return;
  const int s = sorted_slots[j];
  partner[s] = r.partner; dv[3 * (size_t)s] = r.dv[0]; dv[3 * (size_t)s + 1] = r.dv[1]; dv[3 * (size_t)s + 2] = r.dv[2];
}

// ------------------------------------------------------------------ host side
static double *d_kernel_table = nullptr;

static int ensure_sidm_buffers() {
  const size_t n = (size_t)g.maxpart, m = (size_t)g.maxnodes;
  auto al = [](void **p, size_t bytes) { if (*p) return B200_OK; return cudaMalloc(p, bytes + 256) == cudaSuccess ? B200_OK : B200_ERR_ALLOC; };
  B200_TRY(al((void **)&S.snode, (m + 1) * sizeof(SearchNode)));
  B200_TRY(al((void **)&S.snodef, (m + 2) * sizeof(SearchNodeF)));
  B200_TRY(al((void **)&S.last_active, n * sizeof(int)));
  B200_TRY(al((void **)&S.slot_of_sorted, n * sizeof(int)));
  B200_TRY(al((void **)&S.passlist, n * sizeof(int)));
  B200_TRY(al((void **)&S.logpos, (n + 1) * sizeof(int)));
  B200_TRY(al((void **)&S.dt, n * sizeof(float)));
  B200_TRY(al((void **)&S.already, n));
  B200_TRY(al((void **)&S.ptot, n * sizeof(double)));
  B200_TRY(al((void **)&S.x_redo, n * sizeof(int))); B200_TRY(al((void **)&S.x_redo2, n * sizeof(int))); B200_TRY(al((void **)&S.x_want, n * sizeof(int)));
  B200_TRY(al((void **)&S.x_keys, n * sizeof(int))); B200_TRY(al((void **)&S.x_keys2, n * sizeof(int)));
  B200_TRY(al((void **)&S.x_vals, n * sizeof(int))); B200_TRY(al((void **)&S.x_shard, (n + 64) * sizeof(int)));
  B200_TRY(al((void **)&S.groups, (n + m + 1) * sizeof(int2))); B200_TRY(al((void **)&S.gnode, (n + m + 1) * sizeof(int)));
  B200_TRY(al((void **)&S.gflag, (m + 2) * sizeof(int))); B200_TRY(al((void **)&S.gpos, (m + 2) * sizeof(int)));
  B200_TRY(al((void **)&S.order_leaf, n * sizeof(int))); B200_TRY(al((void **)&S.spart, n * sizeof(int)));
  B200_TRY(al((void **)&S.gown, (n + m + 1) * sizeof(int))); B200_TRY(al((void **)&S.gownflag, (n + m + 1) * sizeof(int)));
  if (!d_kernel_table) {
    double K[1002];
    const double PI = 3.14159265358979323846;
    K[1001] = 0;
    for (int i = 0; i <= 1000; i++) {          // begrun.c:968-992
      const double r = ((double)i) / 1000;
      K[i] = r <= 0.5 ? 8 / PI * (1 - 6 * r * r * (1 - r)) : 8 / PI * 2 * (1 - r) * (1 - r) * (1 - r);
    }
    if (cudaMalloc((void **)&d_kernel_table, sizeof(K)) != cudaSuccess) return B200_ERR_ALLOC;
    CUDA_TRY(cudaMemcpy(d_kernel_table, K, sizeof(K), cudaMemcpyHostToDevice));
  }
  return B200_OK;
}

// called by b200_finalize(): the buffers above are sized by the MaxPart of one b200_init()
void sidm_release() {
  void **ptrs[] = {(void **)&S.snode, (void **)&S.snodef, (void **)&S.last_active, (void **)&S.slot_of_sorted, (void **)&S.passlist,
                   (void **)&S.logpos, (void **)&S.rr, (void **)&S.dt, (void **)&S.already, (void **)&S.ptot,
                   (void **)&S.x_redo, (void **)&S.x_redo2, (void **)&S.x_want, (void **)&S.x_keys, (void **)&S.x_keys2, (void **)&S.x_vals, (void **)&S.x_shard, &S.cub_tmp,
                   (void **)&S.groups, (void **)&S.gnode, (void **)&S.gflag, (void **)&S.gpos, (void **)&S.order_leaf, (void **)&S.spart, (void **)&S.gown, (void **)&S.gownflag};
  for (auto pp : ptrs) { if (*pp) cudaFree(*pp); *pp = nullptr; }
  if (S.rx) cudaFree(S.rx); if (S.ro) cudaFree(S.ro);
  S.rx = nullptr; S.ro = nullptr; S.rx_cap = S.ro_cap = 0;
  S.rd = nullptr; S.replay_cap = 0; S.last_nactive = 0; S.last_all = false; S.cub_tmp_bytes = 0; S.ngroups = 0;
}

static int cub_scratch(size_t tb) {
  if (tb <= S.cub_tmp_bytes) return B200_OK;
  if (S.cub_tmp) { cudaStreamSynchronize(sidm_stream()); cudaFree(S.cub_tmp); }
  S.cub_tmp = nullptr; S.cub_tmp_bytes = 0;
  if (tb < ((size_t)64 << 20)) tb = (size_t)64 << 20;      // grow once: cudaMalloc / cudaFree serialise the device
  if (cudaMalloc(&S.cub_tmp, tb + 4096) != cudaSuccess) return B200_ERR_ALLOC;
  S.cub_tmp_bytes = tb + 4096;
  return B200_OK;
}

static SearchCtx search_ctx() {
  SearchCtx C; C.qcap = g.opt_queue_cap; C.M = g.num_nodes; C.snode = S.snode; C.snodef = S.snodef; C.leaf_posm = g.leaf_posm; C.leaf_orig = g.leaf_orig;
  C.nparent = g.nparent; C.leaf_parent = g.leaf_parent; C.orig_leaf = g.orig_leaf;
  C.box = (g.par.PeriodicBoundariesOn && g.par.BoxSize > 0) ? g.par.BoxSize : 0.0;
  C.pad = g.refits_since_build > 0 ? g.d_pad : nullptr;
  static const bool wrapped_always = getenv("B200_PERIODIC_SEARCH_FROM_ROOT") != nullptr;        // A/B: round-1 behaviour
  C.domain = (C.box > 0 && g.ntrees == 1 && !wrapped_always) ? g.d_domain : nullptr;
  return C;
}

int refresh_search_nodes() {
  B200_TRY(ensure_sidm_buffers());
  if (g.search_epoch == g.tree_epoch) return B200_OK;
  g.search_epoch = g.tree_epoch;
  cudaStream_t st = sidm_stream();
  const int m = g.num_nodes;
  k_search_nodes<<<cdiv(m, 256), 256, 0, st>>>(m, g.nodes, g.geom, g.npstart, g.nnp, g.nparent, S.snode, S.snodef);
  // query groups of the warp-shared search
  const int aligned = g.shard_world > 1;
  k_group_flag<<<cdiv(m + 1, 256), 256, 0, st>>>(m, S.snode, g.nparent, S.gflag, aligned);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, S.gflag, S.gpos, m + 1, st);
  B200_TRY(cub_scratch(tb));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(S.cub_tmp, tb, S.gflag, S.gpos, m + 1, st));
  k_group_emit<<<cdiv(m, 256), 256, 0, st>>>(m, S.snode, S.gflag, S.gpos, S.groups, S.gnode, aligned);
  CUDA_TRY(cudaMemcpyAsync(&S.ngroups, S.gpos + m, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  count_launch(5);
  S.nown = 0;
  if (aligned && S.ngroups > 0) {
    k_group_own<<<cdiv(S.ngroups, 256), 256, 0, st>>>(S.ngroups, S.groups, g.shard_world, g.shard_rank, S.gownflag);
    size_t tbo = 0;
    cub::DeviceSelect::Flagged(nullptr, tbo, g.iota, S.gownflag, S.gown, g.d_flags + FL_NPASS, S.ngroups, st);
    B200_TRY(cub_scratch(tbo));
    CUDA_TRY(cub::DeviceSelect::Flagged(S.cub_tmp, tbo, g.iota, S.gownflag, S.gown, g.d_flags + FL_NPASS, S.ngroups, st));
    CUDA_TRY(cudaMemcpyAsync(&S.nown, g.d_flags + FL_NPASS, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    count_launch(3);
  }
  static const bool dbg = getenv("B200_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "libsidm_b200: %d query groups for %d particles (%.1f per warp)\n", S.ngroups, g.n, (double)g.n / (S.ngroups > 0 ? S.ngroups : 1));
  return B200_OK;
}

// one sidm() call for a device-resident active list (d_active == nullptr: every particle in
// index order).  replay arrays are host pointers indexed by buffer slot.
// device counters of the chain are cumulative since sidm_begin(); sidm_collect() brings them to the host (after a sync)
static int sidm_begin(cudaStream_t st) {
  CUDA_TRY(cudaMemsetAsync(g.d_ctr + CT_CAND, 0, (CT_COUNT - CT_CAND) * sizeof(unsigned long long), st));
  CUDA_TRY(cudaMemsetAsync(g.d_flags + FL_ERR_NGB, 0, sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(g.d_nkick + 1, 0, sizeof(int), st));          // scatter-log position
  g.scatlog_n = 0;
  return B200_OK;
}
static int sidm_collect(cudaStream_t st) {
  CUDA_TRY(cudaMemcpyAsync(g.h_ctr + CT_CAND, g.d_ctr + CT_CAND, (CT_COUNT - CT_CAND) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(g.h_flags, g.d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(g.h_flags + FL_COUNT, g.d_nkick + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  const int nlog = g.h_flags[FL_COUNT];
  g.scatlog_n = nlog < g.scatlog_cap ? nlog : g.scatlog_cap;
  g.cnt.sct_pass1 = (int)g.h_ctr[CT_PASS1];
  g.cnt.sct_scattered = (int)g.h_ctr[CT_SCATTERED]; g.cnt.sct_rejected = (int)g.h_ctr[CT_REJECTED];
  g.cnt.ngb_candidates = (long long)g.h_ctr[CT_CAND];
  if (g.h_flags[FL_ERR_NGB]) return B200_ERR_NGBOVERFLOW;
  return B200_OK;
}

int sidm_impl(const int *d_active, int nactive, double time, double vmax, const b200_replay *replay, bool count_only, bool defer_final) {
  if (!g.tree_valid) return B200_ERR_STATE;
  B200_TRY(ensure_sidm_buffers());
  B200_TRY(refresh_search_nodes());
  cudaStream_t st = sidm_stream();
  const int na = nactive;
  if (na <= 0) return B200_OK;
  // small passes (the repair loop, small active sets) are latency-bound: every rank does them completely -
  // same inputs, counter-based random numbers, so same results - instead of paying an exchange per pass
  const bool sharded = g.shard_world > 1 && na >= g.shard_min_work && !g.shard_busy;
  const int B = 256;
  const double sainv = s_a_inverse_at(time);
  // C_Pmax, sidm.c:226-316 (types 0..3)
  const int T = g.par.CrossSectionType;
  double sigma = g.par.CrossSectionInternal, vc = g.par.YukawaVelocity;
  if (g.par.ComovingIntegrationOn) { sigma = sigma / pow(time, T == 1 ? 2.5 : 2.0); vc = vc / sqrt(time); }
  const double ball = (3. / 4. / 3.14159265358979323846) * (g.par.DesNumNgb + g.par.MaxNumNgbDeviation);
  double C_Pmax;
  if (T == 0 || T == 4) C_Pmax = 1.0 * ball * 2 * vmax * sigma;          // sidm.c:267-272, 309-313
  else if (T == 1) C_Pmax = 1.0 * ball * sigma;
  else if (T == 2) {
    if (2.0 * vmax < vc / sqrt(3.0)) { const double beta = 2.0 * vmax / vc, vd = 1.0 / (1.0 + beta * beta); C_Pmax = 1.0 * ball * 2.0 * vmax * vd * vd * sigma; }
    else C_Pmax = 1.0 * ball * (3.0 * sqrt(3.0) / 16.0) * vc * sigma;
  } else C_Pmax = 1.0 * ball * 2 * g.par.CrossSectionVelScale * sigma;

  g.sidm_calls++;
  const uint32_t k0 = (uint32_t)(g.par.Seed & 0xffffffffu) ^ (uint32_t)(g.sidm_calls * 0x9E3779B9u);
  const uint32_t k1 = (uint32_t)(g.par.Seed >> 32) ^ (uint32_t)(g.sidm_calls >> 32) ^ 0x5851F42Du;

  // B200_TIMING=1: device time of the phases of every large pass (development aid)
  static const bool timing = getenv("B200_TIMING") != nullptr;
  static cudaEvent_t tev[8]; static bool tev_ok = false;
  const bool tm = timing && na >= (1 << 20);
  if (tm && !tev_ok) { for (auto &e : tev) cudaEventCreate(&e); tev_ok = true; }
  auto mark = [&](int k) { if (tm) cudaEventRecord(tev[k], st); };
  mark(0);

  int bunch = g.par.BunchSizeSidm > 0 ? g.par.BunchSizeSidm : na;
  size_t replay_off = 0;
  for (int b0 = 0; b0 < na; b0 += bunch) {
    const int nb = (na - b0 < bunch) ? na - b0 : bunch;
    const int *act = d_active ? d_active + b0 : nullptr;
    if (!d_active && b0 > 0) return B200_ERR_ARG;     // bunches need an explicit list
    const int G = cdiv(nb, B);
    // slots: exported-first buffer order
    k_export_flag<<<G, B, 0, st>>>(nb, act, g.posm, g.velh, g.d_domain, g.s_flag, g.ntypes > 1 ? g.ptype : nullptr);
    CUDA_TRY(cudaMemsetAsync(g.s_flag + nb, 0, sizeof(int), st));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, g.s_flag, g.s_pos, nb + 1, st);
    B200_TRY(cub_scratch(tb));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(S.cub_tmp, tb, g.s_flag, g.s_pos, nb + 1, st));
    int *slot_of_active = g.s_repair;     // scratch
    // warp-shared search: when every particle is a query, and - on one rank - for explicit lists that hold a good part of the
    // particles (repair passes after a long step: thread-per-query costs 4 ns per query, the shared search 0.6)
    const bool periodic_box = g.par.PeriodicBoundariesOn && g.par.BoxSize > 0;
    const bool plain_search = !periodic_box || search_ctx().domain != nullptr;      // periodic: the shared searches need box_interior()
    const bool group_ok = plain_search && S.ngroups > 0 && g.opt_group_search;
    const bool group_all = !act && nb == g.n && group_ok;
    const bool group_part = act && !sharded && group_ok && nb > kWarpQueryMax / 4 && (long long)nb * 8 >= g.n && g.par.BunchSizeSidm <= 0;
    if (group_part) CUDA_TRY(cudaMemsetAsync(S.spart, 0xff, (size_t)g.n * sizeof(int), st));
    // processing order: slots sorted along the tree key order (spatial coherence inside a warp)
    k_assign_slots<<<G, B, 0, st>>>(nb, act, g.s_flag, g.s_pos, g.s_slot_part, slot_of_active, g.curtime, g.dvel, time, S.dt, S.already,
                                    g.krank, S.x_keys, act ? S.x_vals : S.slot_of_sorted, g.d_flags, g.s_partner, g.s_dv, g.s_prob, S.ptot,
                                    group_part ? S.spart : nullptr);
    count_launch(4);
    // explicit lists: slots sorted along the tree order, so that the queries of a warp are neighbours.  Small lists are searched
    // one warp per query (k_pass1_warp), where the order of the queries does not matter: no sort (five launches less per repair pass)
    const bool periodic_search = g.par.PeriodicBoundariesOn && g.par.BoxSize > 0 && search_ctx().domain == nullptr;
    const bool sort_slots = act && !group_part && !(nb <= kWarpQueryMax && !periodic_search && g.opt_group_search && !sharded);
    if (sort_slots) {
      size_t tb2 = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tb2, S.x_keys, S.x_keys2, S.x_vals, S.slot_of_sorted, nb, 0, 32, st);
      B200_TRY(cub_scratch(tb2));
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(S.cub_tmp, tb2, S.x_keys, S.x_keys2, S.x_vals, S.slot_of_sorted, nb, 0, 32, st));
      count_launch(4);
    }
    // replay arrays
    const double *d_rr = nullptr, *d_rd = nullptr;
    if (replay && replay->rand && !count_only) {
      const size_t need = (size_t)nb * 4;
      if (S.replay_cap < need) {
        if (S.rr) cudaFree(S.rr);
        if (cudaMalloc((void **)&S.rr, need * sizeof(double)) != cudaSuccess) return B200_ERR_ALLOC;
        S.replay_cap = need; S.rd = S.rr + nb;
      }
      S.rd = S.rr + nb;
      CUDA_TRY(cudaMemcpyAsync(S.rr, replay->rand + replay_off, (size_t)nb * sizeof(double), cudaMemcpyHostToDevice, st));
      d_rr = S.rr;
      if (replay->dir) { CUDA_TRY(cudaMemcpyAsync(S.rd, replay->dir + 3 * replay_off, (size_t)nb * 3 * sizeof(double), cudaMemcpyHostToDevice, st)); d_rd = S.rd; }
      replay_off += nb;
    }
    // type 4: the replayed angle trials (one bunch only)
    const double *d_rx = nullptr; const int *d_ro = nullptr;
    if (replay && replay->extra_off && replay->extra && !count_only && b0 == 0 && nb == na) {
      const int nx = replay->extra_off[nb];
      if (S.rx_cap < (size_t)nx + 1 || S.ro_cap < (size_t)nb + 1) {
        if (S.rx) cudaFree(S.rx); if (S.ro) cudaFree(S.ro);
        S.rx = nullptr; S.ro = nullptr;
        if (cudaMalloc((void **)&S.rx, ((size_t)nx + 1) * 2 * sizeof(double)) != cudaSuccess) return B200_ERR_ALLOC;
        if (cudaMalloc((void **)&S.ro, ((size_t)nb + 1) * sizeof(int)) != cudaSuccess) return B200_ERR_ALLOC;
        S.rx_cap = (size_t)nx + 1; S.ro_cap = (size_t)nb + 1;
      }
      if (nx > 0) CUDA_TRY(cudaMemcpyAsync(S.rx, replay->extra, (size_t)nx * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(S.ro, replay->extra_off, ((size_t)nb + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
      d_rx = S.rx; d_ro = S.ro;
    }
    // this rank's share of the buffer (all of it on one GPU)
    const int *order = (act && !sort_slots) ? S.x_vals : S.slot_of_sorted; int nord = nb;
    const bool group_mode = group_all || group_part;
    const int *global_order = order;                 // all ranks' slots in processing order
    if (group_mode && sharded) {                     // sharded group search: processing order = leaf order
      k_order_leaf<<<G, B, 0, st>>>(nb, g.leaf_orig, slot_of_active, S.order_leaf);
      count_launch();
      global_order = S.order_leaf;
    }
    if (sharded) { B200_TRY(shard_select(global_order, nb, S.x_shard, &nord, st)); order = S.x_shard; }
    mark(1);
    // pass 1
    Pass1 P1;
    P1.ns = nord; P1.order = order; P1.slot_part = g.s_slot_part; P1.C = search_ctx();
    P1.posm = g.posm; P1.velh = g.velh; P1.dt = S.dt; P1.already = S.already; P1.replay_rand = d_rr;
    P1.C_Pmax = C_Pmax; P1.s_a_inverse = sainv; P1.k0 = k0; P1.k1 = k1;
    P1.ngb = g.s_ngb; P1.pmax = g.s_pmax; P1.rnd = g.s_rand; P1.pass = g.s_pass; P1.count_only = count_only; P1.ctr = g.d_ctr;
    // every particle is a query, open boundaries: warp-shared search over the query groups
    const bool grouped = group_mode;
    if (grouped) {
      Pass1G PG;
      PG.ng = sharded ? S.nown : S.ngroups; PG.own = sharded ? S.gown : nullptr; PG.groups = S.groups; PG.gnode = S.gnode; PG.C = P1.C; PG.velh = g.velh; PG.slot_of_part = group_part ? S.spart : slot_of_active; PG.masked = group_part;
      PG.dt = S.dt; PG.already = S.already; PG.replay_rand = d_rr; PG.C_Pmax = C_Pmax; PG.s_a_inverse = sainv; PG.k0 = k0; PG.k1 = k1;
      PG.ngb = g.s_ngb; PG.pmax = g.s_pmax; PG.rnd = g.s_rand; PG.pass = g.s_pass; PG.order_leaf = S.order_leaf; PG.count_only = count_only; PG.ctr = g.d_ctr;
      PG.rank = sharded ? g.shard_rank : 0; PG.world = sharded ? g.shard_world : 1;
      if (PG.ng > 0) k_pass1_group<<<cdiv((long long)PG.ng * 32, 128), 128, 0, st>>>(PG);
      if (!sharded) { order = S.order_leaf; nord = g.n; }  // the pass flags are indexed by leaf position (of all particles)
    } else if (nord > 0) {
      // small query sets: one warp per query (latency), large ones: one thread per query (throughput)
      if (nord <= kWarpQueryMax && plain_search && g.opt_group_search) {
        if (periodic_box) k_pass1_warp<true><<<cdiv((long long)nord * 32, 128), 128, 0, st>>>(P1);
        else k_pass1_warp<false><<<cdiv((long long)nord * 32, 128), 128, 0, st>>>(P1);
      }
      else k_pass1<<<cdiv(nord, 128), 128, 0, st>>>(P1);
    }
    count_launch(2);
    mark(2);
    int npass = 0;
    if (!count_only) {
      // compact the slots that passed the first approximation, keeping the processing order
      size_t tb3 = 0;
      CUDA_TRY(cudaMemsetAsync(g.d_flags + FL_NPASS, 0, sizeof(int), st));
      if (nord > 0) {
        cub::DeviceSelect::Flagged(nullptr, tb3, order, g.s_pass, S.passlist, g.d_flags + FL_NPASS, nord, st);
        B200_TRY(cub_scratch(tb3));
        CUDA_TRY(cub::DeviceSelect::Flagged(S.cub_tmp, tb3, order, g.s_pass, S.passlist, g.d_flags + FL_NPASS, nord, st));
      }
      const bool ref_order = g.par.ReferenceNgbOrder != 0;
      // the scan in tree order needs no scratch sized by the number of slots that passed: the pair kernel is launched for the
      // upper bound and reads the count on the device (no host round trip); the reference-order mode sizes its candidate lists
      const bool lean = !ref_order;
      if (!lean) {
        CUDA_TRY(cudaMemcpyAsync(g.h_flags, g.d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        npass = g.h_flags[FL_NPASS];
      } else npass = nord;
      count_launch(3);
      // reference-order mode: per-slot candidate lists in scratch (option "cand_cap" entries each), about 1.6 GB at most
      const int cand_cap = g.opt_cand_cap;
      const int chunk = ref_order ? (int)(((size_t)131072 * 1024 / cand_cap + 127) / 128 * 128) : npass;
      if (ref_order && npass > 0) {
        const size_t need = (size_t)(npass < chunk ? npass : chunk) * cand_cap;
        if (g.s_cand_cap < need) {
          if (g.s_cand) cudaFree(g.s_cand); if (g.s_candkey) cudaFree(g.s_candkey);
          g.s_cand = nullptr; g.s_candkey = nullptr; g.s_cand_cap = 0;
          if (cudaMalloc((void **)&g.s_cand, need * sizeof(int)) != cudaSuccess) return B200_ERR_ALLOC;
          if (cudaMalloc((void **)&g.s_candkey, need * sizeof(unsigned long long)) != cudaSuccess) return B200_ERR_ALLOC;
          g.s_cand_cap = need;
        }
      }
      for (int c0 = 0; c0 < npass; c0 += chunk) {
        const int nc = (npass - c0 < chunk) ? npass - c0 : chunk;
        Pass2 P2;
        P2.np = nc; P2.np_dev = lean ? g.d_flags + FL_NPASS : nullptr; P2.passlist = S.passlist + c0; P2.slot_part = g.s_slot_part; P2.C = search_ctx();
        P2.ctr_pass1 = sharded ? nullptr : g.d_ctr + CT_PASS1;       // sharded: the gathered headers carry the counts of all ranks
        P2.posm = g.posm; P2.velh = g.velh; P2.dvel = g.dvel; P2.dt = S.dt; P2.rnd = g.s_rand; P2.replay_dir = d_rd;
        P2.sigma = sigma; P2.s_a_inverse = sainv; P2.k0 = k0; P2.k1 = k1; P2.xs_type = T; P2.vc = vc;
        P2.pl_n = g.par.CrossSectionPowLaw; P2.pl_v0 = g.par.CrossSectionVelScale; P2.kernel = d_kernel_table;
        P2.extra = d_rx; P2.extra_off = d_ro;
        P2.partner = g.s_partner; P2.dv = g.s_dv; P2.prob = g.s_prob; P2.ptot = S.ptot;
        P2.ref_order = ref_order; P2.cand = g.s_cand; P2.candkey = g.s_candkey; P2.cand_stride = nc; P2.cand_cap = cand_cap;
        P2.krank = g.krank; P2.lrank = g.lrank; P2.nstart = g.nstart; P2.flags = g.d_flags;
        k_pass2<<<cdiv(nc, 128), 128, 0, st>>>(P2);
        count_launch();
      }
    }
    mark(3);
    if (sharded) {
      // exchange {Ngb, partner, dv} of every slot (replaces the result + confirm hypercube
      // passes of sidm.c:463-553); the two write sweeps below then run identically on all ranks
      const int per_rank = shard_max_blocks(nb, g.shard_world) * kShardBlock;
      const int cap = per_rank / 16 > 4096 ? per_rank / 16 : 4096;
      const size_t cbytes = sizeof(CompactHdr) + compact_ngb_bytes(per_rank) + (size_t)cap * sizeof(PropRec);
      bool dense = !g.opt_compact_exchange || (long long)cbytes > g.shard_cap;
      if (!dense) {
        CUDA_TRY(cudaMemsetAsync(g.shard_send, 0, sizeof(CompactHdr), st));
        if (nord > 0) { k_compact_pack<<<cdiv(nord, B), B, 0, st>>>(nord, order, g.s_ngb, g.s_partner, g.s_dv, g.s_pass, count_only, per_rank, cap, (char *)g.shard_send); count_launch(); }
        B200_TRY(shard_exchange((long long)cbytes, st));
        // every rank reads the same gathered headers, so every rank takes the same decision
        static CompactHdr hh[64];
        if (g.shard_world > 64) dense = true;
        else {
          for (int q = 0; q < g.shard_world; q++)
            CUDA_TRY(cudaMemcpyAsync(&hh[q], (char *)g.shard_recv + (size_t)q * cbytes, sizeof(CompactHdr), cudaMemcpyDeviceToHost, st));
          CUDA_TRY(cudaStreamSynchronize(st));
          for (int q = 0; q < g.shard_world; q++) if (hh[q].big || hh[q].nprop > cap) dense = true;
        }
        if (!dense) {
          k_compact_unpack_ngb<<<cdiv((long long)g.shard_world * per_rank, B), B, 0, st>>>(nb, g.shard_world, per_rank, cbytes, global_order, (const char *)g.shard_recv, g.s_ngb, g.d_ctr);
          k_compact_unpack_props<<<cdiv((long long)g.shard_world * cap, B), B, 0, st>>>(nb, g.shard_world, per_rank, cap, cbytes, global_order, (const char *)g.shard_recv, g.s_partner, g.s_dv);
          count_launch(2);
        }
      }
      if (dense) {
        if (nord > 0) { k_slot_pack<<<cdiv(nord, B), B, 0, st>>>(nord, order, g.s_ngb, g.s_partner, g.s_dv, g.s_pass, count_only, (SlotRec *)g.shard_send); count_launch(); }
        B200_TRY(shard_exchange((long long)per_rank * sizeof(SlotRec), st));
        const long long tot = (long long)g.shard_world * per_rank;
        k_slot_unpack<<<cdiv(tot, B), B, 0, st>>>(nb, g.shard_world, per_rank, global_order, (const SlotRec *)g.shard_recv, g.s_ngb, g.s_partner, g.s_dv, g.d_ctr);
        count_launch();
      }
    }
    mark(4);
    // resolve
    int *confirm = g.s_pass;   // reuse (pass flags are consumed)
    const unsigned long long wbase = g.winner_base;        // entries of earlier calls are smaller than every entry of this one
    g.winner_base += (unsigned long long)nb;
    k_resolve_own<<<G, B, 0, st>>>(nb, g.s_slot_part, g.s_ngb, g.s_dv, g.par.DesNumNgb - g.par.MaxNumNgbDeviation,
                                   g.par.DesNumNgb + g.par.MaxNumNgbDeviation, count_only, g.ngb, g.dvel, confirm, g.s_winner, wbase, g.s_partner, g.d_ctr);
    count_launch();
    if (!count_only) {
      CUDA_TRY(cudaMemsetAsync(confirm + nb, 0, sizeof(int), st));
      size_t tb4 = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb4, confirm, S.logpos, nb + 1, st);
      B200_TRY(cub_scratch(tb4));
      CUDA_TRY(cub::DeviceScan::ExclusiveSum(S.cub_tmp, tb4, confirm, S.logpos, nb + 1, st));
      k_resolve_partner<<<G, B, 0, st>>>(nb, confirm, g.s_partner, g.s_dv, g.s_winner, wbase, g.dvel, S.logpos, g.d_scatlog, g.scatlog_cap, g.d_nkick + 1,
                                         g.s_slot_part, g.posm, g.velh, g.pid, (float)time, g.kick_list, g.d_nkick, g.maxpart);
      k_bump_logbase<<<1, 32, 0, st>>>(g.d_nkick + 1, S.logpos + nb);      // the next bunch / pass appends (no host round trip)
      count_launch(4);
    }
    g.last_nslot = nb;
    mark(5);
    if (tm) {
      cudaEventSynchronize(tev[5]);
      float t[5];
      for (int k = 0; k < 5; k++) cudaEventElapsedTime(&t[k], tev[k], tev[k + 1]);
      fprintf(stderr, "libsidm_b200 timing rank %d: sidm pass of %d slots (%d here, %d groups): slots %.3f | pass1 %.3f | select+pass2 %.3f | exchange %.3f | resolve %.3f ms\n",
              g.shard_rank, nb, nord, S.ngroups, t[0], t[1], t[2], t[3], t[4]);
    }
  }
  g.cnt.sct_ntot += na;
  // totals since the last b200_sidm(): the sum of the reference's SCT lines for this step
  if (defer_final) return B200_OK;                           // the repair loop collects at its own synchronisation points
  return sidm_collect(st);
}

// ------------------------------------------------------------------ k nearest (ngb_treefind)
struct KnnParams { int nq; const int *idx; const int *want; SearchCtx C; const float4 *posm; int k; float *h2; const SearchNode *snode; const float4 *geom; const uint64_t *shi, *slo;
                   const int *nstart; const unsigned char *nlevel; const int *nend; };
constexpr int kKnnMax = 64;

__global__ void __launch_bounds__(128) k_knn(KnnParams P) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < P.nq && (!P.want || P.want[t]);           // want: only the flagged entries of the list
  const int i = valid ? P.idx[t] : 0;
  const float4 p = valid ? P.posm[i] : make_float4(0, 0, 0, 0);
  const int K = P.k;
  float sr = 1.0f;
  if (valid) {
    // starting radius from the local density: descend while the cell holds > 200 particles
    // (forcetree.c:2327-2347)
    int th = P.C.leaf_parent[P.C.orig_leaf[i]];
    while (P.C.nparent[th] >= 0) th = P.C.nparent[th];      // the root of the particle's own tree
    for (;;) {
      const int skip = P.snode[th].skip;
      const float4 gc = P.geom[th];
      const int cnt = P.nend[th] - P.nstart[th] + 1;
      if (cnt <= 200) break;
      const int oct = (p.x > gc.x ? 1 : 0) | (p.y > gc.y ? 2 : 0) | (p.z > gc.z ? 4 : 0);
      int next = -1;
      const int lev = P.nlevel[th];
      for (int c = th + 1; c < skip; c = P.snode[c].skip)
        if (digit_at(P.shi[P.nstart[c]], P.slo[P.nstart[c]], lev) == oct) { next = c; break; }
      if (next < 0 || P.nend[next] - P.nstart[next] + 1 <= 200) break;
      th = next;
    }
    const int cnt_th = P.nend[th] - P.nstart[th] + 1;
    sr = (float)((double)P.geom[th].w * pow((3.0 / (4 * 3.14159265358979323846) * 1.2) * K / ((double)(float)cnt_th), 1.0 / 3));
  }
  float best[kKnnMax];
  float h2max = 0;
  bool done = !valid;
  for (int rep = 0; rep < 200 && !done; rep++) {
    int found = 0;
    for (int a = 0; a < K; a++) best[a] = 3.4e38f;
    const int start = done ? 0 : search_start(P.C, i, p.x, p.y, p.z, sr);
    range_search_fast(P.C, !done, start, p.x, p.y, p.z, sr, [&](int, const float4 &, float r2, bool, int) {
      found++;
      if (r2 < best[K - 1]) {              // keep the K smallest squared distances, sorted
        int a = K - 2;
        while (a >= 0 && best[a] > r2) { best[a + 1] = best[a]; a--; }
        best[a + 1] = r2;
      }
    });
    if (done) continue;
    if (found < K) { if (found > 5) sr = (float)((double)sr * pow((2.1 * (double)(float)K) / found, 1.0 / 3)); else sr *= 2.0f; continue; }
    h2max = best[K - 1];
    if (h2max <= fmul(sr, sr)) { done = true; continue; }
    sr = (float)((double)sr * 1.26);
  }
  if (valid) P.h2[t] = h2max;
}

static int knn_device(const int *d_idx, int nq, int k, float *d_h2, const int *d_want = nullptr) {
  if (k < 1 || k > kKnnMax) return B200_ERR_ARG;
  B200_TRY(refresh_search_nodes());
  KnnParams P; P.nq = nq; P.idx = d_idx; P.want = d_want; P.C = search_ctx(); P.posm = g.posm; P.k = k; P.h2 = d_h2; P.snode = S.snode; P.geom = g.geom;
  P.shi = g.skey_hi; P.slo = g.skey_lo; P.nstart = g.nstart; P.nlevel = g.nlevel; P.nend = g.nend;
  k_knn<<<cdiv(nq, 128), 128, 0, sidm_stream()>>>(P);
  count_launch();
  return B200_OK;
}

// ------------------------------------------------------------------ ensure_neighbours
// sidm.c:862-888 (ensure==1) or init.c:456-478 (ensure==0): flag particles whose neighbour
// count is out of range and not yet converged, update the bisection bracket
__global__ void k_flag_repair(int n, const int *list, int lo, int hi, int ensure_variant, const int *ngb, const float4 *velh, float *left, float *right, int *flag) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int i = list[a];
  int f = 0;
  const int nb = ngb[i];
  if (nb < lo || nb > hi) {
    float L = left[i], R = right[i];
    const float h = velh[i].w;
    if (!(L > 0 && R > 0 && (double)fadd(R, -L) < 1.0e-3 * (double)L)) {
      f = 1;
      if (ensure_variant) {
        if (nb < lo) L = (float)fmax((double)h, (double)L);
        else { if (R != 0) { if (h < R) R = h; } else R = h; }
      } else {
        if (nb < lo) L = h;
        if (nb > hi) R = h;
      }
      left[i] = L; right[i] = R;
    }
  }
  flag[a] = f;
}
// sidm.c:917-929: new smoothing length for the flagged particles (h from k-NN where asked)
struct TypeCount { int n[8]; };
__global__ void k_new_hsml(int nr, const int *redo, const int *ngb, const float *left, const float *right, float4 *velh, int des,
                           int ensure_variant, TypeCount tc, const int *ptype, int *want_knn) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nr) return;
  const int i = redo[a];
  float4 v = velh[i];
  const float L = left[i], R = right[i];
  int knn = 0;
  if (L == 0 || R == 0) {
    if (ensure_variant && R == 0 && ngb[i] < 15 && tc.n[ptype[i] & 7] > des) knn = 1;    // NtypeLocal[P[i].Type], sidm.c:919
    else v.w = (float)((double)v.w * (0.5 + 0.5 * pow(ngb[i] / ((double)des), -1.0 / 3)));
  } else v.w = (float)(0.5 * ((double)L + (double)R));
  velh[i] = v;
  want_knn[a] = knn;
}
__global__ void k_apply_knn(int nw, const int *wlist, const int *want, const float *h2, float4 *velh) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nw || !want[a]) return;
  velh[wlist[a]].w = (float)sqrt((double)h2[a]);
}
__global__ void k_zero_lr(int na, const int *active, float *left, float *right) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= na) return;
  const int i = active ? active[a] : a;
  left[i] = 0; right[i] = 0;
}
__global__ void k_count_out_of_range(int na, const int *active, const int *ngb, int lo, int hi, int *out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= na) return;
  const int i = active ? active[a] : a;
  if (ngb[i] < lo || ngb[i] > hi) atomicAdd(out, 1);
}
__global__ void k_iota(int n, int *a) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = i; }
__global__ void k_set_hsml_from_h2(int n, const float *h2, float4 *velh, float *left, float *right) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  velh[i].w = (float)sqrt((double)h2[i]); left[i] = 0; right[i] = 0;
}

// shared repair loop; ensure_variant 1 = sidm_ensure_neighbours (runs sidm()), 0 = start-up
// (counts only, setup_nbr_sidm)
// `list` (nlist entries): the particles the loop may touch - the active list of the step (sidm.c:862-874 walks the ForceFlag chain)
// or every particle at start-up (init.c:456).  A pass can only flag particles of the pass before it (nobody else's count
// changes), so every pass after the first works on the previous pass's list: no O(N) sweep per pass.
static int repair_loop(int ensure_variant, double time, double vmax, const b200_replay *replay, int maxiter, const int *list, int nlist) {
  cudaStream_t st = sidm_stream();
  const int B = 256;
  const int lo = g.par.DesNumNgb - g.par.MaxNumNgbDeviation, hi = g.par.DesNumNgb + g.par.MaxNumNgbDeviation;
  int iter = 0;
  size_t roff = 0;
  int *redo = S.x_redo;                 // compacted list in list order (the reference's order, sidm.c:862)
  int *want = S.x_want; float *h2 = (float *)S.x_keys2;
  for (;;) {
    k_flag_repair<<<cdiv(nlist, B), B, 0, st>>>(nlist, list, lo, hi, ensure_variant, g.ngb, g.velh, g.left, g.right, g.s_flag);
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, list, g.s_flag, redo, g.d_flags + FL_NREPAIR, nlist, st);
    B200_TRY(cub_scratch(tb));
    CUDA_TRY(cub::DeviceSelect::Flagged(S.cub_tmp, tb, list, g.s_flag, redo, g.d_flags + FL_NREPAIR, nlist, st));
    CUDA_TRY(cudaMemcpyAsync(g.h_flags, g.d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    count_launch(4);
    const int nr = g.h_flags[FL_NREPAIR];
    if (g.h_flags[FL_ERR_NGB]) return B200_ERR_NGBOVERFLOW;
    if (nr == 0) break;
    TypeCount tc;
    for (int t = 0; t < 8; t++) tc.n[t] = t < 6 ? g.type_count[t] : 0;
    k_new_hsml<<<cdiv(nr, B), B, 0, st>>>(nr, redo, g.ngb, g.left, g.right, g.velh, g.par.DesNumNgb, ensure_variant, tc, g.ptype, want);
    count_launch();
    if (ensure_variant) {
      // exact k-th neighbour distance where sidm.c:918-922 asks for it (rare: Ngb < 15 with no upper bracket): the search runs for
      // the flagged entries of the list only - no compaction, no host round trip
      B200_TRY(knn_device(redo, nr, g.par.DesNumNgb, h2, want));
      k_apply_knn<<<cdiv(nr, B), B, 0, st>>>(nr, redo, want, h2, g.velh);
      count_launch();
    }
    b200_replay rp; const b200_replay *rpp = nullptr;
    rp.extra = nullptr; rp.extra_off = nullptr;
    if (replay && replay->rand) { rp.rand = replay->rand + roff; rp.dir = replay->dir ? replay->dir + 3 * roff : nullptr; rpp = &rp; roff += nr; }
    B200_TRY(sidm_impl(redo, nr, time, vmax, rpp, ensure_variant == 0, true));      // collected with the next pass's count
    iter++;
    g.cnt.ensure_repaired += nr;
    list = redo; nlist = nr;                                    // the next pass looks at these only
    redo = (redo == S.x_redo) ? S.x_redo2 : S.x_redo;
    if (iter > maxiter) { fprintf(stderr, "libsidm_b200: failed to converge in ensure_neighbours\n"); return B200_ERR_HSML; }
  }
  g.cnt.ensure_iterations = iter;
  return sidm_collect(st);
}

}  // namespace b200
using namespace b200;

static int stage_active(const int *active, int nactive, const int **d_out, int *n_out) {
  B200_TRY(ensure_sidm_buffers());
  if (!active) { *d_out = nullptr; *n_out = g.n; S.last_all = true; S.last_nactive = g.n; return B200_OK; }
  if (nactive < 0 || nactive > g.n) return B200_ERR_ARG;
  CUDA_TRY(cudaMemcpyAsync(S.last_active, active, (size_t)nactive * sizeof(int), cudaMemcpyHostToDevice, sidm_stream()));
  S.last_all = false; S.last_nactive = nactive;
  *d_out = S.last_active; *n_out = nactive;
  return B200_OK;
}

extern "C" int b200_sidm(const int *active, int nactive, double time, double vmax, const b200_replay *replay) {
  if (!g.ready || g.n <= 0 || !g.tree_valid) return B200_ERR_STATE;
  const int *d; int na;
  B200_TRY(stage_active(active, nactive, &d, &na));
  g.cnt.sct_ntot = g.cnt.sct_pass1 = g.cnt.sct_scattered = g.cnt.sct_rejected = 0; g.cnt.ngb_candidates = 0;
  g.cnt.ensure_iterations = 0; g.cnt.ensure_repaired = 0;
  B200_TRY(sidm_begin(sidm_stream()));
  CUDA_TRY(cudaEventRecord(g.ev_s0, sidm_stream()));
  int rc = sidm_impl(d, na, time, vmax, replay, false, false);
  cudaEventRecord(g.ev_s1, sidm_stream()); cudaEventSynchronize(g.ev_s1);
  cudaEventElapsedTime(&g.cnt.ms_sidm, g.ev_s0, g.ev_s1);
  return rc;
}

extern "C" int b200_setup_nbr_sidm(const int *active, int nactive) {
  if (!g.ready || g.n <= 0 || !g.tree_valid) return B200_ERR_STATE;
  const int *d; int na;
  B200_TRY(stage_active(active, nactive, &d, &na));
  B200_TRY(sidm_begin(sidm_stream()));
  return sidm_impl(d, na, 0.0, 0.0, nullptr, true, false);
}

extern "C" int b200_sidm_ensure_neighbours(int mode, double time, double vmax, const b200_replay *replay) {
  (void)mode;   // restoring / resetting the time line (sidm.c:943-965) is the host driver's job
  if (!g.ready || g.n <= 0 || !g.tree_valid) return B200_ERR_STATE;
  B200_TRY(ensure_sidm_buffers());
  cudaStream_t st = sidm_stream();
  const int lo = g.par.DesNumNgb - g.par.MaxNumNgbDeviation, hi = g.par.DesNumNgb + g.par.MaxNumNgbDeviation;
  const int na = S.last_nactive > 0 ? S.last_nactive : g.n;
  const int *act = (S.last_all || S.last_nactive == 0) ? nullptr : S.last_active;
  CUDA_TRY(cudaEventRecord(g.ev_s0, st));
  // candidates among the active particles, sidm.c:836-846
  CUDA_TRY(cudaMemsetAsync(g.d_flags + FL_NREPAIR, 0, sizeof(int), st));
  k_count_out_of_range<<<cdiv(na, 256), 256, 0, st>>>(na, act, g.ngb, lo, hi, g.d_flags + FL_NREPAIR);
  CUDA_TRY(cudaMemcpyAsync(g.h_flags, g.d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  count_launch();
  int rc = B200_OK;
  if (g.h_flags[FL_NREPAIR] > 0) {
    k_zero_lr<<<cdiv(na, 256), 256, 0, st>>>(na, act, g.left, g.right);     // sidm.c:857-859
    count_launch();
    // all particles: iota[i] = i is what every tree build leaves there (k_keys)
    rc = repair_loop(1, time, vmax, replay, 30, act ? act : g.iota, na);
  }
  cudaEventRecord(g.ev_s1, st); cudaEventSynchronize(g.ev_s1);
  cudaEventElapsedTime(&g.cnt.ms_ensure, g.ev_s0, g.ev_s1);
  return rc;
}

extern "C" int b200_setup_smoothinglengths_sidm(int desired_ngb) {
  if (!g.ready || g.n <= 0 || !g.tree_valid) return B200_ERR_STATE;
  B200_TRY(ensure_sidm_buffers());
  cudaStream_t st = sidm_stream();
  const int n = g.n;
  k_iota<<<cdiv(n, 256), 256, 0, st>>>(n, g.iota);
  float *h2 = (float *)S.x_keys2;
  B200_TRY(knn_device(g.iota, n, desired_ngb, h2));                          // init.c:440-444
  k_set_hsml_from_h2<<<cdiv(n, 256), 256, 0, st>>>(n, h2, g.velh, g.left, g.right);
  count_launch(2);
  S.last_all = true; S.last_nactive = n;
  B200_TRY(sidm_begin(st));
  B200_TRY(sidm_impl(nullptr, n, 0.0, 0.0, nullptr, true, false));           // setup_nbr_sidm(), init.c:446
  return repair_loop(0, 0.0, 0.0, nullptr, 60, g.iota, n);                   // init.c:453-509
}

#include <chrono>
extern "C" int b200_compute_accelerations(int mode, const int *active, int nactive, double time, double vmax) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  static const bool timing = getenv("B200_TIMING") != nullptr;      // host clock at the points where the host has synchronised
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  B200_TRY(b200_predict(time));                    // gravtree.c:72 predict_collisionless_only(All.Time)
  const auto t1 = now();
  B200_TRY(b200_tree_build());                     // gravtree.c:76 force_treebuild()
  const auto t2 = now();
  if (mode != 0) return b200_gravity(active, nactive, time);     // accel.c:62: no SIDM in mode 1
  // gravity_tree() and sidm() + sidm_ensure_neighbours() (accel.c:39-65) are independent once the
  // tree exists: they read the same particles and tree and write disjoint fields (Accel, OldAcc,
  // GravCost | dVel, NgbVelDisp, HsmlVelDisp, Left, Right).  One GPU: the walk is issued without a
  // host sync on the main stream and the SIDM chain - a throughput-bound search followed by a tail
  // of small launches with host round trips - on a second, high-priority stream, so the tail costs
  // no wall time.  Sharded over several GPUs this needs a host whose all-gather callback runs on
  // b200_current_stream() (option "shard_overlap"); the gravity exchange is then issued last.
  // Modes (option "overlap"): 0 = one phase after the other as accel.c:39-65; 1 (default) = the whole SIDM chain next to the
  // walk; 2 = only the SIDM pass next to the walk, the repair loop after it.  Next to the walk every small kernel of the chain
  // waits ~0.1-0.2 ms for thread-block slots: free while a 20 ms walk runs anyway.  Sharded, the exchange of the gravity results
  // is issued as soon as the SIDM pass has released the exchange buffers, so it travels while the repair loop runs
  // (2 GPUs, N = 1e7: 23.1 ms per step in mode 1 against 24.3 in mode 2 or with the exchange issued last).
  int omode = g.opt_overlap < 0 ? 1 : g.opt_overlap;
  if (g.shard_world > 1 && !g.opt_shard_overlap) omode = 0;
  const bool overlap = omode != 0;
  static const bool early_gx = getenv("B200_LATE_GRAVITY_EXCHANGE") == nullptr;      // A/B switch
  if (active && (nactive < 0 || nactive > g.n)) return B200_ERR_ARG;
  if (!overlap) {
    B200_TRY(b200_gravity(active, nactive, time));   // gravtree.c:127-324
    B200_TRY(b200_sidm(active, nactive, time, vmax, nullptr));
    return b200_sidm_ensure_neighbours(mode, time, vmax, nullptr);
  }
  CUDA_TRY(cudaEventRecord(g.ev_fork, g.stream));
  CUDA_TRY(cudaStreamWaitEvent(g.stream_sidm, g.ev_fork, 0));
  // Order of issue: walk first, SIDM chain next to it.  (Measured alternative: SIDM pass alone first,
  // then walk || repair loop - 35.0 vs 34.4 ms on one GPU, 27.6 vs 26.5 ms on two: the chain's many
  // small launches cost the walk about as much as they would cost alone.)
  int rs = B200_OK;
  const int rc = gravity_impl(active, nactive, time, true);
  if (rc == B200_OK) {
    g.overlap_now = true;
    rs = b200_sidm(active, nactive, time, vmax, nullptr);
    // sharded: the SIDM pass has used the exchange buffers and is complete (b200_sidm ends with a host sync of its stream); the
    // exchange of the gravity results is issued now, behind the walk that is still running, instead of after the repair loop -
    // every rank issues its collectives in the same order.  While it is queued the buffers are busy: repair passes run replicated.
    if (rs == B200_OK && g.shard_world > 1 && early_gx) rs = gravity_exchange_early();
    if (rs == B200_OK && omode == 1) rs = b200_sidm_ensure_neighbours(mode, time, vmax, nullptr);
    g.overlap_now = false;
  }
  cudaStreamSynchronize(g.stream_sidm);
  const int rf = gravity_finish();
  // later work on the main stream is ordered after the SIDM chain
  cudaEventRecord(g.ev_join, g.stream_sidm); cudaStreamWaitEvent(g.stream, g.ev_join, 0);
  if (rc != B200_OK) return rc;
  if (rs != B200_OK) return rs;
  if (rf != B200_OK) return rf;
  const auto t3 = now();
  int re = B200_OK;
  if (omode == 2) re = b200_sidm_ensure_neighbours(mode, time, vmax, nullptr);     // on the main stream, the GPU to itself
  if (timing && g.shard_rank == 0)
    fprintf(stderr, "libsidm_b200 timing: step of %d targets, mode %d: predict %.3f | build %.3f | walk || sidm (+ exchanges) %.3f | repair loop after it %.3f | total %.3f ms (walk kernel %.3f)\n",
            active ? nactive : g.n, omode, ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, now()), ms(t0, now()), g.cnt.ms_walk);
  return re;
}

extern "C" int b200_ngb_treefind(const int *idx, int n, int desngb, float *h2_out) {
  if (!g.ready || g.n <= 0 || !g.tree_valid) return B200_ERR_STATE;
  if (!idx || n <= 0 || n > g.n || !h2_out) return B200_ERR_ARG;
  B200_TRY(ensure_sidm_buffers());
  CUDA_TRY(cudaMemcpyAsync(S.x_redo, idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, sidm_stream()));
  float *h2 = (float *)S.x_keys2;
  B200_TRY(knn_device(S.x_redo, n, desngb, h2));
  CUDA_TRY(cudaMemcpyAsync(h2_out, h2, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, sidm_stream()));
  CUDA_TRY(cudaStreamSynchronize(sidm_stream()));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// neighbour lists for tests: the list pass 2 would scan (ordered per ReferenceNgbOrder)
struct ListParams { int nq; const int *idx; SearchCtx C; const float4 *posm, *velh; int cap; int *count, *list; int ref_order;
                    const int *krank, *lrank, *nstart; unsigned long long *keys; };
__global__ void k_ngb_lists(ListParams P) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < P.nq;
  const int i = valid ? P.idx[t] : 0;
  const float4 p = valid ? P.posm[i] : make_float4(0, 0, 0, 0);
  const float h = valid ? P.velh[i].w : 0.0f; const float sr2 = fmul(h, h);
  int *out = P.list + (size_t)(valid ? t : 0) * P.cap; unsigned long long *ck = P.keys + (size_t)(valid ? t : 0) * P.cap;
  int nc = 0;
  const int start = valid ? search_start(P.C, i, p.x, p.y, p.z, h) : 0;
  if (!P.ref_order) {
    range_search_fast(P.C, valid, start, p.x, p.y, p.z, h, [&](int L, const float4 &, float r2, bool, int) { if (r2 < sr2) { if (nc < P.cap) out[nc] = P.C.leaf_orig[L]; nc++; } });
    if (valid) P.count[t] = nc;
    return;
  }
  range_search(P.C, valid, start, p.x, p.y, p.z, h, [&](int L, const float4 &, float, bool bulk, int node) {
    if (nc >= P.cap) { nc++; return; }
    const int o = P.C.leaf_orig[L];
    const unsigned long long key = bulk ? (((unsigned long long)P.nstart[node] << 32) | (unsigned)P.lrank[o]) : ((unsigned long long)P.krank[o] << 32);
    int k = nc - 1;
    while (k >= 0 && ck[k] > key) { ck[k + 1] = ck[k]; out[k + 1] = out[k]; k--; }
    ck[k + 1] = key; out[k + 1] = o; nc++;
  });
  if (!valid) return;
  if (nc > P.cap) { P.count[t] = nc; return; }
  int n = nc;
  for (int a = 0; a < n; a++) {
    const float4 q = P.posm[out[a]];
    if ((P.C.box > 0 ? dist2_per(q.x, q.y, q.z, p.x, p.y, p.z, P.C.box) : dist2_ref(q.x, q.y, q.z, p.x, p.y, p.z)) >= sr2) { out[a] = out[n - 1]; n--; a--; }
  }
  P.count[t] = n;
}

extern "C" int b200_ngb_lists(const int *idx, int n, int cap, int *count_out, int *list_out) {
  if (!g.ready || g.n <= 0 || !g.tree_valid) return B200_ERR_STATE;
  if (!idx || n <= 0 || cap <= 0 || !count_out || !list_out) return B200_ERR_ARG;
  B200_TRY(refresh_search_nodes());
  int *d_idx, *d_cnt, *d_list; unsigned long long *d_keys;
  if (cudaMalloc((void **)&d_idx, (size_t)n * 4) != cudaSuccess) return B200_ERR_ALLOC;
  if (cudaMalloc((void **)&d_cnt, (size_t)n * 4) != cudaSuccess) return B200_ERR_ALLOC;
  if (cudaMalloc((void **)&d_list, (size_t)n * cap * 4) != cudaSuccess) return B200_ERR_ALLOC;
  if (cudaMalloc((void **)&d_keys, (size_t)n * cap * 8) != cudaSuccess) return B200_ERR_ALLOC;
  cudaMemcpyAsync(d_idx, idx, (size_t)n * 4, cudaMemcpyHostToDevice, sidm_stream());
  cudaMemsetAsync(d_list, 0xff, (size_t)n * cap * 4, sidm_stream());
  ListParams P; P.nq = n; P.idx = d_idx; P.C = search_ctx(); P.posm = g.posm; P.velh = g.velh; P.cap = cap; P.count = d_cnt; P.list = d_list;
  P.ref_order = g.par.ReferenceNgbOrder; P.krank = g.krank; P.lrank = g.lrank; P.nstart = g.nstart; P.keys = d_keys;
  k_ngb_lists<<<cdiv(n, 64), 64, 0, sidm_stream()>>>(P);
  count_launch();
  cudaMemcpyAsync(count_out, d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, sidm_stream());
  cudaMemcpyAsync(list_out, d_list, (size_t)n * cap * 4, cudaMemcpyDeviceToHost, sidm_stream());
  cudaError_t e = cudaStreamSynchronize(sidm_stream());
  cudaFree(d_idx); cudaFree(d_cnt); cudaFree(d_list); cudaFree(d_keys);
  if (e != cudaSuccess) { g.last_cuda = (int)e; return B200_ERR_CUDA; }
  return B200_OK;
}

extern "C" int b200_sidm_debug(int nslot, int *slot_particle, double *pmax, double *prob_total, int *partner) {
  if (!g.ready || nslot <= 0 || nslot > g.last_nslot) return B200_ERR_ARG;
  cudaStream_t st = sidm_stream();
  if (slot_particle) cudaMemcpyAsync(slot_particle, g.s_slot_part, (size_t)nslot * 4, cudaMemcpyDeviceToHost, st);
  if (pmax) cudaMemcpyAsync(pmax, g.s_pmax, (size_t)nslot * 8, cudaMemcpyDeviceToHost, st);
  if (prob_total) cudaMemcpyAsync(prob_total, S.ptot, (size_t)nslot * 8, cudaMemcpyDeviceToHost, st);
  if (partner) cudaMemcpyAsync(partner, g.s_partner, (size_t)nslot * 4, cudaMemcpyDeviceToHost, st);
  CUDA_TRY(cudaStreamSynchronize(st));
  return B200_OK;
}

extern "C" int b200_get_scatlog(b200_scatlog *out, int cap, int *n) {
  if (!g.ready || !out || !n) return B200_ERR_ARG;
  const int m = g.scatlog_n < cap ? g.scatlog_n : cap;
  if (m > 0) CUDA_TRY(cudaMemcpy(out, g.d_scatlog, (size_t)m * sizeof(b200_scatlog), cudaMemcpyDeviceToHost));
  *n = m;
  return B200_OK;
}
