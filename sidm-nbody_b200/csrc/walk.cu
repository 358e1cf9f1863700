// walk.cu - force_treeevaluate() for many targets (reference: forcetree.c:786-1377,
// gravtree.c:127-324) and force_treeevaluate_direct() (forcetree.c:1896-1975).
//
// Walk design ("warp-lockstep pre-order stream"):
//   * nodes live in depth-first pre-order, so for every target the reference's walk
//     (open -> nextnode, accept -> sibling) is a strictly increasing sequence of node ids:
//     open = id+1, accept = skip[id].
//   * a warp owns 32 targets that are adjacent along the octant-key order.  The warp visits
//     node `cur` = the smallest id any lane still needs; lanes whose own walk is at `cur`
//     take their OWN open/accept decision with their own OldAcc (so every target gets exactly
//     the interaction set the reference gives it), the others idle until the stream reaches
//     the id they skipped to.  All 32 lanes therefore load the same 64-byte node record
//     (one broadcast transaction instead of 32 gathers) and the same <=8 leaf particles.
//   * interaction arithmetic in float, accumulated in float over a few interactions and
//     flushed into double accumulators (north_star: "float forces accumulated in double").
// Algorithmic HBM bytes: per warp 64 B per node record streamed + 16 B per
// leaf particle streamed; per target 16 B read + 32 B written.
#include <cub/cub.cuh>
#include "ctx.cuh"

namespace b200 {

double s_a_inverse_at(double time);
int ewald_tables(const float4 **out);

struct WalkParams {
  int nt, num_nodes;
  const int *tsorted;      // sorted target list: slot ids (or particle ids when slot_part==null)
  const int *slot_part;    // slot -> particle (null: slot == particle)
  const float4 *posm; const float *oldacc;
  const NodeRec *nodes; const float4 *leaf_posm;
  double *acc; int *cost;
  float theta2, alpha, h_inv; int criterion;
  unsigned long long *ctr;
  // periodic box (forcetree.c:870-877,921-930): nearest image + Ewald correction table
  float box, boxhalf, ewald_fac; const float4 *ewald;
  // one tree per particle type, laid out one after the other in the node array (forcetree.c:90-158, 798-808):
  // tree t = nodes [troot[t], troot[t+1]); softening of an interaction = max(eps of the tree's type, eps of the target's type)
  int ntrees; int troot[7]; int ttype[6]; float eps[6]; const int *ptype;
};

struct EpsTab { double e[8]; };             // SofteningTable by particle type
constexpr int kEwaldN = 64, kEwaldD = 32;        // EN, ED of ewald.c:12-14

// ewald_corr(), ewald.c:171-238: fold into the first octant, trilinear interpolation of the
// (ED+1)^3 table; the table holds {fx, fy, fz, unused} per grid point, already scaled by 1/L^2
__device__ __forceinline__ void ewald_corr(const float4 *tab, float fac, float dx, float dy, float dz, float &cx, float &cy, float &cz) {
  const float sx = dx < 0 ? 1.0f : -1.0f, sy = dy < 0 ? 1.0f : -1.0f, sz = dz < 0 ? 1.0f : -1.0f;
  float u = fabsf(dx) * fac, v = fabsf(dy) * fac, w = fabsf(dz) * fac;
  int i = (int)u, j = (int)v, k = (int)w;
  if (i >= kEwaldD) i = kEwaldD - 1;
  if (j >= kEwaldD) j = kEwaldD - 1;
  if (k >= kEwaldD) k = kEwaldD - 1;
  u -= i; v -= j; w -= k;
  const int S1 = kEwaldD + 1, S2 = S1 * S1;
  const float4 *b = tab + (i * S2 + j * S1 + k);
  const float4 t000 = __ldg(b), t001 = __ldg(b + 1), t010 = __ldg(b + S1), t011 = __ldg(b + S1 + 1);
  const float4 t100 = __ldg(b + S2), t101 = __ldg(b + S2 + 1), t110 = __ldg(b + S2 + S1), t111 = __ldg(b + S2 + S1 + 1);
  const float f1 = (1 - u) * (1 - v) * (1 - w), f2 = (1 - u) * (1 - v) * w, f3 = (1 - u) * v * (1 - w), f4 = (1 - u) * v * w;
  const float f5 = u * (1 - v) * (1 - w), f6 = u * (1 - v) * w, f7 = u * v * (1 - w), f8 = u * v * w;
  cx = sx * (t000.x * f1 + t001.x * f2 + t010.x * f3 + t011.x * f4 + t100.x * f5 + t101.x * f6 + t110.x * f7 + t111.x * f8);
  cy = sy * (t000.y * f1 + t001.y * f2 + t010.y * f3 + t011.y * f4 + t100.y * f5 + t101.y * f6 + t110.y * f7 + t111.y * f8);
  cz = sz * (t000.z * f1 + t001.z * f2 + t010.z * f3 + t011.z * f4 + t100.z * f5 + t101.z * f6 + t110.z * f7 + t111.z * f8);
}
// ewald_pot_corr(), ewald.c:246-285: same trilinear lookup on the potential table (component w)
__device__ __forceinline__ float ewald_pot_corr(const float4 *tab, float fac, float dx, float dy, float dz) {
  float u = fabsf(dx) * fac, v = fabsf(dy) * fac, w = fabsf(dz) * fac;
  int i = (int)u, j = (int)v, k = (int)w;
  if (i >= kEwaldD) i = kEwaldD - 1;
  if (j >= kEwaldD) j = kEwaldD - 1;
  if (k >= kEwaldD) k = kEwaldD - 1;
  u -= i; v -= j; w -= k;
  const int S1 = kEwaldD + 1, S2 = S1 * S1;
  const float4 *b = tab + (i * S2 + j * S1 + k);
  const float p000 = __ldg(&b->w), p001 = __ldg(&(b + 1)->w), p010 = __ldg(&(b + S1)->w), p011 = __ldg(&(b + S1 + 1)->w);
  const float p100 = __ldg(&(b + S2)->w), p101 = __ldg(&(b + S2 + 1)->w), p110 = __ldg(&(b + S2 + S1)->w), p111 = __ldg(&(b + S2 + S1 + 1)->w);
  return p000 * ((1 - u) * (1 - v) * (1 - w)) + p001 * ((1 - u) * (1 - v) * w) + p010 * ((1 - u) * v * (1 - w)) + p011 * ((1 - u) * v * w) +
         p100 * (u * (1 - v) * (1 - w)) + p101 * (u * (1 - v) * w) + p110 * (u * v * (1 - w)) + p111 * (u * v * w);
}
__device__ __forceinline__ float wrap_image(float d, float box, float boxhalf) {
  while (d > boxhalf) d -= box;
  while (d < -boxhalf) d += box;
  return d;
}

constexpr int kFlushEvery = 32;

// softened (r < h) node and particle factors, forcetree.c:1026-1074 / :904-915.  Rare (only cells
// and particles within 2.8 eps of the target), so kept out of line to keep the hot loop short.
__device__ __forceinline__ float2 pn_soft(float r2, float potq, float mass, float pp, float h_inv) {
  const float r = sqrtf(r2), u = r * h_inv;
  if (!(u > 1.0e-4f)) return make_float2(0.f, 0.f);
  float w2, w3, w4;
  soft_w234(u, w2, w3, w4);
  const float wf = soft_force(u);
  const float ri = 1.0f / r;
  const float h2i = h_inv * h_inv, h3i = h2i * h_inv, h4i = h2i * h2i, h5i = h2i * h3i, h6i = h3i * h3i;
  return make_float2(mass * h3i * wf + potq * h6i * w3 * ri + 0.5f * pp * w4 * h4i * ri, w2 * h5i);   // fac, ff
}
__device__ __forceinline__ float pp_soft(float r2, float mass, float h_inv) {
  const float u = sqrtf(r2) * h_inv;
  if (!(u > 1.0e-4f)) return 0.f;
  return mass * h_inv * h_inv * h_inv * soft_force(u);
}

// MODE: 0 = lanes of this warp use different criteria, 1 = all relative (forcetree.c:1097),
// 2 = all BH (forcetree.c:817); the choice is warp-uniform, the arithmetic identical.
template <bool PER, int MODE>
__device__ __forceinline__ void walk_loop(const WalkParams &P, const float4 tp, const bool bh, const float oac, int &no,
                                          double &ax, double &ay, double &az, int &npart, int &nnode, unsigned &wnodes, unsigned &wparts,
                                          const float h_inv, const int M) {
  const float theta2 = P.theta2;
  const float h2 = 1.0f / (h_inv * h_inv);
  const float4 *nodes4 = reinterpret_cast<const float4 *>(P.nodes);
  int cur = __reduce_min_sync(0xffffffffu, no);
  while (cur < M) {
    float fx = 0, fy = 0, fz = 0;
    // float partial sums over <= kFlushEvery cells, then one flush into the double accumulators
    int it = 0;
    for (; it < kFlushEvery && cur < M; it++) {
      const float4 *nd = nodes4 + 4 * (size_t)cur;
      const float4 A = __ldg(nd);              // s.xyz, mass
      const float4 Bv = __ldg(nd + 1);         // oc, bmax2, pinfo, skip
#ifdef WALK_EARLY_Q
      // request the second half with the first (volatile asm: the compiler may not sink it into the
      // accept branch): one dependent cache access per visit instead of two
      float4 Cv, Dv;                           // Q11 Q22 Q33 Q12 | Q13 Q23 P len2
      asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(Cv.x), "=f"(Cv.y), "=f"(Cv.z), "=f"(Cv.w) : "l"(nd + 2));
      asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(Dv.x), "=f"(Dv.y), "=f"(Dv.z), "=f"(Dv.w) : "l"(nd + 3));
#else
      const float4 Cv = __ldg(nd + 2);         // Q11 Q22 Q33 Q12
      const float4 Dv = __ldg(nd + 3);         // Q13 Q23 P len2
#endif
      float dx = A.x - tp.x, dy = A.y - tp.y, dz = A.z - tp.z;
      if (PER) { dx = wrap_image(dx, P.box, P.boxhalf); dy = wrap_image(dy, P.box, P.boxhalf); dz = wrap_image(dz, P.box, P.boxhalf); }
      const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      // forcetree.c:967 / :1253-1257
      bool crit;
      if (MODE == 1) crit = (Bv.x > oac * r2 * r2 * r2) || (r2 < Bv.y);
      else if (MODE == 2) crit = Dv.w > r2 * theta2;
      else crit = bh ? (Dv.w > r2 * theta2) : ((Bv.x > oac * r2 * r2 * r2) || (r2 < Bv.y));
      const bool act = (no == cur);
      const bool acc = act && !crit, open = act && crit;
      if (acc) {
        const float qx = fmaf(Dv.x, dz, fmaf(Cv.w, dy, Cv.x * dx));
        const float qy = fmaf(Dv.y, dz, fmaf(Cv.y, dy, Cv.w * dx));
        const float qz = fmaf(Cv.z, dz, fmaf(Dv.y, dy, Dv.x * dx));
        const float t = fmaf(dz, qz, fmaf(dy, qy, dx * qx));               // y^T Q y = 2 potq
        // fac = m/r^3 + (15 potq/r^2 - 1.5 P)/r^5 ; ff = -3/r^5   (forcetree.c:1262-1301)
        const float ri = rsqrt_fast(r2), r2i = ri * ri, r3i = r2i * ri, r5i = r3i * r2i;
        float fac = fmaf(fmaf(t * r2i, 7.5f, -1.5f * Dv.z), r5i, A.w * r3i);
        float ff = -3.0f * r5i;
        if (r2 < h2) { const float2 v = pn_soft(r2, 0.5f * t, A.w, Dv.z, h_inv); fac = v.x; ff = v.y; }
        fx = fmaf(ff, qx, fmaf(dx, fac, fx));
        fy = fmaf(ff, qy, fmaf(dy, fac, fy));
        fz = fmaf(ff, qz, fmaf(dz, fac, fz));
        if (PER) {                             // forcetree.c:1076-1082
          float ex, ey, ez;
          ewald_corr(P.ewald, P.ewald_fac, dx, dy, dz, ex, ey, ez);
          fx += A.w * ex; fy += A.w * ey; fz += A.w * ez;
        }
        nnode++;
      }
      no = acc ? __float_as_int(Bv.w) : (open ? cur + 1 : no);   // accept: jump over the subtree; open: first child
      if (__any_sync(0xffffffffu, open)) {
        const int pinfo = __float_as_int(Bv.z);
        const int np = pinfo & 15;
        const float4 *lp = P.leaf_posm + (pinfo >> 4);
        wparts += np;
        for (int k = 0; k < np; k++) {
          const float4 q = __ldg(lp + k);
          if (open) {
            float px = q.x - tp.x, py = q.y - tp.y, pz = q.z - tp.z;
            if (PER) { px = wrap_image(px, P.box, P.boxhalf); py = wrap_image(py, P.box, P.boxhalf); pz = wrap_image(pz, P.box, P.boxhalf); }
            const float pr2 = fmaf(pz, pz, fmaf(py, py, px * px));
            const float ri = rsqrt_fast(pr2);
            float fac = q.w * ri * ri * ri;                                  // m/r^3, forcetree.c:1135-1186
            if (pr2 < h2) fac = pp_soft(pr2, q.w, h_inv);
            fx = fmaf(px, fac, fx); fy = fmaf(py, fac, fy); fz = fmaf(pz, fac, fz);
            if (PER && pr2 * h_inv * h_inv > 1.0e-8f) {                      // u > 1e-4, forcetree.c:921-930
              float ex, ey, ez;
              ewald_corr(P.ewald, P.ewald_fac, px, py, pz, ex, ey, ez);
              fx += q.w * ex; fy += q.w * ey; fz += q.w * ez;
            }
          }
        }
        if (open) npart += np;
      }
      cur = __reduce_min_sync(0xffffffffu, no);
    }
    ax += (double)fx; ay += (double)fy; az += (double)fz;
    wnodes += it;                              // node records this warp streamed (I_n)
  }
}

#ifndef WALK_THREADS
#define WALK_THREADS 128
#endif
#ifndef WALK_MINBLOCKS
#define WALK_MINBLOCKS 10
#endif
template <bool PER, bool MULTI>
__global__ void __launch_bounds__(WALK_THREADS, WALK_MINBLOCKS) k_walk(WalkParams P) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < P.nt;
  int slot = 0, part = 0;
  if (valid) { slot = P.tsorted[t]; part = P.slot_part ? P.slot_part[slot] : slot; }
  float4 tp = make_float4(0, 0, 0, 0); float oa = 0;
  if (valid) { tp = P.posm[part]; oa = P.oldacc[part]; }
  const bool bh = (P.criterion == 0) || (oa == 0.0f);          // forcetree.c:801
  const float oac = oa * P.alpha;                               // forcetree.c:1129
  int no = valid ? 0 : 0x7fffffff;
  double ax = 0, ay = 0, az = 0;
  int npart = 0, nnode = 0;
  unsigned wnodes = 0, wparts = 0;
  const bool all_rel = __all_sync(0xffffffffu, !valid || !bh), all_bh = __all_sync(0xffffffffu, !valid || bh);
  if (!MULTI) {
    if (all_rel) walk_loop<PER, 1>(P, tp, bh, oac, no, ax, ay, az, npart, nnode, wnodes, wparts, P.h_inv, P.num_nodes);
    else if (all_bh) walk_loop<PER, 2>(P, tp, bh, oac, no, ax, ay, az, npart, nnode, wnodes, wparts, P.h_inv, P.num_nodes);
    else walk_loop<PER, 0>(P, tp, bh, oac, no, ax, ay, az, npart, nnode, wnodes, wparts, P.h_inv, P.num_nodes);
  } else {
    const float eps_t = valid ? P.eps[P.ptype[part] & 7] : 0.f;
    for (int t = 0; t < P.ntrees; t++) {                       // forcetree.c:798-808: every tree in turn
      const float h_inv = 1.0f / (2.8f * fmaxf(P.eps[P.ttype[t]], eps_t));
      no = valid ? P.troot[t] : 0x7fffffff;
      walk_loop<PER, 0>(P, tp, bh, oac, no, ax, ay, az, npart, nnode, wnodes, wparts, h_inv, P.troot[t + 1]);
    }
  }
  if (valid) {
    P.acc[3 * (size_t)slot] = ax; P.acc[3 * (size_t)slot + 1] = ay; P.acc[3 * (size_t)slot + 2] = az;
    P.cost[2 * (size_t)slot] = npart; P.cost[2 * (size_t)slot + 1] = nnode;
  }
  // DIAG counters (forcetree.c:65-66) and the per-warp list lengths of SURVEY.md 8d
  unsigned long long sp = npart, sn = nnode;
  for (int o = 16; o > 0; o >>= 1) { sp += __shfl_down_sync(0xffffffffu, sp, o); sn += __shfl_down_sync(0xffffffffu, sn, o); }
  if (lane == 0) {
    atomicAdd(&P.ctr[CT_PART], sp); atomicAdd(&P.ctr[CT_NODE], sn);
    atomicAdd(&P.ctr[CT_LIST_NODES], (unsigned long long)wnodes); atomicAdd(&P.ctr[CT_LIST_PARTS], (unsigned long long)wparts);
  }
}

// ------------------------------------------------------------------ packed walk over sibling pairs
// Same interaction lists as k_walk (every lane takes its own open/accept decision on every cell it reaches), other
// order of work: when a lane opens a cell, ALL child cells of that cell will be visited by the warp, so they are
// processed together, two at a time with the packed fp32 instructions of sm_100a (FADD2 / FMUL2 / FFMA2: one issue
// slot for two cells).  The warp keeps a LIFO of pending sibling groups {first pair | number of cells, mask of the
// lanes that opened the parent} in shared memory; a group is popped, its <= 4 cell pairs are loaded by broadcast
// (one 128-byte record per pair), every lane of the mask decides and accumulates, opened cells contribute their
// direct particles at once and push their own child group.  The quadrupole arrives pre-scaled (-3 Q, -1.5 P).
// Open boundaries, one tree.  Results differ from k_walk only in the order of the float partial sums.
constexpr int kPairStack = 160;
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

struct PairOut { double ax, ay, az; int npart, nnode; unsigned wnodes, wparts; bool overflow; };
// the softened forms as real calls: rare, and out of the hot loop's register allocation
__device__ __noinline__ float2 pn_soft_call(float r2, float potq, float mass, float pp, float h_inv) { return pn_soft(r2, potq, mass, pp, h_inv); }
__device__ __noinline__ float pp_soft_call(float r2, float mass, float h_inv) { return pp_soft(r2, mass, h_inv); }

template <int MODE>
__device__ __forceinline__ void walk_pairs_body(const WalkParams &P, const PairRec *__restrict__ pairs, uint2 *stk, float4 *stage, const int lane, const bool valid,
                                                const float4 tp, const bool bh, const float oac, PairOut &R) {
  const float2 ntx = f2(-tp.x, -tp.x), nty = f2(-tp.y, -tp.y), ntz = f2(-tp.z, -tp.z), oac2 = f2(oac, oac);
  const float theta2 = P.theta2, h_inv = P.h_inv, h2 = 1.0f / (h_inv * h_inv);
  const float inf = __int_as_float(0x7f800000);
  const unsigned lbit = 1u << lane;
  float2 fx = f2(0, 0), fy = f2(0, 0), fz = f2(0, 0);
  double ax = 0, ay = 0, az = 0;
  int npart = 0, nvisit = 0, nopen = 0, pending = 0;            // accepted cells = cells this lane decided on - cells it opened
  unsigned wnodes = 0, wparts = 0;
  const unsigned vm = __ballot_sync(0xffffffffu, valid);
  int sp = 0;
  if (vm) { stk[0] = make_uint2(1u, vm); sp = 1; }              // the root: pair 0, one cell
  bool overflow = false;
  const float4 *recs = reinterpret_cast<const float4 *>(pairs);
  while (sp > 0) {
    const uint2 e = stk[--sp];
    __syncwarp();                                               // every lane has read the entry before anyone overwrites the slot
    const bool in = (e.y & lbit) != 0;
    const int k = (int)(e.x & 15u);
    // the whole group (<= 4 pair records = 512 bytes) in one coalesced load, staged in shared memory
    if (lane < 4 * (k + 1)) stage[lane] = __ldg(recs + 8 * (size_t)(e.x >> 4) + lane);      // 8 float4 per pair, (k+1)/2 pairs
    __syncwarp();
    const float4 *rec = stage;
    if (in) nvisit += k;
    wnodes += k;
    for (int j = 0; j < k; j += 2, rec += 8) {
      const float4 A = rec[0], B = rec[1], Cv = rec[2], D = rec[3];
      const float4 H = rec[7];                                  // -1.5 P (x2), len2 (x2)
      const float2 dx = __fadd2_rn(f2(A.x, A.y), ntx), dy = __fadd2_rn(f2(A.z, A.w), nty), dz = __fadd2_rn(f2(B.x, B.y), ntz);
      float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
      bool c0, c1;                                              // forcetree.c:967 / :1253-1257
      if (MODE == 2) { c0 = H.z > r2.x * theta2; c1 = H.w > r2.y * theta2; }
      else {
        const float2 o = __fmul2_rn(__fmul2_rn(__fmul2_rn(oac2, r2), r2), r2);
        c0 = (Cv.x > o.x) || (r2.x < Cv.z); c1 = (Cv.y > o.y) || (r2.y < Cv.w);
        if (MODE == 0 && bh) { c0 = H.z > r2.x * theta2; c1 = H.w > r2.y * theta2; }
      }
      const bool in1 = in && (j + 1 < k);
      const bool a0 = in && !c0, a1 = in1 && !c1, o0 = in && c0, o1 = in1 && c1;
      if (a0 || a1) {
        const float4 E = rec[4], F = rec[5], G = rec[6];
        const float2 q11 = f2(E.x, E.y), q22 = f2(E.z, E.w), q33 = f2(F.x, F.y), q12 = f2(F.z, F.w), q13 = f2(G.x, G.y), q23 = f2(G.z, G.w);
        const float2 qx = __ffma2_rn(q13, dz, __ffma2_rn(q12, dy, __fmul2_rn(q11, dx)));
        const float2 qy = __ffma2_rn(q23, dz, __ffma2_rn(q22, dy, __fmul2_rn(q12, dx)));
        const float2 qz = __ffma2_rn(q33, dz, __ffma2_rn(q23, dy, __fmul2_rn(q13, dx)));
        const float2 tt = __ffma2_rn(dz, qz, __ffma2_rn(dy, qy, __fmul2_rn(dx, qx)));          // -3 y^T Q y
        // a half this lane does not accept gets r^2 = inf: 1/r = 0 and every term of it vanishes
        if (!a0) r2.x = inf;
        if (!a1) r2.y = inf;
        // fac = m/r^3 + (15 potq/r^2 - 1.5 P)/r^5, ff = -3/r^5 (forcetree.c:1262-1301) with -3 folded into Q
        const float2 ri = f2(rsqrt_fast(r2.x), rsqrt_fast(r2.y));
        const float2 r2i = __fmul2_rn(ri, ri), r3i = __fmul2_rn(r2i, ri);
        float2 r5i = __fmul2_rn(r3i, r2i);
        float2 fac = __ffma2_rn(__ffma2_rn(__fmul2_rn(tt, r2i), f2(-2.5f, -2.5f), f2(H.x, H.y)), r5i, __fmul2_rn(f2(B.z, B.w), r3i));
        if (r2.x < h2 || r2.y < h2) {                            // softened cell (rare): forcetree.c:1026-1074
          if (r2.x < h2) { const float2 v = pn_soft_call(r2.x, tt.x * (-1.0f / 6), B.z, H.x * (-2.0f / 3), h_inv); fac.x = v.x; r5i.x = v.y * (-1.0f / 3); }
          if (r2.y < h2) { const float2 v = pn_soft_call(r2.y, tt.y * (-1.0f / 6), B.w, H.y * (-2.0f / 3), h_inv); fac.y = v.x; r5i.y = v.y * (-1.0f / 3); }
        }
        fx = __ffma2_rn(r5i, qx, __ffma2_rn(dx, fac, fx));
        fy = __ffma2_rn(r5i, qy, __ffma2_rn(dy, fac, fy));
        fz = __ffma2_rn(r5i, qz, __ffma2_rn(dz, fac, fz));
      }
      const unsigned m0 = __ballot_sync(0xffffffffu, o0), m1 = __ballot_sync(0xffffffffu, o1);
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        const unsigned mo = hh ? m1 : m0;
        if (!mo) continue;                                      // warp-uniform
        const bool open = hh ? o1 : o0;
        const int cinfo = __float_as_int(hh ? D.y : D.x), pinfo = __float_as_int(hh ? D.w : D.z);
        const int np = pinfo & 15;
        const float4 *lp = P.leaf_posm + (pinfo >> 4);
        wparts += np;
#pragma unroll 1
        for (int q = 0; q < np; q++) {
          const float4 pq = __ldg(lp + q);
          if (open) {
            const float px = pq.x - tp.x, py = pq.y - tp.y, pz = pq.z - tp.z;
            const float pr2 = fmaf(pz, pz, fmaf(py, py, px * px));
            const float pri = rsqrt_fast(pr2);
            float pf = pq.w * pri * pri * pri;                  // m/r^3, forcetree.c:1135-1186
            if (pr2 < h2) pf = pp_soft_call(pr2, pq.w, h_inv);
            fx.x = fmaf(px, pf, fx.x); fy.x = fmaf(py, pf, fy.x); fz.x = fmaf(pz, pf, fz.x);
          }
        }
        if (open) { npart += np; nopen++; }
        if (cinfo & 15) {                                       // child cells: one more pending group (every lane writes the same words)
          if (sp < kPairStack) { stk[sp] = make_uint2((unsigned)cinfo, mo); sp++; } else overflow = true;
          // its records towards L1 while the rest of this group is processed (one 128-byte line per pair)
          if (lane < (((cinfo & 15) + 1) >> 1)) asm volatile("prefetch.global.L1 [%0];" ::"l"(recs + 8 * (size_t)((unsigned)cinfo >> 4) + 8 * lane));
        }
      }
    }
    pending += k;
    if (pending >= 24) {                                        // float partial sums -> double accumulators
      ax += (double)fx.x + (double)fx.y; ay += (double)fy.x + (double)fy.y; az += (double)fz.x + (double)fz.y;
      fx = f2(0, 0); fy = f2(0, 0); fz = f2(0, 0); pending = 0;
    }
  }
  ax += (double)fx.x + (double)fx.y; ay += (double)fy.x + (double)fy.y; az += (double)fz.x + (double)fz.y;
  R.ax = ax; R.ay = ay; R.az = az; R.npart = npart; R.nnode = nvisit - nopen; R.wnodes = wnodes; R.wparts = wparts; R.overflow = overflow;
}

template <int MINB>
__global__ void __launch_bounds__(WALK_THREADS, MINB) k_walk_pairs(WalkParams P, const PairRec *__restrict__ pairs) {
  __shared__ uint2 s_stack[WALK_THREADS / 32][kPairStack];
  __shared__ float4 s_stage[WALK_THREADS / 32][32];
  const int lane = threadIdx.x & 31;
  uint2 *stk = s_stack[threadIdx.x >> 5];
  float4 *stage = s_stage[threadIdx.x >> 5];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < P.nt;
  int slot = 0, part = 0;
  if (valid) { slot = P.tsorted[t]; part = P.slot_part ? P.slot_part[slot] : slot; }
  float4 tp = make_float4(0, 0, 0, 0); float oa = 0;
  if (valid) { tp = P.posm[part]; oa = P.oldacc[part]; }
  const bool bh = (P.criterion == 0) || (oa == 0.0f);          // forcetree.c:801
  const float oac = oa * P.alpha;                               // forcetree.c:1129
  const bool all_rel = __all_sync(0xffffffffu, !valid || !bh), all_bh = __all_sync(0xffffffffu, !valid || bh);
  PairOut R;
  if (all_rel) walk_pairs_body<1>(P, pairs, stk, stage, lane, valid, tp, bh, oac, R);
  else if (all_bh) walk_pairs_body<2>(P, pairs, stk, stage, lane, valid, tp, bh, oac, R);
  else walk_pairs_body<0>(P, pairs, stk, stage, lane, valid, tp, bh, oac, R);
  if (valid) {
    P.acc[3 * (size_t)slot] = R.ax; P.acc[3 * (size_t)slot + 1] = R.ay; P.acc[3 * (size_t)slot + 2] = R.az;
    P.cost[2 * (size_t)slot] = R.npart; P.cost[2 * (size_t)slot + 1] = R.nnode;
  }
  unsigned long long spp = R.npart, sn = R.nnode;
  for (int o = 16; o > 0; o >>= 1) { spp += __shfl_down_sync(0xffffffffu, spp, o); sn += __shfl_down_sync(0xffffffffu, sn, o); }
  if (lane == 0) {
    atomicAdd(&P.ctr[CT_PART], spp); atomicAdd(&P.ctr[CT_NODE], sn);
    atomicAdd(&P.ctr[CT_LIST_NODES], (unsigned long long)R.wnodes); atomicAdd(&P.ctr[CT_LIST_PARTS], (unsigned long long)R.wparts);
    if (R.overflow) atomicAdd(&P.ctr[CT_WALK_OVF], 1ull);
  }
}

static float h_inv_of_type1();
static void fill_trees(WalkParams &P);

// ------------------------------------------------------------------ potential walk
// force_treeevaluate_potential(), forcetree.c:1389-1755: the same lock-step walk and the same open/accept
// decisions as the force walk, scalar accumulator.  Leaf: -m/r, or m/h * knlpot(r/h) inside the softening
// radius (no u > 1e-4 guard here: the target's own particle contributes -m/eps, which compute_potential()
// adds back, potential.c:135).  Cell: -M/r + (-3 potq/r^2 + P/2)/r^3, softened form below h.  Open boundaries.
template <bool PER, int MODE>
__device__ __forceinline__ void walk_loop_pot(const WalkParams &P, const float4 tp, const bool bh, const float oac, int &no, double &pot,
                                              const float h_inv, const int M) {
  const float theta2 = P.theta2;
  const float h2 = 1.0f / (h_inv * h_inv);
  const float h3i = h_inv * h_inv * h_inv, h5i = h3i * h_inv * h_inv;
  const float4 *nodes4 = reinterpret_cast<const float4 *>(P.nodes);
  int cur = __reduce_min_sync(0xffffffffu, no);
  while (cur < M) {
    float f = 0;
    for (int it = 0; it < kFlushEvery && cur < M; it++) {
      const float4 *nd = nodes4 + 4 * (size_t)cur;
      const float4 A = __ldg(nd), Bv = __ldg(nd + 1), Cv = __ldg(nd + 2), Dv = __ldg(nd + 3);
      float dx = A.x - tp.x, dy = A.y - tp.y, dz = A.z - tp.z;
      if (PER) { dx = wrap_image(dx, P.box, P.boxhalf); dy = wrap_image(dy, P.box, P.boxhalf); dz = wrap_image(dz, P.box, P.boxhalf); }
      const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      bool crit;
      if (MODE == 1) crit = (Bv.x > oac * r2 * r2 * r2) || (r2 < Bv.y);
      else if (MODE == 2) crit = Dv.w > r2 * theta2;
      else crit = bh ? (Dv.w > r2 * theta2) : ((Bv.x > oac * r2 * r2 * r2) || (r2 < Bv.y));
      const bool act = (no == cur);
      const bool acc = act && !crit, open = act && crit;
      if (acc) {
        const float qx = fmaf(Dv.x, dz, fmaf(Cv.w, dy, Cv.x * dx));
        const float qy = fmaf(Dv.y, dz, fmaf(Cv.y, dy, Cv.w * dx));
        const float qz = fmaf(Cv.z, dz, fmaf(Dv.y, dy, Dv.x * dx));
        const float potq = 0.5f * fmaf(dz, qz, fmaf(dy, qy, dx * qx));     // 1/2 y^T Q y
        if (r2 >= h2) {
          const float ri = rsqrt_fast(r2), r2i = ri * ri, r3i = r2i * ri;
          f += fmaf(r3i, fmaf(-3.0f * potq, r2i, 0.5f * Dv.z), -A.w * ri);   // forcetree.c:1694-1695
        } else {
          const float u = sqrtf(r2) * h_inv;
          float w2, w3, w4;
          soft_w234(u, w2, w3, w4);
          f += A.w * h_inv * soft_pot(u) + potq * w2 * h5i + 0.5f * Dv.z * soft_force(u) * h3i;   // :1722-1723
        }
        if (PER) f += A.w * ewald_pot_corr(P.ewald, P.ewald_fac, dx, dy, dz);        // :1726-1728
      }
      no = acc ? __float_as_int(Bv.w) : (open ? cur + 1 : no);
      if (__any_sync(0xffffffffu, open)) {
        const int pinfo = __float_as_int(Bv.z);
        const int np = pinfo & 15;
        const float4 *lp = P.leaf_posm + (pinfo >> 4);
        for (int k = 0; k < np; k++) {
          const float4 q = __ldg(lp + k);
          if (open) {
            float px = q.x - tp.x, py = q.y - tp.y, pz = q.z - tp.z;
            if (PER) { px = wrap_image(px, P.box, P.boxhalf); py = wrap_image(py, P.box, P.boxhalf); pz = wrap_image(pz, P.box, P.boxhalf); }
            const float pr2 = fmaf(pz, pz, fmaf(py, py, px * px));
            if (pr2 >= h2) f -= q.w * rsqrt_fast(pr2);                       // forcetree.c:1625
            else f += q.w * h_inv * soft_pot(sqrtf(pr2) * h_inv);            // :1629-1631
            if (PER) f += q.w * ewald_pot_corr(P.ewald, P.ewald_fac, px, py, pz);   // :1633-1635 (the target itself included)
          }
        }
      }
      cur = __reduce_min_sync(0xffffffffu, no);
    }
    pot += (double)f;
  }
}

template <bool PER>
__global__ void __launch_bounds__(128) k_walk_pot(WalkParams P) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < P.nt;
  int slot = 0, part = 0;
  if (valid) { slot = P.tsorted[t]; part = P.slot_part ? P.slot_part[slot] : slot; }
  float4 tp = make_float4(0, 0, 0, 0); float oa = 0;
  if (valid) { tp = P.posm[part]; oa = P.oldacc[part]; }
  const bool bh = (P.criterion == 0) || (oa == 0.0f);           // forcetree.c:1404
  const float oac = oa * P.alpha;
  int no = valid ? 0 : 0x7fffffff;
  double pot = 0;
  const bool all_rel = __all_sync(0xffffffffu, !valid || !bh), all_bh = __all_sync(0xffffffffu, !valid || bh);
  const float eps_t = valid ? P.eps[P.ptype[part] & 7] : 0.f;
  for (int t = 0; t < P.ntrees; t++) {                         // forcetree.c:1397-1409: every tree in turn
    const float h_inv = P.ntrees > 1 ? 1.0f / (2.8f * fmaxf(P.eps[P.ttype[t]], eps_t)) : P.h_inv;
    no = valid ? P.troot[t] : 0x7fffffff;
    if (all_rel) walk_loop_pot<PER, 1>(P, tp, bh, oac, no, pot, h_inv, P.troot[t + 1]);
    else if (all_bh) walk_loop_pot<PER, 2>(P, tp, bh, oac, no, pot, h_inv, P.troot[t + 1]);
    else walk_loop_pot<PER, 0>(P, tp, bh, oac, no, pot, h_inv, P.troot[t + 1]);
  }
  if (valid) P.acc[slot] = pot;                                  // raw potential per target slot (GravDataPotential)
}

// compute_potential(), potential.c:131-168: float Potential <- raw; += m/eps (self energy); *G; Lambda / comoving terms
__global__ void k_pot_epilogue(int n, const double *raw, const float4 *posm, float *potential, const int *ptype, EpsTab tab, double G,
                               int comoving, int periodic, double H, double O0, double OL) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = posm[i];
  float v = (float)raw[i];
  v = (float)((double)v + (double)p.w / tab.e[ptype[i] & 7]);      // P[i].Mass / All.SofteningTable[P[i].Type]
  double r2 = 0;
  r2 += (double)fmul(p.x, p.x); r2 += (double)fmul(p.y, p.y); r2 += (double)fmul(p.z, p.z);
  // products and sums individually rounded like the reference's C expressions (no DFMA contraction)
  if (comoving) {
    const double fac = 0.5 * O0 * H * H;
    v = periodic ? (float)__dmul_rn(G, (double)v) : (float)__dsub_rn(__dmul_rn(G, (double)v), __dmul_rn(fac, r2));        // potential.c:141-150
  } else {
    const double fac = -0.5 * OL * H * H;
    v = (float)__dmul_rn((double)v, G);
    if (fac != 0) v = (float)__dadd_rn((double)v, __dmul_rn(fac, r2));
  }
  potential[i] = v;
}

static int potential_walk(const int *d_sorted, int nt, bool with_slots) {
  const bool per = g.par.PeriodicBoundariesOn && g.par.BoxSize > 0;
  WalkParams P;
  P.nt = nt; P.num_nodes = g.num_nodes; P.tsorted = d_sorted; P.slot_part = with_slots ? g.d_active : nullptr;
  P.posm = g.posm; P.oldacc = g.oldacc; P.nodes = g.nodes; P.leaf_posm = g.leaf_posm;
  P.acc = g.d_acc; P.cost = g.d_cost;
  P.theta2 = (float)(g.par.ErrTolTheta * g.par.ErrTolTheta); P.alpha = (float)g.par.ErrTolForceAcc;
  P.h_inv = h_inv_of_type1(); P.criterion = g.par.TypeOfOpeningCriterion; P.ctr = g.d_ctr;
  fill_trees(P);
  P.box = (float)g.par.BoxSize; P.boxhalf = (float)(g.par.BoxSize / 2); P.ewald_fac = per ? (float)(kEwaldN / g.par.BoxSize) : 0.f; P.ewald = nullptr;
  if (per) { B200_TRY(ewald_tables(&P.ewald)); }
  if (nt > 0) {
    if (per) k_walk_pot<true><<<cdiv(nt, 128), 128, 0, g.stream>>>(P); else k_walk_pot<false><<<cdiv(nt, 128), 128, 0, g.stream>>>(P);
    count_launch();
  }
  return B200_OK;
}

// ------------------------------------------------------------------ Ewald tables
// ewald_init() / ewald_force(), ewald.c:35-162, 332-381: correction force of the periodic images
// of a unit point mass in a unit box (alpha = 2, |n|,|h| <= 4 per axis), tabulated on the
// (ED+1)^3 grid of the first octant, then scaled by 1/L^2 (ewald.c:145-155).  One thread per
// grid point, double precision, stored as float like the reference's tables.
__global__ void k_ewald_table(float4 *tab, double inv_box2, double inv_box) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int S1 = kEwaldD + 1;
  if (idx >= S1 * S1 * S1) return;
  const int i = idx / (S1 * S1), j = (idx / S1) % S1, k = idx % S1;
  const double PI = 3.14159265358979323846, alpha = 2.0;
  const double x[3] = {(double)i / kEwaldN, (double)j / kEwaldN, (double)k / kEwaldN};
  double f[3] = {0, 0, 0};
  const double r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  // ewald_psi(), ewald.c:291-325: potential correction; the origin holds the constant 2.8372975 (ewald.c:102-103)
  double psi = 2.8372975;
  if (idx != 0) {
    double sum1 = 0, sum2 = 0;
    for (int n0 = -4; n0 <= 4; n0++) for (int n1 = -4; n1 <= 4; n1++) for (int n2 = -4; n2 <= 4; n2++) {
      const double d[3] = {x[0] - n0, x[1] - n1, x[2] - n2};
      const double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      sum1 += erfc(alpha * r) / r;
    }
    for (int h0 = -4; h0 <= 4; h0++) for (int h1 = -4; h1 <= 4; h1++) for (int h2_ = -4; h2_ <= 4; h2_++) {
      const int hh = h0 * h0 + h1 * h1 + h2_ * h2_;
      if (hh > 0) sum2 += 1 / (PI * hh) * exp(-PI * PI * hh / (alpha * alpha)) * cos(2 * PI * (x[0] * h0 + x[1] * h1 + x[2] * h2_));
    }
    psi = PI / (alpha * alpha) - sum1 - sum2 + 1 / sqrt(r2);
  }
  if (r2 != 0) {
    for (int a = 0; a < 3; a++) f[a] += x[a] / (r2 * sqrt(r2));
    for (int n0 = -4; n0 <= 4; n0++) for (int n1 = -4; n1 <= 4; n1++) for (int n2 = -4; n2 <= 4; n2++) {
      const double d[3] = {x[0] - n0, x[1] - n1, x[2] - n2};
      const double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      const double val = erfc(alpha * r) + 2 * alpha * r / sqrt(PI) * exp(-alpha * alpha * r * r);
      for (int a = 0; a < 3; a++) f[a] -= d[a] / (r * r * r) * val;
    }
    for (int h0 = -4; h0 <= 4; h0++) for (int h1 = -4; h1 <= 4; h1++) for (int h2_ = -4; h2_ <= 4; h2_++) {
      const int hh = h0 * h0 + h1 * h1 + h2_ * h2_;
      if (hh > 0) {
        const double hdotx = x[0] * h0 + x[1] * h1 + x[2] * h2_;
        const double val = 2.0 / ((double)hh) * exp(-PI * PI * hh / (alpha * alpha)) * sin(2 * PI * hdotx);
        f[0] -= h0 * val; f[1] -= h1 * val; f[2] -= h2_ * val;
      }
    }
  }
  // the reference stores the unit-box value as float, then divides the float by L^2 (a double)
  tab[idx] = make_float4((float)((double)(float)f[0] * inv_box2), (float)((double)(float)f[1] * inv_box2), (float)((double)(float)f[2] * inv_box2),
                         (float)((double)(float)psi * inv_box));
}

int ewald_tables(const float4 **out) {
  const int S1 = kEwaldD + 1, n = S1 * S1 * S1;
  if (!g.d_ewald && cudaMalloc((void **)&g.d_ewald, (size_t)n * sizeof(float4)) != cudaSuccess) return B200_ERR_ALLOC;
  if (g.ewald_box != g.par.BoxSize) {
    k_ewald_table<<<cdiv(n, 128), 128, 0, g.stream>>>(g.d_ewald, 1.0 / (g.par.BoxSize * g.par.BoxSize), 1.0 / g.par.BoxSize);
    count_launch();
    g.ewald_box = g.par.BoxSize;
  }
  *out = g.d_ewald;
  return B200_OK;
}

// sorted target list for an explicit active list: order the slots along the key order so a
// warp's 32 targets are spatial neighbours
__global__ void k_target_keys(int nt, const int *active, const int *krank, int *keys, int *vals) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nt) { keys[t] = krank[active[t]]; vals[t] = t; }
}

int prepare_targets(const int *active_host, int nactive, int **d_sorted_out) {
  if (!active_host) { *d_sorted_out = g.sidx; return B200_OK; }   // all particles, key order, slot == particle
  CUDA_TRY(cudaMemcpyAsync(g.d_active, active_host, (size_t)nactive * sizeof(int), cudaMemcpyHostToDevice, g.stream));
  static const bool keep_order = getenv("B200_KEEP_TARGET_ORDER") != nullptr;    // experiment: targets grouped into warps as listed
  if (keep_order) { CUDA_TRY(cudaMemcpyAsync(g.d_tsorted, g.iota, (size_t)nactive * sizeof(int), cudaMemcpyDeviceToDevice, g.stream)); *d_sorted_out = g.d_tsorted; return B200_OK; }
  k_target_keys<<<cdiv(nactive, 256), 256, 0, g.stream>>>(nactive, g.d_active, g.krank, g.d_tkeys, g.d_tvals2);
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, g.d_tkeys, g.d_tkeys2, g.d_tvals2, g.d_tsorted, nactive, 0, 32, g.stream);
  if (tb > g.cub_tmp_bytes) {
    if (g.cub_tmp) cudaFree(g.cub_tmp);
    g.cub_tmp = nullptr; g.cub_tmp_bytes = 0;
    if (cudaMalloc(&g.cub_tmp, tb + 4096) != cudaSuccess) return B200_ERR_ALLOC;
    g.cub_tmp_bytes = tb + 4096;
  }
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(g.cub_tmp, tb, g.d_tkeys, g.d_tkeys2, g.d_tvals2, g.d_tsorted, nactive, 0, 32, g.stream));
  count_launch(5);
  *d_sorted_out = g.d_tsorted;
  return B200_OK;
}

static void fill_trees(WalkParams &P) {
  P.ntrees = g.ntrees; P.ptype = g.ptype;
  for (int t = 0; t < 7; t++) P.troot[t] = g.tree_root[t];
  for (int t = 0; t < 6; t++) { P.ttype[t] = g.tree_type[t]; P.eps[t] = (float)g.par.SofteningTable[t]; }
  if (g.ntrees == 1) { P.troot[0] = 0; P.troot[1] = g.num_nodes; }
}

static float h_inv_of_type1() {
  // epsilon = max(eps_tree, eps_target) (forcetree.c:800); one collisionless type => its own eps
  double eps = g.par.SofteningTable[g.tree_type[0]];
  if (!(eps > 0)) for (int t = 0; t < 6; t++) if (g.par.SofteningTable[t] > eps) eps = g.par.SofteningTable[t];
  return (float)(1.0 / (2.8 * eps));
}

// walk for nt targets; d_sorted = sorted slot list; with_slots: slot->particle through d_active
static int walk_read_counters() {
  cudaEventElapsedTime(&g.cnt.ms_walk, g.ev0, g.ev1);
  g.cnt.part_interactions = (long long)g.h_ctr[CT_PART]; g.cnt.node_interactions = (long long)g.h_ctr[CT_NODE];
  g.cnt.list_nodes = (long long)g.h_ctr[CT_LIST_NODES]; g.cnt.list_parts = (long long)g.h_ctr[CT_LIST_PARTS];
  if (g.h_ctr[CT_WALK_OVF]) {
    fprintf(stderr, "libsidm_b200: %llu warps overflowed the %d-entry group stack of the packed walk; b200_set_option(\"walk_pairs\", 0) selects the stack-free walk\n",
            g.h_ctr[CT_WALK_OVF], kPairStack);
    return B200_ERR_STATE;
  }
  return B200_OK;
}

int walk_impl(const int *d_sorted, int nt, bool with_slots, bool defer_sync) {
  WalkParams P;
  P.nt = nt; P.num_nodes = g.num_nodes; P.tsorted = d_sorted; P.slot_part = with_slots ? g.d_active : nullptr;
  P.posm = g.posm; P.oldacc = g.oldacc; P.nodes = g.nodes; P.leaf_posm = g.leaf_posm;
  P.acc = g.d_acc; P.cost = g.d_cost;
  P.theta2 = (float)(g.par.ErrTolTheta * g.par.ErrTolTheta); P.alpha = (float)g.par.ErrTolForceAcc;
  P.h_inv = h_inv_of_type1(); P.criterion = g.par.TypeOfOpeningCriterion; P.ctr = g.d_ctr;
  fill_trees(P);
  CUDA_TRY(cudaMemsetAsync(g.d_ctr, 0, kWalkCounters * sizeof(unsigned long long), g.stream));
  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  const bool per = g.par.PeriodicBoundariesOn && g.par.BoxSize > 0;
  P.box = (float)g.par.BoxSize; P.boxhalf = (float)(g.par.BoxSize / 2); P.ewald_fac = per ? (float)(kEwaldN / g.par.BoxSize) : 0.f; P.ewald = nullptr;
  if (per) { B200_TRY(ewald_tables(&P.ewald)); }
  if (nt > 0) {
    const int GW = cdiv(nt, WALK_THREADS);
    if (g.pairs_valid && g.opt_walk_pairs && g.ntrees == 1 && !per) {
      switch (g.opt_walkp_minb) {                 // registers per thread: 64 / 80 / 96 / 128 (occupancy A/B, option "walkp_minb")
        case 8: k_walk_pairs<8><<<GW, WALK_THREADS, 0, g.stream>>>(P, g.pairs); break;
        case 5: k_walk_pairs<5><<<GW, WALK_THREADS, 0, g.stream>>>(P, g.pairs); break;
        case 4: k_walk_pairs<4><<<GW, WALK_THREADS, 0, g.stream>>>(P, g.pairs); break;
        default: k_walk_pairs<6><<<GW, WALK_THREADS, 0, g.stream>>>(P, g.pairs); break;
      }
    }
    else if (g.ntrees > 1) { if (per) k_walk<true, true><<<GW, WALK_THREADS, 0, g.stream>>>(P); else k_walk<false, true><<<GW, WALK_THREADS, 0, g.stream>>>(P); }
    else if (per) k_walk<true, false><<<GW, WALK_THREADS, 0, g.stream>>>(P);
    else k_walk<false, false><<<GW, WALK_THREADS, 0, g.stream>>>(P);
    count_launch();
  }
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  CUDA_TRY(cudaMemcpyAsync(g.h_ctr, g.d_ctr, kWalkCounters * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g.stream));
  g.cnt.num_targets = nt;
  g.cnt.num_lists = (nt + 31) / 32;
  if (defer_sync) { g.walk_pending = true; return B200_OK; }      // gravity_finish() reads the counters
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  return walk_read_counters();
}

// gravtree.c:230-324: Accel <- (float)Acc; OldAcc = |Accel| before G (relative criterion);
// Accel <- G*Accel + OmegaLambda*H^2*PosPred, or the comoving combination.
struct EpiParams {
  int nt; const int *list; const int *slot_part; const double *acc; const int *cost;
  const float4 *posm; const float *velpred; float *accel, *oldacc, *gravcost;
  int criterion, comoving, periodic; double G, H, O0, OL, time;
};
__global__ void k_grav_epilogue(EpiParams E) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E.nt) return;
  const int s = E.list ? E.list[k] : k;
  const int p = E.slot_part ? E.slot_part[s] : s;
  float a[3];
  for (int k = 0; k < 3; k++) a[k] = (float)E.acc[3 * (size_t)s + k];
  const float4 pp = E.posm[p];
  const float pos[3] = {pp.x, pp.y, pp.z};
  if (!E.comoving) {
    if (E.criterion == 1)
      E.oldacc[p] = (float)sqrt((double)fadd(fadd(fmul(a[0], a[0]), fmul(a[1], a[1])), fmul(a[2], a[2])));
    const double fac1 = E.OL * E.H * E.H;
    for (int k = 0; k < 3; k++) E.accel[3 * (size_t)p + k] = (float)__dadd_rn(__dmul_rn(E.G, (double)a[k]), __dmul_rn(fac1, (double)pos[k]));
  } else {
    if (E.criterion == 1) {
      const double fac3 = 0.5 * E.H * E.H * E.O0 / E.G;
      double a2 = 0;
      for (int k = 0; k < 3; k++) { const double x = E.periodic ? (double)a[k] : __dadd_rn((double)a[k], __dmul_rn(fac3, (double)pos[k])); a2 = __dadd_rn(a2, __dmul_rn(x, x)); }
      E.oldacc[p] = (float)sqrt(a2);
    }
    const double t = E.time;
    const double s_a = sqrt(E.O0 + t * (1 - E.O0 - E.OL) + t * t * t * E.OL);
    const double fac1 = E.G / (E.H * t * t * s_a), fac2 = -1.5 / t, fac3 = 0.5 * E.H * E.O0 / (t * t * s_a);
    for (int k = 0; k < 3; k++) {
      double v = __dadd_rn(__dmul_rn(fac1, (double)a[k]), __dmul_rn(fac2, (double)E.velpred[3 * (size_t)p + k]));
      if (!E.periodic) v = __dadd_rn(v, __dmul_rn(fac3, (double)pos[k]));
      E.accel[3 * (size_t)p + k] = (float)v;
    }
  }
  E.gravcost[p] = (float)(E.cost[2 * (size_t)s] + E.cost[2 * (size_t)s + 1]);
}

// multi-GPU: {Accel, OldAcc} of this rank's targets -> send buffer, and back from all ranks
__global__ void k_grav_pack(int nw, const int *work, const int *slot_part, const float *accel, const float *oldacc, float4 *send) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nw) return;
  const int s = work[k]; const int p = slot_part ? slot_part[s] : s;
  send[k] = make_float4(accel[3 * (size_t)p], accel[3 * (size_t)p + 1], accel[3 * (size_t)p + 2], oldacc[p]);
}
__global__ void k_grav_unpack(int nt, int world, int per_rank, const int *sorted, const int *slot_part, const float4 *recv, float *accel, float *oldacc) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)world * per_rank) return;
  const int q = (int)(g / per_rank), k = (int)(g % per_rank);
  const long long j = ((long long)(k >> kShardShift) * world + q) * kShardBlock + (k & (kShardBlock - 1));
  if (j >= nt) return;
  const int s = sorted[j]; const int p = slot_part ? slot_part[s] : s;
  const float4 v = recv[g];
  accel[3 * (size_t)p] = v.x; accel[3 * (size_t)p + 1] = v.y; accel[3 * (size_t)p + 2] = v.z; oldacc[p] = v.w;
}

// deferred multi-GPU exchange of {Accel, OldAcc}
static struct { bool pending = false; int nt = 0, nw = 0; const int *work = nullptr, *sorted = nullptr, *slot_part = nullptr; } GX;
static int gravity_exchange() {
  if (!GX.pending) return B200_OK;
  GX.pending = false;
  const int per_rank = shard_max_blocks(GX.nt, g.shard_world) * kShardBlock;
  if (GX.nw > 0) { k_grav_pack<<<cdiv(GX.nw, 256), 256, 0, g.stream>>>(GX.nw, GX.work, GX.slot_part, g.accel, g.oldacc, (float4 *)g.shard_send); count_launch(); }
  B200_TRY(shard_exchange((long long)per_rank * sizeof(float4), g.stream));
  const long long tot = (long long)g.shard_world * per_rank;
  k_grav_unpack<<<cdiv(tot, 256), 256, 0, g.stream>>>(GX.nt, g.shard_world, per_rank, GX.sorted, GX.slot_part, (const float4 *)g.shard_recv, g.accel, g.oldacc);
  count_launch();
  return B200_OK;
}

int gravity_exchange_early() {
  if (!GX.pending) return B200_OK;
  g.shard_busy = true;
  return gravity_exchange();
}

int gravity_finish() {
  B200_TRY(gravity_exchange());
  g.shard_busy = false;
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  if (g.walk_pending) { g.walk_pending = false; return walk_read_counters(); }
  return B200_OK;
}

int gravity_impl(const int *active, int nactive, double time, bool defer_sync) {
  if (!g.tree_valid) return B200_ERR_STATE;
  const int nt = active ? nactive : g.n;
  if (nt <= 0) return B200_OK;
  int *d_sorted = nullptr;
  B200_TRY(prepare_targets(active, nt, &d_sorted));
  const int *work = d_sorted; int nw = nt;
  const bool sharded = g.shard_world > 1 && nt >= g.shard_min_work;
  if (sharded) { B200_TRY(shard_select(d_sorted, nt, g.d_shard_list, &nw, g.stream)); work = g.d_shard_list; }
  B200_TRY(walk_impl(work, nw, active != nullptr, defer_sync));
  EpiParams E;
  // all particles on one rank: run the epilogue in particle order (coalesced) instead of key order
  E.nt = nw; E.list = (!active && !sharded) ? nullptr : work; E.slot_part = active ? g.d_active : nullptr; E.acc = g.d_acc; E.cost = g.d_cost;
  E.posm = g.posm; E.velpred = g.velpred; E.accel = g.accel; E.oldacc = g.oldacc; E.gravcost = g.gravcost;
  E.criterion = g.par.TypeOfOpeningCriterion; E.comoving = g.par.ComovingIntegrationOn;
  E.periodic = g.par.PeriodicBoundariesOn && g.par.BoxSize > 0;
  E.G = g.par.G; E.H = g.par.Hubble; E.O0 = g.par.Omega0; E.OL = g.par.OmegaLambda; E.time = time;
  if (nw > 0) { k_grav_epilogue<<<cdiv(nw, 256), 256, 0, g.stream>>>(E); count_launch(); }
  if (sharded) {
    // all-gather of the partial results (the reduce step of gravtree.c:208-222 becomes a gather:
    // every target is evaluated completely by exactly one rank).  When the SIDM chain runs next to
    // the walk the exchange is issued by gravity_finish(), after the SIDM collectives: collectives
    // of one communicator run in issue order, and this one has to wait for the walk.
    GX.pending = true; GX.nt = nt; GX.nw = nw; GX.work = work; GX.sorted = d_sorted; GX.slot_part = E.slot_part;
    if (!defer_sync) B200_TRY(gravity_exchange());
  }
  if (defer_sync) return B200_OK;
  return gravity_finish();
}

// ------------------------------------------------------------------ direct summation
// forcetree.c:1896-1975 in double, spline evaluated analytically; one thread per target,
// sources tiled through shared memory.
__device__ __forceinline__ double soft_force_d(double u) {
  if (u <= 0.5) return 32.0 * (1.0 / 3 - 6.0 / 5 * u * u + u * u * u);
  return 64.0 * (1.0 / 3 - 3.0 / 4 * u + 3.0 / 5 * u * u - u * u * u / 6) - 1.0 / 15 / (u * u * u);
}
__global__ void __launch_bounds__(128) k_direct(int nt, const int *targets, int n, const float4 *posm, double h_inv_one, double *acc,
                                                double box, const float4 *ewald, float ewald_fac, const int *ptype, EpsTab tab, int ntypes) {
  __shared__ float4 tile[128];
  __shared__ int ttile[128];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  float4 tp = make_float4(0, 0, 0, 0);
  double eps_t = 0;
  if (t < nt) { tp = posm[targets[t]]; eps_t = tab.e[ptype[targets[t]] & 7]; }
  double ax = 0, ay = 0, az = 0;
  for (int base = 0; base < n; base += 128) {
    const int j = base + threadIdx.x;
    tile[threadIdx.x] = j < n ? posm[j] : make_float4(0, 0, 0, 0);
    ttile[threadIdx.x] = j < n ? (ptype[j] & 7) : 0;
    __syncthreads();
    const int lim = min(128, n - base);
    for (int k = 0; k < lim; k++) {
      const float4 q = tile[k];
      double dx = (double)q.x - (double)tp.x, dy = (double)q.y - (double)tp.y, dz = (double)q.z - (double)tp.z;
      if (box > 0) {                     // forcetree.c:1931-1938
        const double bh = 0.5 * box;
        while (dx > bh) dx -= box; while (dy > bh) dy -= box; while (dz > bh) dz -= box;
        while (dx < -bh) dx += box; while (dy < -bh) dy += box; while (dz < -bh) dz += box;
      }
      const double r2 = dx * dx + dy * dy + dz * dz;
      // forcetree.c:1910: epsilon = max(eps of the source's type, eps of the target's type)
      const double h_inv = ntypes > 1 ? 1.0 / (2.8 * fmax(tab.e[ttile[k]], eps_t)) : h_inv_one;
      const double r = sqrt(r2), u = r * h_inv;
      double fac = 0;
      if (u >= 1) fac = (double)q.w / (r2 * r);
      else if (u > 1.0e-4) fac = (double)q.w * h_inv * h_inv * h_inv * soft_force_d(u);
      ax += dx * fac; ay += dy * fac; az += dz * fac;
      if (box > 0 && u > 1.0e-4) {       // forcetree.c:1963-1971
        float ex, ey, ez;
        ewald_corr(ewald, ewald_fac, (float)dx, (float)dy, (float)dz, ex, ey, ez);
        ax += (double)q.w * ex; ay += (double)q.w * ey; az += (double)q.w * ez;
      }
    }
    __syncthreads();
  }
  if (t < nt) { acc[3 * (size_t)t] = ax; acc[3 * (size_t)t + 1] = ay; acc[3 * (size_t)t + 2] = az; }
}

int direct_impl(const int *targets, int nt, double *acc_out) {
  if (nt <= 0) return B200_OK;
  if (nt > g.maxpart) return B200_ERR_ARG;
  CUDA_TRY(cudaMemcpyAsync(g.d_active, targets, (size_t)nt * sizeof(int), cudaMemcpyHostToDevice, g.stream));
  const bool per = g.par.PeriodicBoundariesOn && g.par.BoxSize > 0;
  const float4 *ew = nullptr;
  if (per) { B200_TRY(ewald_tables(&ew)); }
  EpsTab tab;
  for (int t = 0; t < 8; t++) tab.e[t] = t < 6 ? g.par.SofteningTable[t] : 0.0;
  k_direct<<<cdiv(nt, 128), 128, 0, g.stream>>>(nt, g.d_active, g.n, g.posm, (double)h_inv_of_type1(), g.d_acc,
                                                per ? g.par.BoxSize : 0.0, ew, per ? (float)(kEwaldN / g.par.BoxSize) : 0.f, g.ptype, tab, g.ntypes);
  count_launch();
  CUDA_TRY(cudaMemcpyAsync(acc_out, g.d_acc, (size_t)nt * 3 * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
using namespace b200;

extern "C" int b200_gravity(const int *active, int nactive, double time) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  if (active && (nactive < 0 || nactive > g.n)) return B200_ERR_ARG;
  return gravity_impl(active, nactive, time);
}

extern "C" int b200_direct(const int *targets, int n, double *acc_out) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  if (!targets || !acc_out) return B200_ERR_ARG;
  return direct_impl(targets, n, acc_out);
}

extern "C" int b200_walk_raw(const int *targets, int n, double *acc_out, int *cost_out) {
  if (!g.ready || g.n <= 0 || !g.tree_valid) return B200_ERR_STATE;
  if (!targets || n <= 0 || n > g.n) return B200_ERR_ARG;
  int *d_sorted = nullptr;
  B200_TRY(prepare_targets(targets, n, &d_sorted));
  B200_TRY(walk_impl(d_sorted, n, true));
  if (acc_out) CUDA_TRY(cudaMemcpyAsync(acc_out, g.d_acc, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  if (cost_out) CUDA_TRY(cudaMemcpyAsync(cost_out, g.d_cost, (size_t)n * 2 * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  return B200_OK;
}

extern "C" int b200_potential_raw(const int *targets, int n, double *pot_out) {
  if (!g.ready || g.n <= 0 || !g.tree_valid) return B200_ERR_STATE;
  if (!targets || n <= 0 || n > g.n || !pot_out) return B200_ERR_ARG;
  int *d_sorted = nullptr;
  B200_TRY(prepare_targets(targets, n, &d_sorted));
  B200_TRY(potential_walk(d_sorted, n, true));
  CUDA_TRY(cudaMemcpyAsync(pot_out, g.d_acc, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_compute_potential(float *pot_out) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  B200_TRY(b200_tree_build());                                   // potential.c:47 force_treebuild()
  B200_TRY(potential_walk(g.sidx, g.n, false));
  EpsTab tab;
  for (int t = 0; t < 8; t++) tab.e[t] = t < 6 ? g.par.SofteningTable[t] : 0.0;
  k_pot_epilogue<<<cdiv(g.n, 256), 256, 0, g.stream>>>(g.n, g.d_acc, g.posm, g.potential, g.ptype, tab, g.par.G, g.par.ComovingIntegrationOn,
                                                     (g.par.PeriodicBoundariesOn && g.par.BoxSize > 0) ? 1 : 0, g.par.Hubble, g.par.Omega0, g.par.OmegaLambda);
  count_launch();
  if (pot_out) CUDA_TRY(cudaMemcpyAsync(pot_out, g.potential, (size_t)g.n * sizeof(float), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
