// tree_logic.h - the arithmetic of the octree build and of one gravitational interaction,
// written once as host+device inline functions.  The CUDA kernels (tree_build.cu, walk.cu)
// call these per thread / per lane; tests/hostcheck compiles the very same functions with
// g++ to check the construction logic against the reference on machines without a GPU.
// (That host build is a test fixture; the product library has no CPU path.)
//
// Geometry contract (what makes the tree identical to the reference's, forcetree.c:166-345):
//   root  : centre = (float)((xmax+xmin)/2), len = (float)(1.01*max extent), both evaluated
//           in double from float coordinates (forcetree.c:179-212);
//   child : len_c = len_p/2 ; centre_c = centre_p (+/-) len_c/2, evaluated in FLOAT and
//           rounded at every level (forcetree.c:300-306);
//   octant: bit k set iff pos[k] > centre[k], strict (forcetree.c:254-256);
//   leaves: exactly one particle.
// A particle's path through the tree therefore depends only on its own position and the
// root box, so every particle can compute its own 3-bit-per-level key independently; the
// internal nodes are exactly the key prefixes shared by >= 2 particles (plus the root).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

constexpr int kLevelsPerWord = 21;          // 63 bits of a 64-bit word
constexpr int kMaxLevels = 42;              // two words; beyond this particles count as coincident

struct RootBox { float cx, cy, cz, len; };

// float add that the compiler may not fuse or re-associate (keeps host and device identical)
B200_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
B200_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}

// ---- periodic boxes (neighbour searches, sidm.cu)
// ngb_periodic(), forcetree.c:1999-2006: a float coordinate difference wrapped into [-Box/2, Box/2] with double Box / BoxHalf,
// rounded back to float
B200_HD float wrap_periodic(float x, double box) {
  const double bh = 0.5 * box;
  while ((double)x > bh) x = (float)((double)x - box);
  while ((double)x < -bh) x = (float)((double)x + box);
  return x;
}
// Is the search cube [lo, hi] so far from the faces of the box [0, Box)^3 that no periodic image of any particle can lie in it?
// `dom` = min xyz, max xyz of the particle coordinates: between two do_box_wrapping() calls (run.c:135) particles may stick out
// of the box by e = max(0, -min, max - Box); their images lie within e of the opposite face.  A cube that keeps more than e
// (plus a rounding margin) from every face contains no image, every coordinate difference to a particle inside it is below
// Box/2 in magnitude, and wrap_periodic() leaves it unchanged: the open-boundary search gives the wrapped search's result.
B200_HD bool cube_clear_of_faces(const float dom[6], double box, float lox, float loy, float loz, float hix, float hiy, float hiz) {
  double e = 0.0;
  for (int k = 0; k < 3; k++) { e = fmax(e, -(double)dom[k]); e = fmax(e, (double)dom[3 + k] - box); }
  e = e * 1.000001 + 1.0e-6 * box;
  return (double)lox > e && (double)loy > e && (double)loz > e && (double)hix < box - e && (double)hiy < box - e && (double)hiz < box - e;
}

// root box from the bounding box of float coordinates (forcetree.c:179-212)
B200_HD RootBox make_root(const double mn[3], const double mx[3]) {
  double len = mx[0] - mn[0];
  if (mx[1] - mn[1] > len) len = mx[1] - mn[1];
  if (mx[2] - mn[2] > len) len = mx[2] - mn[2];
  len *= 1.01;
  RootBox r;
  r.cx = (float)((mx[0] + mn[0]) / 2);
  r.cy = (float)((mx[1] + mn[1]) / 2);
  r.cz = (float)((mx[2] + mn[2]) / 2);
  r.len = (float)len;
  return r;
}

// one level of descent: returns the octant and moves (c, len) to the child cell
B200_HD int descend(float x, float y, float z, float &cx, float &cy, float &cz, float &len) {
  const int bx = x > cx, by = y > cy, bz = z > cz;
  len = len * 0.5f;                 // exact
  const float q = len * 0.5f;       // exact
  cx = bx ? fadd(cx, q) : fadd(cx, -q);
  cy = by ? fadd(cy, q) : fadd(cy, -q);
  cz = bz ? fadd(cz, q) : fadd(cz, -q);
  return bx | (by << 1) | (bz << 2);
}

// child cell of a known octant (same arithmetic as descend)
B200_HD void child_cell(int oct, float &cx, float &cy, float &cz, float &len) {
  len = len * 0.5f;
  const float q = len * 0.5f;
  cx = (oct & 1) ? fadd(cx, q) : fadd(cx, -q);
  cy = (oct & 2) ? fadd(cy, q) : fadd(cy, -q);
  cz = (oct & 4) ? fadd(cz, q) : fadd(cz, -q);
}

// 42-level key: hi = octants of levels 0..20 (level 0 in the top 3 of 63 bits), lo = 21..41
B200_HD void make_key(float x, float y, float z, const RootBox &rb, uint64_t &hi, uint64_t &lo) {
  float cx = rb.cx, cy = rb.cy, cz = rb.cz, len = rb.len;
  uint64_t h = 0, l = 0;
  for (int lev = 0; lev < kLevelsPerWord; lev++) h = (h << 3) | (uint64_t)descend(x, y, z, cx, cy, cz, len);
  for (int lev = 0; lev < kLevelsPerWord; lev++) l = (l << 3) | (uint64_t)descend(x, y, z, cx, cy, cz, len);
  hi = h; lo = l;
}

B200_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}

// number of leading octree levels two keys share (0..42)
B200_HD int common_levels(uint64_t ahi, uint64_t alo, uint64_t bhi, uint64_t blo) {
  uint64_t x = ahi ^ bhi;
  if (x) return (clz64(x) - 1) / 3;
  x = alo ^ blo;
  if (x) return kLevelsPerWord + (clz64(x) - 1) / 3;
  return kMaxLevels;
}

B200_HD int digit_at(uint64_t hi, uint64_t lo, int level) {
  return level < kLevelsPerWord ? (int)((hi >> (3 * (kLevelsPerWord - 1 - level))) & 7)
                                : (int)((lo >> (3 * (kMaxLevels - 1 - level))) & 7);
}

B200_HD bool key_less(uint64_t ahi, uint64_t alo, uint64_t bhi, uint64_t blo) {
  return ahi < bhi || (ahi == bhi && alo < blo);
}

// ---------------------------------------------------------------------------------------
// Node record the walk streams: 64 bytes, one 64-byte-aligned chunk per node, nodes stored
// in depth-first pre-order so that "open" is always id+1 and "accept" jumps to `skip`.
struct __attribute__((aligned(16))) NodeRec {
  float sx, sy, sz, mass;          // centre of mass, mass                  (NODE.s, .mass)
  float oc, bmax2;                 // oc = mass*len^4 as the reference stores it (NODE.oc), NODE.bmax2
  int   pinfo;                     // (first leaf slot << 4) | number of direct particles
  int   skip;                      // next node in pre-order outside this subtree
  float q11, q22, q33, q12;        // raw second moments about the c.o.m.   (NODE.Q11..)
  float q13, q23, p, len2;         // ... trace P; len*len (BH test, forcetree.c:967)
};
// the first 32 bytes decide open/accept for the relative criterion (the BH test needs len2)

// Record of the packed walk: two sibling cells, component c of cell h at f[2*c + h] (so that one 16-byte load yields
// two (cell0, cell1) register pairs for the sm_100a f32x2 instructions).  Components:
//   0 sx 1 sy 2 sz 3 mass | 4 oc 5 bmax2 6 cinfo 7 pinfo | 8 -3 Q11 9 -3 Q22 10 -3 Q33 11 -3 Q12 | 12 -3 Q13 13 -3 Q23 14 -1.5 P 15 len2
// cinfo = (pair index of the first child cell << 4) | number of child cells; pinfo as in NodeRec.
struct __attribute__((aligned(16))) PairRec { float f[32]; };

// raw moments of a node about its geometric centre, in double (forcetree.c:433-571)
struct Moments { double m, s[3], r[6]; };   // r: xx yy zz xy xz yz

B200_HD void moments_zero(Moments &a) { a.m = 0; for (int k = 0; k < 3; k++) a.s[k] = 0; for (int k = 0; k < 6; k++) a.r[k] = 0; }

// add one particle; rel = pos - centre evaluated in float like the reference (":476")
B200_HD void moments_add_particle(Moments &a, float x, float y, float z, float mass, float cx, float cy, float cz) {
  const double rx = (double)fadd(x, -cx), ry = (double)fadd(y, -cy), rz = (double)fadd(z, -cz);
  const double m = mass;
  a.m += m;
  a.s[0] += m * rx; a.s[1] += m * ry; a.s[2] += m * rz;
  a.r[0] += m * rx * rx; a.r[1] += m * ry * ry; a.r[2] += m * rz * rz;
  a.r[3] += m * rx * ry; a.r[4] += m * rx * rz; a.r[5] += m * ry * rz;
}

// add a child whose moments are about its own centre, shifted by d = centre_child - centre
B200_HD void moments_add_child(Moments &a, const Moments &c, double dx, double dy, double dz) {
  a.m += c.m;
  a.s[0] += c.s[0] + c.m * dx; a.s[1] += c.s[1] + c.m * dy; a.s[2] += c.s[2] + c.m * dz;
  a.r[0] += c.r[0] + 2 * dx * c.s[0] + c.m * dx * dx;
  a.r[1] += c.r[1] + 2 * dy * c.s[1] + c.m * dy * dy;
  a.r[2] += c.r[2] + 2 * dz * c.s[2] + c.m * dz * dz;
  a.r[3] += c.r[3] + dx * c.s[1] + dy * c.s[0] + c.m * dx * dy;
  a.r[4] += c.r[4] + dx * c.s[2] + dz * c.s[0] + c.m * dx * dz;
  a.r[5] += c.r[5] + dy * c.s[2] + dz * c.s[1] + c.m * dy * dz;
}

// final float node fields from the raw moments (forcetree.c:527-570)
B200_HD void moments_finish(const Moments &a, float cx, float cy, float cz, float len, NodeRec &n) {
  double s[3] = {0, 0, 0};
  if (a.m != 0) { s[0] = a.s[0] / a.m; s[1] = a.s[1] / a.m; s[2] = a.s[2] / a.m; }
  const double q11 = a.r[0] - a.m * s[0] * s[0];
  const double q22 = a.r[1] - a.m * s[1] * s[1];
  const double q33 = a.r[2] - a.m * s[2] * s[2];
  const double q12 = a.r[3] - a.m * s[0] * s[1];
  const double q13 = a.r[4] - a.m * s[0] * s[2];
  const double q23 = a.r[5] - a.m * s[1] * s[2];
  const double pp = (a.r[0] + a.r[1] + a.r[2]) - a.m * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  const double c[3] = {cx, cy, cz};
  double sa[3];
  for (int k = 0; k < 3; k++) sa[k] = s[k] + c[k];
  n.sx = (float)sa[0]; n.sy = (float)sa[1]; n.sz = (float)sa[2];
  n.mass = (float)a.m;
  n.q11 = (float)q11; n.q22 = (float)q22; n.q33 = (float)q33;
  n.q12 = (float)q12; n.q13 = (float)q13; n.q23 = (float)q23; n.p = (float)pp;
  const float l2 = fmul(len, len);
  n.len2 = l2;
  n.oc = fmul(fmul(n.mass, l2), l2);            // nop->mass * oc * oc, left to right
  double b2 = 0;
  for (int k = 0; k < 3; k++) { const double dx = fabs(sa[k] - c[k]) + 0.5 * (double)len; b2 += dx * dx; }
  n.bmax2 = (float)b2;
}

// ---------------------------------------------------------------------------------------
// Spline-softened kernels of forcetree.c:1763-1793 evaluated analytically (the reference
// interpolates 10001-entry tables of the same polynomials; lerp error ~1e-8 relative).
B200_HD float soft_force(float u) {            // knlforce
  if (u <= 0.5f) return 32.0f * (1.0f / 3 - 1.2f * u * u + u * u * u);
  return 64.0f * (1.0f / 3 - 0.75f * u + 0.6f * u * u - u * u * u / 6) - 1.0f / (15 * u * u * u);
}
B200_HD float soft_pot(float u) {              // knlpot, forcetree.c:1778,1787
  const float u2 = u * u;
  if (u <= 0.5f) return u2 * (16.0f / 3 + u2 * (-9.6f + 6.4f * u)) - 2.8f;
  return 1.0f / (15 * u) + u2 * (32.0f / 3 + u * (-16.0f + u * (9.6f - 32.0f / 15 * u))) - 3.2f;
}
B200_HD void soft_w234(float u, float &w2, float &w3, float &w4) {   // knlW2, knlW3, knlW4
  if (u <= 0.5f) {
    w2 = -76.8f + 96.0f * u; w3 = 96.0f; w4 = 19.2f * u * (5 * u - 4);
  } else {
    const float u2 = u * u, iu2 = 1.0f / u2;
    w2 = 76.8f + 0.2f * iu2 * iu2 / u - 48.0f / u - 32 * u;
    w3 = -32 - iu2 * iu2 * iu2 + 48 * iu2;
    w4 = -48 + 0.2f * iu2 * iu2 + 76.8f * u - 32 * u2;
  }
}

// particle-particle term (forcetree.c:1135-1186): adds to (ax,ay,az); dx = source - target
B200_HD void pp_force(float dx, float dy, float dz, float mass, float h_inv, float &ax, float &ay, float &az) {
  const float r2 = dx * dx + dy * dy + dz * dz;
  const float r = sqrtf(r2);
  const float u = r * h_inv;
  float fac;
  if (u >= 1.0f) {
    const float ri = 1.0f / r;
    fac = mass * ri * ri * ri;
  } else {
    if (!(u > 1.0e-4f)) return;
    fac = mass * h_inv * h_inv * h_inv * soft_force(u);
  }
  ax += dx * fac; ay += dy * fac; az += dz * fac;
}

// particle-node term: monopole + quadrupole, softened below h (forcetree.c:1262-1373)
B200_HD void pn_force(float dx, float dy, float dz, float r2, const NodeRec &n, float h_inv, float &ax, float &ay, float &az) {
  const float r = sqrtf(r2);
  const float u = r * h_inv;
  const float q11dx = n.q11 * dx, q12dy = n.q12 * dy, q13dz = n.q13 * dz;
  const float q12dx = n.q12 * dx, q22dy = n.q22 * dy, q23dz = n.q23 * dz;
  const float q13dx = n.q13 * dx, q23dy = n.q23 * dy, q33dz = n.q33 * dz;
  const float potq = 0.5f * (q11dx * dx + q22dy * dy + q33dz * dz) + q12dx * dy + q13dx * dz + q23dy * dz;
  float fac, ff;
  if (u >= 1.0f) {
    const float ri = 1.0f / r, r2i = ri * ri, r3i = r2i * ri, r5i = r2i * r3i;
    fac = n.mass * r3i + (15 * potq * r2i - 1.5f * n.p) * r5i;
    ff = -3 * r5i;
  } else {
    if (!(u > 1.0e-4f)) return;
    float w2, w3, w4;
    soft_w234(u, w2, w3, w4);
    const float wf = soft_force(u);
    const float ri = 1.0f / r;
    const float h2i = h_inv * h_inv, h3i = h2i * h_inv, h4i = h2i * h2i, h5i = h2i * h3i, h6i = h3i * h3i;
    fac = n.mass * h3i * wf + potq * h6i * w3 * ri + 0.5f * n.p * w4 * h4i * ri;
    ff = w2 * h5i;
  }
  ax += dx * fac + ff * (q11dx + q12dy + q13dz);
  ay += dy * fac + ff * (q12dx + q22dy + q23dz);
  az += dz * fac + ff * (q13dx + q23dy + q33dz);
}

// fast forms used by the GPU walk: one MUFU.RSQ instead of an IEEE sqrt + division (2 ulp, far
// inside the 1e-4 tolerance); the softened branch (r < h) is rare and keeps the spline.
B200_HD float rsqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
  float y;                                   // one MUFU.RSQ, no denormal fix-up (r2 >= h^2 here)
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.0f / sqrtf(x);
#endif
}
B200_HD void pp_force_fast(float dx, float dy, float dz, float mass, float h_inv, float h2, float &ax, float &ay, float &az) {
  const float r2 = dx * dx + dy * dy + dz * dz;
  float fac;
  if (r2 >= h2) {
    const float ri = rsqrt_fast(r2);
    fac = mass * ri * ri * ri;
  } else {
    const float u = sqrtf(r2) * h_inv;
    if (!(u > 1.0e-4f)) return;
    fac = mass * h_inv * h_inv * h_inv * soft_force(u);
  }
  ax += dx * fac; ay += dy * fac; az += dz * fac;
}
B200_HD void pn_force_fast(float dx, float dy, float dz, float r2, float mass, float q11, float q22, float q33, float q12, float q13,
                           float q23, float pp, float h_inv, float h2, float &ax, float &ay, float &az) {
  const float qx = q11 * dx + q12 * dy + q13 * dz;
  const float qy = q12 * dx + q22 * dy + q23 * dz;
  const float qz = q13 * dx + q23 * dy + q33 * dz;
  const float potq = 0.5f * (dx * qx + dy * qy + dz * qz);          // 1/2 y^T Q y
  float fac, ff;
  if (r2 >= h2) {
    const float ri = rsqrt_fast(r2), r2i = ri * ri, r3i = r2i * ri, r5i = r2i * r3i;
    fac = mass * r3i + (15 * potq * r2i - 1.5f * pp) * r5i;
    ff = -3 * r5i;
  } else {
    const float r = sqrtf(r2), u = r * h_inv;
    if (!(u > 1.0e-4f)) return;
    float w2, w3, w4;
    soft_w234(u, w2, w3, w4);
    const float wf = soft_force(u);
    const float ri = 1.0f / r;
    const float h2i = h_inv * h_inv, h3i = h2i * h_inv, h4i = h2i * h2i, h5i = h2i * h3i, h6i = h3i * h3i;
    fac = mass * h3i * wf + potq * h6i * w3 * ri + 0.5f * pp * w4 * h4i * ri;
    ff = w2 * h5i;
  }
  ax += dx * fac + ff * qx; ay += dy * fac + ff * qy; az += dz * fac + ff * qz;
}

// opening tests.  true = open the cell.
B200_HD bool open_bh(float len2, float r2, float theta2) { return len2 > r2 * theta2; }            // forcetree.c:967
B200_HD bool open_rel(float oc, float bmax2, float r2, float oac) { return oc > oac * r2 * r2 * r2 || r2 < bmax2; }  // :1253-1257

}  // namespace b200
