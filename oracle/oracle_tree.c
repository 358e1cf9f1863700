/* oracle_tree.c - CPU restatement of the reference's octree, tree walk, direct sum and
 * neighbour search.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restates: forcetree.c:166-399 (build by sequential insertion), :408-422 (walk threading),
 * :433-571 (moments), :786-1377 (walks), :1763-1793 (softening tables), :1896-1975 (direct),
 * :2163-2297 (range search), :2311-2414 + :2625-2700 (k nearest).
 * Precision follows the reference: float storage, double arithmetic on float operands,
 * float-only arithmetic where the reference uses float temporaries.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "oracle.h"

#define KLEN 10000

struct otree {
  int n, cap, m;                /* particles, node capacity, nodes used */
  const float *pos, *mass;      /* borrowed, [n][3], [n] */
  double eps;
  /* node arrays (index 0 = root) */
  float *len, (*ctr)[3], (*com)[3], *nmass, (*q)[7], *oc, *bmax2;
  int (*kid)[8];                /* >=0 particle, <= -2 node -(k+2), -1 empty */
  int *count, *first, *up, *sib;
  /* particle arrays */
  int *chain;                   /* next[]: linked list through the particles of a node */
  int *pup;                     /* father[] of particles */
  int *thread;                  /* nextnode[] over the combined index space: particle i -> i, node k -> n+k */
  int randoms;
  float dmin[3], dmax[3];
  double kf[KLEN + 1], kw2[KLEN + 1], kw3[KLEN + 1], kw4[KLEN + 1], kp[KLEN + 1];
};

#define ISNODE(c) ((c) <= -2)
#define NODEOF(c) (-(c) - 2)
#define ASKID(k) (-(k) - 2)

static int octant_of(const float *x, const float *c)
{ return (x[0] > c[0] ? 1 : 0) | (x[1] > c[1] ? 2 : 0) | (x[2] > c[2] ? 4 : 0); }   /* forcetree.c:254-256 */

static int new_node(otree *t, int parent, int oct)
{
  int k = t->m++;
  if (t->m > t->cap) abort();
  if (parent < 0) return k;
  t->len[k] = t->len[parent] / 2;                                   /* forcetree.c:298 */
  for (int j = 0; j < 3; j++)                                       /* forcetree.c:300-306, float */
    t->ctr[k][j] = (oct & (1 << j)) ? t->ctr[parent][j] + t->len[k] / 2 : t->ctr[parent][j] - t->len[k] / 2;
  for (int j = 0; j < 8; j++) t->kid[k][j] = -1;
  t->up[k] = parent; t->count[k] = 0; t->sib[k] = -1;
  return k;
}

/* softening tables, forcetree.c:1763-1793 */
static void fill_tables(otree *t)
{
  for (int i = 0; i <= KLEN; i++) {
    double u = ((double)i) / KLEN;
    if (u <= 0.5) {
      t->kf[i] = 32 * (1.0 / 3 - 6.0 / 5 * pow(u, 2) + pow(u, 3));
      t->kp[i] = 16.0 / 3 * pow(u, 2) - 48.0 / 5 * pow(u, 4) + 32.0 / 5 * pow(u, 5) - 14.0 / 5;            /* knlpot, forcetree.c:1778 */
      t->kw2[i] = -384.0 / 5 + 96.0 * u; t->kw3[i] = 96.0; t->kw4[i] = 96.0 / 5 * u * (5 * u - 4);
    } else {
      t->kf[i] = 64 * (1.0 / 3 - 3.0 / 4 * u + 3.0 / 5 * pow(u, 2) - pow(u, 3) / 6) - 1.0 / 15 / pow(u, 3);
      t->kp[i] = 1.0 / 15 / u + 32.0 / 3 * pow(u, 2) - 16.0 * pow(u, 3) + 48.0 / 5 * pow(u, 4) - 32.0 / 15 * pow(u, 5) - 16.0 / 5;   /* :1787 */
      t->kw2[i] = 384.0 / 5 + 1 / (5.0 * pow(u, 5)) - 48.0 / u - 32 * u;
      t->kw3[i] = -32 - 1 / pow(u, 6) + 48 / pow(u, 2);
      t->kw4[i] = -48 + 1 / (5 * pow(u, 4)) + 384.0 / 5 * u - 32 * pow(u, 2);
    }
  }
}

static void thread_walk(otree *t, int item, int *last)      /* forcetree.c:408-422 */
{
  if (*last >= 0) t->thread[*last] = item;
  *last = item;
  if (item >= t->n) {
    int k = item - t->n;
    for (int j = 0; j < 8; j++) {
      int c = t->kid[k][j];
      if (c >= 0) thread_walk(t, c, last);
      else if (ISNODE(c)) thread_walk(t, t->n + NODEOF(c), last);
    }
  }
}

/* moments of one node by looping over its particle chain, forcetree.c:433-571 (flag=0) */
static void node_moments(otree *t, int k)
{
  double s[3] = {0, 0, 0}, mass = 0, q11 = 0, q22 = 0, q33 = 0, q12 = 0, q13 = 0, q23 = 0, pp = 0, extmax = 0;
  int p = t->first[k];
  for (int c = 0; c < t->count[k]; c++, p = t->chain[p]) {
    double rel[3];
    float m = t->mass[p];
    for (int j = 0; j < 3; j++) {
      rel[j] = t->pos[3 * p + j] - t->ctr[k][j];            /* float subtraction, then widened */
      s[j] += m * rel[j];
      if (extmax < fabs(rel[j])) extmax = fabs(rel[j]);
    }
    mass += m;
    q11 += m * rel[0] * rel[0]; q22 += m * rel[1] * rel[1]; q33 += m * rel[2] * rel[2];
    q12 += m * rel[0] * rel[1]; q13 += m * rel[0] * rel[2]; q23 += m * rel[1] * rel[2];
    pp += m * (rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]);
  }
  if (extmax > t->len[k]) t->len[k] = extmax;               /* forcetree.c:513-519 (SIDM block) */
  if (mass) for (int j = 0; j < 3; j++) s[j] /= mass;
  q11 -= mass * s[0] * s[0]; q22 -= mass * s[1] * s[1]; q33 -= mass * s[2] * s[2];
  q12 -= mass * s[0] * s[1]; q23 -= mass * s[1] * s[2]; q13 -= mass * s[0] * s[2];
  pp -= mass * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  for (int j = 0; j < 3; j++) { s[j] += t->ctr[k][j]; t->com[k][j] = s[j]; }
  t->nmass[k] = mass;
  t->q[k][0] = q11; t->q[k][1] = q22; t->q[k][2] = q33; t->q[k][3] = q12; t->q[k][4] = q13; t->q[k][5] = q23; t->q[k][6] = pp;
  t->oc[k] = t->len[k] * t->len[k];
  t->oc[k] = t->nmass[k] * t->oc[k] * t->oc[k];
  double r2 = 0;
  for (int j = 0; j < 3; j++) { double dx = fabs(s[j] - t->ctr[k][j]) + 0.5 * t->len[k]; r2 += dx * dx; }
  t->bmax2[k] = r2;
}

otree *otree_build(int n, const float *pos, const float *mass, double eps)
{
  otree *t = calloc(1, sizeof(otree));
  t->n = n; t->pos = pos; t->mass = mass; t->eps = eps;
  t->cap = 4 * n + 64;
  t->len = malloc(sizeof(float) * t->cap); t->ctr = malloc(sizeof(float[3]) * t->cap); t->com = malloc(sizeof(float[3]) * t->cap);
  t->nmass = malloc(sizeof(float) * t->cap); t->q = malloc(sizeof(float[7]) * t->cap); t->oc = malloc(sizeof(float) * t->cap);
  t->bmax2 = malloc(sizeof(float) * t->cap); t->kid = malloc(sizeof(int[8]) * t->cap);
  t->count = malloc(sizeof(int) * t->cap); t->first = malloc(sizeof(int) * t->cap); t->up = malloc(sizeof(int) * t->cap);
  t->sib = malloc(sizeof(int) * t->cap);
  t->chain = malloc(sizeof(int) * n); t->pup = malloc(sizeof(int) * n); t->thread = malloc(sizeof(int) * (n + t->cap));
  fill_tables(t);

  /* bounding box and root cell, forcetree.c:179-212 */
  double lo[3], hi[3];
  for (int j = 0; j < 3; j++) lo[j] = hi[j] = pos[j];
  for (int i = 1; i < n; i++)
    for (int j = 0; j < 3; j++) {
      if (pos[3 * i + j] > hi[j]) hi[j] = pos[3 * i + j];
      if (pos[3 * i + j] < lo[j]) lo[j] = pos[3 * i + j];
    }
  for (int j = 0; j < 3; j++) { t->dmin[j] = lo[j]; t->dmax[j] = hi[j]; }
  double ext = hi[0] - lo[0];
  for (int j = 1; j < 3; j++) if (hi[j] - lo[j] > ext) ext = hi[j] - lo[j];
  ext *= 1.01;
  int root = new_node(t, -1, 0);
  for (int j = 0; j < 3; j++) t->ctr[root][j] = (hi[j] + lo[j]) / 2;
  t->len[root] = ext;
  for (int j = 0; j < 8; j++) t->kid[root][j] = -1;
  t->up[root] = -1; t->sib[root] = -1;
  t->kid[root][octant_of(pos, t->ctr[root])] = 0;
  t->pup[0] = root; t->count[root] = 1; t->first[root] = 0; t->chain[0] = -1;

  /* insert the remaining particles one at a time, forcetree.c:241-345 */
  for (int i = 1; i < n; i++) {
    int at = root;
    for (;;) {
      t->count[at]++;
      int oct = octant_of(pos + 3 * i, t->ctr[at]);
      int c = t->kid[at][oct];
      if (ISNODE(c)) { at = NODEOF(c); continue; }
      if (c < 0) {                                        /* empty slot: attach as a leaf */
        t->kid[at][oct] = i; t->pup[i] = at;
        int p = t->first[at];
        for (int s = 0; s < t->count[at] - 2; s++) p = t->chain[p];   /* last particle of this node */
        t->chain[i] = t->chain[p]; t->chain[p] = i;
        break;
      }
      /* slot holds one particle: split it into a new node and retry there */
      int resident = c;
      int k = new_node(t, at, oct);
      t->kid[at][oct] = ASKID(k);
      t->first[k] = resident; t->count[k] = 1;
      int ro = octant_of(pos + 3 * resident, t->ctr[k]);
      if (t->len[k] < 1.0e-3 * eps) t->randoms++;         /* the reference draws rand() here (:320-326) */
      t->kid[k][ro] = resident; t->pup[resident] = k;
      at = k;
    }
  }

  for (int k = 0; k < t->m; k++) {                        /* forcetree.c:356-372 */
    node_moments(t, k);
    int nn = -1;
    for (int j = 7; j >= 0; j--) {
      int c = t->kid[k][j];
      if (c == -1) continue;
      if (ISNODE(c)) t->sib[NODEOF(c)] = nn;
      nn = ISNODE(c) ? t->n + NODEOF(c) : c;
    }
  }
  int last = -1;
  thread_walk(t, t->n + root, &last);
  t->thread[last] = -1;
  for (int k = 0; k < t->m; k++)                          /* forcetree.c:379-394 */
    if (t->sib[k] < 0) {
      int f = k, nn = t->sib[k];
      while (nn < 0) { f = t->up[f]; if (f < 0) break; nn = t->sib[f]; }
      t->sib[k] = nn;
    }
  return t;
}

void otree_free(otree *t)
{
  if (!t) return;
  free(t->len); free(t->ctr); free(t->com); free(t->nmass); free(t->q); free(t->oc); free(t->bmax2); free(t->kid);
  free(t->count); free(t->first); free(t->up); free(t->sib); free(t->chain); free(t->pup); free(t->thread); free(t);
}
int otree_num_nodes(const otree *t) { return t->m; }
int otree_random_subnodes(const otree *t) { return t->randoms; }
void otree_domain(const otree *t, float *mn, float *mx) { for (int j = 0; j < 3; j++) { mn[j] = t->dmin[j]; mx[j] = t->dmax[j]; } }

void otree_dump(const otree *t, float *center, float *len, float *mass, float *s, float *Q, float *oc, float *bmax2, int *count)
{
  for (int k = 0; k < t->m; k++) {
    for (int j = 0; j < 3; j++) { center[3 * k + j] = t->ctr[k][j]; s[3 * k + j] = t->com[k][j]; }
    for (int j = 0; j < 7; j++) Q[7 * k + j] = t->q[k][j];
    len[k] = t->len[k]; mass[k] = t->nmass[k]; oc[k] = t->oc[k]; bmax2[k] = t->bmax2[k]; count[k] = t->count[k];
  }
}
void otree_chain(const otree *t, int *order)
{ int p = t->first[0]; for (int c = 0; c < t->n && p >= 0; c++, p = t->chain[p]) order[c] = p; }

/* ------------------------------------------------------------------ walks */

static double lerp_tab(const double *tab, double u, int *ii, double *ff)
{ *ii = (int)(u * KLEN); *ff = (u - ((double)*ii) / KLEN) * KLEN; return tab[*ii] + (tab[*ii + 1] - tab[*ii]) * *ff; }

/* one target; forcetree.c:817-1089 (BH) and :1097-1377 (relative), dt = 0 */
/* keep != 0: add to acc / cost instead of starting from zero - the second and later trees of a target when there is
 * one tree per particle type (forcetree.c:798-808: the accumulators are zeroed once, then every tree is walked) */
static void walk_one(const otree *t, const oparams *p, const float *tp, float oldacc, double *acc, int *cost, int keep)
{
  const int bh = (p->criterion == 0 || oldacc == 0);
  const double h = 2.8 * p->eps, h_inv = 1 / h;
  const double h2_inv = h_inv * h_inv, h3_inv = h2_inv * h_inv, h4_inv = h2_inv * h2_inv, h5_inv = h2_inv * h3_inv, h6_inv = h3_inv * h3_inv;
  const double oac = oldacc * p->alpha;
  if (!keep) { acc[0] = acc[1] = acc[2] = 0; cost[0] = cost[1] = 0; }
  int item = t->n;                                          /* root */
  while (item >= 0) {
    if (item < t->n) {                                      /* a particle */
      int j = item;
      double dx = t->pos[3 * j] - tp[0], dy = t->pos[3 * j + 1] - tp[1], dz = t->pos[3 * j + 2] - tp[2];  /* float sub */
      double r2 = dx * dx + dy * dy + dz * dz, r = sqrt(r2), u = r * h_inv, fac;
      cost[0]++;
      if (u >= 1) { double ri = 1 / r; fac = t->mass[j] * ri * ri * ri; acc[0] += dx * fac; acc[1] += dy * fac; acc[2] += dz * fac; }
      else {
        int ii; double ff; double wf = lerp_tab(t->kf, u, &ii, &ff);
        if (u > 1.0e-4) { fac = t->mass[j] * h_inv * h_inv * h_inv * wf; acc[0] += dx * fac; acc[1] += dy * fac; acc[2] += dz * fac; }
      }
      item = t->thread[item];
      continue;
    }
    int k = item - t->n;
    double dx = (double)t->com[k][0] - tp[0], dy = (double)t->com[k][1] - tp[1], dz = (double)t->com[k][2] - tp[2];
    double r2 = dx * dx + dy * dy + dz * dz;
    int open = bh ? (t->len[k] * t->len[k] > r2 * p->theta * p->theta)
                  : (t->oc[k] > oac * r2 * r2 * r2 || r2 < t->bmax2[k]);
    if (open) { item = t->thread[item]; continue; }
    cost[1]++;
    const float *Q = t->q[k];
    double r = sqrt(r2), u = r * h_inv;
    double q11dx = Q[0] * dx, q12dy = Q[3] * dy, q13dz = Q[4] * dz, q12dx = Q[3] * dx, q22dy = Q[1] * dy, q23dz = Q[5] * dz,
           q13dx = Q[4] * dx, q23dy = Q[5] * dy, q33dz = Q[2] * dz;
    double potq = 0.5 * (q11dx * dx + q22dy * dy + q33dz * dz) + q12dx * dy + q13dx * dz + q23dy * dz;
    double fac, ff;
    if (u >= 1) {
      double ri = 1 / r, r2i = ri * ri, r3i = r2i * ri, r5i = r2i * r3i;
      fac = t->nmass[k] * r3i + (15 * potq * r2i - 1.5 * Q[6]) * r5i;
      acc[0] += dx * fac; acc[1] += dy * fac; acc[2] += dz * fac;
      ff = -3 * r5i;
      acc[0] += ff * (q11dx + q12dy + q13dz); acc[1] += ff * (q12dx + q22dy + q23dz); acc[2] += ff * (q13dx + q23dy + q33dz);
    } else {
      int ii; double f2;
      double wf = lerp_tab(t->kf, u, &ii, &f2);
      double w2 = t->kw2[ii] + (t->kw2[ii + 1] - t->kw2[ii]) * f2, w3 = t->kw3[ii] + (t->kw3[ii + 1] - t->kw3[ii]) * f2,
             w4 = t->kw4[ii] + (t->kw4[ii + 1] - t->kw4[ii]) * f2;
      if (u > 1.0e-4) {
        double ri = 1 / r;
        fac = t->nmass[k] * h2_inv * h_inv * wf + potq * h6_inv * w3 * ri + 0.5 * Q[6] * w4 * h4_inv * ri;
        acc[0] += dx * fac; acc[1] += dy * fac; acc[2] += dz * fac;
        ff = w2 * h5_inv;
        acc[0] += ff * (q11dx + q12dy + q13dz); acc[1] += ff * (q12dx + q22dy + q23dz); acc[2] += ff * (q13dx + q23dy + q33dz);
      }
    }
    item = t->sib[k];
  }
}

void otree_force(const otree *t, const oparams *p, int nt, const int *targets, const float *oldacc, double *acc, int *cost)
{
  for (int i = 0; i < nt; i++) {
    int c[2];
    walk_one(t, p, t->pos + 3 * targets[i], oldacc ? oldacc[targets[i]] : 0.0f, acc + 3 * i, c, 0);
    if (cost) { cost[2 * i] = c[0]; cost[2 * i + 1] = c[1]; }
  }
}
/* the same walk for targets given by position (particles of another type's tree, forcetree.c:798-808); oldacc and the
 * results are per target entry */
void otree_force_at(const otree *t, const oparams *p, int nt, const float *xyz, const float *oldacc, double *acc, int *cost, int keep)
{
  for (int i = 0; i < nt; i++) walk_one(t, p, xyz + 3 * i, oldacc ? oldacc[i] : 0.0f, acc + 3 * i, cost + 2 * i, keep);
}

/* potential of one target: forcetree.c:1417-1577 (BH) and :1585-1755 (relative criterion) */
static double pot_one(const otree *t, const oparams *p, const float *tp, float oldacc, double pot)
{
  const int bh = (p->criterion == 0 || oldacc == 0);        /* forcetree.c:1404 */
  const double h = 2.8 * p->eps, h_inv = 1 / h;
  const double h2_inv = h_inv * h_inv, h3_inv = h2_inv * h_inv, h5_inv = h2_inv * h3_inv;
  const double oac = oldacc * p->alpha;
  int item = t->n;
  while (item >= 0) {
    if (item < t->n) {
      int j = item;
      double dx = t->pos[3 * j] - tp[0], dy = t->pos[3 * j + 1] - tp[1], dz = t->pos[3 * j + 2] - tp[2];
      double r2 = dx * dx + dy * dy + dz * dz, r = sqrt(r2), u = r * h_inv;
      if (u >= 1) pot -= t->mass[j] / r;
      else { int ii; double ff; double wp = lerp_tab(t->kp, u, &ii, &ff); pot += t->mass[j] * h_inv * wp; }
      item = t->thread[item];
      continue;
    }
    int k = item - t->n;
    double dx = t->com[k][0] - tp[0], dy = t->com[k][1] - tp[1], dz = t->com[k][2] - tp[2];   /* float - float: no node drift here (forcetree.c:1477,1645) */
    double r2 = dx * dx + dy * dy + dz * dz;
    int open = bh ? (t->len[k] * t->len[k] > r2 * p->theta * p->theta)
                  : (t->oc[k] > oac * r2 * r2 * r2 || r2 < t->bmax2[k]);
    if (open) { item = t->thread[item]; continue; }
    const float *Q = t->q[k];
    double r = sqrt(r2), u = r * h_inv;
    double q11dx = Q[0] * dx, q12dx = Q[3] * dx, q22dy = Q[1] * dy, q13dx = Q[4] * dx, q23dy = Q[5] * dy, q33dz = Q[2] * dz;
    double potq = 0.5 * (q11dx * dx + q22dy * dy + q33dz * dz) + q12dx * dy + q13dx * dz + q23dy * dz;
    if (u >= 1) {
      double ri = 1 / r, r2i = ri * ri, r3i = r2i * ri;
      pot += -t->nmass[k] * ri + r3i * (-3 * potq * r2i + 0.5 * Q[6]);
    } else {
      int ii; double f2;
      double wf = lerp_tab(t->kf, u, &ii, &f2);
      double wp = t->kp[ii] + (t->kp[ii + 1] - t->kp[ii]) * f2, w2 = t->kw2[ii] + (t->kw2[ii + 1] - t->kw2[ii]) * f2;
      pot += t->nmass[k] * h_inv * wp + potq * w2 * h5_inv + 0.5 * Q[6] * wf * h2_inv * h_inv;
    }
    item = t->sib[k];
  }
  (void)h3_inv;
  return pot;
}

/* raw potentials as force_treeevaluate_potential() leaves them in GravDataPotential (forcetree.c:1389) */
void otree_potential(const otree *t, const oparams *p, int nt, const int *targets, const float *oldacc, double *pot)
{
  for (int i = 0; i < nt; i++) pot[i] = pot_one(t, p, t->pos + 3 * targets[i], oldacc ? oldacc[targets[i]] : 0.0f, 0.0);
}
void otree_potential_at(const otree *t, const oparams *p, int nt, const float *xyz, const float *oldacc, double *pot, int keep)
{
  for (int i = 0; i < nt; i++) pot[i] = pot_one(t, p, xyz + 3 * i, oldacc ? oldacc[i] : 0.0f, keep ? pot[i] : 0.0);
}
/* potential.c:131-168 without comoving integration and Lambda: float Potential = raw; += m/eps (self energy); *= G */
void opot_epilogue(const oparams *p, int nt, const double *pot, const float *mass, float *out)
{
  for (int i = 0; i < nt; i++) {
    float v = pot[i];
    v += mass[i] / p->eps;
    v *= p->G;
    out[i] = v;
  }
}

void otree_direct(const otree *t, const oparams *p, int nt, const int *targets, double *acc)   /* forcetree.c:1896-1975 */
{
  const double h = 2.8 * p->eps, h_inv = 1 / h;
  for (int i = 0; i < nt; i++) {
    const float *tp = t->pos + 3 * targets[i];
    double a[3] = {0, 0, 0};
    for (int j = 0; j < t->n; j++) {
      double dx = (double)t->pos[3 * j] - tp[0], dy = (double)t->pos[3 * j + 1] - tp[1], dz = (double)t->pos[3 * j + 2] - tp[2];
      double r2 = dx * dx + dy * dy + dz * dz, r = sqrt(r2), u = r * h_inv, fac;
      if (u >= 1) { double ri = 1 / r; fac = t->mass[j] * ri * ri * ri; }
      else {
        int ii; double ff; double wf = lerp_tab(t->kf, u, &ii, &ff);
        if (!(u > 1.0e-4)) continue;
        fac = t->mass[j] * h_inv * h_inv * h_inv * wf;
      }
      a[0] += dx * fac; a[1] += dy * fac; a[2] += dz * fac;
    }
    acc[3 * i] = a[0]; acc[3 * i + 1] = a[1]; acc[3 * i + 2] = a[2];
  }
}

void ograv_epilogue(const oparams *p, int nt, const double *acc, float *accel, float *oldacc)   /* gravtree.c:230-324 */
{
  for (int i = 0; i < nt; i++) {
    float a[3] = {(float)acc[3 * i], (float)acc[3 * i + 1], (float)acc[3 * i + 2]};
    if (p->criterion == 1) oldacc[i] = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    for (int j = 0; j < 3; j++) accel[3 * i + j] = p->G * a[j];
  }
}

/* ------------------------------------------------------------------ neighbours */

typedef struct { const otree *t; float lo[3], hi[3]; int *list; int n, cap; } search;

static void range_rec(search *s, int item)                  /* forcetree.c:2224-2297 */
{
  const otree *t = s->t;
  if (item < t->n) {
    for (int j = 0; j < 3; j++) { if (t->pos[3 * item + j] < s->lo[j]) return; if (t->pos[3 * item + j] > s->hi[j]) return; }
    if (s->n < s->cap) s->list[s->n++] = item;
    return;
  }
  int k = item - t->n, j;
  for (j = 0; j < 3; j++) {
    if ((t->ctr[k][j] + 0.5 * t->len[k]) < s->lo[j]) return;
    if ((t->ctr[k][j] - 0.5 * t->len[k]) > s->hi[j]) return;
  }
  for (j = 0; j < 3; j++) {
    if ((t->ctr[k][j] + 0.5 * t->len[k]) > s->hi[j]) break;
    if ((t->ctr[k][j] - 0.5 * t->len[k]) < s->lo[j]) break;
  }
  if (j >= 3) {                                             /* cell completely inside: take its chain */
    int p = t->first[k];
    for (int c = 0; c < t->count[k]; c++, p = t->chain[p]) if (s->n < s->cap) s->list[s->n++] = p;
    return;
  }
  for (j = 0; j < 8; j++) {
    int c = t->kid[k][j];
    if (c >= 0) range_rec(s, c); else if (ISNODE(c)) range_rec(s, t->n + NODEOF(c));
  }
}

static float dist2f(const float *a, const float *b)        /* float arithmetic, forcetree.c:2195-2204 */
{ float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2]; return dx * dx + dy * dy + dz * dz; }

int ongb_variable(const otree *t, const float xyz[3], float h, int *list, float *r2, int cap)   /* forcetree.c:2163-2218 */
{
  search s; s.t = t; s.list = list; s.n = 0; s.cap = cap;
  for (int j = 0; j < 3; j++) { s.lo[j] = xyz[j] - h; s.hi[j] = xyz[j] + h; }
  float sr2 = h * h;
  range_rec(&s, t->n);
  int n = s.n;
  for (int i = 0; i < n; i++) {
    float d2 = dist2f(t->pos + 3 * list[i], xyz);
    r2[i] = d2;
    if (d2 >= sr2) { list[i] = list[n - 1]; n--; i--; }     /* swap-remove */
  }
  return n;
}

static int cmpf(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }

float ongb_treefind(const otree *t, const float xyz[3], int desngb)   /* forcetree.c:2311-2414, hguess = 0 */
{
  int th = 0;
  while (t->count[th] > 200) {
    int c = t->kid[th][octant_of(xyz, t->ctr[th])];
    if (ISNODE(c) && t->count[NODEOF(c)] > 200) th = NODEOF(c); else break;
  }
  float sr = t->len[th] * pow((3.0 / (4 * 3.14159265358979323846) * 1.2) * desngb / ((float)(t->count[th])), 1.0 / 3);
  int cap = t->n; int *list = malloc(sizeof(int) * cap); float *r2 = malloc(sizeof(float) * cap);
  float h2max = 0;
  for (;;) {
    search s; s.t = t; s.list = list; s.n = 0; s.cap = cap;
    for (int j = 0; j < 3; j++) { s.lo[j] = xyz[j] - sr; s.hi[j] = xyz[j] + sr; }
    float sr2 = sr * sr;
    range_rec(&s, t->n);
    if (s.n < desngb) { if (s.n > 5) sr *= pow((2.1 * (float)desngb) / s.n, 1.0 / 3); else sr *= 2.0; continue; }
    for (int i = 0; i < s.n; i++) r2[i] = dist2f(t->pos + 3 * list[i], xyz);
    qsort(r2, s.n, sizeof(float), cmpf);                    /* the reference quick-selects (:2625); same k-th value */
    h2max = r2[desngb - 1];
    if (h2max <= sr2) break;
    sr *= 1.26;
  }
  free(list); free(r2);
  return h2max;
}

const float *otree_positions(const otree *t) { return t->pos; }

/* ---- global quantities (global.c:18-135) -------------------------------------------------------
 * Particle order, float products and double sums exactly as the reference's expressions. */
static void onorm3(double *v) { v[3] = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
void oglobal_quantities(int n, const float *pp, const float *vp, const float *mass, const float *potential,
                        const int *type, osysstate *S)
{
  memset(S, 0, sizeof(*S));
  for (int i = 0; i < n; i++) {
    const int t = type ? type[i] : 1;
    const float m = mass[i];
    const float *x = pp + 3 * (size_t)i, *v = vp + 3 * (size_t)i;
    if (t < 0 || t > 4) continue;
    S->MassComp[t] += m;                                                         /* :33 */
    S->EnergyPotComp[t] += 0.5 * m * potential[i];                               /* :35 */
    S->EnergyKinComp[t] += 0.5 * m * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);  /* :40-42 */
    for (int j = 0; j < 3; j++) {                                                /* :44-48 */
      S->MomentumComp[t][j] += m * v[j];
      S->CenterOfMassComp[t][j] += m * x[j];
    }
    S->AngMomentumComp[t][0] += m * (x[1] * v[2] - x[2] * v[1]);                 /* :50-55 */
    S->AngMomentumComp[t][1] += m * (x[2] * v[0] - x[0] * v[2]);
    S->AngMomentumComp[t][2] += m * (x[0] * v[1] - x[1] * v[0]);
  }
  for (int i = 0; i < 5; i++) {                                                  /* :71-93 */
    S->EnergyTotComp[i] = S->EnergyKinComp[i] + S->EnergyPotComp[i] + S->EnergyIntComp[i];
    S->Mass += S->MassComp[i]; S->EnergyKin += S->EnergyKinComp[i]; S->EnergyPot += S->EnergyPotComp[i];
    S->EnergyInt += S->EnergyIntComp[i]; S->EnergyTot += S->EnergyTotComp[i];
    for (int j = 0; j < 3; j++) {
      S->Momentum[j] += S->MomentumComp[i][j]; S->AngMomentum[j] += S->AngMomentumComp[i][j];
      S->CenterOfMass[j] += S->CenterOfMassComp[i][j];
    }
  }
  for (int i = 0; i < 5; i++) for (int j = 0; j < 3; j++) if (S->MassComp[i] > 0) S->CenterOfMassComp[i][j] /= S->MassComp[i];
  for (int j = 0; j < 3; j++) if (S->Mass > 0) S->CenterOfMass[j] /= S->Mass;    /* :95-102 */
  for (int i = 0; i < 5; i++) { onorm3(S->CenterOfMassComp[i]); onorm3(S->MomentumComp[i]); onorm3(S->AngMomentumComp[i]); }
  onorm3(S->CenterOfMass); onorm3(S->Momentum); onorm3(S->AngMomentum);          /* :104-130 */
}
