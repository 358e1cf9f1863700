/* oracle_sidm.c - CPU restatement of the reference's SIDM scatter step for one MPI rank.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restates: sidm_rand.c:17-46 + sidm_rand.h:8-37 (MT19937 stream, Marsaglia direction),
 * begrun.c:968-992 (SPH kernel table), sidm.c:57-627 (sidm(), CROSS_SECTION_TYPE 0..4, one
 * bunch), sidm.c:814-968 (sidm_ensure_neighbours, mode 0), sidm.c:970-990 (getvmax).
 * GSL is not vendored by the reference; its gsl_rng_mt19937 is the published MT19937
 * (Matsumoto & Nishimura) with the 2002 seeding, restated here.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "oracle.h"

/* ------------------------------------------------------------------ MT19937 */
struct orng { unsigned int mt[624]; int at; long drawn; };

orng *orng_new(unsigned long seed, int warmup)
{
  orng *r = malloc(sizeof(orng));
  if (seed == 0) seed = 4357;
  r->mt[0] = (unsigned int)(seed & 0xffffffffUL);
  for (int i = 1; i < 624; i++) r->mt[i] = 1812433253U * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned int)i;
  r->at = 624; r->drawn = 0;
  if (warmup) { for (int i = 0; i < 1000001; i++) orng_uniform(r); r->drawn = 0; }   /* sidm_rand.c:29-36 */
  return r;
}
void orng_free(orng *r) { free(r); }
long orng_count(const orng *r) { return r->drawn; }

double orng_uniform(orng *r)
{
  if (r->at >= 624) {
    for (int k = 0; k < 624; k++) {
      unsigned int y = (r->mt[k] & 0x80000000U) | (r->mt[(k + 1) % 624] & 0x7fffffffU);
      r->mt[k] = r->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    r->at = 0;
  }
  unsigned int y = r->mt[r->at++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680U; y ^= (y << 15) & 0xefc60000U; y ^= y >> 18;
  r->drawn++;
  return y / 4294967296.0;                                  /* sidm_rand.h:8-11 via gsl_rng_uniform */
}

static void random_direction(orng *r, double n[3])          /* sidm_rand.h:24-37 */
{
  double y1, y2, r2;
  do { y1 = 1.0 - 2.0 * orng_uniform(r); y2 = 1.0 - 2.0 * orng_uniform(r); r2 = y1 * y1 + y2 * y2; } while (r2 > 1.0);
  double sq = sqrt(1.0 - r2);
  n[0] = 2.0 * y1 * sq; n[1] = 2.0 * y2 * sq; n[2] = 1.0 - 2.0 * r2;
}

/* ------------------------------------------------------------------ SPH kernel table */
#define KT 1000
static double Wtab[KT + 2], Rtab[KT + 2];
static int Wready = 0;
static void kernel_table(void)                              /* begrun.c:968-992 */
{
  const double PI = 3.14159265358979323846;
  for (int i = 0; i <= KT + 1; i++) Rtab[i] = ((double)i) / KT;
  Wtab[KT + 1] = 0;
  for (int i = 0; i <= KT; i++) {
    if (Rtab[i] <= 0.5) Wtab[i] = 8 / PI * (1 - 6 * Rtab[i] * Rtab[i] * (1 - Rtab[i]));
    else Wtab[i] = 8 / PI * 2 * (1 - Rtab[i]) * (1 - Rtab[i]) * (1 - Rtab[i]);
  }
  Wready = 1;
}

double ogetvmax(int n, const float *vel)                    /* sidm.c:970-990 */
{
  double v2, vm = 0.0;
  for (int j = 0; j < n; j++) {
    v2 = vel[3 * j] * vel[3 * j] + vel[3 * j + 1] * vel[3 * j + 1] + vel[3 * j + 2] * vel[3 * j + 2];   /* float expr */
    if (vm < v2) vm = v2;
  }
  return sqrt(vm);
}

extern const float *otree_positions(const otree *t);

/* ------------------------------------------------------------------ sidm() */
void osidm_pass(const otree *t, const oparams *p, int nactive, const int *active,
                const float *vel, const float *mass, const float *hsml, const float *dt,
                float *dvel, int *ngbcount, double vmax, orng *rng, osidm_out *out)
{
  if (!Wready) kernel_table();
  const float *pos = otree_positions(t);
  float dmin[3], dmax[3];
  otree_domain(t, dmin, dmax);
  int n = nactive;
  memset(out, 0, sizeof(*out));
  out->nslot = n;
  out->slot_particle = malloc(sizeof(int) * (n + 1)); out->rand = malloc(sizeof(double) * (n + 1));
  out->dir = calloc(3 * (size_t)(n + 1), sizeof(double)); out->pmax = malloc(sizeof(double) * (n + 1));
  out->prob = calloc(n + 1, sizeof(double)); out->partner = malloc(sizeof(int) * (n + 1)); out->ngb = malloc(sizeof(int) * (n + 1));
  out->log_i = malloc(sizeof(int) * (n + 1)); out->log_j = malloc(sizeof(int) * (n + 1)); out->log_dv = malloc(sizeof(float) * 3 * (size_t)(n + 1));
  out->extra_off = calloc(n + 2, sizeof(int));
  int xcap = 1024, xn = 0; out->extra = malloc(sizeof(double) * 2 * xcap);
  float (*dv)[3] = calloc(n + 1, sizeof(float[3]));
  int *already = malloc(sizeof(int) * (n + 1));
  int *confirm = malloc(sizeof(int) * (n + 1));
  int *place_of = malloc(sizeof(int) * (n + 1));

  /* buffer order, sidm.c:141-200: particles within h of the domain box edge ("exported",
   * Type|=8) come first, then the rest, each group in active-list order */
  int nexport = 0;
  unsigned char *exp = calloc(n + 1, 1);
  for (int a = 0; a < n; a++) {
    int i = active[a], j;
    for (j = 0; j < 3; j++) {
      if (pos[3 * i + j] < (dmin[j] + hsml[i])) break;      /* float expressions */
      if (pos[3 * i + j] > (dmax[j] - hsml[i])) break;
    }
    if (j != 3) { exp[a] = 1; nexport++; }
  }
  int ne = 0, ni = 0;
  for (int a = 0; a < n; a++) {
    int place = exp[a] ? ne++ : nexport + ni++;
    place_of[a] = place;
    out->slot_particle[place] = active[a];
    already[place] = (dvel[3 * active[a]] != 0.0f);        /* ID = 0 marks "already scattered", sidm.c:189-192 */
    confirm[place] = -1;
  }
  free(exp);

  const double ball = 1.0 * (3. / 4. / 3.14159265358979323846) * (p->des_ngb + p->max_dev);
  double C_Pmax;                                            /* sidm.c:276-314, non-comoving branch */
  if (p->xs_type == 1) C_Pmax = ball * p->sigma;
  else if (p->xs_type == 2) {
    if (2.0 * vmax < p->vc / sqrt(3.0)) { double beta = 2.0 * vmax / p->vc, v_dep = 1.0 / (1.0 + beta * beta); C_Pmax = ball * 2.0 * vmax * v_dep * v_dep * p->sigma; }
    else C_Pmax = ball * (3.0 * sqrt(3.0) / 16.0) * p->vc * p->sigma;
  } else if (p->xs_type == 3) C_Pmax = ball * 2 * p->pl_v0 * p->sigma;
  else C_Pmax = ball * 2 * vmax * p->sigma;
  const double sigma = p->sigma;
  int cap = t ? 0 : 0; (void)cap;
  int lcap = 65536; int *list = malloc(sizeof(int) * lcap); float *r2l = malloc(sizeof(float) * lcap);
  int pass1 = 0;

  for (int s = 0; s < n; s++) {                             /* sidm.c:319-460 */
    int i = out->slot_particle[s];
    out->extra_off[s] = xn; out->extra_off[s + 1] = xn;
    int numngb = ongb_variable(t, pos + 3 * i, hsml[i], list, r2l, lcap);
    out->ngb[s] = numngb;
    double dt_h0 = dt[i];                                   /* s_a_inverse = 1 */
    double h = 1.0 * hsml[i], hinv = 1.0 / h, hinv3 = hinv * hinv * hinv;
    out->partner[s] = -1;
    double P_max = C_Pmax * mass[i] * hinv3 * dt_h0;
    out->pmax[s] = P_max;
    double rnd = orng_uniform(rng);
    out->rand[s] = rnd;
    if (P_max < rnd) continue;
    if (already[s]) continue;
    pass1++;
    double Prob = 0.0, wk = 0.0;
    for (int k = 0; k < numngb; k++) {
      int j = list[k];
      double r = sqrt(r2l[k]);
      if (dvel[3 * j] != 0.0f) continue;
      if (r < h) {
        double u = r * hinv; int ii = (int)(u * KT);
        wk = hinv3 * (Wtab[ii] + (Wtab[ii + 1] - Wtab[ii]) * (u - Rtab[ii]) * KT);
      }
      double rvx = vel[3 * i] - vel[3 * j], rvy = vel[3 * i + 1] - vel[3 * j + 1], rvz = vel[3 * i + 2] - vel[3 * j + 2];   /* float sub */
      double rv = sqrt(rvx * rvx + rvy * rvy + rvz * rvz);
      if (p->xs_type == 1) Prob += 0.5 * mass[j] * wk * sigma * dt_h0;                       /* sidm.c:374 */
      else if (p->xs_type == 2) { double beta = rv / p->vc, v_dep = 1.0 / (1.0 + beta * beta); Prob += 0.5 * mass[j] * wk * rv * v_dep * v_dep * sigma * dt_h0; }
      else if (p->xs_type == 3) Prob += 0.5 * mass[j] * wk * rv * pow(rv / p->pl_v0, p->pl_n) * sigma * dt_h0;
      else Prob += 0.5 * mass[j] * wk * rv * sigma * dt_h0;                                  /* sidm.c:372 (types 0 and 4) */
      if (Prob < rnd) continue;
      out->partner[s] = j;                                  /* SidmTarget[i] = j: stays set even if type 4 rejects the angle */
      double rmass = mass[j] / (mass[i] + mass[j]);         /* float expr widened */
      double nx[3];
      if (p->xs_type == 4) {                                /* sidm.c:391-427: angular dependence by rejection */
        double beta = rv / p->vc;
        rnd = orng_uniform(rng);                            /* overwrites the slot's uniform: later neighbours compare with it */
        double ucos = orng_uniform(rng);
        if (xn >= xcap) { xcap *= 2; out->extra = realloc(out->extra, sizeof(double) * 2 * xcap); }
        out->extra[2 * xn] = rnd; out->extra[2 * xn + 1] = ucos; xn++; out->extra_off[s + 1] = xn;
        double cosO = 2.0 * ucos - 1.0;
        double sin22 = 0.5 * (1.0 - cosO);
        double denom = 1.0 + beta * beta * sin22;
        if (rnd >= 1 / (denom * denom) || rv == 0.0) continue;
        double rvv[3] = {rvx, rvy, rvz}, np_[3];
        random_direction(rng, nx);
        np_[0] = rvv[1] * nx[2] - rvv[2] * nx[1]; np_[1] = rvv[2] * nx[0] - rvv[0] * nx[2]; np_[2] = rvv[0] * nx[1] - rvv[1] * nx[0];   /* perp(), sidm.c:29-52 */
        double oo = sqrt(np_[0] * np_[0] + np_[1] * np_[1] + np_[2] * np_[2]);
        np_[0] /= oo; np_[1] /= oo; np_[2] /= oo;
        double sinO = 1.0 - cosO * cosO > 0.0 ? sqrt(1.0 - cosO * cosO) : 0.0;
        out->dir[3 * s] = nx[0]; out->dir[3 * s + 1] = nx[1]; out->dir[3 * s + 2] = nx[2];
        dv[s][0] = rmass * (-rvx + cosO * rvv[0] + sinO * rv * np_[0]);
        dv[s][1] = rmass * (-rvy + cosO * rvv[1] + sinO * rv * np_[1]);
        dv[s][2] = rmass * (-rvz + cosO * rvv[2] + sinO * rv * np_[2]);
        confirm[s] = 0;
        break;
      }
      random_direction(rng, nx);
      out->dir[3 * s] = nx[0]; out->dir[3 * s + 1] = nx[1]; out->dir[3 * s + 2] = nx[2];
      dv[s][0] = rmass * (-rvx + rv * nx[0]); dv[s][1] = rmass * (-rvy + rv * nx[1]); dv[s][2] = rmass * (-rvz + rv * nx[2]);
      confirm[s] = 0;
      break;
    }
    out->prob[s] = Prob;
  }

  /* results back to the particles in active-list order, sidm.c:495-537 */
  int scattered = 0, rejected = 0;
  for (int a = 0; a < n; a++) {
    int i = active[a], s = place_of[a];
    ngbcount[i] = out->ngb[s];
    if (ngbcount[i] < (p->des_ngb - p->max_dev) || ngbcount[i] > (p->des_ngb + p->max_dev)) {
      confirm[s] = -1;
      if (dv[s][0] != 0.0) rejected++;
    } else if (dv[s][0] != 0.0) {
      dvel[3 * i] = dv[s][0]; dvel[3 * i + 1] = dv[s][1]; dvel[3 * i + 2] = dv[s][2];
      scattered++;
    } else confirm[s] = -1;
  }
  /* partners, in buffer order: the last writer wins, sidm.c:559-601 */
  for (int s = 0; s < n; s++)
    if (confirm[s] == 0) {
      int j = out->partner[s];
      dvel[3 * j] = -dv[s][0]; dvel[3 * j + 1] = -dv[s][1]; dvel[3 * j + 2] = -dv[s][2];
      out->log_i[out->nlog] = out->slot_particle[s]; out->log_j[out->nlog] = j;
      memcpy(out->log_dv + 3 * out->nlog, dv[s], sizeof(float) * 3);
      out->nlog++;
    }
  out->sct[0] = n; out->sct[1] = pass1; out->sct[2] = scattered; out->sct[3] = rejected;
  free(dv); free(already); free(confirm); free(place_of); free(list); free(r2l);
}

void osidm_out_free(osidm_out *o)
{
  free(o->slot_particle); free(o->rand); free(o->dir); free(o->pmax); free(o->prob); free(o->partner); free(o->ngb);
  free(o->log_i); free(o->log_j); free(o->log_dv); free(o->extra); free(o->extra_off);
  memset(o, 0, sizeof(*o));
}

/* ------------------------------------------------------------------ sidm_ensure_neighbours(0) */
int osidm_ensure(const otree *t, const oparams *p, int n, const float *vel, const float *mass, float *hsml,
                 const float *dt, float *dvel, int *ngbcount, float *left, float *right, double vmax, orng *rng)
{
  const float *pos = otree_positions(t);
  const int lo = p->des_ngb - p->max_dev, hi = p->des_ngb + p->max_dev;
  int candidates = 0;
  for (int i = 0; i < n; i++) if (ngbcount[i] < lo || ngbcount[i] > hi) candidates++;   /* sidm.c:836-843 (all active) */
  if (!candidates) return 0;
  for (int i = 0; i < n; i++) left[i] = right[i] = 0;       /* sidm.c:857-859 */
  int *redo = malloc(sizeof(int) * n);
  int iter = 0;
  for (;;) {
    int nr = 0;
    for (int i = 0; i < n; i++) {                           /* sidm.c:862-888 */
      if (ngbcount[i] < lo || ngbcount[i] > hi) {
        if (left[i] > 0 && right[i] > 0) if ((right[i] - left[i]) < 1.0e-3 * left[i]) continue;
        redo[nr++] = i;
        if (ngbcount[i] < lo) left[i] = hsml[i] > left[i] ? hsml[i] : left[i];
        else { if (right[i] != 0) { if (hsml[i] < right[i]) right[i] = hsml[i]; } else right[i] = hsml[i]; }
      }
    }
    if (nr == 0) break;
    for (int a = 0; a < nr; a++) {                          /* sidm.c:917-929 */
      int i = redo[a];
      if (left[i] == 0 || right[i] == 0) {
        if (right[i] == 0 && ngbcount[i] < 15 && n > p->des_ngb) hsml[i] = sqrt(ongb_treefind(t, pos + 3 * i, p->des_ngb));
        else hsml[i] = hsml[i] * (0.5 + 0.5 * pow(ngbcount[i] / ((double)p->des_ngb), -1.0 / 3));
      } else hsml[i] = 0.5 * (left[i] + right[i]);
    }
    osidm_out o;
    osidm_pass(t, p, nr, redo, vel, mass, hsml, dt, dvel, ngbcount, vmax, rng, &o);
    osidm_out_free(&o);
    iter++;
    if (iter > 30) { free(redo); return -1155; }
  }
  free(redo);
  return iter;
}

/* ------------------------------------------------------------------ find_timesteps(), timestep.c:17-334 */
int ofind_timesteps(const oparams *p, const otimestep *ts, int nactive, const int *active, int mode, double time, double vmax,
                    const float *accel, const float *curtime, float *maxpred, const float *hsml, const float *mass, const double *jitter)
{
  const double PI = 3.14159265358979323846;
  const double C_Grho = (3. / 4. / PI) * (p->des_ngb + p->max_dev);                     /* :45 */
  const double s_a = 1;                                                                /* :97 */
  double C_max;                                                                        /* :100-130 */
  if (p->xs_type == 1) C_max = 1.0 * (3. / 4. / PI) * (p->des_ngb + p->max_dev) * p->sigma;
  else if (p->xs_type == 2) {
    if (2.0 * vmax < p->vc / sqrt(3.0)) { double v_dep = 1.0 / (1.0 + 2.0 * vmax / p->vc); C_max = 1.0 * (3. / 4. / PI) * (p->des_ngb + p->max_dev) * 2.0 * vmax * v_dep * v_dep * p->sigma; }
    else C_max = 1.0 * (3. / 4. / PI) * (p->des_ngb + p->max_dev) * (3.0 * sqrt(3.0) / 16.0) * p->vc * p->sigma;
  } else if (p->xs_type == 3) C_max = 1.0 * (3. / 4. / PI) * (p->des_ngb + p->max_dev) * 2 * p->pl_v0 * p->sigma;
  else C_max = 1.0 * (3. / 4. / PI) * (p->des_ngb + p->max_dev) * 2 * vmax * p->sigma;
  int clamped = 0;
  for (int a = 0; a < nactive; a++) {
    int i = active[a];
    double ac = sqrt(accel[3 * i] * accel[3 * i] + accel[3 * i + 1] * accel[3 * i + 1] + accel[3 * i + 2] * accel[3 * i + 2]);   /* float expr, :138 */
    double dtold = 2 * (curtime[i] + maxpred[i] - 2 * time);                            /* float sum first, :142 */
    double dt;
    if (ts->crit == 0) dt = sqrt(2 * ts->eta * p->eps / ac * s_a); else dt = ts->velscale / ac;   /* :153-160 */
    double h = hsml[i], hinv = 1.0 / h, hinv3 = hinv * hinv * hinv;                     /* :247-265 */
    double dt_sidm = ts->probtol / (C_max * mass[i] * hinv3);
    if (dt_sidm < dt) dt = dt_sidm;
    double dt_Grho = ts->dyntol / sqrt(C_Grho * p->G * mass[i] * hinv3);
    if (dt_Grho < dt) dt = dt_Grho;
    if (dt > 1.3 * dtold) { if (mode != 2) dt = 1.3 * dtold; }                          /* TIMESTEP_INCREASE_FACTOR, allvars.h:85 */
    int c = 0;
    if (dt >= ts->dtmax) { dt = ts->dtmax * (1.00 + 0.02 * jitter[a]); c = 1; }
    if (dt < ts->dtmin) { dt = ts->dtmin; dt *= 1.0 + 0.02 * jitter[a]; c = 1; }
    clamped += c;
    maxpred[i] = curtime[i] + 0.5 * dt;                                                 /* :315 */
  }
  return clamped;
}

/* ------------------------------------------------------------------ reflect(), reflection.c:7-33 (all float arithmetic) */
int oreflect(int nactive, const int *active, double radius, const float *pos, float *vel)
{
  const float r_ref2 = (float)(radius * radius);
  int n = 0;
  for (int a = 0; a < nactive; a++) {
    int i = active[a];
    float r2 = pos[3 * i] * pos[3 * i] + pos[3 * i + 1] * pos[3 * i + 1] + pos[3 * i + 2] * pos[3 * i + 2];
    if (r2 > r_ref2) {
      float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
      float rv = x * vel[3 * i] + y * vel[3 * i + 1] + z * vel[3 * i + 2];
      if (rv > 0) {
        float r2inv2 = 2.0f / r2;
        vel[3 * i] -= rv * x * r2inv2; vel[3 * i + 1] -= rv * y * r2inv2; vel[3 * i + 2] -= rv * z * r2inv2;
        n++;
      }
    }
  }
  return n;
}
