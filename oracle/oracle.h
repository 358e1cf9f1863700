/* oracle.h - CPU restatement of the reference's hot path (tree gravity + SIDM scatter).
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared with; it is
 * never linked into, imported by, or executed from the product (sidm-nbody_b200/).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so
 * every function here is checked against the unmodified reference compiled single-rank
 * (oracle/_ref, see oracle/Makefile) in tests/test_oracle_vs_reference.py, and against the
 * fixtures generated from it by tests/golden/make_golden.py.
 *
 * Each function cites the reference file:line it restates.  The algorithms follow the
 * reference (sequential insertion tree, pointer walks, per-particle RNG stream); data
 * layout and code are this repo's own.
 */
#ifndef ORACLE_H
#define ORACLE_H

typedef struct otree otree;

typedef struct oparams {
  double theta;            /* All.ErrTolTheta */
  double alpha;            /* All.ErrTolForceAcc */
  int    criterion;        /* All.TypeOfOpeningCriterion */
  double eps;              /* Plummer softening: of the particle type, or max(tree type, target type) in a forest */
  double G;
  int    des_ngb, max_dev; /* All.DesNumNgb, All.MaxNumNgbDeviation */
  double sigma;            /* All.CrossSectionInternal */
  int    xs_type;          /* the reference's compile-time CROSS_SECTION_TYPE, 0..4 (sidm.c:226-316, 366-439) */
  double vc;               /* All.YukawaVelocity        (types 2, 4) */
  double pl_n, pl_v0;      /* All.CrossSectionPowLaw, All.CrossSectionVelScale (type 3) */
} oparams;

/* ---- tree (forcetree.c:90-571) ---- */
otree *otree_build(int n, const float *pos, const float *mass, double eps);
void   otree_free(otree *t);
int    otree_num_nodes(const otree *t);
int    otree_random_subnodes(const otree *t);   /* times forcetree.c:320-326 would have used rand() */
void   otree_dump(const otree *t, float *center, float *len, float *mass, float *s, float *Q /*[m][7]*/,
                  float *oc, float *bmax2, int *count);
void   otree_chain(const otree *t, int *order);  /* particles in next[] order (forcetree.c:274-279) */
void   otree_domain(const otree *t, float *mn, float *mx);

/* ---- forces (forcetree.c:786-1377, 1896-1975) ---- */
void   otree_force(const otree *t, const oparams *p, int nt, const int *targets, const float *oldacc,
                   double *acc, int *cost /*[nt][2] particle,node*/);
void   otree_direct(const otree *t, const oparams *p, int nt, const int *targets, double *acc);
/* several particle types = one tree per type (forcetree.c:90-158, 798-808, 1397-1409): the same walks for targets given
 * by position, with p->eps = max(eps of the tree's type, eps of the target's type); keep != 0 adds to acc / cost / pot
 * (second and later trees of a target, in ascending type order like the reference's loop) */
void   otree_force_at(const otree *t, const oparams *p, int nt, const float *xyz, const float *oldacc, double *acc, int *cost, int keep);
void   otree_potential_at(const otree *t, const oparams *p, int nt, const float *xyz, const float *oldacc, double *pot, int keep);
/* gravtree.c:230-324 epilogue for non-comoving runs */
void   ograv_epilogue(const oparams *p, int nt, const double *acc, float *accel, float *oldacc);

/* ---- neighbours (forcetree.c:2163-2414) ---- */
int    ongb_variable(const otree *t, const float xyz[3], float h, int *list, float *r2, int cap);
float  ongb_treefind(const otree *t, const float xyz[3], int desngb);

/* ---- SIDM (sidm.c:57-627, 814-990; sidm_rand.c, sidm_rand.h) ---- */
typedef struct orng orng;
orng  *orng_new(unsigned long seed, int warmup);   /* init_rand(): MT19937 + 1000001 discarded draws */
void   orng_free(orng *r);
double orng_uniform(orng *r);
long   orng_count(const orng *r);

typedef struct osidm_out {
  /* per buffer slot */
  int    *slot_particle;   /* slot -> particle index */
  double *rand;            /* the uniform drawn at sidm.c:341 */
  double *dir;             /* [nslot][3] random_direction() result where a partner was found */
  double *pmax;            /* sidm.c:338 */
  double *prob;            /* cumulative probability when the loop ended (sidm.c:352-383) */
  int    *partner;         /* SidmTarget-1, or -1 */
  int    *ngb;             /* numngb */
  int     nslot;
  int     sct[4];          /* ntot, pass1, scattered, rejected (sidm.c:614-620) */
  /* scatter log (sidm.c:571-601): id1 id2 per event + dv */
  int     nlog; int *log_i; int *log_j; float *log_dv;
  /* CROSS_SECTION_TYPE 4 only: the pairs (rand, uniform of cosO) drawn at sidm.c:393-394 for every
   * crossing of slot s are extra[2*extra_off[s] .. 2*extra_off[s+1]) */
  double *extra; int *extra_off;
} osidm_out;

/* one sidm() call for the active list (particle indices, list order).  Reads/writes the
 * particle arrays like the reference: vel (Vel), hsml, dvel, ngbcount. */
void   osidm_pass(const otree *t, const oparams *p, int nactive, const int *active,
                  const float *vel, const float *mass, const float *hsml, const float *dt /*per particle: 2(t-t_i)*/,
                  float *dvel, int *ngbcount, double vmax, orng *rng, osidm_out *out);
void   osidm_out_free(osidm_out *o);
double ogetvmax(int n, const float *vel);          /* sidm.c:970-990 */
/* sidm.c:814-968 for mode 0 with every particle on the same step: repairs hsml until all
 * active neighbour counts are in range; returns passes done or -1155. */
int    osidm_ensure(const otree *t, const oparams *p, int n, const float *vel, const float *mass, float *hsml,
                    const float *dt, float *dvel, int *ngbcount, float *left, float *right, double vmax, orng *rng);
/* force_treeevaluate_potential(), forcetree.c:1389-1755 and the epilogue of compute_potential(), potential.c:131-168 */
void   otree_potential(const otree *t, const oparams *p, int nt, const int *targets, const float *oldacc, double *pot);
void   opot_epilogue(const oparams *p, int nt, const double *pot, const float *mass, float *out);
/* reflect(), reflection.c:7-33: specular reflection of outgoing active particles beyond the radius; returns how many */
int    oreflect(int nactive, const int *active, double radius, const float *pos, float *vel);
/* find_timesteps(mode), timestep.c:17-334, collisionless particles of type 1, no comoving integration,
 * steps that hit Max/MinSizeTimestep take the uniform jitter[a] (the reference: drand48()).  Writes
 * maxpred[i] = CurrentTime + dt/2 for the active particles; returns the number of clamped steps. */
typedef struct otimestep { int crit; double eta, velscale, probtol, dyntol, dtmax, dtmin; } otimestep;
int    ofind_timesteps(const oparams *p, const otimestep *ts, int nactive, const int *active, int mode, double time, double vmax,
                       const float *accel, const float *curtime, float *maxpred, const float *hsml, const float *mass, const double *jitter);
/* compute_global_quantities_of_system(), global.c:18-135, one rank, no gas: per-type and total mass, energies,
 * momentum, angular momentum, centre of mass.  Field order of `struct state_of_system` (allvars.h:517-537). */
typedef struct osysstate {
  double Mass, EnergyKin, EnergyPot, EnergyInt, EnergyTot, Momentum[4], AngMomentum[4], CenterOfMass[4];
  double MassComp[5], EnergyKinComp[5], EnergyPotComp[5], EnergyIntComp[5], EnergyTotComp[5];
  double MomentumComp[5][4], AngMomentumComp[5][4], CenterOfMassComp[5][4];
} osysstate;
void   oglobal_quantities(int n, const float *pospred, const float *velpred, const float *mass, const float *potential,
                          const int *type, osysstate *out);
#endif
