/* oracle/ref_harness.c - drives the UNMODIFIED reference (junkoda/sidm-nbody) as a
 * single-rank shared library so that tests can pin this repo's oracle and CUDA path
 * against the real thing.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or
 * executed from the product path (sidm-nbody_b200/).  Only tests/, the smoke check and
 * bench.py's cpu_baseline / --impl reference legs may load the library built from this.
 *
 * How it reaches the reference's private state: this translation unit textually
 * includes the reference's forcetree.c from where it lies under /root/reference (the
 * Makefile passes -I$(REF)); nothing is copied into the repo.  The tree arrays
 * (`nodes`, `next`, `nextnode`, `father`, `trees`, forcetree.c:27-75) are `static`
 * there, so this is the only way to dump them without touching the sources.
 *
 * The harness replaces main.c (which refuses NTask<=1, main.c:39-44), begrun() and
 * init() (which need a parameter file and an IC file): ref_setup() fills the same
 * `All` fields begrun()/init() would (begrun.c:16-60, init.c:20-180) from a plain
 * struct, ref_set_particles() fills P[1..N] the way init.c:76-100 does.
 */
#ifndef B200_SHIM
#include "forcetree.c"          /* the reference's own file, unmodified */
#else
/* drop-in build: the reference driver linked against sidm-nbody_b200/shim/b200_shim.c instead of
 * gravtree.c / forcetree.c / sidm.c (GPU box only; proves the boundary on the reference's own
 * accel.c:compute_accelerations) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <mpi.h>
#include "allvars.h"
#include "proto.h"
#endif
#include "sidm_rand.h"
#include <gsl/gsl_rng.h>

/* ------------------------------------------------------------------ config */

typedef struct ref_cfg {
  int    MaxPart;
  int    BufferSizeMB;
  double TreeAllocFactor;
  double ErrTolTheta;
  double ErrTolForceAcc;
  int    TypeOfOpeningCriterion;
  int    ComovingIntegrationOn;
  double MaxNodeMove;
  double TreeUpdateFrequency;
  double G;
  double SofteningHalo;
  int    DesNumNgb;
  int    MaxNumNgbDeviation;
  double CrossSectionInternal;
  double ProbabilityTol;
  int    Seed1, Seed2;
  double BoxSize;
  double Omega0, OmegaLambda, Hubble;
  double Time;
  double YukawaVelocity, CrossSectionPowLaw, CrossSectionVelScale;
} ref_cfg;

static int ref_ready = 0;

int ref_sizeof_particle(void) { return (int)sizeof(struct particle_data); }
#ifndef B200_SHIM
int ref_sizeof_node(void)     { return (int)sizeof(struct NODE); }
#endif

int ref_setup(const ref_cfg *c)
{
  if (ref_ready) {
    /* the reference never frees (allocate.c:168-185); re-setup only re-points params */
    if (c->MaxPart > All.MaxPart) return -1;
  }
  ThisTask = 0; NTask = 1; PTask = 0;
  if (!ref_ready) {
    memset(&All, 0, sizeof(All));
    All.MaxPart = c->MaxPart;
    All.MaxPartSph = 0;
    All.BufferSize = c->BufferSizeMB;
    All.PartAllocFactor = 1.0;
    All.TreeAllocFactor = c->TreeAllocFactor;
  }
  All.ErrTolTheta = c->ErrTolTheta;
  All.ErrTolForceAcc = c->ErrTolForceAcc;
  All.TypeOfOpeningCriterion = c->TypeOfOpeningCriterion;
  All.ComovingIntegrationOn = c->ComovingIntegrationOn;
  All.MaxNodeMove = c->MaxNodeMove;
  All.TreeUpdateFrequency = c->TreeUpdateFrequency;
  All.DomainUpdateFrequency = 1e30;
  All.G = c->G;
  All.SofteningHalo = All.SofteningHaloMaxPhys = c->SofteningHalo;
  for (int t = 0; t < 6; t++) All.SofteningTable[t] = All.SofteningTableMaxPhys[t] = 0;
  All.SofteningTable[1] = All.SofteningTableMaxPhys[1] = c->SofteningHalo;
  All.DesNumNgb = c->DesNumNgb;
  All.MaxNumNgbDeviation = c->MaxNumNgbDeviation;
  All.CrossSectionInternal = c->CrossSectionInternal;
  All.CrossSection = c->CrossSectionInternal;
  All.ProbabilityTol = c->ProbabilityTol;
#if (CROSS_SECTION_TYPE == 2 || CROSS_SECTION_TYPE == 4)
  All.YukawaVelocity = c->YukawaVelocity;
#elif (CROSS_SECTION_TYPE == 3)
  All.CrossSectionPowLaw = c->CrossSectionPowLaw; All.CrossSectionVelScale = c->CrossSectionVelScale;
#endif
  All.Seed1 = c->Seed1; All.Seed2 = c->Seed2;
  All.BoxSize = c->BoxSize; All.BoxHalf = c->BoxSize / 2;
  All.Omega0 = c->Omega0; All.OmegaLambda = c->OmegaLambda; All.Hubble = c->Hubble;
  All.Time = All.TimeBegin = c->Time;
  All.MinSizeTimestep = 0; All.MaxSizeTimestep = 1e30;
#ifdef REFLECTIONBOUNDARY
  All.ReflectionRadius = 1e30;
#endif
  if (!ref_ready) {
    set_sph_kernel();            /* begrun.c:22 */
    allocate_commbuffers();      /* begrun.c:24 */
    allocate_memory();           /* read_ic.c does this after counting particles */
    force_treeallocate(All.TreeAllocFactor * All.MaxPart, All.MaxPart);  /* init.c:120 */
    ngb_treeallocate(MAX_NGB);   /* init.c:130 */
    /* log files the diagnostics print into (begrun.c open_outputfiles) */
    FdInfo = fopen("/dev/null", "w"); FdEnergy = fopen("/dev/null", "w");
    FdTimings = fopen("/dev/null", "w"); FdCPU = fopen("/dev/null", "w");
    ref_ready = 1;
  }
#ifdef PERIODIC
  if (All.BoxSize > 0) ewald_init();
#endif
  return 0;
}

void ref_init_rand(int seed) { init_rand(seed, 0); }

/* P[1..n] filled as read_ic.c + init.c:76-100,124-128 leave them at start-up */
void ref_set_particles(int n, const float *pos, const float *vel, const float *mass, const int *id)
{
  NumPart = n; N_gas = 0;
  All.TotNumPart = n; All.TotN_gas = 0; All.TotN_halo = n;
  for (int i = 1; i <= n; i++) {
    struct particle_data *p = &P[i];
    memset(p, 0, sizeof(*p));
    for (int k = 0; k < 3; k++) {
      p->Pos[k] = p->PosPred[k] = pos[3 * (i - 1) + k];
      p->Vel[k] = p->VelPred[k] = vel[3 * (i - 1) + k];
    }
    p->Mass = mass[i - 1];
    p->ID = id[i - 1];
    p->Type = 1;
    p->GravCost = 1;
    p->CurrentTime = p->MaxPredTime = All.Time;
    p->ForceFlag = i + 1;
  }
  P[n].ForceFlag = 1; IndFirstUpdate = 1; NumForceUpdate = n; NumSphUpdate = 0;
  for (int t = 0; t < 6; t++) Ntype[t] = NtypeLocal[t] = 0;
  Ntype[1] = NtypeLocal[1] = n;
  All.NumForcesSinceLastTreeConstruction = All.TreeUpdateFrequency * All.TotNumPart;
  NoCostFlag = 1;
}

/* ------------------------------------------------------- field get / set */

enum { F_POS, F_VEL, F_MASS, F_ID, F_TYPE, F_CURTIME, F_MAXPRED, F_POSPRED, F_VELPRED,
       F_ACCEL, F_POT, F_GRAVCOST, F_OLDACC, F_FORCEFLAG, F_LEFT, F_RIGHT, F_NGB, F_HSML, F_DVEL };

static void *field_ptr(struct particle_data *p, int f, int *nbytes)
{
  switch (f) {
  case F_POS: *nbytes = 12; return p->Pos;
  case F_VEL: *nbytes = 12; return p->Vel;
  case F_MASS: *nbytes = 4; return &p->Mass;
  case F_ID: *nbytes = 4; return &p->ID;
  case F_TYPE: *nbytes = 4; return &p->Type;
  case F_CURTIME: *nbytes = 4; return &p->CurrentTime;
  case F_MAXPRED: *nbytes = 4; return &p->MaxPredTime;
  case F_POSPRED: *nbytes = 12; return p->PosPred;
  case F_VELPRED: *nbytes = 12; return p->VelPred;
  case F_ACCEL: *nbytes = 12; return p->Accel;
  case F_POT: *nbytes = 4; return &p->Potential;
  case F_GRAVCOST: *nbytes = 4; return &p->GravCost;
  case F_OLDACC: *nbytes = 4; return &p->OldAcc;
  case F_FORCEFLAG: *nbytes = 4; return &p->ForceFlag;
  case F_LEFT: *nbytes = 4; return &p->Left;
  case F_RIGHT: *nbytes = 4; return &p->Right;
  case F_NGB: *nbytes = 4; return &p->NgbVelDisp;
  case F_HSML: *nbytes = 4; return &p->HsmlVelDisp;
  case F_DVEL: *nbytes = 12; return p->dVel;
  }
  *nbytes = 0; return 0;
}

void ref_get_field(int f, void *out)
{
  int nb; char *o = (char *)out;
  for (int i = 1; i <= NumPart; i++) { void *s = field_ptr(&P[i], f, &nb); memcpy(o, s, nb); o += nb; }
}
void ref_set_field(int f, const void *in)
{
  int nb; const char *s = (const char *)in;
  for (int i = 1; i <= NumPart; i++) { void *d = field_ptr(&P[i], f, &nb); memcpy(d, s, nb); s += nb; }
}
/* whole AoS image, for the drop-in tests (same bytes the C-ABI sees) */
void ref_get_particles_raw(void *out) { memcpy(out, &P[1], (size_t)NumPart * sizeof(struct particle_data)); }
void ref_set_particles_raw(const void *in, int n)
{ NumPart = n; memcpy(&P[1], in, (size_t)n * sizeof(struct particle_data)); }

double ref_get_time(void) { return All.Time; }
void   ref_set_time(double t) { All.Time = t; }
void   ref_set_vmax(double v) { vmax = v; }
double ref_get_vmax_global(void) { return vmax; }
double ref_getvmax(void) { vmax = getvmax(); return vmax; }
int    ref_num_active(void) { return NumForceUpdate; }
void   ref_get_active(int *out)   /* 0-based indices in list order */
{ int i = IndFirstUpdate; for (int c = 0; c < NumForceUpdate; c++, i = P[i].ForceFlag) out[c] = i - 1; }
void   ref_set_active(const int *idx, int n)   /* 0-based, list order as given */
{
  NumForceUpdate = n; NumSphUpdate = 0; IndFirstUpdate = n ? idx[0] + 1 : 0;
  for (int c = 0; c < n; c++) P[idx[c] + 1].ForceFlag = idx[(c + 1) % n] + 1;
}
void   ref_get_domain(float *mn, float *mx)
{ for (int k = 0; k < 3; k++) { mn[k] = DomainMin[1][k]; mx[k] = DomainMax[1][k]; } }
void   ref_get_cpu(double *out)
{
  out[0] = All.CPU_Gravity; out[1] = All.CPU_TreeConstruction; out[2] = All.CPU_TreeWalk;
  out[3] = All.CPU_EnsureNgb; out[4] = All.CPU_CommSum; out[5] = All.CPU_Predict; out[6] = All.CPU_TimeLine;
}

/* ------------------------------------------------------- timeline helper */

/* All particles at CurrentTime=tcur with MaxPredTime=tnext.  construct_timetree()
 * (timeline.c:127) would turn N equal keys into an N-deep chain and blow the C stack in
 * the recursive find_next_time_walk(); any binary search tree over equal keys is valid,
 * so build a balanced one here, then let the reference's own find_next_time() derive
 * All.Time and the active list (ascending index). */
static int balanced(int lo, int hi)
{
  if (lo > hi) return 0;
  int mid = lo + (hi - lo) / 2;
  PTimeTree[mid].left = balanced(lo, mid - 1);
  PTimeTree[mid].right = balanced(mid + 1, hi);
  return mid;
}
void ref_all_active(double tcur, double tnext)
{
  for (int i = 1; i <= NumPart; i++) { P[i].CurrentTime = tcur; P[i].MaxPredTime = tnext; }
  TimeTreeRoot = balanced(1, NumPart);
  All.Time = tcur;
  find_next_time();
}

#ifndef B200_SHIM
/* ------------------------------------------------------- tree dump */

int ref_tree_first(void) { return trees[1] - All.MaxPart; }
int ref_tree_numnodes(void) { return numnodestree[1]; }

/* per node (index = node - All.MaxPart - first): 24 floats + 13 ints */
void ref_dump_nodes(float *f, int *ii)
{
  int n = numnodestree[1];
  for (int k = 0; k < n; k++) {
    struct NODE *nd = &nodes[trees[1] + k];
    float *o = f + 24 * k;
    o[0] = nd->len; o[1] = nd->mass;
    o[2] = nd->s[0]; o[3] = nd->s[1]; o[4] = nd->s[2];
    o[5] = nd->center[0]; o[6] = nd->center[1]; o[7] = nd->center[2];
    o[8] = nd->Q11; o[9] = nd->Q22; o[10] = nd->Q33; o[11] = nd->Q12; o[12] = nd->Q13; o[13] = nd->Q23;
    o[14] = nd->P; o[15] = nd->vs[0]; o[16] = nd->vs[1]; o[17] = nd->vs[2];
    o[18] = nd->tilu; o[19] = nd->oc; o[20] = nd->hmax;
#ifdef BMAX
    o[21] = nd->bmax2;
#else
    o[21] = 0;
#endif
    o[22] = o[23] = 0;
    int *q = ii + 13 * k;
    for (int j = 0; j < 8; j++) q[j] = nd->suns[j] >= All.MaxPart ? nd->suns[j] - trees[1] + (1 << 30) : nd->suns[j];
    q[8] = nd->sibling >= All.MaxPart ? nd->sibling - trees[1] + (1 << 30) : nd->sibling;
    q[9] = nd->partind; q[10] = nd->count; q[11] = nd->cost;
    q[12] = father[trees[1] + k] >= 0 ? father[trees[1] + k] - trees[1] : -1;
  }
}
void ref_dump_next(int *out) { for (int i = 0; i < NumPart; i++) out[i] = next[i]; }
void ref_dump_particle_father(int *out)
{ for (int i = 0; i < NumPart; i++) out[i] = father[i] - trees[1]; }

/* ------------------------------------------------------- force evaluation */

/* tree force for targets idx[] (0-based) using P[].PosPred / OldAcc, exactly the
 * per-target call of gravtree.c:189-190; cost = (particle, node) interaction counts
 * from the reference's own -DDIAG counters (forcetree.c:65-66). */
void ref_force_tree(int n, const int *idx, double *acc, int *cost)
{
  for (int t = 0; t < n; t++) {
    struct particle_data *p = &P[idx[t] + 1];
    for (int k = 0; k < 3; k++) GravDataIn[0].Pos[k] = p->PosPred[k];
    GravDataIn[0].Type = p->Type;
    GravDataIn[0].OldAcc = p->OldAcc;
    int c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    for (int tr = 0; tr < 6; tr++) { c0 += treecost[tr]; c1 += treecost_quadru[tr]; }      /* one tree per particle type */
    force_treeevaluate(0, 1.0);
    for (int k = 0; k < 3; k++) acc[3 * t + k] = GravDataResult[0].Acc[k];
    for (int tr = 0; tr < 6; tr++) { d0 += treecost[tr]; d1 += treecost_quadru[tr]; }
    if (cost) { cost[2 * t] = d0 - c0; cost[2 * t + 1] = d1 - c1; }
  }
}
void ref_potential(int n, const int *idx, double *pot)     /* forcetree.c:1389, tree already built */
{
  for (int t = 0; t < n; t++) {
    struct particle_data *p = &P[idx[t] + 1];
    for (int k = 0; k < 3; k++) GravDataIn[0].Pos[k] = p->PosPred[k];
    GravDataIn[0].Type = p->Type;
    GravDataIn[0].OldAcc = p->OldAcc;
    force_treeevaluate_potential(0);
    pot[t] = GravDataPotential[0];
  }
}
void ref_force_direct(int n, const int *idx, double *acc)
{
  for (int t = 0; t < n; t++) {
    struct particle_data *p = &P[idx[t] + 1];
    for (int k = 0; k < 3; k++) GravDataIn[0].Pos[k] = p->PosPred[k];
    GravDataIn[0].Type = p->Type;
    GravDataIn[0].OldAcc = p->OldAcc;
    force_treeevaluate_direct(0, 1.0);
    for (int k = 0; k < 3; k++) acc[3 * t + k] = GravDataResult[0].Acc[k];
  }
}

/* ------------------------------------------------------- neighbour search */

int ref_ngb_variable(const float *xyz, float h, int *list, float *r2, int cap)
{
  int *nl; float *rl; float x[3] = { xyz[0], xyz[1], xyz[2] };
  int n = ngb_treefind_variable(x, h, 1, &nl, &rl);
  for (int i = 0; i < n && i < cap; i++) { list[i] = nl[i]; r2[i] = rl[i]; }
  return n;
}
float ref_ngb_treefind(const float *xyz, int desngb, float hguess)
{
  int *nl; float *rl; float x[3] = { xyz[0], xyz[1], xyz[2] };
  return ngb_treefind(x, desngb, hguess, 1, &nl, &rl);
}

#endif /* !B200_SHIM */

/* ------------------------------------------------------- hot-path entry points */

int ref_treebuild(void) { return force_treebuild(); }   /* uses P[].PosPred, forcetree.c:90 */

void ref_gravity_tree(void) { gravity_tree(); }
void ref_determine_interior(void) { determine_interior(); }
void ref_sidm(void) { sidm(); }
void ref_setup_nbr_sidm(void) { setup_nbr_sidm(); }
void ref_sidm_ensure_neighbours(int mode) { sidm_ensure_neighbours(mode); }
void ref_setup_smoothinglengths_sidm(int desngb) { setup_smoothinglengths_sidm(desngb); }
void ref_compute_accelerations(int mode) { compute_accelerations(mode); }
void ref_advance(void) { advance(); }
#ifdef SIDM
int  ref_n_scat_particles(void) { return n_scat_particles; }   /* predict.c:258,268-269: particles that carried a kick into advance() */
#endif
/* several particle types (one tree per type, forcetree.c:90-158): softening per type; types go in through ref_set_field(F_TYPE) */
void ref_set_softening(int type, double eps) { All.SofteningTable[type] = All.SofteningTableMaxPhys[type] = eps; }
/* global.c:18: SysState as 102 doubles (allvars.h:517-537) */
int ref_global_quantities(double *out) { compute_global_quantities_of_system(); memcpy(out, &SysState, sizeof(SysState)); return (int)(sizeof(SysState) / sizeof(double)); }
/* savepositions(), io.c:16: one file <dir><base>_<num> in format 1 */
void ref_savepositions(int num, const char *dir, const char *base, const double *mass_table, double hubble_param)
{
  strcpy(All.OutputDir, dir); strcpy(All.SnapshotFileBase, base);
  All.NumFilesPerSnapshot = 1; All.CoolingOn = 0; All.StarformationOn = 0; All.MultiPhaseModelOn = 0;
  for (int t = 0; t < 6; t++) All.MassTable[t] = mass_table[t];
  All.HubbleParam = hubble_param;
  savepositions(num);
}
#ifdef REFLECTIONBOUNDARY
void ref_reflect(double radius) { All.ReflectionRadius = radius; reflect(); All.ReflectionRadius = 1e30; }   /* reflection.c:7 */
#endif
void ref_compute_potential(void) { compute_potential(); }   /* potential.c:18: rebuilds the tree, all particles */
/* timestep.c:17 with the accuracy parameters of the parameter file set here; mode 2 = start-up (no growth limit) */
void ref_find_timesteps(int mode, int crit, double eta, double velscale, double probtol, double dyntol, double dtmax, double dtmin)
{
  All.TypeOfTimestepCriterion = crit; All.ErrTolIntAccuracy = eta; All.ErrTolVelScale = velscale;
  All.ProbabilityTol = probtol; All.ErrTolDynamicalAccuracy = dyntol; All.MaxSizeTimestep = dtmax; All.MinSizeTimestep = dtmin;
  find_timesteps(mode);
}
/* k iterations of the main loop of run.c:34-150 (no statistics / snapshots / domain decomposition):
 * find_next_time(), compute_accelerations(0), advance(), reflect(), update_node_sidm(), find_timesteps(0).
 * time_out[s], nactive_out[s] = All.Time and NumForceUpdate of iteration s. */
void ref_run_steps(int k, double *time_out, int *nactive_out)
{
  for (int s = 0; s < k; s++) {
    find_next_time();
    time_out[s] = All.Time; nactive_out[s] = NumForceUpdate;
    compute_accelerations(0);
    advance();
#ifdef REFLECTIONBOUNDARY
    reflect();
#endif
    update_node_sidm();
    find_timesteps(0);
    All.NumCurrentTiStep++;
  }
}
void ref_force_rebuild_next(void)
{ All.NumForcesSinceLastTreeConstruction = 1 << 30; }
void ref_set_snapcount(int c) { All.SnapshotFileCount = c; }
