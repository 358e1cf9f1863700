/* b200_shim.c - the reference's own hot-path symbols, re-implemented on libsidm_b200.so.
 *
 * Compile this file with the reference's headers (-I<reference>/nbody, the same -D flags as
 * the rest of the build) and link it INSTEAD OF gravtree.c, forcetree.c and sidm.c; the driver
 * (main.c run.c accel.c timeline.c timestep.c predict.c begrun.c init.c io.c restart.c ...)
 * compiles unchanged and calls these functions where it used to call the CPU ones:
 *
 *   gravity_tree()            gravtree.c:18     accel.c:39
 *   sidm()                    sidm.c:57         accel.c:63, sidm.c:931 (repair loop - here on the GPU)
 *   sidm_ensure_neighbours()  sidm.c:814        accel.c:64
 *   setup_nbr_sidm()          sidm.c:630        init.c:446,502
 *   getvmax()                 sidm.c:970        init.c:74, begrun.c:106, run.c:122
 *   update_node_sidm()        sidm.c:992        run.c (no-op: the GPU rebuilds the tree every step)
 *   force_treeallocate/build/free, force_costevaluate/resetcost/getcost_*   forcetree.h:9-27
 *   ngb_treeallocate/build/free, ngb_update_nodes, ngb_treefind             forcetree.h:30-40
 *   set_softenings()          gravtree.c:425    (lives in the replaced file, restated here)
 *   compute_potential()       potential.c:18    only with -DB200_SHIM_ACCEL, linked instead of potential.c
 *   compute_accelerations()   accel.c:27        only with -DB200_SHIM_ACCEL, linked instead of accel.c:
 *                                               one upload, ONE library call for gravity + sidm + repair
 *                                               loop (walk and SIDM chain overlapped on two CUDA streams),
 *                                               one download - the fast form of the drop-in
 *
 * Error convention: a non-zero return of the C ABI becomes endrun(code) like the CPU code
 * (endrun.c:19-30).  Timers: elapsed device time goes into the same All.CPU_* fields.
 * What differs from the CPU code on purpose: the tree is rebuilt on every gravity_tree()
 * call (TreeUpdateFrequency is ignored; the dynamic node updates of forcetree.c:935-954
 * vanish), P[i].GravCost receives the target's interaction count, and random numbers come
 * from the library's counter-based generator seeded with All.Seed1 + All.Seed2*ThisTask.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <mpi.h>

#include "allvars.h"
#include "proto.h"
#include "sidm_b200.h"

static int shim_ready = 0;
static int *active_list = 0;
static int active_cap = 0;

static void b200_check(int rc, const char *what)
{
  if (rc != B200_OK) {
    printf("task %d: %s failed in libsidm_b200 with code %d\n", ThisTask, what, rc);
    endrun(rc);
  }
}

static void fill_params(b200_params *p)
{
  int t;
  memset(p, 0, sizeof(*p));
  p->device = ThisTask;                       /* one rank per GPU of the box */
  p->MaxPart = All.MaxPart;
  p->TreeAllocFactor = All.TreeAllocFactor;
  p->ErrTolTheta = All.ErrTolTheta;
  p->ErrTolForceAcc = All.ErrTolForceAcc;
  p->TypeOfOpeningCriterion = All.TypeOfOpeningCriterion;
  p->ComovingIntegrationOn = All.ComovingIntegrationOn;
  p->G = All.G;
  for (t = 0; t < 6; t++) p->SofteningTable[t] = All.SofteningTable[t];
  p->BoxSize = All.BoxSize;
  p->PeriodicBoundariesOn = All.PeriodicBoundariesOn;
  p->Omega0 = All.Omega0; p->OmegaLambda = All.OmegaLambda; p->Hubble = All.Hubble;
  p->DesNumNgb = All.DesNumNgb;
  p->MaxNumNgbDeviation = All.MaxNumNgbDeviation;
#ifdef SIDM
  p->CrossSectionInternal = All.CrossSectionInternal;
  p->CrossSectionType = CROSS_SECTION_TYPE;
#if (CROSS_SECTION_TYPE == 2) || (CROSS_SECTION_TYPE == 4)
  p->YukawaVelocity = All.YukawaVelocity;
#elif (CROSS_SECTION_TYPE == 3)
  p->CrossSectionPowLaw = All.CrossSectionPowLaw;
  p->CrossSectionVelScale = All.CrossSectionVelScale;
#endif
  p->Seed = (unsigned long long)(All.Seed1 + All.Seed2 * ThisTask);
  p->BunchSizeSidm = 0;
#endif
}

static void fill_layout(b200_layout *l)
{
  l->stride = (int)sizeof(struct particle_data);
  l->Pos = (int)offsetof(struct particle_data, Pos);
  l->Vel = (int)offsetof(struct particle_data, Vel);
  l->Mass = (int)offsetof(struct particle_data, Mass);
  l->ID = (int)offsetof(struct particle_data, ID);
  l->Type = (int)offsetof(struct particle_data, Type);
  l->CurrentTime = (int)offsetof(struct particle_data, CurrentTime);
  l->PosPred = (int)offsetof(struct particle_data, PosPred);
  l->VelPred = (int)offsetof(struct particle_data, VelPred);
  l->Accel = (int)offsetof(struct particle_data, Accel);
  l->GravCost = (int)offsetof(struct particle_data, GravCost);
  l->OldAcc = (int)offsetof(struct particle_data, OldAcc);
  l->Left = (int)offsetof(struct particle_data, Left);
  l->Right = (int)offsetof(struct particle_data, Right);
  l->NgbVelDisp = (int)offsetof(struct particle_data, NgbVelDisp);
  l->HsmlVelDisp = (int)offsetof(struct particle_data, HsmlVelDisp);
  l->dVel = (int)offsetof(struct particle_data, dVel);
  l->MaxPredTime = (int)offsetof(struct particle_data, MaxPredTime);
  l->Potential = (int)offsetof(struct particle_data, Potential);
}

/* walk the ForceFlag-linked active list (timeline.c:20-80) into a 0-based index array */
static int gather_active(void)
{
  int i, c;
  if (active_cap < NumForceUpdate + 1) {
    free(active_list);
    active_cap = All.MaxPart + 1;
    active_list = (int *)malloc(sizeof(int) * active_cap);
    if (!active_list) endrun(3);
  }
  for (i = IndFirstUpdate, c = 0; c < NumForceUpdate; i = P[i].ForceFlag, c++) active_list[c] = i - 1;
  return NumForceUpdate;
}

static void sync_params_and_particles(void)
{
  b200_params p;
  b200_layout l;
  fill_params(&p);
  b200_check(b200_set_params(&p), "b200_set_params");
  fill_layout(&l);
  b200_check(b200_bind_particles(&P[1], NumPart, &l, 1), "b200_bind_particles");
  b200_check(b200_upload(), "b200_upload");
}

/* ---------------------------------------------------------------- forcetree.h surface */

void force_treeallocate(int maxnodes, int maxpart)      /* forcetree.c:1797 */
{
  b200_params p;
  (void)maxnodes; (void)maxpart;
  if (shim_ready) return;
  fill_params(&p);
  b200_check(b200_init(&p), "b200_init");
  shim_ready = 1;
}
void force_treefree(void) { b200_finalize(); shim_ready = 0; }

int force_treebuild(void)                               /* forcetree.c:90 (uses P[].PosPred) */
{
  b200_counters c;
  int t, i;
  sync_params_and_particles();
  b200_check(b200_tree_build(), "b200_tree_build");
  b200_get_counters(&c);
  for (t = 0; t < 6; t++) NtypeLocal[t] = 0;
  for (i = 1; i <= NumPart; i++) NtypeLocal[P[i].Type & 7]++;
  MPI_Allreduce(NtypeLocal, Ntype, 5, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  return c.num_nodes;
}
void force_costevaluate(void) {}                        /* GravCost already holds per-target counts */
void force_resetcost(void) {}
int  force_getcost_single(void) { b200_counters c; b200_get_counters(&c); return (int)c.part_interactions; }
int  force_getcost_quadru(void) { b200_counters c; b200_get_counters(&c); return (int)c.node_interactions; }

void ngb_treeallocate(int npart) { (void)npart; }
void ngb_treefree(void) {}
void ngb_treebuild(void) { force_treebuild(); }         /* forcetree.c:2438: same tree */
void ngb_update_nodes(void) {}                          /* forcetree.c:2486: tree is rebuilt instead */

float ngb_treefind(float xyz[3], int desngb, float hguess, int parttype, int **ngblistback, float **r2listback)
{                                                       /* forcetree.c:2311; only P[i].PosPred callers exist */
  int i, idx = -1;
  float h2 = 0;
  (void)hguess; (void)parttype;
  i = (int)(((char *)xyz - (char *)&P[1].PosPred[0]) / (long)sizeof(struct particle_data));
  if (i >= 0 && i < NumPart && xyz == P[i + 1].PosPred) idx = i;
  if (idx < 0) { printf("ngb_treefind: only particle positions are supported by the GPU path\n"); endrun(9003); }
  b200_check(b200_ngb_treefind(&idx, 1, desngb, &h2), "b200_ngb_treefind");
  if (ngblistback) *ngblistback = 0;
  if (r2listback) *r2listback = 0;
  return h2;
}

/* SPH / potential entry points of the replaced file: not on this path (configs have no gas) */
int  ngb_treefind_pairs(float xyz[3], float hsml, int **a, float **b) { (void)xyz; (void)hsml; (void)a; (void)b; endrun(9006); return 0; }
int  ngb_treefind_variable(float xyz[3], float h, int t, int **a, float **b) { (void)xyz; (void)h; (void)t; (void)a; (void)b; endrun(9006); return 0; }
void force_treeevaluate_potential(int target) { (void)target; endrun(9006); }
void update_node_of_scat_particle(int i) { (void)i; }

/* ---------------------------------------------------------------- gravtree.c */

void set_softenings(void)                               /* gravtree.c:425-458 */
{
  const double soft[5] = { All.SofteningGas, All.SofteningHalo, All.SofteningDisk, All.SofteningBulge, All.SofteningStars };
  const double maxp[5] = { All.SofteningGasMaxPhys, All.SofteningHaloMaxPhys, All.SofteningDiskMaxPhys,
                           All.SofteningBulgeMaxPhys, All.SofteningStarsMaxPhys };
  int t;
  for (t = 0; t < 5; t++) All.SofteningTable[t] = (soft[t] * All.Time > maxp[t]) ? maxp[t] / All.Time : soft[t];
  All.MinGasHsml = All.MinGasHsmlFractional * All.SofteningTable[0];
}

void gravity_tree(void)                                 /* gravtree.c:18-419 */
{
  b200_counters c;
  int n, ntot;
  double t0 = second(), t1;
  if (All.ComovingIntegrationOn) set_softenings();
  MPI_Allreduce(&NumForceUpdate, &ntot, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  All.NumForcesSinceLastDomainDecomp += ntot;
  All.NumForcesSinceLastTreeConstruction = 0;
  if (ThisTask == 0) printf("Tree construction.\n");
  sync_params_and_particles();
  b200_check(b200_predict(All.Time), "b200_predict");               /* gravtree.c:72 */
  b200_check(b200_tree_build(), "b200_tree_build");                 /* gravtree.c:76 */
  n = gather_active();
  b200_check(b200_gravity(active_list, n, All.Time), "b200_gravity"); /* gravtree.c:127-324 */
  b200_check(b200_download(), "b200_download");
  b200_get_counters(&c);
  All.CPU_TreeConstruction += 1e-3 * (c.ms_build + c.ms_predict);
  All.CPU_TreeWalk += 1e-3 * c.ms_walk;
  All.CPU_CommSum += 1e-3 * (c.ms_upload + c.ms_download);
  All.TotNumOfForces += ntot;
  NoCostFlag = 0;
  t1 = second();
  (void)t0; (void)t1;
}

/* ---------------------------------------------------------------- sidm.c */

#ifdef SIDM
double getvmax(void)                                    /* sidm.c:970-990 */
{
  double v = 0;
  sync_params_and_particles();
  b200_check(b200_getvmax(&v), "b200_getvmax");
#ifdef FINDNBRLOG
  if (ThisTask == 0) fprintf(stdout, "Vmax= %g Processor %d\n", v, ThisTask);
#endif
  return v;
}

static void print_sct(void)
{
#ifdef FINDNBRLOG
  b200_counters c;
  b200_get_counters(&c);
  if (ThisTask == 0) fprintf(stdout, "SCT %d %d %d %d\n", c.sct_ntot, c.sct_pass1, c.sct_scattered, c.sct_rejected);
#endif
}

void sidm(void)                                         /* sidm.c:57-627; tree + particles are on the device
                                                           since gravity_tree() of this step (accel.c:39,63) */
{
  int n = gather_active();
  b200_check(b200_sidm(active_list, n, All.Time, vmax, 0), "b200_sidm");
  print_sct();
}

void setup_nbr_sidm(void)                               /* sidm.c:630-805 */
{
  int n = gather_active();
  sync_params_and_particles();
  b200_check(b200_tree_build(), "b200_tree_build");
  b200_check(b200_setup_nbr_sidm(active_list, n), "b200_setup_nbr_sidm");
  b200_check(b200_download(), "b200_download");
}

void sidm_ensure_neighbours(int mode)                   /* sidm.c:814-968 */
{
  b200_counters c;
  double save;
  int i;
  b200_check(b200_sidm_ensure_neighbours(mode, All.Time, vmax, 0), "b200_sidm_ensure_neighbours");
  b200_check(b200_download(), "b200_download");
  b200_get_counters(&c);
  print_sct();
  All.CPU_CommSum += 1e-3 * c.ms_download;
  if (c.ensure_iterations > 0) {
    if (mode == 0) {                                    /* sidm.c:943-955: restore the time line */
      save = All.TimeStep;
      find_next_time();
      All.TimeStep = save;
    } else {                                            /* sidm.c:956-965 */
      for (i = 1; i <= NumPart; i++) P[i].ForceFlag = i + 1;
      P[NumPart].ForceFlag = 1; IndFirstUpdate = 1; NumForceUpdate = NumPart; NumSphUpdate = N_gas;
    }
  }
}

void update_node_sidm(void) {}                          /* sidm.c:992-997 */

#ifdef B200_SHIM_ACCEL
/* compute_potential(), potential.c:18-180 (linked instead of potential.c): potential of all particles for
 * the energy statistics (run.c:51-60 -> global.c) */
void compute_potential(void)
{
  double t0 = second(), t1;
  if (All.ComovingIntegrationOn) set_softenings();
  if (ThisTask == 0) { printf("Start computation of potential for all particles...\n"); fflush(stdout); }
  sync_params_and_particles();
  b200_check(b200_compute_potential(0), "b200_compute_potential");
  b200_check(b200_download(), "b200_download");
  All.NumForcesSinceLastTreeConstruction = All.TreeUpdateFrequency * All.TotNumPart;   /* potential.c:49 */
  NoCostFlag = 1;
  if (ThisTask == 0) { printf("potential done.\n"); fflush(stdout); }
  t1 = second();
  All.CPU_Potential += timediff(t0, t1);
}

/* compute_global_quantities_of_system(), global.c:18-135 (linked instead of global.c): the sums come from the
 * device state that compute_potential() / compute_accelerations() left there; b200_sysstate has the field order
 * of struct state_of_system (allvars.h:517-537). */
void compute_global_quantities_of_system(void)
{
  b200_sysstate s;
  sync_params_and_particles();                          /* the host may have advanced P[] since the last device call */
  b200_check(b200_compute_global_quantities(&s), "b200_compute_global_quantities");
  memcpy(&SysState, &s, sizeof(SysState) < sizeof(s) ? sizeof(SysState) : sizeof(s));
}

/* savepositions(), io.c:16-590 (linked instead of io.c): one format-1 file written from the device state */
size_t my_fwrite(void *ptr, size_t size, size_t nmemb, FILE *stream)          /* io.c:594-605, used by ewald.c:136 */
{
  size_t nwritten = fwrite(ptr, size, nmemb, stream);
  if (nwritten != nmemb) { printf("I/O error (fwrite) on task=%d has occured.\n", ThisTask); fflush(stdout); endrun(777); }
  return nwritten;
}
size_t my_fread(void *ptr, size_t size, size_t nmemb, FILE *stream)           /* io.c:611-622, used by read_ic.c, ewald.c */
{
  size_t nread = fread(ptr, size, nmemb, stream);
  if (nread != nmemb) { printf("I/O error (fread) on task=%d has occured.\n", ThisTask); fflush(stdout); endrun(778); }
  return nread;
}
void savepositions(int num)
{
  char buf[500];
  double t0 = second(), t1;
  if (ThisTask == 0) printf("\nwriting snapshot file... \n");
  if (num < 0) num = 1000 + num;                                                /* io.c:77-78 */
  if (All.NumFilesPerSnapshot != 1 || NTask != 1 || All.TotN_gas > 0) {
    printf("savepositions: the device writer covers one file, one task, no gas\n"); endrun(9003);
  }
  sprintf(buf, "%s%s_%03d", All.OutputDir, All.SnapshotFileBase, num);         /* io.c:96 */
  sync_params_and_particles();
  b200_check(b200_savepositions(buf, All.Time, All.MassTable, All.HubbleParam, 0), "b200_savepositions");
  if (ThisTask == 0) printf("done with snapshot.\n");
  t1 = second();
  All.CPU_Snapshot += timediff(t0, t1);
}

/* accel.c:27-132 for collisionless runs, as one coarse call.  Same order of effects as the CPU code:
 * gravity_tree() bookkeeping (gravtree.c:42-60), forces, determine_interior(), sidm() +
 * sidm_ensure_neighbours(mode) when mode == 0, timers into All.CPU_Gravity / CPU_EnsureNgb. */
void compute_accelerations(int mode)
{
  b200_counters c;
  int n, ntot, i;
  double save;
  if (ThisTask == 0) { printf("Start force computation...\n"); fflush(stdout); }
  if (All.TotN_gas > 0) { printf("compute_accelerations: gas particles are not on the GPU path\n"); endrun(9006); }
  if (All.ComovingIntegrationOn) set_softenings();
  MPI_Allreduce(&NumForceUpdate, &ntot, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  All.NumForcesSinceLastDomainDecomp += ntot;
  All.NumForcesSinceLastTreeConstruction = 0;
  if (ThisTask == 0) printf("Tree construction.\n");
  sync_params_and_particles();
  n = gather_active();
  b200_check(b200_compute_accelerations(mode, active_list, n, All.Time, vmax), "b200_compute_accelerations");
  b200_check(b200_download(), "b200_download");
  b200_get_counters(&c);
  All.CPU_TreeConstruction += 1e-3 * (c.ms_build + c.ms_predict);
  All.CPU_TreeWalk += 1e-3 * c.ms_walk;
  All.CPU_CommSum += 1e-3 * (c.ms_upload + c.ms_download);
  All.CPU_Gravity += 1e-3 * (c.ms_build + c.ms_predict + c.ms_walk);
  All.TotNumOfForces += ntot;
  NoCostFlag = 0;
  determine_interior();                                 /* accel.c:60 */
  if (mode == 0) {
    All.CPU_EnsureNgb += 1e-3 * (c.ms_sidm + c.ms_ensure);
    print_sct();
    if (c.ensure_iterations > 0) {                      /* sidm.c:943-955: restore the time line */
      save = All.TimeStep;
      find_next_time();
      All.TimeStep = save;
    }
  }
  (void)i;
  if (ThisTask == 0) { printf("force computation done.\n"); fflush(stdout); }
}
#endif
#endif
