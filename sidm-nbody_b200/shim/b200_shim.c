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
 *                                               one download - the fast form of the drop-in.  Between steps only
 *                                               the particles the driver advanced go up and only the active
 *                                               particles (+ kicked partners) come down (b200_upload_active /
 *                                               b200_download_active; B200_SHIM_FULL_COPY=1 in the environment
 *                                               restores whole-array copies)
 *   setup_smoothinglengths_sidm()  init.c:431   only with -DB200_SHIM_ACCEL (init.o's own definition weakened with
 *                                               objcopy): one batched library call instead of
 *                                               NumPart ngb_treefind() round trips
 *
 * Several tasks (NTask > 1, one GPU per task; -DB200_SHIM_ACCEL form): every task keeps the particles the reference's own
 * domain decomposition gave it (domain.c is linked unchanged and keeps moving whole particles between tasks); the
 * shim sends each task's rows up and replicates them on all GPUs (b200_bind_rows / b200_upload_rows), hands the library
 * the global active list (the tasks' lists one after the other), and brings each task's own rows back.  The library
 * deals the work out over the GPUs and all-gathers the results through b200_comm.c (NCCL over NVLink, or MPI through
 * the host).  Task-dependent quantities of the reference become global ones, because every GPU must compute the same
 * thing: vmax is the maximum over ALL particles (sidm.c:970-990 takes the local one), the generator seed is task 0's.
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
#ifdef SCATTERLOG
#include "sidm.h"                 /* struct scatlog */
typedef char b200_scatlog_has_the_layout_of_struct_scatlog[sizeof(b200_scatlog) == sizeof(struct scatlog) ? 1 : -1];
#endif

static int shim_ready = 0;
static int *active_list = 0;
static int active_cap = 0;

/* several tasks: global particle order = the tasks' P[] one after the other */
#define B200_MAX_TASKS 64
static int rows_of_task[B200_MAX_TASKS];
static int first_row = 0, n_global = 0;
int b200_comm_init(int rank, int world, int use_nccl);
int b200_comm_allgather(long long bytes, void *user);
int b200_comm_uses_nccl(void);

static void rows_refresh(void)
{
  int q;
  MPI_Allgather(&NumPart, 1, MPI_INT, rows_of_task, 1, MPI_INT, MPI_COMM_WORLD);
  for (q = 0, first_row = 0, n_global = 0; q < NTask; q++) { if (q < ThisTask) first_row += rows_of_task[q]; n_global += rows_of_task[q]; }
}

/* partial transfers of the fast path: the device mirrors P[] except for what advance() / reflect() / find_timesteps()
 * did to the particles of the previous force computation */
static int  dev_mirrors_host = 0;
static int *prev_active = 0;
static int  prev_n = 0;

void force_treeallocate(int maxnodes, int maxpart);
static void rng_state_io(int write);

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
  {
    const int ndev = b200_device_count();
    p->device = ndev > 0 ? ThisTask % ndev : 0;         /* one task per GPU of the box (tasks share GPUs if there are fewer) */
  }
  p->MaxPart = NTask > 1 ? (int)(All.TotNumPart + 64) : All.MaxPart;   /* several tasks: every GPU holds all particles */
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
  p->Seed = (unsigned long long)(All.Seed1 + All.Seed2 * (NTask > 1 ? 0 : ThisTask));   /* begrun.c:44; replicated work needs one seed */
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

/* several tasks: the global active list every GPU works on = the tasks' lists one after the other, in global row numbers */
static int gather_active_global(int **list)
{
  static int *glist = 0; static int gcap = 0;
  int counts[B200_MAX_TASKS], q, tot = 0, at = 0, i;
  int n = gather_active();
  if (NTask == 1) { *list = active_list; return n; }
  MPI_Allgather(&n, 1, MPI_INT, counts, 1, MPI_INT, MPI_COMM_WORLD);
  for (q = 0; q < NTask; q++) tot += counts[q];
  if (gcap < tot + 1) { free(glist); gcap = (int)All.TotNumPart + 64; glist = (int *)malloc(sizeof(int) * gcap); if (!glist) endrun(3); }
  for (q = 0; q < NTask; q++) {
    if (q == ThisTask) for (i = 0; i < n; i++) glist[at + i] = first_row + active_list[i];
    if (counts[q] > 0) MPI_Bcast(glist + at, counts[q], MPI_INT, q, MPI_COMM_WORLD);
    at += counts[q];
  }
  *list = glist;
  return tot;
}

static void sync_params_and_particles(void)
{
  b200_params p;
  b200_layout l;
  prev_n = 0;                                   /* whole array goes up: nothing is pending from the previous step */
  if (!shim_ready) force_treeallocate(0, 0);    /* init.c:120 asks for the statistics before it allocates the tree (init.c:123) */
  fill_params(&p);
  b200_check(b200_set_params(&p), "b200_set_params");
  fill_layout(&l);
  if (NTask == 1) {
    b200_check(b200_bind_particles(&P[1], NumPart, &l, 1), "b200_bind_particles");
    b200_check(b200_upload(), "b200_upload");
  } else {
    rows_refresh();
    b200_check(b200_bind_rows(&P[1], first_row, NumPart, n_global, &l, 1), "b200_bind_rows");
    b200_check(b200_upload_rows(rows_of_task), "b200_upload_rows");      /* own rows over PCIe, replicated over NVLink */
  }
}

static void download_particles(void)
{
  if (NTask == 1) b200_check(b200_download(), "b200_download");
  else b200_check(b200_download_shard(0, first_row, NumPart), "b200_download_shard");
}

/* ---------------------------------------------------------------- forcetree.h surface */

void force_treeallocate(int maxnodes, int maxpart)      /* forcetree.c:1797 */
{
  b200_params p;
  (void)maxnodes; (void)maxpart;
  if (shim_ready) return;
  fill_params(&p);
  b200_check(b200_init(&p), "b200_init");
  if (NTask > 1) {
    /* exchange buffers: one task's rows of the particle array, or its share of the 32-byte per-slot records */
    const long long rows = All.MaxPart, slots = ((All.TotNumPart + 31) / 32 + NTask - 1) / NTask * 32;
    long long cap = rows * (long long)sizeof(struct particle_data);
    if (slots * 32 > cap) cap = slots * 32;
    if (NTask > B200_MAX_TASKS) endrun(9003);
    b200_check(b200_set_shard(ThisTask, NTask, 0, 0, cap, b200_comm_allgather, 0), "b200_set_shard");
    if (b200_comm_init(ThisTask, NTask, b200_device_count() >= NTask) != 0) { printf("task %d: communicator set-up failed\n", ThisTask); endrun(9001); }
    b200_check(b200_set_option("shard_overlap", 1), "b200_set_option");     /* b200_comm_allgather honours b200_current_stream() */
    if (getenv("B200_SHARD_MIN_WORK")) b200_check(b200_set_option("shard_min_work", atoi(getenv("B200_SHARD_MIN_WORK"))), "b200_set_option");
    if (ThisTask == 0) printf("libsidm_b200 on %d tasks, %d GPU(s), all-gather through %s\n", NTask, b200_device_count(), b200_comm_uses_nccl() ? "NCCL" : "MPI (host staged)");
  }
  shim_ready = 1;
  if (RestartFlag == 1) rng_state_io(0);
}
void force_treefree(void) { b200_finalize(); shim_ready = 0; }

/* The generator state of the path (two call counters, b200_get_rng_state) next to the reference's restart files, which do not
 * hold its MT19937 state (restart.c:37-154): <OutputDir><RestartFile>.<task>.b200rng is rewritten after every force
 * computation and read back when the run is started with RestartFlag = 1 (begrun.c:57), so that a restarted run draws the
 * same random numbers as the uninterrupted one. */
static void rng_state_io(int write)
{
#ifdef SIDM
  char name[400];
  unsigned long long st[2];
  FILE *f;
  sprintf(name, "%s%s.%d.b200rng", All.OutputDir, All.RestartFile, ThisTask);
  if (write) {
    if (b200_get_rng_state(st) != B200_OK) return;
    if ((f = fopen(name, "wb"))) { fwrite(st, sizeof(st), 1, f); fclose(f); }
  } else if ((f = fopen(name, "rb"))) {
    if (fread(st, sizeof(st), 1, f) == 1) b200_check(b200_set_rng_state(st), "b200_set_rng_state");
    fclose(f);
  }
#endif
}

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
  int n, ntot, *glist;
  double t0 = second(), t1;
  if (All.ComovingIntegrationOn) set_softenings();
  MPI_Allreduce(&NumForceUpdate, &ntot, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  All.NumForcesSinceLastDomainDecomp += ntot;
  All.NumForcesSinceLastTreeConstruction = 0;
  if (ThisTask == 0) printf("Tree construction.\n");
  sync_params_and_particles();
  b200_check(b200_predict(All.Time), "b200_predict");               /* gravtree.c:72 */
  b200_check(b200_tree_build(), "b200_tree_build");                 /* gravtree.c:76 */
  n = gather_active_global(&glist);
  b200_check(b200_gravity(glist, n, All.Time), "b200_gravity");     /* gravtree.c:127-324 */
  download_particles();
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
double getvmax(void)                                    /* sidm.c:970-990; called before force_treeallocate() (init.c:74): host loop */
{
  int j, q;
  double v2, v = 0.0, all[B200_MAX_TASKS];
  for (j = 1; j <= NumPart; j++) {
    v2 = P[j].Vel[0] * P[j].Vel[0] + P[j].Vel[1] * P[j].Vel[1] + P[j].Vel[2] * P[j].Vel[2];
    if (v < v2) v = v2;
  }
  v = sqrt(v);
  if (NTask > 1) {                                      /* every GPU works on all particles: the global maximum */
    MPI_Allgather(&v, 1, MPI_DOUBLE, all, 1, MPI_DOUBLE, MPI_COMM_WORLD);
    for (q = 0; q < NTask; q++) if (all[q] > v) v = all[q];
  }
#ifdef FINDNBRLOG
  if (ThisTask == 0) fprintf(stdout, "Vmax= %g Processor %d\n", v, ThisTask);
#endif
  return v;
}

/* -DSCATTERLOG (sidm.c:96-104, 571-601, 622-624): the records of this call appended to sct_<snapshot count>.<task> */
static void write_scatterlog(void)
{
#ifdef SCATTERLOG
  static b200_scatlog *buf = 0;
  static int cap = 0;
  char filename[128];
  FILE *fp;
  int n = 0;
  if (!buf) { cap = 1 << 16; buf = (b200_scatlog *)malloc(sizeof(b200_scatlog) * cap); if (!buf) endrun(3); }
  b200_check(b200_get_scatlog(buf, cap, &n), "b200_get_scatlog");
  if (n == cap) {                                        /* more than the buffer holds: take all the library keeps (2^20) */
    free(buf); cap = 1 << 20; buf = (b200_scatlog *)malloc(sizeof(b200_scatlog) * cap); if (!buf) endrun(3);
    b200_check(b200_get_scatlog(buf, cap, &n), "b200_get_scatlog");
  }
  sprintf(filename, "sct_%03d.%d", All.SnapshotFileCount, ThisTask);
  fp = fopen(filename, "ab");
  if (!fp) return;
  if (n > 0) fwrite(buf, sizeof(struct scatlog), n, fp);   /* b200_scatlog has the layout of struct scatlog (sidm.h:1-10) */
  fclose(fp);
#endif
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
  int *glist, n = gather_active_global(&glist);
  b200_check(b200_sidm(glist, n, All.Time, vmax, 0), "b200_sidm");
  write_scatterlog();
  print_sct();
}

void setup_nbr_sidm(void)                               /* sidm.c:630-805 */
{
  int *glist, n;
  sync_params_and_particles();
  n = gather_active_global(&glist);
  b200_check(b200_tree_build(), "b200_tree_build");
  b200_check(b200_setup_nbr_sidm(glist, n), "b200_setup_nbr_sidm");
  download_particles();
}

void sidm_ensure_neighbours(int mode)                   /* sidm.c:814-968 */
{
  b200_counters c;
  double save;
  int i;
  b200_check(b200_sidm_ensure_neighbours(mode, All.Time, vmax, 0), "b200_sidm_ensure_neighbours");
  download_particles();
  b200_get_counters(&c);
  print_sct();
  dev_mirrors_host = 0;
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
  download_particles();
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
  if (All.TotN_gas > 0) { printf("savepositions: gas blocks are not on the GPU path\n"); endrun(9003); }
  sync_params_and_particles();
  if (All.NumFilesPerSnapshot <= 1) {
    /* several tasks: every GPU holds all particles, task 0 writes the one file (the reference sends the blocks to task 0, io.c:390-470) */
    sprintf(buf, "%s%s_%03d", All.OutputDir, All.SnapshotFileBase, num);       /* io.c:96 */
    if (ThisTask == 0) b200_check(b200_savepositions(buf, All.Time, All.MassTable, All.HubbleParam, 0), "b200_savepositions");
  } else {
    /* io.c:78-103: file i holds the particles of the tasks [i*nprocgroup, (i+1)*nprocgroup); its first task writes it - here from
     * its own GPU's copy of those rows, all files at the same time */
    const int nprocgroup = NTask / All.NumFilesPerSnapshot;
    if (nprocgroup < 1 || (NTask % nprocgroup)) { printf("Fatal error.\nNumber of processors must be a multiple of All.NumFilesPerSnapshot.\n"); endrun(213); }
    if (ThisTask % nprocgroup == 0) {
      int q, rows = 0;
      for (q = ThisTask; q < ThisTask + nprocgroup; q++) rows += rows_of_task[q];
      sprintf(buf, "%s%s_%03d.%d", All.OutputDir, All.SnapshotFileBase, num, ThisTask / nprocgroup);      /* io.c:94 */
      b200_check(b200_savepositions_part(buf, All.Time, All.MassTable, All.HubbleParam, first_row, rows, All.NumFilesPerSnapshot, 0), "b200_savepositions_part");
    }
  }
  MPI_Barrier(MPI_COMM_WORLD);
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
  int n, ntot, i, partial, *glist;
  double save;
  if (ThisTask == 0) { printf("Start force computation...\n"); fflush(stdout); }
  if (All.TotN_gas > 0) { printf("compute_accelerations: gas particles are not on the GPU path\n"); endrun(9006); }
  if (All.ComovingIntegrationOn) set_softenings();
  MPI_Allreduce(&NumForceUpdate, &ntot, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  All.NumForcesSinceLastDomainDecomp += ntot;
  All.NumForcesSinceLastTreeConstruction = 0;
  if (ThisTask == 0) printf("Tree construction.\n");
  if (NTask > 1) rows_refresh();
  n = gather_active_global(&glist);
  /* small active sets: move only what changed.  Up: the particles of the previous force computation (advance(), reflect(),
   * find_timesteps() touched nothing else); down: this call's active particles and the partners they kicked. */
  partial = NTask == 1 && dev_mirrors_host && !getenv("B200_SHIM_FULL_COPY") && (long long)4 * (prev_n + n) < NumPart;
  if (partial) {
    b200_params p;
    fill_params(&p);
    b200_check(b200_set_params(&p), "b200_set_params");
    b200_check(b200_upload_active(prev_active, prev_n), "b200_upload_active");
  } else sync_params_and_particles();
  b200_check(b200_compute_accelerations(mode, glist, n, All.Time, vmax), "b200_compute_accelerations");
  if (partial) b200_check(b200_download_active(glist, n, 0), "b200_download_active");
  else download_particles();
  if (NTask == 1) {
    if (!prev_active) { prev_active = (int *)malloc(sizeof(int) * (All.MaxPart + 1)); if (!prev_active) endrun(3); }
    memcpy(prev_active, glist, sizeof(int) * n);
    prev_n = n;
  }
  b200_get_counters(&c);
  if (mode == 0) write_scatterlog();
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
  if (mode == 0) rng_state_io(1);
  dev_mirrors_host = 1;
  if (ThisTask == 0) { printf("force computation done.\n"); fflush(stdout); }
}

/* setup_smoothinglengths_sidm(), init.c:431-512: k-th neighbour distance of every particle, then the count / bisection
 * iteration - one batched library call (the per-particle ngb_treefind() above stays as the fall-back of the
 * symbol-by-symbol form) */
void setup_smoothinglengths_sidm(int desired_ngb)
{
  sync_params_and_particles();
  b200_check(b200_tree_build(), "b200_tree_build");
  b200_check(b200_setup_smoothinglengths_sidm(desired_ngb), "b200_setup_smoothinglengths_sidm");
  download_particles();
}
#endif
#endif
