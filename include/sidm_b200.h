/* sidm_b200.h - C ABI of libsidm_b200.so: the B200 (sm_100a) implementation of the
 * per-step hot path of junkoda/sidm-nbody ("sidm-gadget"): Barnes-Hut tree gravity
 * (gravity_tree / force_treebuild / force_treeevaluate) and the SIDM scatter step
 * (sidm / ngb_treefind_variable / sidm_ensure_neighbours).
 *
 * The reference has no plugin API: its driver (run.c, accel.c, init.c) calls plain C
 * functions that work on the globals of allvars.h.  The drop-in therefore has two
 * layers:
 *   1. this coarse C ABI (plain pointers and sizes, no torch/CUDA types), one call per
 *      reference entry point; and
 *   2. sidm-nbody_b200/shim/b200_shim.c, which re-implements the reference's own symbols
 *      (gravity_tree(), sidm(), sidm_ensure_neighbours(), force_treebuild(), ...) on top
 *      of it and is linked with the unmodified driver instead of gravtree.c /
 *      forcetree.c / sidm.c.  INTEGRATION.md shows the link line.
 *
 * Conventions (mirroring the reference):
 *   - every function returns 0 on success, else an error code.  Where the reference has
 *     an endrun() code for the same condition the same number is returned (1 tree nodes
 *     exhausted forcetree.c:233-239, 3 allocation failure forcetree.c:1805-1809, 78
 *     neighbour list overflow forcetree.c:2184-2188, 1155 smoothing-length iteration
 *     failed sidm.c:936-939); CUDA failures return B200_ERR_CUDA (+ cudaError in
 *     b200_last_cuda_error()).  The library never calls exit().
 *   - particle indices are 0-based positions in the bound particle array (the
 *     reference's P[i+1]).
 *   - one CUDA device per process; calls are blocking; not thread-safe (the reference is
 *     single-threaded per rank with static state, forcetree.c:54,1991-1994).
 *   - the library owns all device memory; the host particle array stays owned by the
 *     caller and is the source of truth between calls.
 *   - there is NO CPU fallback: without a CUDA device b200_init() fails with
 *     B200_ERR_NODEVICE and nothing else works.
 */
#ifndef SIDM_B200_H
#define SIDM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK             0
#define B200_ERR_NODES      1     /* forcetree.c:233-239  endrun(1)  */
#define B200_ERR_ALLOC      3     /* forcetree.c:1805-1809 endrun(3) */
#define B200_ERR_NGBOVERFLOW 78   /* forcetree.c:2184-2188 endrun(78) */
#define B200_ERR_HSML       1155  /* sidm.c:936-939 endrun(1155)     */
#define B200_ERR_CUDA       9001
#define B200_ERR_NODEVICE   9002
#define B200_ERR_ARG        9003
#define B200_ERR_STATE      9004
#define B200_ERR_COINCIDENT 9005  /* >1 particle at one position to 42 octree levels;
                                     the reference randomises such pairs with rand()
                                     (forcetree.c:320-326), which cannot be reproduced */
#define B200_ERR_TYPES      9006  /* particle types changed behind the library's back (set them through b200_set_field / b200_upload) */

#define B200_ERR_IO         9007  /* snapshot file could not be opened / written (io.c:98-102 endrun(10), io.c:594-605) */

/* The `All` fields the path reads (allvars.h:164-420).  Filled by the shim from `All`. */
typedef struct b200_params {
  int    device;                   /* CUDA device ordinal (rank-local)                 */
  int    MaxPart;                  /* capacity in particles   (All.MaxPart)            */
  double TreeAllocFactor;          /* node capacity = factor*MaxPart (All.TreeAllocFactor) */
  /* gravity */
  double ErrTolTheta;              /* BH opening angle          forcetree.c:967        */
  double ErrTolForceAcc;           /* relative criterion alpha  forcetree.c:1129       */
  int    TypeOfOpeningCriterion;   /* 0 BH, 1 relative          forcetree.c:801        */
  int    ComovingIntegrationOn;    /* gravtree.c:42-49,252-298                         */
  double G;                        /* gravtree.c:318                                   */
  double SofteningTable[6];        /* Plummer-equivalent eps per type forcetree.c:800  */
  double BoxSize;                  /* >0 with PeriodicBoundariesOn: Ewald path         */
  int    PeriodicBoundariesOn;
  double Omega0, OmegaLambda, Hubble;
  /* SIDM */
  int    DesNumNgb;                /* sidm.c:512                                       */
  int    MaxNumNgbDeviation;
  double CrossSectionInternal;     /* sidm.c:278-280,372                               */
  int    CrossSectionType;         /* compile-time CROSS_SECTION_TYPE of the reference;
                                      0 hard sphere, 1 ~1/v, 2 Yukawa-like, 3 power law,
                                      4 Yukawa with angular dependence (sidm.c:391-439)   */
  double YukawaVelocity, CrossSectionPowLaw, CrossSectionVelScale;
  unsigned long long Seed;         /* All.Seed1 + All.Seed2*ThisTask  begrun.c:44      */
  int    BunchSizeSidm;            /* slots per sidm() bunch (allocate.c:64); <=0: one
                                      bunch holds every active particle                */
  int    ReferenceNgbOrder;        /* 1: scan neighbours in the reference's list order
                                      (tree order + next[] chains + swap-remove filter,
                                      forcetree.c:2163-2297); 0: ascending tree order   */
} b200_params;

/* Byte layout of the caller's array-of-structs (struct particle_data, allvars.h:422-460,
 * depends on the reference's -D flags, so offsets are passed, not assumed). */
typedef struct b200_layout {
  int stride;
  int Pos, Vel, Mass, ID, Type, CurrentTime, PosPred, VelPred, Accel, GravCost, OldAcc;
  int Left, Right, NgbVelDisp, HsmlVelDisp, dVel;
  int MaxPredTime;                 /* > 0: offset of P[].MaxPredTime (b200_find_timesteps writes it);
                                      0 = field not bound                                            */
  int Potential;                   /* > 0: offset of P[].Potential (b200_compute_potential writes it) */
} b200_layout;

/* Optional replay of the reference's random numbers: the
 * uniform each buffered particle consumed at sidm.c:341 and the unit vector
 * random_direction() returned for it (sidm_rand.h:24-37), both indexed by buffer slot. */
typedef struct b200_replay {
  const double *rand;              /* [nslot]            */
  const double *dir;               /* [nslot][3], used only where a scatter happens */
  /* CROSS_SECTION_TYPE 4 only (may be NULL): the (rand, cosO-uniform) pairs the reference drew at
   * sidm.c:393-394 for slot s are extra[2*extra_off[s] .. 2*extra_off[s+1]); honoured by b200_sidm(),
   * not by the repair passes */
  const double *extra;
  const int    *extra_off;         /* [nslot+1]          */
} b200_replay;

typedef struct b200_scatlog {      /* struct scatlog, sidm.h:1-10 */
  float time; int id1, id2; float Hsml1, Hsml2;
  float x1[3], x2[3], v1[3], v2[3], dv[3];
} b200_scatlog;

typedef struct b200_counters {
  /* tree */
  int       num_nodes;             /* internal nodes (numnodestree, forcetree.c:60)   */
  int       max_level;
  /* last gravity call: the reference's -DDIAG quantities (forcetree.c:65-66) */
  long long part_interactions;     /* treecost        */
  long long node_interactions;     /* treecost_quadru */
  long long list_nodes;            /* node records a 32-target warp streamed (I_n sum) */
  long long list_parts;            /* particle records a warp streamed       (I_p sum) */
  long long num_targets;
  long long num_lists;             /* interaction lists walked (one per warp of 32 targets)         */
  /* since the last b200_sidm(), repair passes included: the step's SCT lines summed (sidm.c:614-620) */
  int       sct_ntot, sct_pass1, sct_scattered, sct_rejected;
  long long ngb_candidates;        /* cube candidates examined (C in SURVEY 8d)       */
  int       ensure_iterations;     /* passes of the last sidm_ensure_neighbours       */
  int       ensure_repaired;       /* particles re-done, summed over those passes     */
  /* device milliseconds of the last calls (CUDA events) */
  float     ms_upload, ms_predict, ms_build, ms_walk, ms_sidm, ms_ensure, ms_download;
  long long kernel_launches;       /* cumulative launches of this library's kernels   */
} b200_counters;

/* ---- life cycle ------------------------------------------------------------------ */
int  b200_init(const b200_params *p);          /* allocate.c / force_treeallocate():1797 */
int  b200_set_params(const b200_params *p);    /* re-read All (time-dependent softenings) */
void b200_finalize(void);
int  b200_last_cuda_error(void);
/* run on the caller's CUDA stream (a cudaStream_t); default is the legacy default stream */
int  b200_set_stream(void *cuda_stream);
const char *b200_version(void);
/* run-time switches of the current context: legal between b200_init and b200_finalize (B200_ERR_STATE otherwise), and every
 * one returns to its default at the next b200_init.  "overlap": how b200_compute_accelerations(0) uses its two CUDA streams: 0 = the phases one after the
 * other as accel.c:39-65 does; 1 = gravity walk and the whole SIDM chain (pass + repair loop) next to each other; 2 = walk and
 * SIDM pass next to each other, the repair loop's many small launches after the walk.  Default (-1): 1; sharded over several
 * GPUs the exchange of the gravity results is then issued behind the running walk and travels while the repair loop runs.
 * "shard_overlap" (default 0): see
 * b200_set_shard.  "group_search" (default 1): warp-shared neighbour search for all-active passes.  "shard_min_work"
 * (default 262144): work lists shorter than this are done completely by every rank instead of being sharded
 * (the repair passes and small active sets are latency-bound; an exchange per pass costs more than it saves).
 * "compact_exchange" (default 1): SIDM results travel between ranks as {uint16 count per slot + the few
 * scatter proposals} instead of 32-byte records.  "cand_cap" (default 1024): candidates kept per slot in the
 * ReferenceNgbOrder parity mode (the search cube of a particle in the outskirts can clip the dense centre);
 * more than that returns B200_ERR_NGBOVERFLOW like the reference's endrun(78).  "walk_pairs" (default 0): 1 selects the packed
 * sibling-pair form of the gravity walk (f32x2 arithmetic over two child cells at a time, same interaction lists; open boundaries,
 * one particle type), "walkp_minb" (8 / 6 / 4) its occupancy variant.  "tree_reuse" (default 0 = a full build at every
 * b200_tree_build / b200_compute_accelerations, i.e. the reference run with TreeUpdateFrequency 0): k > 1 builds the tree at
 * every k-th request and REFITS it in between - same topology, leaves and all multipole moments from the current predicted
 * positions (the counterpart of the reference's dynamic tree updates, gravtree.c:63-96, forcetree.c:935-954,2486-2549, which
 * drift the cells instead).  Neighbour searches on a refitted tree are exact (cell tests widened by the largest displacement
 * since the build); forces come from a different, equally valid tree: as close to direct summation as the fresh tree's, about
 * 1e-3 relative rms away from them (tests/test_gpu_reuse.py) - outside the 1e-4 the default path keeps to the reference's tree,
 * hence opt-in.  A refit needs an unchanged particle set: full uploads, new types or
 * b200_set_soa positions force a build; one particle type, open boundaries, tree-order neighbour scans. */
int  b200_set_option(const char *name, int value);
/* Generator state for restarts ("next" row f3; the reference's restart files, restart.c:37-154, do not save its
 * MT19937 state, so a restarted reference run draws different scatterings).  Here every random number is a function
 * of (Seed, call counter, particle index): state[0] = sidm() calls so far, state[1] = find_timesteps() calls so far.
 * Saving the two words with the particle data and setting them after b200_init makes a restarted run bit-identical
 * to the uninterrupted one. */
int  b200_get_rng_state(unsigned long long *state);
int  b200_set_rng_state(const unsigned long long *state);

/* ---- particle state -------------------------------------------------------------- */
/* Bind the host AoS (the reference's &P[1]).  pin!=0 page-locks it for async DMA. */
int  b200_bind_particles(void *base, int num_part, const b200_layout *layout, int pin);
int  b200_upload(void);                        /* host AoS -> device (all fields)        */
int  b200_download(void);                      /* device -> host AoS (fields the path writes:
                                                  PosPred VelPred Accel GravCost OldAcc Left
                                                  Right NgbVelDisp HsmlVelDisp dVel)      */
int  b200_download_to(void *dst);              /* same, into another array of the bound layout */
/* Partial transfers for small active sets (the reference moves 20 bytes in / 24 bytes out per ACTIVE particle,
 * gravtree.c:149-166,230-238; allvars.h:547-559).  b200_upload_active: the fields the host driver changes between two
 * force computations - advance() (predict.c:245-345: Pos, Vel, VelPred = Vel, dVel = 0, CurrentTime), reflect() (Vel),
 * find_timesteps() (MaxPredTime) - of the listed particles (normally the previous step's active list), read from the bound
 * array, 36 bytes each in one H2D copy.  b200_download_active: the fields the path writes (PosPred VelPred Accel OldAcc
 * GravCost dVel NgbVelDisp HsmlVelDisp Left Right, 76 bytes each, one D2H copy) of the listed particles AND of every
 * particle that received a partner kick since the last download (the partner of a scattering need not be active,
 * sidm.c:559-601), scattered into `dst` (NULL: the bound array).  A host that changes anything else uses b200_upload(). */
int  b200_upload_active(const int *idx, int n);
int  b200_download_active(const int *idx, int n, void *dst);
/* Multi-GPU variants (after b200_set_shard): each rank owns host rows [first, first+count) of the
 * global particle order (the reference's per-rank P[] after its domain decomposition).  Upload
 * copies only those rows over PCIe and replicates them to every rank with one all-gather over
 * NVLink; download returns only the rank's own rows.  `rows_per_rank` >= count is the common slice
 * length of the all-gather (ceil(N/world)); the shard buffers must hold rows_per_rank*stride bytes. */
int  b200_upload_shard(int first, int count, int rows_per_rank);
int  b200_download_shard(void *dst, int first, int count);
/* The same for a host array that holds ONLY the rank's own rows with any number of rows per rank (the reference's per-task P[]
 * after DomainDecomposition(), domain.c:31-884): b200_bind_rows binds rows [first, first+count) of a global order of n_global
 * particles = the tasks' arrays one after the other; b200_upload_rows(counts[world]) sends them up and replicates them;
 * b200_download_shard(NULL, first, count) brings the own rows back into the bound array. */
int  b200_bind_rows(void *base, int first, int count, int n_global, const b200_layout *layout, int pin);
int  b200_upload_rows(const int *counts);
/* Structure-of-arrays alternative used by the tests / bench (host pointers, float32/int32;
 * any pointer may be NULL = keep current).  pos/vel are [n][3]. */
int  b200_set_soa(int num_part, const float *pos, const float *vel, const float *mass,
                  const int *id, const float *curtime, const float *accel, const float *oldacc,
                  const float *hsml, const float *dvel);
int  b200_get_soa(float *pospred, float *velpred, float *accel, float *oldacc, float *gravcost,
                  float *hsml, int *ngb, float *dvel, float *left, float *right);

/* ---- sharding over the GPUs of one box (replaces domain.c + the hypercube
 * exchanges of gravtree.c:171-222 and sidm.c:204-553) -------------------------------------
 * Every rank holds all particles and builds the same tree; rank r evaluates the 32-entry
 * blocks b (b % world == r) of each work list sorted along the tree order, then the ranks
 * all-gather the per-target results.  The library packs into `send`, calls
 * fn(bytes_per_rank, user) - which must all-gather send[0..bytes) of every rank into
 * recv[rank*bytes ..] ordered on the stream b200_current_stream() returns at that moment
 * (ncclAllGather(..., stream) / torch.distributed under that stream) - and unpacks `recv`.
 * A host whose callback honours b200_current_stream() sets the option "shard_overlap" to 1: the
 * SIDM chain then runs on its own stream next to the gravity walk also when sharded (its
 * collectives are issued first, the gravity exchange last).  send/recv are device pointers owned by the caller; cap_bytes = size of `send`
 * (recv holds world*cap_bytes).  world == 1 switches sharding off. */
typedef int (*b200_allgather_fn)(long long bytes_per_rank, void *user);
int  b200_set_shard(int rank, int world, void *send, void *recv, long long cap_bytes,
                    b200_allgather_fn fn, void *user);
void *b200_current_stream(void);               /* cudaStream_t of the collective being requested */
/* for a C host without CUDA code of its own (the shim): send == recv == NULL in b200_set_shard lets the library allocate the
 * exchange buffers (cap_bytes and world * cap_bytes); b200_shard_buffers returns them for the callback's ncclAllGather */
int  b200_shard_buffers(void **send, void **recv, long long *cap_bytes);
int  b200_device_count(void);                  /* CUDA devices visible to this process (0: none) */

/* ---- the hot path ---------------------------------------------------------------- */
/* predict_collisionless_only(time), predict.c:106: PosPred, VelPred for all particles. */
int  b200_predict(double time);
/* force_treebuild(), forcetree.c:90: octree over PosPred + multipole moments.        */
int  b200_tree_build(void);
/* gravity_tree(), gravtree.c:127-324 for the given active list (NULL = all, in index
 * order): walk, Accel (pre-G) -> OldAcc, Accel*G (+ Lambda / comoving terms). */
int  b200_gravity(const int *active, int nactive, double time);
/* sidm(), sidm.c:57-627, for the active list, in list order. */
int  b200_sidm(const int *active, int nactive, double time, double vmax,
               const b200_replay *replay);
/* setup_nbr_sidm(), sidm.c:630-805: neighbour counts only. */
int  b200_setup_nbr_sidm(const int *active, int nactive);
/* sidm_ensure_neighbours(mode), sidm.c:814-968 (repair loop; calls the sidm pass on the
 * repaired set; replay arrays are indexed by [pass][slot] flattened, may be NULL). */
int  b200_sidm_ensure_neighbours(int mode, double time, double vmax, const b200_replay *replay);
/* setup_smoothinglengths_sidm(desngb), init.c:431-512. */
int  b200_setup_smoothinglengths_sidm(int desired_ngb);
/* compute_accelerations(mode), accel.c:27-132: predict + build + gravity
 * (+ sidm + ensure_neighbours when mode==0). */
int  b200_compute_accelerations(int mode, const int *active, int nactive, double time, double vmax);
/* advance(), predict.c:245-345 ("next" row f1 of SURVEY.md section 8): kick-drift of the active
 * particles with Accel*dt + dVel, dVel cleared; *num_scattered (may be NULL) = n_scat_particles. */
int  b200_advance(const int *active, int nactive, double time, int *num_scattered);
/* reflect(), reflection.c:7-33 (-DREFLECTIONBOUNDARY, run.c:106; row f1): specular reflection of the active
 * particles that are outside `radius` and moving outwards.  Acts on Pos / Vel as advance() left them. */
int  b200_reflect(const int *active, int nactive, double radius, int *num_reflected);
/* find_timesteps(mode), timestep.c:17-334 for collisionless particles ("next" row f1): new time step of
 * every active particle from the acceleration criterion (TypeOfTimestepCriterion 0 or 1), the SIDM
 * probability limit ProbabilityTol/(C_max m h^-3) and the G*rho limit (timestep.c:247-265), the 1.3*dtold
 * growth limit (not for mode 2) and the Max/MinSizeTimestep clamps; P[i].MaxPredTime = CurrentTime + dt/2.
 * The time-line tree itself (timeline.c delete_node / insert_node / construct_timetree) stays with the
 * host driver, which reads MaxPredTime back (b200_download, or maxpred_out per active entry).
 * The reference jitters clamped steps with drand48() (timestep.c:283,309); here the uniform comes from
 * `jitter` (one per active entry, in list order) or, if NULL, from the particle's counter-based stream.
 * *num_clamped (may be NULL) = number of steps that hit a clamp. */
typedef struct b200_timestep_params {
  int    TypeOfTimestepCriterion;          /* 0: sqrt(2 eta eps / |a|), 1: ErrTolVelScale / |a|   */
  double ErrTolIntAccuracy, ErrTolVelScale, ProbabilityTol, ErrTolDynamicalAccuracy;
  double MaxSizeTimestep, MinSizeTimestep;
} b200_timestep_params;
int  b200_find_timesteps(const int *active, int nactive, int mode, double time, double vmax,
                         const b200_timestep_params *tp, const double *jitter, float *maxpred_out, int *num_clamped);
/* compute_potential(), potential.c:18-180 ("next" row f2): rebuilds the tree over the predicted positions,
 * walks it for ALL particles with force_treeevaluate_potential() (forcetree.c:1389-1755, same opening
 * decisions as the force walk), adds the self energy back and applies G and the Lambda / comoving terms.
 * P[].Potential on the device (downloaded when the layout binds it) and, if not NULL, pot_out[n].
 * Periodic boxes add ewald_pot_corr() (ewald.c:246-285) per interaction. */
int  b200_compute_potential(float *pot_out);
/* compute_global_quantities_of_system(), global.c:18-135 ("next" row f2): mass, kinetic and potential energy,
 * momentum, angular momentum and centre of mass per particle type and in total, from PosPred / VelPred / Mass /
 * Potential / Type on the device (P[].Potential as the last b200_compute_potential left it).  The struct has the
 * layout of the reference's `struct state_of_system` (allvars.h:517-537) so that a drop-in can copy it into
 * SysState.  Products are formed in float like the reference's expressions, sums in double (tree order instead
 * of particle order: equal to rounding of the double sums).  EnergyInt is 0 (no gas on this path). */
typedef struct b200_sysstate {
  double Mass, EnergyKin, EnergyPot, EnergyInt, EnergyTot, Momentum[4], AngMomentum[4], CenterOfMass[4];
  double MassComp[5], EnergyKinComp[5], EnergyPotComp[5], EnergyIntComp[5], EnergyTotComp[5];
  double MomentumComp[5][4], AngMomentumComp[5][4], CenterOfMassComp[5][4];
} b200_sysstate;
int  b200_compute_global_quantities(b200_sysstate *out);
/* savepositions(), io.c:16-590 ("next" row f3): one snapshot file in GADGET format 1 (NumFilesPerSnapshot = 1) written
 * from the device state - header (struct io_header_1, allvars.h:727-746), then PosPred, VelPred, ID and the masses of
 * the types whose mass_table entry is 0, particles in TYPE order and particle order inside a type, every block
 * between 4-byte length markers; positions wrapped into [0, BoxSize] in periodic runs (io.c:275-283).  The blocks
 * are gathered on the device and streamed to the file through pinned buffers; the host AoS is not touched.
 * `time` = All.Time, `mass_table` = All.MassTable[6] (NULL: all 0), `hubble_param` = All.HubbleParam; BoxSize / Omega0 /
 * OmegaLambda / ComovingIntegrationOn come from b200_params.  npart_out[6] (may be NULL) = header1.npart.  Byte-identical
 * to the file the reference writes from the same state.  Gas particles (type 0: u / rho / hsml blocks) are not on
 * this path: B200_ERR_ARG.  The 84 fill bytes of the header are zero (the reference leaves what read_ic() put there). */
int  b200_savepositions(const char *path, double time, const double *mass_table, double hubble_param, int *npart_out);
/* one file of a snapshot split over NumFilesPerSnapshot > 1 files (io.c:90-103,127-160, file <base>_<num>.<i>): the rows
 * [first, first+count) of the particle order (the particles of the tasks of one file group) in type order; header npart = the
 * file's counts, npartTotal = the system's, num_files as given */
int  b200_savepositions_part(const char *path, double time, const double *mass_table, double hubble_param,
                             int first, int count, int num_files, int *npart_out);
/* read_ic() + the start-up loop of init() (read_ic.c:32-481, init.c:76-100) for one format-1 file without gas, straight
 * into the device state: types from the header's block ranges, masses from MassTable or the mass block, PosPred = Pos,
 * VelPred = Vel, CurrentTime = header time, Accel = dVel = OldAcc = Potential = 0, GravCost = 1, Hsml = 0.  The file is
 * streamed through pinned buffers; the particle count becomes the file's (<= MaxPart).  Block markers are checked
 * (B200_ERR_IO).  time_out, mass_table_out[6], npart_out[6] may be NULL.  A snapshot split over several files (NumFilesPerSnapshot
 * > 1, read_ic.c:62-75) is read when `path` names no file but path.0 does: path.0 .. path.<num_files-1>, their particles one
 * after the other. */
int  b200_load_snapshot(const char *path, double *time_out, double *mass_table_out, int *npart_out);
/* raw double potentials of the given targets as forcetree.c:1389 leaves them in GravDataPotential */
int  b200_potential_raw(const int *targets, int n, double *pot_out);
/* host -> device copy into a named internal buffer (see b200_device_buffer), e.g. "maxpred" */
int  b200_set_field(const char *name, const void *host, long long nbytes);
/* getvmax(), sidm.c:970-990. */
int  b200_getvmax(double *vmax);
/* ngb_treefind(xyz, desngb, 0, type), forcetree.c:2311: exact k-th neighbour distance^2
 * for the given particle indices. */
int  b200_ngb_treefind(const int *idx, int n, int desngb, float *h2_out);

/* ---- parity / debug accessors (used by tests through the same ABI) ------------------ */
/* force_treeevaluate_direct(), forcetree.c:1896: direct summation for the targets. */
int  b200_direct(const int *targets, int n, double *acc_out /*[n][3]*/);
/* tree walk without the G / OldAcc epilogue: raw double accelerations + per-target
 * (particle, node) interaction counts, as gravtree.c:189-190 leaves GravDataResult. */
int  b200_walk_raw(const int *targets, int n, double *acc_out /*[n][3]*/, int *cost_out /*[n][2]*/);
/* Tree dump in this library's node order (depth-first pre-order).  Any pointer may be
 * NULL.  center/len define the geometry the reference builds at forcetree.c:241-345. */
int  b200_get_tree(int *num_nodes, float *center /*[m][3]*/, float *len, float *mass,
                   float *s /*[m][3]*/, float *Q /*[m][7]: Q11 Q22 Q33 Q12 Q13 Q23 P*/,
                   float *oc, float *bmax2, int *count, int *level);
/* ngb_treefind_variable() lists (forcetree.c:2163) for the given particles with their
 * current HsmlVelDisp, in the scan order selected by ReferenceNgbOrder. */
int  b200_ngb_lists(const int *idx, int n, int cap, int *count_out, int *list_out /*[n][cap]*/);
/* per-slot cumulative probabilities of the last b200_sidm() call: P_max, total Prob over
 * the neighbour list, chosen partner (-1 none) -- sidm.c:338-383. */
int  b200_sidm_debug(int nslot, int *slot_particle, double *pmax, double *prob_total, int *partner);
int  b200_get_scatlog(b200_scatlog *out, int cap, int *n);
int  b200_get_counters(b200_counters *c);
/* device pointer + element count of an internal buffer, for NCCL plumbing from the host
 * language (names: "posm", "velh", "accel", "dvel", ...).  Returns B200_ERR_ARG if unknown. */
int  b200_device_buffer(const char *name, void **dptr, long long *nbytes);

#ifdef __cplusplus
}
#endif
#endif
