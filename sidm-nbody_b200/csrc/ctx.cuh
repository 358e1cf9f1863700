// ctx.cuh - library-private state of libsidm_b200.so (one CUDA device per process).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sidm_b200.h"
#include "tree_logic.h"

namespace b200 {

// work lists shorter than Ctx::shard_min_work are not sharded (every rank does all of it); option "shard_min_work"
constexpr int kShardMinWorkDefault = 1 << 18;

struct Ctx {
  bool ready = false;
  b200_params par{};
  int n = 0;                 // particles currently held
  int maxpart = 0, maxnodes = 0;
  int last_cuda = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // second, high-priority stream: inside b200_compute_accelerations() the SIDM chain (search,
  // scatter, repair loop - many small latency-bound launches with host round trips) runs here
  // while the gravity walk fills the machine on `stream` (both only read the tree)
  cudaStream_t stream_sidm = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
  int opt_overlap = -1;            // b200_set_option("overlap", 0|1|2); -1 = default = 1 (sidm.cu, b200_compute_accelerations)
  bool shard_busy = false;         // the exchange buffers hold a collective that is still queued behind the walk: SIDM passes issued now run replicated
  bool opt_shard_overlap = false;  // b200_set_option("shard_overlap", 1): the host's all-gather callback runs on
                                   // b200_current_stream(), so the two-stream overlap is also safe when sharded
  int shard_min_work = kShardMinWorkDefault;
  bool opt_compact_exchange = true;   // option "compact_exchange": 2.5-byte instead of 32-byte slot records between ranks
  int opt_queue_cap = 320;         // = kQCap of sidm.cu; option "queue_cap" lets tests force the queue-overflow fallback
  cudaStream_t coll_stream = nullptr;   // stream the pending collective has to be ordered on
  int opt_cand_cap = 1024;        // option "cand_cap": per-slot candidate capacity of the reference-order mode (parity runs)
  bool opt_group_search = true;    // b200_set_option("group_search", 0|1): warp-shared neighbour search for all-active passes
  bool overlap_now = false;        // true while the SIDM chain is being issued on stream_sidm
  bool walk_pending = false;       // a deferred walk whose counters / timing are still to be read

  // host binding (the reference's &P[1])
  char *h_base = nullptr; b200_layout lay{}; bool pinned = false; bool have_aos = false;
  int h_first = 0, h_count = 0;    // the bound host array holds the rows [h_first, h_first + h_count) of the particle order (b200_bind_rows)
  char *d_aos = nullptr; size_t aos_cap = 0;

  // ---- particle state, original index order (struct particle_data fields, allvars.h:422-460)
  float4 *posm = nullptr;    // PosPred.xyz, Mass
  float4 *velh = nullptr;    // Vel.xyz, HsmlVelDisp
  float  *pos0 = nullptr;    // Pos [n][3]
  float  *velpred = nullptr; // VelPred [n][3]
  float  *accel = nullptr;   // Accel [n][3]
  float  *dvel = nullptr;    // dVel [n][3]
  float  *maxpred = nullptr;  // MaxPredTime (written by b200_find_timesteps)
  float  *potential = nullptr; // Potential (written by b200_compute_potential)
  float  *curtime = nullptr, *oldacc = nullptr, *gravcost = nullptr, *left = nullptr, *right = nullptr;
  int    *ngb = nullptr, *pid = nullptr, *ptype = nullptr;
  // particle types present (one tree per type, forcetree.c:90-158): refreshed by the tree build when dirty
  bool types_dirty = true; int type_count[6] = {0, 0, 0, 0, 0, 0}; int ntypes = 1;
  unsigned char *stype = nullptr;  // several types: type of every sorted particle
  int ntrees = 1; int tree_type[6] = {1, 0, 0, 0, 0, 0}; int tree_root[7] = {0, 0, 0, 0, 0, 0, 0};   // roots in node order, [ntrees] = num_nodes

  // ---- tree (rebuilt by b200_tree_build)
  bool tree_valid = false;
  // tree reuse (option "tree_reuse" = k > 0): a full build every k-th build request, in between a REFIT - same topology (order,
  // cells, leaf order, level lists, search records), leaves and moments from the current positions (tree_build.cu tree_refit_impl)
  bool topo_valid = false; int opt_tree_reuse = 0, refits_since_build = 0;
  float *d_pad = nullptr;          // [2] largest coordinate displacement of a particle since the full build (searches inflate their cell tests by it), and the last refit's
  unsigned long long tree_epoch = 0, search_epoch = ~0ull;   // bumped by every tree build
  double *d_bbox = nullptr;        // [6] min xyz, max xyz (double), written by the bbox kernels
  RootBox *d_root = nullptr;       // root cell
  float *d_domain = nullptr;       // DomainMin[3], DomainMax[3]  (forcetree.c:192-198)
  uint64_t *key_hi = nullptr, *key_lo = nullptr;   // original order
  uint64_t *skey_hi = nullptr, *skey_lo = nullptr; // sorted
  uint64_t *key_tmp = nullptr;
  int *sidx = nullptr, *sidx_tmp = nullptr;        // sorted position -> original index
  int *krank = nullptr;                            // original index -> sorted position
  int *iota = nullptr;
  signed char *clev = nullptr;     // common levels with the next sorted particle
  int *nodestart = nullptr;        // [n+1] exclusive scan of nodes starting at each sorted particle
  void *cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
  int *d_flags = nullptr;          // [16] device-side status words (error flags, node count, ...)
  int *h_flags = nullptr;          // pinned mirror
  // per node (pre-order id)
  NodeRec *nodes = nullptr;
  float4 *geom = nullptr;          // centre xyz, len
  int *nstart = nullptr, *nend = nullptr, *nparent = nullptr, *npstart = nullptr;   // npstart [m+1]
  unsigned char *nlevel = nullptr, *nnp = nullptr, *nnchild = nullptr;
  int *ndp = nullptr;              // [m][8] sorted indices of a node's direct particles
  int *narrive = nullptr;
  int *nminidx = nullptr;          // min original index below the node (next[] order, forcetree.c:274-279)
  int *nlstart = nullptr;          // first position of the node's particles in next[] order
  Moments *nmom = nullptr;
  int *lev_off = nullptr;          // [72] cells per level -> offsets of the level lists (k_b5_levels)
  // sibling-pair records of the packed walk (walk.cu k_walk_pairs): the child cells of a node are stored next to each
  // other, two per 128-byte record, components interleaved for the f32x2 instructions
  PairRec *pairs = nullptr; int *gbase = nullptr; bool pairs_valid = false;
  bool opt_walk_pairs = false;     // b200_set_option("walk_pairs", 1): the packed sibling-pair walk (measured slower than k_walk: profiles/walk_pairs_r2_ncu_summary.txt)
  int opt_walkp_minb = 6;          // resident 128-thread blocks per SM the packed walk is compiled for (8, 6, 5 or 4)
  // leaf order (particles sorted by (parent node, octant)): every subtree is a contiguous range
  float4 *leaf_posm = nullptr;
  int *leaf_orig = nullptr, *orig_leaf = nullptr, *leaf_parent = nullptr;
  int *lrank = nullptr;            // original index -> rank in the reference's next[] chain
  int num_nodes = 0, max_level = 0;

  // ---- work buffers for the hot calls
  int *d_shard_list = nullptr;
  int *d_active = nullptr, *d_tsorted = nullptr, *d_tkeys = nullptr, *d_tkeys2 = nullptr, *d_tvals2 = nullptr;
  double *d_acc = nullptr;         // [n][3] raw accelerations per target slot
  int *d_cost = nullptr;           // [n][2]
  unsigned long long *d_ctr = nullptr;   // [CT_COUNT] device counters
  unsigned long long *h_ctr = nullptr;

  // ---- sidm work buffers
  int *s_slot_part = nullptr;      // [n] buffer slot -> particle
  int *s_flag = nullptr, *s_pos = nullptr;
  int *s_ngb = nullptr, *s_partner = nullptr, *s_pass = nullptr, *s_passlist = nullptr;
  double *s_rand = nullptr, *s_dir = nullptr, *s_pmax = nullptr, *s_prob = nullptr;
  float *s_dv = nullptr;           // [n][3]
  unsigned long long *s_winner = nullptr;   // per particle: winner_base + last buffer slot that chose it as partner (never reset: the base grows with every call)
  unsigned long long winner_base = 1;
  int *s_cand = nullptr; unsigned long long *s_candkey = nullptr; size_t s_cand_cap = 0;
  int *s_repair = nullptr;
  b200_scatlog *d_scatlog = nullptr; int scatlog_cap = 0; int scatlog_n = 0;
  int last_nslot = 0;
  // particles that received a partner kick since the last download (b200_download_active sends them along)
  int *kick_list = nullptr; int *d_nkick = nullptr;
  // pinned / device staging of the partial transfers (b200_upload_active / b200_download_active)
  char *h_stage = nullptr, *d_stage = nullptr; size_t stage_cap = 0;
  unsigned long long sidm_calls = 0;   // counter-based RNG: the key of a sidm() call is (Seed, sidm_calls); b200_get/set_rng_state
  unsigned long long ts_calls = 0;     // same for the Max/MinSizeTimestep jitter of find_timesteps()

  // ---- sharding across the GPUs of one box (b200_set_shard): this rank works on the 32-entry
  // blocks b of every sorted work list with b % world == rank; results are all-gathered
  int shard_rank = 0, shard_world = 1;
  void *shard_send = nullptr, *shard_recv = nullptr; long long shard_cap = 0;
  b200_allgather_fn shard_fn = nullptr; void *shard_user = nullptr; bool shard_own = false;

  float4 *d_ewald = nullptr; double ewald_box = -1;   // (ED+1)^3 correction table, scaled for ewald_box

  b200_counters cnt{};
};

// sharded work lists are dealt out in blocks of kShardBlock consecutive entries of the key order
// (= the 32 targets one warp of the walk shares an interaction list between)
constexpr int kShardBlock = 32, kShardShift = 5;
// number of entries of a sorted list of length nt that rank r of `world` owns
inline int shard_count(int nt, int world, int r) {
  const int nblk = (nt + kShardBlock - 1) / kShardBlock;
  long long c = 0;
  for (int q = r; q < nblk; q += world) { const int lo = q * kShardBlock; c += (nt - lo < kShardBlock) ? nt - lo : kShardBlock; }
  return (int)c;
}
inline int shard_max_blocks(int nt, int world) { const int nblk = (nt + kShardBlock - 1) / kShardBlock; return (nblk + world - 1) / world; }
int shard_select(const int *d_in, int nt, int *d_out, int *n_own, cudaStream_t st);
int shard_exchange(long long bytes_per_rank, cudaStream_t st);

extern Ctx g;

#define CUDA_TRY(x)                                                                      \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      g.last_cuda = (int)e_;                                                             \
      fprintf(stderr, "libsidm_b200: %s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return B200_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

#define B200_TRY(x) do { int r_ = (x); if (r_ != B200_OK) return r_; } while (0)

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }
inline void count_launch(int k = 1) { g.cnt.kernel_launches += k; }

// device status words (d_flags)
enum { FL_ERR_COINCIDENT = 0, FL_NUM_NODES = 1, FL_MAX_LEVEL = 2, FL_ERR_NGB = 3, FL_NPASS = 4,
       FL_NREPAIR = 5, FL_NSCATLOG = 6, FL_NEXPORT = 7, FL_MULTITYPE = 8, FL_TROOT0 = 9 /* .. 15 */, FL_COUNT = 16 };
// device counters (d_ctr)
enum { CT_PART = 0, CT_NODE = 1, CT_LIST_NODES = 2, CT_LIST_PARTS = 3, CT_WALK_OVF = 4, CT_CAND = 5, CT_PASS1 = 6,
       CT_SCATTERED = 7, CT_REJECTED = 8, CT_COUNT = 9 };
constexpr int kWalkCounters = 5;   // [CT_PART, CT_CAND): written by the walk; [CT_CAND, CT_COUNT): by the SIDM chain

// implemented across the .cu files
int tree_build_impl();
int walk_impl(const int *d_targets_sorted, int nt, bool with_slots, bool defer_sync = false);
int gravity_impl(const int *active, int nactive, double time, bool defer_sync = false);
int gravity_finish();
int gravity_exchange_early();     // sharded: issue the exchange of {Accel, OldAcc} behind the running walk; the buffers stay busy until gravity_finish()
inline cudaStream_t sidm_stream() { return g.overlap_now ? g.stream_sidm : g.stream; }
int direct_impl(const int *targets, int n, double *acc_out);
int sidm_impl(const int *d_active, int nactive, double time, double vmax, const b200_replay *replay, bool count_only, bool defer_final = false);
int prepare_targets(const int *active_host, int nactive, int **d_sorted_out);
void sidm_release();
void snapshot_release();

}  // namespace b200
