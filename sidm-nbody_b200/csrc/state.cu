// state.cu - life cycle, particle state (AoS <-> device SoA), prediction, counters.
// Reference roles: allocate.c:14-185 (buffers), predict.c:106-150 (predict_collisionless_only),
// sidm.c:970-990 (getvmax).
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "ctx.cuh"

namespace b200 {
Ctx g;
double s_a_inverse_at(double time);

template <class T>
static int dalloc(T **p, size_t count) {
  if (*p) return B200_OK;
  cudaError_t e = cudaMalloc((void **)p, count * sizeof(T) + 256);
  if (e != cudaSuccess) { g.last_cuda = (int)e; return B200_ERR_ALLOC; }
  return B200_OK;
}
template <class T>
static void dfree(T **p) { if (*p) cudaFree(*p); *p = nullptr; }

static int alloc_all() {
  const size_t n = (size_t)g.maxpart, m = (size_t)g.maxnodes;
  B200_TRY(dalloc(&g.posm, n)); B200_TRY(dalloc(&g.velh, n));
  B200_TRY(dalloc(&g.pos0, 3 * n)); B200_TRY(dalloc(&g.velpred, 3 * n));
  B200_TRY(dalloc(&g.accel, 3 * n)); B200_TRY(dalloc(&g.dvel, 3 * n));
  B200_TRY(dalloc(&g.curtime, n)); B200_TRY(dalloc(&g.oldacc, n)); B200_TRY(dalloc(&g.gravcost, n));
  B200_TRY(dalloc(&g.left, n)); B200_TRY(dalloc(&g.right, n)); B200_TRY(dalloc(&g.maxpred, n)); B200_TRY(dalloc(&g.potential, n));
  B200_TRY(dalloc(&g.ngb, n)); B200_TRY(dalloc(&g.pid, n)); B200_TRY(dalloc(&g.ptype, n));
  B200_TRY(dalloc(&g.d_bbox, 8)); B200_TRY(dalloc(&g.d_root, 8)); B200_TRY(dalloc(&g.d_domain, 48));   // root cell and DomainMin/Max per particle type
  B200_TRY(dalloc(&g.key_hi, n)); B200_TRY(dalloc(&g.key_lo, n));
  B200_TRY(dalloc(&g.skey_hi, n)); B200_TRY(dalloc(&g.skey_lo, n)); B200_TRY(dalloc(&g.key_tmp, n));
  B200_TRY(dalloc(&g.sidx, n)); B200_TRY(dalloc(&g.sidx_tmp, n)); B200_TRY(dalloc(&g.krank, n));
  B200_TRY(dalloc(&g.iota, n));
  B200_TRY(dalloc(&g.clev, n + 1)); B200_TRY(dalloc(&g.nodestart, n + 2));
  B200_TRY(dalloc(&g.d_flags, (size_t)FL_COUNT)); B200_TRY(dalloc(&g.d_ctr, (size_t)CT_COUNT));
  B200_TRY(dalloc(&g.nodes, m + 1)); B200_TRY(dalloc(&g.geom, m + 1));
  B200_TRY(dalloc(&g.nstart, m + 1)); B200_TRY(dalloc(&g.nend, m + 1)); B200_TRY(dalloc(&g.nparent, m + 1));
  B200_TRY(dalloc(&g.npstart, m + 2));
  B200_TRY(dalloc(&g.nlevel, m + 1)); B200_TRY(dalloc(&g.nnp, m + 1)); B200_TRY(dalloc(&g.nnchild, m + 1));
  B200_TRY(dalloc(&g.ndp, 8 * (m + 1))); B200_TRY(dalloc(&g.narrive, m + 1));
  B200_TRY(dalloc(&g.nminidx, m + 1)); B200_TRY(dalloc(&g.nlstart, m + 1));
  B200_TRY(dalloc(&g.nmom, m + 1)); B200_TRY(dalloc(&g.lev_off, (size_t)72)); B200_TRY(dalloc(&g.d_pad, (size_t)4));
  B200_TRY(dalloc(&g.pairs, m + 4)); B200_TRY(dalloc(&g.gbase, m + 2));
  B200_TRY(dalloc(&g.leaf_posm, n)); B200_TRY(dalloc(&g.leaf_orig, n)); B200_TRY(dalloc(&g.orig_leaf, n)); B200_TRY(dalloc(&g.leaf_parent, n));
  B200_TRY(dalloc(&g.lrank, n));
  B200_TRY(dalloc(&g.d_shard_list, n + 64));
  B200_TRY(dalloc(&g.d_active, n)); B200_TRY(dalloc(&g.d_tsorted, n)); B200_TRY(dalloc(&g.d_tkeys, n));
  B200_TRY(dalloc(&g.d_tkeys2, n)); B200_TRY(dalloc(&g.d_tvals2, n));
  B200_TRY(dalloc(&g.d_acc, 3 * n)); B200_TRY(dalloc(&g.d_cost, 2 * n));
  B200_TRY(dalloc(&g.s_slot_part, n)); B200_TRY(dalloc(&g.s_flag, n + 1)); B200_TRY(dalloc(&g.s_pos, n + 1));
  B200_TRY(dalloc(&g.s_ngb, n)); B200_TRY(dalloc(&g.s_partner, n)); B200_TRY(dalloc(&g.s_pass, n + 1));
  B200_TRY(dalloc(&g.s_passlist, n));
  B200_TRY(dalloc(&g.s_rand, n)); B200_TRY(dalloc(&g.s_dir, 3 * n)); B200_TRY(dalloc(&g.s_pmax, n));
  B200_TRY(dalloc(&g.s_prob, n)); B200_TRY(dalloc(&g.s_dv, 3 * n)); B200_TRY(dalloc(&g.s_winner, n));
  B200_TRY(dalloc(&g.s_repair, n));
  B200_TRY(dalloc(&g.kick_list, n)); B200_TRY(dalloc(&g.d_nkick, (size_t)4));
  g.scatlog_cap = 1 << 20;
  B200_TRY(dalloc(&g.d_scatlog, (size_t)g.scatlog_cap));
  return B200_OK;
}

}  // namespace b200
using namespace b200;

extern "C" const char *b200_version(void) { return "sidm_b200 0.1 (sm_100a)"; }
extern "C" int b200_last_cuda_error(void) { return g.last_cuda; }
extern "C" int b200_set_stream(void *cuda_stream) { g.stream = (cudaStream_t)cuda_stream; return B200_OK; }
extern "C" void *b200_current_stream(void) { return (void *)g.coll_stream; }
extern "C" int b200_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
extern "C" int b200_shard_buffers(void **send, void **recv, long long *cap_bytes) {
  if (g.shard_world <= 1) return B200_ERR_STATE;
  if (send) *send = g.shard_send; if (recv) *recv = g.shard_recv; if (cap_bytes) *cap_bytes = g.shard_cap;
  return B200_OK;
}
extern "C" int b200_set_option(const char *name, int value) {
  if (!name) return B200_ERR_ARG;
  if (!g.ready) return B200_ERR_STATE;          // options belong to a context: b200_init resets them to their defaults
  if (!strcmp(name, "overlap")) { g.opt_overlap = value < 0 ? -1 : (value > 2 ? 2 : value); return B200_OK; }
  if (!strcmp(name, "group_search")) { g.opt_group_search = value != 0; return B200_OK; }
  if (!strcmp(name, "walkp_minb")) { g.opt_walkp_minb = value; return B200_OK; }
  if (!strcmp(name, "walk_pairs")) { g.opt_walk_pairs = value != 0; g.tree_valid = false; g.topo_valid = false; return B200_OK; }
  if (!strcmp(name, "tree_reuse")) { g.opt_tree_reuse = value < 0 ? 0 : value; return B200_OK; }
  if (!strcmp(name, "shard_overlap")) { g.opt_shard_overlap = value != 0; return B200_OK; }
  if (!strcmp(name, "shard_min_work")) { g.shard_min_work = value; return B200_OK; }
  if (!strcmp(name, "compact_exchange")) { g.opt_compact_exchange = value != 0; return B200_OK; }
  if (!strcmp(name, "cand_cap")) { g.opt_cand_cap = value < 64 ? 64 : (value > (1 << 20) ? (1 << 20) : value); return B200_OK; }
  if (!strcmp(name, "queue_cap")) { g.opt_queue_cap = value < 2 ? 2 : (value > 320 ? 320 : value); return B200_OK; }
  return B200_ERR_ARG;
}

// restart support (restart.c:37-154 dumps All and P[] but not the generator state): the random numbers of this
// path are a pure function of (Seed, call counter, particle index), so two counters are the whole generator state
extern "C" int b200_get_rng_state(unsigned long long *state) {
  if (!state) return B200_ERR_ARG;
  state[0] = g.sidm_calls; state[1] = g.ts_calls;
  return B200_OK;
}
extern "C" int b200_set_rng_state(const unsigned long long *state) {
  if (!state) return B200_ERR_ARG;
  g.sidm_calls = state[0]; g.ts_calls = state[1];
  return B200_OK;
}

extern "C" int b200_set_shard(int rank, int world, void *send, void *recv, long long cap_bytes, b200_allgather_fn fn, void *user) {
  if (world < 1 || rank < 0 || rank >= world) return B200_ERR_ARG;
  if (world > 1 && (!fn || cap_bytes <= 0 || (send == nullptr) != (recv == nullptr))) return B200_ERR_ARG;
  if (g.shard_own) { cudaFree(g.shard_send); cudaFree(g.shard_recv); g.shard_own = false; g.shard_send = g.shard_recv = nullptr; }
  if (world > 1 && !send) {                   // a C host without CUDA of its own (the shim): the library owns the exchange buffers
    if (!g.ready) return B200_ERR_STATE;
    if (cudaMalloc(&send, (size_t)cap_bytes) != cudaSuccess || cudaMalloc(&recv, (size_t)cap_bytes * world) != cudaSuccess) return B200_ERR_ALLOC;
    g.shard_own = true;
  }
  g.shard_rank = rank; g.shard_world = world; g.shard_send = send; g.shard_recv = recv; g.shard_cap = cap_bytes;
  g.shard_fn = fn; g.shard_user = user;
  g.search_epoch = ~0ull;        // the query groups of the SIDM search depend on the sharding (sidm.cu)
  return B200_OK;
}

// own[k] = in[(k/32*world + rank)*32 + k%32] : this rank's blocks of a sorted work list
__global__ void k_shard_select(int nt, int world, int rank, const int *in, int *out, int nown_max) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nown_max) return;
  const long long j = ((long long)(k >> kShardShift) * world + rank) * kShardBlock + (k & (kShardBlock - 1));
  if (j < nt) out[k] = in[j];
}
namespace b200 {
int shard_select(const int *d_in, int nt, int *d_out, int *n_own, cudaStream_t st) {
  const int nown = shard_count(nt, g.shard_world, g.shard_rank);
  *n_own = nown;
  if (nown > 0) {
    k_shard_select<<<cdiv(nown, 256), 256, 0, st>>>(nt, g.shard_world, g.shard_rank, d_in, d_out, nown);
    count_launch();
  }
  return B200_OK;
}
int shard_exchange(long long bytes_per_rank, cudaStream_t st) {
  if (bytes_per_rank > g.shard_cap) return B200_ERR_ARG;
  g.coll_stream = st;                    // what b200_current_stream() tells the callback
  const int rc = g.shard_fn(bytes_per_rank, g.shard_user);
  return rc == 0 ? B200_OK : B200_ERR_STATE;
}
}

extern "C" int b200_set_params(const b200_params *p) {
  if (!p) return B200_ERR_ARG;
  const int dev = g.par.device, mp = g.par.MaxPart; const double taf = g.par.TreeAllocFactor;
  const bool was = g.ready;
  g.par = *p;
  if (was) { g.par.device = dev; g.par.MaxPart = mp; g.par.TreeAllocFactor = taf; }
  if (g.par.CrossSectionType < 0 || g.par.CrossSectionType > 4) return B200_ERR_ARG;
  return B200_OK;
}

extern "C" int b200_init(const b200_params *p) {
  if (!p || p->MaxPart <= 0) return B200_ERR_ARG;
  if (g.ready) return B200_ERR_STATE;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g.last_cuda = (int)e;
    fprintf(stderr, "libsidm_b200: no CUDA device (%s); this library has no CPU path\n", cudaGetErrorString(e));
    return B200_ERR_NODEVICE;
  }
  if (p->device < 0 || p->device >= ndev) return B200_ERR_ARG;
  CUDA_TRY(cudaSetDevice(p->device));
  B200_TRY(b200_set_params(p));
  g.maxpart = p->MaxPart;
  double taf = p->TreeAllocFactor > 0 ? p->TreeAllocFactor : 0.8;
  g.maxnodes = (int)(taf * (double)g.maxpart) + 64;
  g.stream = nullptr;   // legacy default stream: what torch's current stream is unless the host changes it
  CUDA_TRY(cudaEventCreate(&g.ev0)); CUDA_TRY(cudaEventCreate(&g.ev1));
  CUDA_TRY(cudaEventCreate(&g.ev_s0)); CUDA_TRY(cudaEventCreate(&g.ev_s1));
  CUDA_TRY(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&g.ev_join, cudaEventDisableTiming));
  {
    int plo = 0, phi = 0;                 // numerically lowest = highest priority
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&plo, &phi));
    CUDA_TRY(cudaStreamCreateWithPriority(&g.stream_sidm, cudaStreamNonBlocking, phi));
  }
  // every option starts from its default in a new context (b200_set_option is legal only between b200_init and b200_finalize's
  // successor: a setting must not leak from one run of a process into the next)
  g.opt_shard_overlap = false; g.shard_min_work = kShardMinWorkDefault; g.opt_compact_exchange = true; g.opt_queue_cap = 320;
  g.opt_cand_cap = 1024; g.opt_group_search = true; g.opt_tree_reuse = 0; g.opt_walk_pairs = false; g.opt_walkp_minb = 6;
  g.opt_overlap = getenv("B200_NO_OVERLAP") ? 0 : (getenv("B200_OVERLAP") ? atoi(getenv("B200_OVERLAP")) : -1); g.overlap_now = false; g.walk_pending = false;
  g.winner_base = 1;
  int rc = alloc_all();
  if (rc != B200_OK) { b200_finalize(); return rc; }
  CUDA_TRY(cudaMallocHost((void **)&g.h_flags, (FL_COUNT + 4) * sizeof(int)));   // + the scatter-log position (sidm_collect)
  CUDA_TRY(cudaMallocHost((void **)&g.h_ctr, CT_COUNT * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(g.d_flags, 0, FL_COUNT * sizeof(int), g.stream));
  CUDA_TRY(cudaMemsetAsync(g.d_ctr, 0, CT_COUNT * sizeof(unsigned long long), g.stream));
  CUDA_TRY(cudaMemsetAsync(g.d_nkick, 0, 4 * sizeof(int), g.stream));
  CUDA_TRY(cudaMemsetAsync(g.s_winner, 0, (size_t)g.maxpart * sizeof(unsigned long long), g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  memset(&g.cnt, 0, sizeof(g.cnt));
  g.n = 0; g.tree_valid = false; g.topo_valid = false; g.refits_since_build = 0; g.sidm_calls = 0; g.ts_calls = 0;
  g.ready = true;
  return B200_OK;
}

extern "C" void b200_finalize(void) {
  if (g.ready) cudaStreamSynchronize(g.stream);
  if (g.pinned && g.h_base) cudaHostUnregister(g.h_base);
  g.pinned = false; g.h_base = nullptr; g.have_aos = false;
  dfree(&g.d_aos); g.aos_cap = 0;
  dfree(&g.posm); dfree(&g.velh); dfree(&g.pos0); dfree(&g.velpred); dfree(&g.accel); dfree(&g.dvel);
  dfree(&g.maxpred); dfree(&g.potential);
  if (g.stype) cudaFree(g.stype); g.stype = nullptr; g.types_dirty = true; g.ntypes = 1; g.ntrees = 1;
  dfree(&g.curtime); dfree(&g.oldacc); dfree(&g.gravcost); dfree(&g.left); dfree(&g.right);
  dfree(&g.ngb); dfree(&g.pid); dfree(&g.ptype);
  dfree(&g.d_bbox); dfree(&g.d_root); dfree(&g.d_domain);
  dfree(&g.key_hi); dfree(&g.key_lo); dfree(&g.skey_hi); dfree(&g.skey_lo); dfree(&g.key_tmp);
  dfree(&g.sidx); dfree(&g.sidx_tmp); dfree(&g.krank); dfree(&g.iota); dfree(&g.clev); dfree(&g.nodestart);
  dfree(&g.cub_tmp); g.cub_tmp_bytes = 0;
  dfree(&g.d_flags); dfree(&g.d_ctr);
  dfree(&g.nodes); dfree(&g.geom); dfree(&g.nstart); dfree(&g.nend); dfree(&g.nparent); dfree(&g.npstart);
  dfree(&g.nlevel); dfree(&g.nnp); dfree(&g.nnchild); dfree(&g.ndp); dfree(&g.narrive);
  dfree(&g.nminidx); dfree(&g.nlstart); dfree(&g.nmom); dfree(&g.lev_off); dfree(&g.d_pad); dfree(&g.pairs); dfree(&g.gbase); g.pairs_valid = false;
  dfree(&g.leaf_posm); dfree(&g.leaf_orig); dfree(&g.orig_leaf); dfree(&g.leaf_parent); dfree(&g.lrank);
  dfree(&g.d_shard_list);
  if (g.shard_own) { cudaFree(g.shard_send); cudaFree(g.shard_recv); g.shard_own = false; g.shard_send = g.shard_recv = nullptr; g.shard_world = 1; g.shard_rank = 0; }
  dfree(&g.d_active); dfree(&g.d_tsorted); dfree(&g.d_tkeys); dfree(&g.d_tkeys2); dfree(&g.d_tvals2);
  dfree(&g.d_acc); dfree(&g.d_cost);
  dfree(&g.s_slot_part); dfree(&g.s_flag); dfree(&g.s_pos); dfree(&g.s_ngb); dfree(&g.s_partner);
  dfree(&g.s_pass); dfree(&g.s_passlist); dfree(&g.s_rand); dfree(&g.s_dir); dfree(&g.s_pmax);
  dfree(&g.s_prob); dfree(&g.s_dv); dfree(&g.s_winner); dfree(&g.s_repair);
  dfree(&g.s_cand); dfree(&g.s_candkey); g.s_cand_cap = 0;
  dfree(&g.d_scatlog); dfree(&g.kick_list); dfree(&g.d_nkick);
  if (g.h_stage) cudaFreeHost(g.h_stage); g.h_stage = nullptr; dfree(&g.d_stage); g.stage_cap = 0;
  sidm_release(); snapshot_release();
  g.search_epoch = ~0ull; g.tree_epoch = 0;
  dfree(&g.d_ewald); g.ewald_box = -1;
  if (g.h_flags) cudaFreeHost(g.h_flags); g.h_flags = nullptr;
  if (g.h_ctr) cudaFreeHost(g.h_ctr); g.h_ctr = nullptr;
  if (g.ev0) cudaEventDestroy(g.ev0); g.ev0 = nullptr;
  if (g.ev1) cudaEventDestroy(g.ev1); g.ev1 = nullptr;
  if (g.ev_s0) cudaEventDestroy(g.ev_s0); g.ev_s0 = nullptr;
  if (g.ev_s1) cudaEventDestroy(g.ev_s1); g.ev_s1 = nullptr;
  if (g.ev_fork) cudaEventDestroy(g.ev_fork); g.ev_fork = nullptr;
  if (g.ev_join) cudaEventDestroy(g.ev_join); g.ev_join = nullptr;
  if (g.stream_sidm) cudaStreamDestroy(g.stream_sidm); g.stream_sidm = nullptr;
  g.overlap_now = false; g.walk_pending = false;
  g.stream = nullptr;
  g.ready = false; g.n = 0; g.tree_valid = false; g.topo_valid = false;
}

// ----------------------------------------------------------------------------- SoA I/O

static int h2d(void *dst, const void *src, size_t bytes) {
  if (!src || !bytes) return B200_OK;
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g.stream));
  return B200_OK;
}
static int d2h(void *dst, const void *src, size_t bytes) {
  if (!dst || !bytes) return B200_OK;
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g.stream));
  return B200_OK;
}

__global__ void k_pack_soa(int n, const float *pos, const float *vel, const float *mass, const float *hsml,
                           float4 *posm, float4 *velh, float *pos0, float *velpred, int has_pos, int has_vel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = posm[i], b = velh[i];
  if (has_pos) { a.x = pos[3 * i]; a.y = pos[3 * i + 1]; a.z = pos[3 * i + 2]; pos0[3 * i] = a.x; pos0[3 * i + 1] = a.y; pos0[3 * i + 2] = a.z; }
  if (mass) a.w = mass[i];
  if (has_vel) { b.x = vel[3 * i]; b.y = vel[3 * i + 1]; b.z = vel[3 * i + 2]; velpred[3 * i] = b.x; velpred[3 * i + 1] = b.y; velpred[3 * i + 2] = b.z; }
  if (hsml) b.w = hsml[i];
  posm[i] = a; velh[i] = b;
}

__global__ void k_fill_i(int n, int *a, int v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = v; }
__global__ void k_fill_f(int n, float *a, float v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = v; }

extern "C" int b200_set_soa(int n, const float *pos, const float *vel, const float *mass, const int *id,
                            const float *curtime, const float *accel, const float *oldacc,
                            const float *hsml, const float *dvel) {
  if (!g.ready) return B200_ERR_STATE;
  if (n <= 0 || n > g.maxpart) return B200_ERR_ARG;
  const bool fresh = (n != g.n);
  g.n = n;
  if (fresh || pos || mass) g.tree_valid = false;   // the tree depends on PosPred and Mass only
  if (fresh || pos) g.topo_valid = false;           // new particles / arbitrary new positions: no refit of the old topology
  const int B = 256, G = cdiv(n, B);
  if (fresh) {   // start-up state of init.c:76-100
    CUDA_TRY(cudaMemsetAsync(g.posm, 0, n * sizeof(float4), g.stream));
    CUDA_TRY(cudaMemsetAsync(g.velh, 0, n * sizeof(float4), g.stream));
    CUDA_TRY(cudaMemsetAsync(g.accel, 0, 3 * n * sizeof(float), g.stream));
    CUDA_TRY(cudaMemsetAsync(g.dvel, 0, 3 * n * sizeof(float), g.stream));
    CUDA_TRY(cudaMemsetAsync(g.curtime, 0, n * sizeof(float), g.stream));
    CUDA_TRY(cudaMemsetAsync(g.oldacc, 0, n * sizeof(float), g.stream));
    CUDA_TRY(cudaMemsetAsync(g.left, 0, n * sizeof(float), g.stream));
    CUDA_TRY(cudaMemsetAsync(g.right, 0, n * sizeof(float), g.stream));
    CUDA_TRY(cudaMemsetAsync(g.ngb, 0, n * sizeof(int), g.stream));
    CUDA_TRY(cudaMemsetAsync(g.pid, 0, n * sizeof(int), g.stream));
    k_fill_i<<<G, B, 0, g.stream>>>(n, g.ptype, 1);
    g.types_dirty = true;
    k_fill_f<<<G, B, 0, g.stream>>>(n, g.gravcost, 1.0f);
    count_launch(2);
  }
  // stage through scratch (d_acc is large enough for [n][3] floats twice over)
  float *st_pos = (float *)g.d_acc, *st_vel = st_pos + 3 * (size_t)n;
  float *st_mass = (float *)g.d_cost, *st_h = st_mass + n;
  B200_TRY(h2d(st_pos, pos, 3 * n * sizeof(float)));
  B200_TRY(h2d(st_vel, vel, 3 * n * sizeof(float)));
  B200_TRY(h2d(st_mass, mass, n * sizeof(float)));
  B200_TRY(h2d(st_h, hsml, n * sizeof(float)));
  k_pack_soa<<<G, B, 0, g.stream>>>(n, st_pos, st_vel, mass ? st_mass : nullptr, hsml ? st_h : nullptr,
                                    g.posm, g.velh, g.pos0, g.velpred, pos != nullptr, vel != nullptr);
  count_launch();
  B200_TRY(h2d(g.pid, id, n * sizeof(int)));
  B200_TRY(h2d(g.curtime, curtime, n * sizeof(float)));
  B200_TRY(h2d(g.accel, accel, 3 * n * sizeof(float)));
  B200_TRY(h2d(g.oldacc, oldacc, n * sizeof(float)));
  B200_TRY(h2d(g.dvel, dvel, 3 * n * sizeof(float)));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

__global__ void k_unpack_soa(int n, const float4 *posm, const float4 *velh, float *pospred, float *hsml) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pospred) { float4 a = posm[i]; pospred[3 * i] = a.x; pospred[3 * i + 1] = a.y; pospred[3 * i + 2] = a.z; }
  if (hsml) hsml[i] = velh[i].w;
}

extern "C" int b200_get_soa(float *pospred, float *velpred, float *accel, float *oldacc, float *gravcost,
                            float *hsml, int *ngb, float *dvel, float *left, float *right) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  const int n = g.n;
  float *st_pos = (float *)g.d_acc; float *st_h = (float *)g.d_cost;
  if (pospred || hsml) {
    k_unpack_soa<<<cdiv(n, 256), 256, 0, g.stream>>>(n, g.posm, g.velh, pospred ? st_pos : nullptr, hsml ? st_h : nullptr);
    count_launch();
  }
  B200_TRY(d2h(pospred, st_pos, 3 * n * sizeof(float)));
  B200_TRY(d2h(hsml, st_h, n * sizeof(float)));
  B200_TRY(d2h(velpred, g.velpred, 3 * n * sizeof(float)));
  B200_TRY(d2h(accel, g.accel, 3 * n * sizeof(float)));
  B200_TRY(d2h(oldacc, g.oldacc, n * sizeof(float)));
  B200_TRY(d2h(gravcost, g.gravcost, n * sizeof(float)));
  B200_TRY(d2h(ngb, g.ngb, n * sizeof(int)));
  B200_TRY(d2h(dvel, g.dvel, 3 * n * sizeof(float)));
  B200_TRY(d2h(left, g.left, n * sizeof(float)));
  B200_TRY(d2h(right, g.right, n * sizeof(float)));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// ----------------------------------------------------------------------------- AoS I/O

extern "C" int b200_bind_particles(void *base, int num_part, const b200_layout *lay, int pin) {
  if (!g.ready) return B200_ERR_STATE;
  if (!base || !lay || num_part <= 0 || num_part > g.maxpart || lay->stride <= 0 || (lay->stride & 3) || lay->stride > 188) return B200_ERR_ARG;
  if (g.pinned && g.h_base && g.h_base != (char *)base) { cudaHostUnregister(g.h_base); g.pinned = false; }
  if (num_part != g.n) g.topo_valid = false;
  g.h_base = (char *)base; g.lay = *lay; g.n = num_part; g.tree_valid = false; g.h_first = 0; g.h_count = num_part;
  const size_t bytes = (size_t)g.maxpart * lay->stride;
  if (g.aos_cap < bytes) {
    dfree(&g.d_aos);
    if (cudaMalloc((void **)&g.d_aos, bytes) != cudaSuccess) return B200_ERR_ALLOC;
    g.aos_cap = bytes;
  }
  if (pin && !g.pinned) {
    // page-lock the caller's array once (the reference allocates P once, allocate.c:127-160)
    cudaError_t e = cudaHostRegister(base, (size_t)num_part * lay->stride, cudaHostRegisterDefault);
    if (e == cudaSuccess) g.pinned = true; else (void)cudaGetLastError();
  }
  g.have_aos = true;
  return B200_OK;
}

// The host array of one rank of a distributed run (the reference's per-task P[], domain.c): `base` holds the rows
// [first, first+count) of the global particle order = the tasks' arrays one after the other.
extern "C" int b200_bind_rows(void *base, int first, int count, int n_global, const b200_layout *lay, int pin) {
  if (!g.ready) return B200_ERR_STATE;
  if (!base || !lay || count < 0 || first < 0 || n_global <= 0 || first + count > n_global || n_global > g.maxpart || lay->stride <= 0 || (lay->stride & 3) || lay->stride > 188) return B200_ERR_ARG;
  if (g.pinned && g.h_base && (g.h_base != (char *)base || g.h_count != count)) { cudaHostUnregister(g.h_base); g.pinned = false; }
  g.topo_valid = false;
  g.h_base = (char *)base; g.lay = *lay; g.n = n_global; g.tree_valid = false; g.h_first = first; g.h_count = count;
  const size_t bytes = (size_t)g.maxpart * lay->stride;
  if (g.aos_cap < bytes) {
    dfree(&g.d_aos);
    if (cudaMalloc((void **)&g.d_aos, bytes) != cudaSuccess) return B200_ERR_ALLOC;
    g.aos_cap = bytes;
  }
  if (pin && !g.pinned && count > 0) {
    cudaError_t e = cudaHostRegister(base, (size_t)count * lay->stride, cudaHostRegisterDefault);
    if (e == cudaSuccess) g.pinned = true; else (void)cudaGetLastError();
  }
  g.have_aos = true;
  return B200_OK;
}

struct Lay { int stride, Pos, Vel, Mass, ID, Type, CurrentTime, PosPred, VelPred, Accel, GravCost, OldAcc, Left, Right, Ngb, Hsml, dVel, MaxPred, Pot; };

__device__ __forceinline__ float ldf(const char *p, int off) { return *(const float *)(p + off); }
__device__ __forceinline__ int ldi(const char *p, int off) { return *(const int *)(p + off); }

// AoS <-> SoA through shared memory: a block moves kAosRows records; the array-of-structs side is read / written as one
// contiguous stream of 16-byte words, the field accesses go to shared memory (a 124-byte record is 31 words: the lanes of a
// warp hit 32 different banks), the structure-of-arrays side is coalesced as before.
constexpr int kAosRows = 256;
__device__ __forceinline__ void aos_block_load(const char *aos, int stride, int r0, int nrec, int *sm) {
  const size_t off = (size_t)r0 * stride;                    // kAosRows * stride is a multiple of 16 (stride % 4 == 0)
  const int words = nrec * (stride >> 2);
  const int4 *src4 = reinterpret_cast<const int4 *>(aos + off);
  int4 *dst4 = reinterpret_cast<int4 *>(sm);
  const int quads = words >> 2;
  for (int q = threadIdx.x; q < quads; q += blockDim.x) dst4[q] = src4[q];
  for (int w = (quads << 2) + threadIdx.x; w < words; w += blockDim.x) sm[w] = reinterpret_cast<const int *>(aos + off)[w];
}
__device__ __forceinline__ void aos_block_store(char *aos, int stride, int r0, int nrec, const int *sm) {
  const size_t off = (size_t)r0 * stride;
  const int words = nrec * (stride >> 2);
  int4 *dst4 = reinterpret_cast<int4 *>(aos + off);
  const int4 *src4 = reinterpret_cast<const int4 *>(sm);
  const int quads = words >> 2;
  for (int q = threadIdx.x; q < quads; q += blockDim.x) dst4[q] = src4[q];
  for (int w = (quads << 2) + threadIdx.x; w < words; w += blockDim.x) reinterpret_cast<int *>(aos + off)[w] = sm[w];
}

__global__ void __launch_bounds__(kAosRows) k_unpack_aos(int n, const char *aos, Lay L, float4 *posm, float4 *velh, float *pos0, float *velpred,
                             float *accel, float *dvel, float *curtime, float *oldacc, float *gravcost,
                             float *left, float *right, int *ngb, int *pid, int *ptype, float *maxpred, float *potential) {
  extern __shared__ __align__(16) int sm_aos[];
  const int r0 = blockIdx.x * kAosRows, nrec = min(kAosRows, n - r0);
  aos_block_load(aos, L.stride, r0, nrec, sm_aos);
  __syncthreads();
  if ((int)threadIdx.x >= nrec) return;
  const int i = r0 + threadIdx.x;
  const char *p = reinterpret_cast<const char *>(sm_aos) + (size_t)threadIdx.x * L.stride;
  if (L.Pot > 0) potential[i] = ldf(p, L.Pot);
  if (L.MaxPred > 0) maxpred[i] = ldf(p, L.MaxPred);
  float4 a, b;
  a.x = ldf(p, L.PosPred); a.y = ldf(p, L.PosPred + 4); a.z = ldf(p, L.PosPred + 8); a.w = ldf(p, L.Mass);
  b.x = ldf(p, L.Vel); b.y = ldf(p, L.Vel + 4); b.z = ldf(p, L.Vel + 8); b.w = ldf(p, L.Hsml);
  posm[i] = a; velh[i] = b;
  for (int k = 0; k < 3; k++) {
    pos0[3 * (size_t)i + k] = ldf(p, L.Pos + 4 * k);
    velpred[3 * (size_t)i + k] = ldf(p, L.VelPred + 4 * k);
    accel[3 * (size_t)i + k] = ldf(p, L.Accel + 4 * k);
    dvel[3 * (size_t)i + k] = ldf(p, L.dVel + 4 * k);
  }
  curtime[i] = ldf(p, L.CurrentTime); oldacc[i] = ldf(p, L.OldAcc); gravcost[i] = ldf(p, L.GravCost);
  left[i] = ldf(p, L.Left); right[i] = ldf(p, L.Right);
  ngb[i] = ldi(p, L.Ngb); pid[i] = ldi(p, L.ID); ptype[i] = ldi(p, L.Type);
}

// the device image keeps the fields the path does not write (Pos Vel Mass ID Type CurrentTime ForceFlag) as uploaded
__global__ void __launch_bounds__(kAosRows) k_pack_aos(int first, int n, char *aos, Lay L, const float4 *posm, const float4 *velh, const float *velpred,
                           const float *accel, const float *dvel, const float *oldacc, const float *gravcost,
                           const float *left, const float *right, const int *ngb, const float *maxpred, const float *potential) {
  extern __shared__ __align__(16) int sm_aos[];
  const int r0 = first + blockIdx.x * kAosRows, nrec = min(kAosRows, first + n - r0);     // rows [first, first + n)
  aos_block_load(aos, L.stride, r0, nrec, sm_aos);
  __syncthreads();
  if ((int)threadIdx.x < nrec) {
    const int i = r0 + threadIdx.x;
    char *p = reinterpret_cast<char *>(sm_aos) + (size_t)threadIdx.x * L.stride;
    if (L.Pot > 0) *(float *)(p + L.Pot) = potential[i];
    if (L.MaxPred > 0) *(float *)(p + L.MaxPred) = maxpred[i];
    const float4 a = posm[i];
    *(float *)(p + L.PosPred) = a.x; *(float *)(p + L.PosPred + 4) = a.y; *(float *)(p + L.PosPred + 8) = a.z;
    for (int k = 0; k < 3; k++) {
      *(float *)(p + L.VelPred + 4 * k) = velpred[3 * (size_t)i + k];
      *(float *)(p + L.Accel + 4 * k) = accel[3 * (size_t)i + k];
      *(float *)(p + L.dVel + 4 * k) = dvel[3 * (size_t)i + k];
    }
    *(float *)(p + L.OldAcc) = oldacc[i]; *(float *)(p + L.GravCost) = gravcost[i];
    *(float *)(p + L.Left) = left[i]; *(float *)(p + L.Right) = right[i];
    *(int *)(p + L.Ngb) = ngb[i]; *(float *)(p + L.Hsml) = velh[i].w;
  }
  __syncthreads();
  aos_block_store(aos, L.stride, r0, nrec, sm_aos);
}

static Lay to_lay(const b200_layout &l) {
  Lay L{l.stride, l.Pos, l.Vel, l.Mass, l.ID, l.Type, l.CurrentTime, l.PosPred, l.VelPred, l.Accel,
        l.GravCost, l.OldAcc, l.Left, l.Right, l.NgbVelDisp, l.HsmlVelDisp, l.dVel, l.MaxPredTime, l.Potential};
  return L;
}

extern "C" int b200_upload(void) {
  if (!g.ready || !g.have_aos) return B200_ERR_STATE;
  const int n = g.n;
  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  CUDA_TRY(cudaMemcpyAsync(g.d_aos, g.h_base, (size_t)n * g.lay.stride, cudaMemcpyHostToDevice, g.stream));
  k_unpack_aos<<<cdiv(n, kAosRows), kAosRows, (size_t)kAosRows * g.lay.stride, g.stream>>>(n, g.d_aos, to_lay(g.lay), g.posm, g.velh, g.pos0, g.velpred, g.accel,
                                                   g.dvel, g.curtime, g.oldacc, g.gravcost, g.left, g.right, g.ngb, g.pid, g.ptype, g.maxpred, g.potential);
  count_launch();
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  cudaEventElapsedTime(&g.cnt.ms_upload, g.ev0, g.ev1);
  g.tree_valid = false; g.types_dirty = true; g.topo_valid = false;
  return B200_OK;
}

extern "C" int b200_download_to(void *dst) {
  if (!g.ready || !g.have_aos) return B200_ERR_STATE;
  char *keep = g.h_base;
  if (dst) g.h_base = (char *)dst;
  const int rc = b200_download();
  g.h_base = keep;
  return rc;
}

extern "C" int b200_download(void) {
  if (!g.ready || !g.have_aos) return B200_ERR_STATE;
  const int n = g.n;
  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  k_pack_aos<<<cdiv(n, kAosRows), kAosRows, (size_t)kAosRows * g.lay.stride, g.stream>>>(0, n, g.d_aos, to_lay(g.lay), g.posm, g.velh, g.velpred, g.accel, g.dvel,
                                                 g.oldacc, g.gravcost, g.left, g.right, g.ngb, g.maxpred, g.potential);
  count_launch();
  CUDA_TRY(cudaMemcpyAsync(g.h_base, g.d_aos, (size_t)n * g.lay.stride, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.d_nkick, 0, sizeof(int), g.stream));        // every partner kick has gone to the host
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  cudaEventElapsedTime(&g.cnt.ms_download, g.ev0, g.ev1);
  return B200_OK;
}

// every rank's rows -> all ranks: own rows over PCIe, one all-gather over NVLink, compacted into the device image of the
// whole array.  counts[q] = rows of rank q (their sum = the particle number), rank q's rows start at the sum of the counts before it.
static int upload_rows_impl(const int *counts) {
  const int n = g.n, W = g.shard_world; const size_t st = (size_t)g.lay.stride;
  long long tot = 0; int rows_per_rank = 0, first = 0;
  for (int q = 0; q < W; q++) { if (counts[q] < 0) return B200_ERR_ARG; if (q < g.shard_rank) first += counts[q]; tot += counts[q]; if (counts[q] > rows_per_rank) rows_per_rank = counts[q]; }
  const int count = counts[g.shard_rank];
  if (tot != n || first < g.h_first || first + count > g.h_first + g.h_count) return B200_ERR_ARG;
  const long long bytes = (long long)rows_per_rank * (long long)st;
  if (bytes > g.shard_cap) return B200_ERR_ARG;
  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  if (count > 0) CUDA_TRY(cudaMemcpyAsync((char *)g.shard_send, g.h_base + (size_t)(first - g.h_first) * st, (size_t)count * st, cudaMemcpyHostToDevice, g.stream));
  B200_TRY(shard_exchange(bytes, g.stream));             // every rank's rows -> all ranks, over NVLink
  long long f = 0;
  for (int q = 0; q < W; q++) {
    if (counts[q] > 0) CUDA_TRY(cudaMemcpyAsync(g.d_aos + (size_t)f * st, (char *)g.shard_recv + (size_t)q * bytes, (size_t)counts[q] * st, cudaMemcpyDeviceToDevice, g.stream));
    f += counts[q];
  }
  k_unpack_aos<<<cdiv(n, kAosRows), kAosRows, (size_t)kAosRows * g.lay.stride, g.stream>>>(n, g.d_aos, to_lay(g.lay), g.posm, g.velh, g.pos0, g.velpred, g.accel,
                                                   g.dvel, g.curtime, g.oldacc, g.gravcost, g.left, g.right, g.ngb, g.pid, g.ptype, g.maxpred, g.potential);
  count_launch();
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  cudaEventElapsedTime(&g.cnt.ms_upload, g.ev0, g.ev1);
  g.tree_valid = false; g.types_dirty = true; g.topo_valid = false;
  return B200_OK;
}

extern "C" int b200_upload_rows(const int *counts) {
  if (!g.ready || !g.have_aos) return B200_ERR_STATE;
  if (!counts) return B200_ERR_ARG;
  if (g.shard_world <= 1) return counts[0] == g.n ? b200_upload() : B200_ERR_ARG;
  return upload_rows_impl(counts);
}

extern "C" int b200_upload_shard(int first, int count, int rows_per_rank) {
  if (!g.ready || !g.have_aos) return B200_ERR_STATE;
  if (g.shard_world <= 1) return b200_upload();
  const int n = g.n;
  if (first < 0 || count < 0 || first + count > n || rows_per_rank < count || (long long)rows_per_rank * g.shard_world < n || g.shard_world > 64) return B200_ERR_ARG;
  int counts[64];
  for (int q = 0; q < g.shard_world; q++) { const long long f = (long long)q * rows_per_rank; counts[q] = f >= n ? 0 : (int)((n - f < rows_per_rank) ? n - f : rows_per_rank); }
  if ((long long)g.shard_rank * rows_per_rank != first && count > 0) return B200_ERR_ARG;
  return upload_rows_impl(counts);
}

extern "C" int b200_download_shard(void *dst, int first, int count) {
  if (!g.ready || !g.have_aos) return B200_ERR_STATE;
  const int n = g.n; const size_t st = (size_t)g.lay.stride;
  if (first < 0 || count < 0 || first + count > n) return B200_ERR_ARG;
  char *out = dst ? (char *)dst : g.h_base;
  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  // only the rows that go down are packed (1/world of the image); the range starts at a multiple of 4 rows so that the
  // 16-byte accesses of aos_block_load/store stay aligned for every stride (packing a row more than once is harmless)
  const int first_al = first & ~3, count_al = first + count - first_al;
  if (count > 0) k_pack_aos<<<cdiv(count_al, kAosRows), kAosRows, (size_t)kAosRows * g.lay.stride, g.stream>>>(first_al, count_al, g.d_aos, to_lay(g.lay), g.posm, g.velh, g.velpred, g.accel, g.dvel,
                                                 g.oldacc, g.gravcost, g.left, g.right, g.ngb, g.maxpred, g.potential);
  count_launch();
  if (!dst && (first < g.h_first || first + count > g.h_first + g.h_count)) return B200_ERR_ARG;
  // `dst` (if given) is a whole-array image; the bound array may hold the rank's rows only (b200_bind_rows)
  if (count > 0) CUDA_TRY(cudaMemcpyAsync(out + (size_t)(first - (dst ? 0 : g.h_first)) * st, g.d_aos + (size_t)first * st, (size_t)count * st, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  cudaEventElapsedTime(&g.cnt.ms_download, g.ev0, g.ev1);
  return B200_OK;
}

// ----------------------------------------------------------------------------- partial transfers (small active sets)
// The reference moves 20 bytes in and 24 bytes out per ACTIVE particle (gravtree.c:149-166, 230-238; allvars.h:547-559).
// Up: what the host driver changes between two force computations - advance() (predict.c:245-345: Pos, Vel, VelPred = Vel,
// dVel = 0, CurrentTime), reflect() (Vel), find_timesteps() (MaxPredTime) - 32 bytes + index per listed particle.
// Down: what the path writes - PosPred VelPred Accel OldAcc GravCost dVel NgbVelDisp HsmlVelDisp Left Right, 72 bytes +
// index - for the listed particles and for every particle that received a partner kick since the last download (the
// partner of a scattering need not be active, sidm.c:559-601).
constexpr int kUpWords = 9, kDownWords = 19;

static int ensure_stage(size_t bytes) {
  if (bytes <= g.stage_cap) return B200_OK;
  if (g.h_stage) cudaFreeHost(g.h_stage);
  if (g.d_stage) cudaFree(g.d_stage);
  g.h_stage = nullptr; g.d_stage = nullptr; g.stage_cap = 0;
  // grow rarely: pinned allocations cost milliseconds and the active sets of successive steps differ in size
  bytes = 4 * bytes + ((size_t)16 << 20);
  const size_t most = (size_t)g.maxpart * kDownWords * 4 + (size_t)g.maxpart * 4 + 4096;
  if (bytes > most) bytes = most;
  if (cudaMallocHost((void **)&g.h_stage, bytes) != cudaSuccess) return B200_ERR_ALLOC;
  if (cudaMalloc((void **)&g.d_stage, bytes) != cudaSuccess) return B200_ERR_ALLOC;
  g.stage_cap = bytes;
  return B200_OK;
}

__global__ void k_scatter_up(int n, const int *rec, int nmax, float *pos0, float4 *posm, float4 *velh, float *velpred, float *dvel,
                             float *curtime, float *maxpred) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int *r = rec + (size_t)a * kUpWords;
  const int i = r[0];
  if (i < 0 || i >= nmax) return;
  float4 v = velh[i];
  for (int k = 0; k < 3; k++) { pos0[3 * (size_t)i + k] = __int_as_float(r[1 + k]); velpred[3 * (size_t)i + k] = __int_as_float(r[4 + k]); dvel[3 * (size_t)i + k] = 0.0f; }
  v.x = __int_as_float(r[4]); v.y = __int_as_float(r[5]); v.z = __int_as_float(r[6]);
  velh[i] = v;
  curtime[i] = __int_as_float(r[7]); maxpred[i] = __int_as_float(r[8]);
}
__global__ void k_gather_down(int n, const int *idx, int nk, const int *kicked, int *rec, const float4 *posm, const float4 *velh, const float *velpred,
                              const float *accel, const float *oldacc, const float *gravcost, const float *dvel, const int *ngb,
                              const float *left, const float *right) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n + nk) return;
  const int i = a < n ? idx[a] : kicked[a - n];
  int *r = rec + (size_t)a * kDownWords;
  const float4 p = posm[i];
  r[0] = i;
  r[1] = __float_as_int(p.x); r[2] = __float_as_int(p.y); r[3] = __float_as_int(p.z);
  for (int k = 0; k < 3; k++) {
    r[4 + k] = __float_as_int(velpred[3 * (size_t)i + k]); r[7 + k] = __float_as_int(accel[3 * (size_t)i + k]); r[12 + k] = __float_as_int(dvel[3 * (size_t)i + k]);
  }
  r[10] = __float_as_int(oldacc[i]); r[11] = __float_as_int(gravcost[i]);
  r[15] = ngb[i]; r[16] = __float_as_int(velh[i].w); r[17] = __float_as_int(left[i]); r[18] = __float_as_int(right[i]);
}

extern "C" int b200_upload_active(const int *idx, int n) {
  if (!g.ready || !g.have_aos || g.n <= 0) return B200_ERR_STATE;
  if (n < 0 || n > g.n || (n > 0 && !idx)) return B200_ERR_ARG;
  if (n == 0) return B200_OK;
  B200_TRY(ensure_stage((size_t)n * kUpWords * 4));
  CUDA_TRY(cudaStreamSynchronize(g.stream));             // the staging buffer may still be in flight
  int *h = (int *)g.h_stage;
  const b200_layout &L = g.lay;
  for (int a = 0; a < n; a++) {
    const int i = idx[a];
    if (i < 0 || i >= g.n) return B200_ERR_ARG;
    const char *p = g.h_base + (size_t)i * L.stride;
    int *r = h + (size_t)a * kUpWords;
    r[0] = i;
    memcpy(r + 1, p + L.Pos, 12); memcpy(r + 4, p + L.Vel, 12); memcpy(r + 7, p + L.CurrentTime, 4);
    if (L.MaxPredTime > 0) memcpy(r + 8, p + L.MaxPredTime, 4); else r[8] = 0;
  }
  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  CUDA_TRY(cudaMemcpyAsync(g.d_stage, h, (size_t)n * kUpWords * 4, cudaMemcpyHostToDevice, g.stream));
  k_scatter_up<<<cdiv(n, 256), 256, 0, g.stream>>>(n, (const int *)g.d_stage, g.n, g.pos0, g.posm, g.velh, g.velpred, g.dvel, g.curtime, g.maxpred);
  count_launch();
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  cudaEventElapsedTime(&g.cnt.ms_upload, g.ev0, g.ev1);
  g.tree_valid = false;
  return B200_OK;
}

extern "C" int b200_download_active(const int *idx, int n, void *dst) {
  if (!g.ready || !g.have_aos || g.n <= 0) return B200_ERR_STATE;
  if (n < 0 || n > g.n || (n > 0 && !idx)) return B200_ERR_ARG;
  char *out = dst ? (char *)dst : g.h_base;
  int nk = 0;
  CUDA_TRY(cudaMemcpyAsync(&nk, g.d_nkick, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  if (nk > g.maxpart) nk = g.maxpart;
  const int tot = n + nk;
  if (tot == 0) return B200_OK;
  B200_TRY(ensure_stage((size_t)tot * kDownWords * 4 + (size_t)n * 4));
  int *d_rec = (int *)g.d_stage, *d_idx = d_rec + (size_t)tot * kDownWords;
  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  if (n > 0) CUDA_TRY(cudaMemcpyAsync(d_idx, idx, (size_t)n * 4, cudaMemcpyHostToDevice, g.stream));
  k_gather_down<<<cdiv(tot, 256), 256, 0, g.stream>>>(n, d_idx, nk, g.kick_list, d_rec, g.posm, g.velh, g.velpred, g.accel, g.oldacc, g.gravcost, g.dvel,
                                                      g.ngb, g.left, g.right);
  count_launch();
  CUDA_TRY(cudaMemcpyAsync(g.h_stage, d_rec, (size_t)tot * kDownWords * 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaMemsetAsync(g.d_nkick, 0, sizeof(int), g.stream));
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  cudaEventElapsedTime(&g.cnt.ms_download, g.ev0, g.ev1);
  const b200_layout &L = g.lay;
  const int *h = (const int *)g.h_stage;
  for (int a = 0; a < tot; a++) {
    const int *r = h + (size_t)a * kDownWords;
    char *p = out + (size_t)r[0] * L.stride;
    memcpy(p + L.PosPred, r + 1, 12); memcpy(p + L.VelPred, r + 4, 12); memcpy(p + L.Accel, r + 7, 12);
    memcpy(p + L.OldAcc, r + 10, 4); memcpy(p + L.GravCost, r + 11, 4); memcpy(p + L.dVel, r + 12, 12);
    memcpy(p + L.NgbVelDisp, r + 15, 4); memcpy(p + L.HsmlVelDisp, r + 16, 4); memcpy(p + L.Left, r + 17, 4); memcpy(p + L.Right, r + 18, 4);
  }
  return B200_OK;
}

// ----------------------------------------------------------------------------- predict

// predict_collisionless_only(), predict.c:106-150: PosPred = Pos + Vel*dt_h0 and
// VelPred = Vel + Accel*dt, float operands, double arithmetic, stored as float.
__global__ void k_predict(int n, double time, double s_a_inverse, const float *pos0, const float *curtime,
                          const float *accel, float4 *posm, const float4 *velh, float *velpred) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double dt = time - (double)curtime[i];
  const double dth = dt * s_a_inverse;
  float4 a = posm[i]; const float4 v = velh[i];
  a.x = (float)((double)pos0[3 * i] + (double)v.x * dth);
  a.y = (float)((double)pos0[3 * i + 1] + (double)v.y * dth);
  a.z = (float)((double)pos0[3 * i + 2] + (double)v.z * dth);
  posm[i] = a;
  velpred[3 * i] = (float)((double)v.x + (double)accel[3 * i] * dt);
  velpred[3 * i + 1] = (float)((double)v.y + (double)accel[3 * i + 1] * dt);
  velpred[3 * i + 2] = (float)((double)v.z + (double)accel[3 * i + 2] * dt);
}

namespace b200 {
double s_a_inverse_at(double time) {   // gravtree.c:42-49
  if (!g.par.ComovingIntegrationOn) return 1.0;
  const double O0 = g.par.Omega0, OL = g.par.OmegaLambda;
  return 1.0 / (g.par.Hubble * sqrt(O0 + time * (1 - O0 - OL) + time * time * time * OL));
}
}

extern "C" int b200_predict(double time) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  k_predict<<<cdiv(g.n, 256), 256, 0, g.stream>>>(g.n, time, s_a_inverse_at(time), g.pos0, g.curtime, g.accel, g.posm, g.velh, g.velpred);
  count_launch();
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  cudaEventElapsedTime(&g.cnt.ms_predict, g.ev0, g.ev1);
  g.tree_valid = false;
  return B200_OK;
}

// ----------------------------------------------------------------------------- advance
// advance(), predict.c:245-345: leap-frog update of the active particles,
//   dt = 2*(time - CurrentTime); Pos += Vel*dt/2; Vel += Accel*dt + dVel; dVel = 0;
//   Pos += Vel*dt/2; CurrentTime = time + dt/2  (dt_h0 = dt/S(a) for the drifts when comoving)
__global__ void k_advance(int na, const int *active, double time, double s_a_inverse, float *pos0, float4 *velh, float *velpred,
                          const float *accel, float *dvel, float *curtime, int *nscat) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= na) return;
  const int i = active ? active[a] : a;
  const double dt = 2 * (time - (double)curtime[i]);
  const double dth = dt * s_a_inverse;
  if (dvel[3 * (size_t)i] != 0.0f) atomicAdd(nscat, 1);
  float4 v = velh[i];
  float vv[3] = {v.x, v.y, v.z};
  for (int j = 0; j < 3; j++) {
    float x = pos0[3 * (size_t)i + j];
    x = (float)((double)x + 0.5 * (double)vv[j] * dth);
    vv[j] = (float)((double)vv[j] + (double)accel[3 * (size_t)i + j] * dt + (double)dvel[3 * (size_t)i + j]);
    velpred[3 * (size_t)i + j] = vv[j];
    dvel[3 * (size_t)i + j] = 0;
    x = (float)((double)x + 0.5 * (double)vv[j] * dth);
    pos0[3 * (size_t)i + j] = x;
  }
  v.x = vv[0]; v.y = vv[1]; v.z = vv[2];
  velh[i] = v;
  curtime[i] = (float)(time + 0.5 * dt);
}

extern "C" int b200_advance(const int *active, int nactive, double time, int *num_scattered) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  const int na = active ? nactive : g.n;
  if (na < 0 || na > g.n) return B200_ERR_ARG;
  if (na == 0) return B200_OK;
  const int *d_act = nullptr;
  if (active) {
    CUDA_TRY(cudaMemcpyAsync(g.d_active, active, (size_t)na * sizeof(int), cudaMemcpyHostToDevice, g.stream));
    d_act = g.d_active;
  }
  CUDA_TRY(cudaMemsetAsync(g.d_flags + FL_NSCATLOG, 0, sizeof(int), g.stream));
  k_advance<<<cdiv(na, 256), 256, 0, g.stream>>>(na, d_act, time, s_a_inverse_at(time), g.pos0, g.velh, g.velpred, g.accel, g.dvel,
                                                 g.curtime, g.d_flags + FL_NSCATLOG);
  count_launch();
  if (num_scattered) {
    CUDA_TRY(cudaMemcpyAsync(g.h_flags + FL_NSCATLOG, g.d_flags + FL_NSCATLOG, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    *num_scattered = g.h_flags[FL_NSCATLOG];
  }
  CUDA_TRY(cudaGetLastError());
  g.tree_valid = false;
  return B200_OK;
}

extern "C" int b200_set_field(const char *name, const void *host, long long nbytes) {
  void *d = nullptr; long long nb = 0;
  B200_TRY(b200_device_buffer(name, &d, &nb));
  if (!host || nbytes <= 0 || nbytes > nb || !d) return B200_ERR_ARG;
  CUDA_TRY(cudaMemcpyAsync(d, host, (size_t)nbytes, cudaMemcpyHostToDevice, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  if (!strcmp(name, "ptype")) { g.types_dirty = true; g.tree_valid = false; g.topo_valid = false; }
  return B200_OK;
}

// ----------------------------------------------------------------------------- reflect
// reflection.c:7-33: float arithmetic throughout, products and sums individually rounded
__global__ void k_reflect(int na, const int *active, float r_ref2, const float *pos0, float4 *velh, int *count) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= na) return;
  const int i = active ? active[a] : a;
  const float x = pos0[3 * (size_t)i], y = pos0[3 * (size_t)i + 1], z = pos0[3 * (size_t)i + 2];
  const float r2 = fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z));
  if (!(r2 > r_ref2)) return;
  float4 v = velh[i];
  const float rv = fadd(fadd(fmul(x, v.x), fmul(y, v.y)), fmul(z, v.z));
  if (!(rv > 0)) return;
  const float r2inv2 = __fdiv_rn(2.0f, r2);
  v.x = fadd(v.x, -fmul(fmul(rv, x), r2inv2)); v.y = fadd(v.y, -fmul(fmul(rv, y), r2inv2)); v.z = fadd(v.z, -fmul(fmul(rv, z), r2inv2));
  velh[i] = v;
  atomicAdd(count, 1);
}
extern "C" int b200_reflect(const int *active, int nactive, double radius, int *num_reflected) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  const int na = active ? nactive : g.n;
  if (na < 0 || na > g.n) return B200_ERR_ARG;
  if (num_reflected) *num_reflected = 0;
  if (na == 0) return B200_OK;
  const int *d_act = nullptr;
  if (active) { CUDA_TRY(cudaMemcpyAsync(g.d_active, active, (size_t)na * sizeof(int), cudaMemcpyHostToDevice, g.stream)); d_act = g.d_active; }
  CUDA_TRY(cudaMemsetAsync(g.d_flags + FL_NSCATLOG, 0, sizeof(int), g.stream));
  k_reflect<<<cdiv(na, 256), 256, 0, g.stream>>>(na, d_act, (float)(radius * radius), g.pos0, g.velh, g.d_flags + FL_NSCATLOG);
  count_launch();
  CUDA_TRY(cudaMemcpyAsync(g.h_flags + FL_NSCATLOG, g.d_flags + FL_NSCATLOG, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  if (num_reflected) *num_reflected = g.h_flags[FL_NSCATLOG];
  return B200_OK;
}

// ----------------------------------------------------------------------------- find_timesteps
// timestep.c:17-334 for collisionless particles.  Types as in the C code: Accel, CurrentTime,
// MaxPredTime, HsmlVelDisp, Mass are floats, |a|^2 and CurrentTime + MaxPredTime are formed in float,
// everything else in double.  One thread per active particle.
struct TsParams {
  int na; const int *active; int mode, crit, comoving; double time, s_a, hubble_a, a3inv, C_max, C_Grho, G;
  double eta, velscale, probtol, dyntol, dtmax, dtmin; double soft[6];
  const float *accel, *curtime; float *maxpred; const float4 *velh, *posm; const int *ptype;
  const double *jitter; uint32_t k0, k1; float *out; int *nclamped;
};
__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void k_find_timesteps(TsParams T) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= T.na) return;
  const int i = T.active ? T.active[a] : a;
  const float a0 = T.accel[3 * (size_t)i], a1 = T.accel[3 * (size_t)i + 1], a2 = T.accel[3 * (size_t)i + 2];
  const double ac = sqrt((double)fadd(fadd(fmul(a0, a0), fmul(a1, a1)), fmul(a2, a2)));      // timestep.c:138-140
  const float ct = T.curtime[i], mp = T.maxpred[i];
  const double dtold = 2 * ((double)fadd(ct, mp) - 2 * T.time);                                // :142
  double dt;
  if (T.crit == 0) dt = sqrt(2 * T.eta * T.soft[T.ptype[i] & 7] / ac * T.s_a);                // :156
  else dt = T.velscale / ac;                                                                  // :159
  {                                                                                           // :247-265 (SIDM)
    const double h = (double)T.velh[i].w, hinv = 1.0 / h, hinv3 = hinv * hinv * hinv;
    const double m = (double)T.posm[i].w;
    const double dt_sidm = T.probtol / (T.C_max * m * hinv3);
    if (dt_sidm < dt) dt = dt_sidm;
    double dt_Grho;
    if (T.comoving) dt_Grho = T.dyntol * T.hubble_a * T.time / sqrt(T.C_Grho * T.G * m * hinv3 * T.a3inv);
    else dt_Grho = T.dyntol / sqrt(T.C_Grho * T.G * m * hinv3);
    if (dt_Grho < dt) dt = dt_Grho;
  }
  if (dt > 1.3 * dtold && T.mode != 2) dt = 1.3 * dtold;                                      // TIMESTEP_INCREASE_FACTOR, :268-272
  bool clamped = false;
  double u = 0;
  if (dt >= T.dtmax || dt < T.dtmin) {
    clamped = true;
    u = T.jitter ? T.jitter[a] : (double)mix32((uint32_t)i * 0x9E3779B9u ^ T.k0 ^ mix32(T.k1)) / 4294967296.0;
  }
  if (dt >= T.dtmax) dt = T.dtmax * (1.00 + 0.02 * u);                                        // :274-283
  if (dt < T.dtmin) { dt = T.dtmin; dt *= 1.0 + 0.02 * u; }                                   // :285-310
  const float np = (float)((double)ct + 0.5 * dt);                                            // :315
  T.maxpred[i] = np;
  if (T.out) T.out[a] = np;
  if (clamped) atomicAdd(T.nclamped, 1);
}

extern "C" int b200_find_timesteps(const int *active, int nactive, int mode, double time, double vmax,
                                   const b200_timestep_params *tp, const double *jitter, float *maxpred_out, int *num_clamped) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  if (!tp || (tp->TypeOfTimestepCriterion != 0 && tp->TypeOfTimestepCriterion != 1)) return B200_ERR_ARG;
  const int na = active ? nactive : g.n;
  if (na < 0 || na > g.n) return B200_ERR_ARG;
  if (na == 0) { if (num_clamped) *num_clamped = 0; return B200_OK; }
  TsParams T;
  T.na = na; T.active = nullptr; T.mode = mode; T.crit = tp->TypeOfTimestepCriterion; T.comoving = g.par.ComovingIntegrationOn;
  T.time = time; T.G = g.par.G;
  if (active) { CUDA_TRY(cudaMemcpyAsync(g.d_active, active, (size_t)na * sizeof(int), cudaMemcpyHostToDevice, g.stream)); T.active = g.d_active; }
  // constants of timestep.c:45-131
  const double ball = (3. / 4. / 3.14159265358979323846) * (g.par.DesNumNgb + g.par.MaxNumNgbDeviation);
  const int X = g.par.CrossSectionType;
  const double sig = g.par.CrossSectionInternal;
  double vc = g.par.YukawaVelocity, s_a = 1, a3inv = 1, hubble_a = 0;
  double C_max;
  if (T.comoving) {                     // timestep.c:47-93, operations in the reference's order
    hubble_a = g.par.Hubble * sqrt(g.par.Omega0 / pow(time, 3) + (1 - g.par.Omega0 - g.par.OmegaLambda) / pow(time, 2) + g.par.OmegaLambda);
    s_a = g.par.Hubble * sqrt(g.par.Omega0 + time * (1 - g.par.Omega0 - g.par.OmegaLambda) + pow(time, 3) * g.par.OmegaLambda);
    a3inv = 1 / (time * time * time);
    if (X == 1) C_max = 1.0 * ball * sig / pow(time, 2.5) / s_a;
    else if (X == 2) {
      vc = g.par.YukawaVelocity / sqrt(time);
      if (2.0 * vmax < vc / sqrt(3.0)) { const double beta = 2.0 * vmax / vc, v_dep = 1.0 / (1.0 + beta * beta); C_max = 1.0 * ball * 2.0 * vmax * v_dep * v_dep * sig / pow(time, 2) / s_a; }
      else C_max = 1.0 * ball * (3.0 * sqrt(3.0) / 16.0) * vc * sig / pow(time, 2) / s_a;
    } else if (X == 3) C_max = 1.0 * ball * sig / pow(time, 2.0) * 2 * g.par.CrossSectionVelScale / s_a;
    else C_max = 1.0 * ball * 2 * vmax * sig / pow(time, 2) / s_a;
  } else {                              // timestep.c:95-130
    if (X == 1) C_max = 1.0 * ball * sig;
    else if (X == 2) {
      // NB this branch of the reference uses v_dep = 1/(1 + 2 vmax/vc), not 1/(1 + beta^2) (timestep.c:108)
      if (2.0 * vmax < vc / sqrt(3.0)) { const double v_dep = 1.0 / (1.0 + 2.0 * vmax / vc); C_max = 1.0 * ball * 2.0 * vmax * v_dep * v_dep * sig; }
      else C_max = 1.0 * ball * (3.0 * sqrt(3.0) / 16.0) * vc * sig;
    } else if (X == 3) C_max = 1.0 * ball * 2 * g.par.CrossSectionVelScale * sig;
    else C_max = 1.0 * ball * 2 * vmax * sig;
  }
  T.C_max = C_max; T.C_Grho = ball; T.s_a = s_a; T.hubble_a = hubble_a; T.a3inv = a3inv;
  T.eta = tp->ErrTolIntAccuracy; T.velscale = tp->ErrTolVelScale; T.probtol = tp->ProbabilityTol; T.dyntol = tp->ErrTolDynamicalAccuracy;
  T.dtmax = tp->MaxSizeTimestep; T.dtmin = tp->MinSizeTimestep;
  for (int t = 0; t < 6; t++) T.soft[t] = g.par.SofteningTable[t];
  T.accel = g.accel; T.curtime = g.curtime; T.maxpred = g.maxpred; T.velh = g.velh; T.posm = g.posm; T.ptype = g.ptype;
  const unsigned long long calls = ++g.ts_calls;
  T.k0 = (uint32_t)(g.par.Seed ^ (calls * 0x9E3779B97F4A7C15ull)); T.k1 = (uint32_t)(calls >> 7) ^ 0x51ED270Bu;
  T.jitter = nullptr; T.out = nullptr; T.nclamped = g.d_flags + FL_NSCATLOG;
  if (jitter) { CUDA_TRY(cudaMemcpyAsync(g.d_acc, jitter, (size_t)na * sizeof(double), cudaMemcpyHostToDevice, g.stream)); T.jitter = g.d_acc; }
  if (maxpred_out) T.out = (float *)g.d_cost;
  CUDA_TRY(cudaMemsetAsync(T.nclamped, 0, sizeof(int), g.stream));
  k_find_timesteps<<<cdiv(na, 256), 256, 0, g.stream>>>(T);
  count_launch();
  if (maxpred_out) CUDA_TRY(cudaMemcpyAsync(maxpred_out, g.d_cost, (size_t)na * sizeof(float), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaMemcpyAsync(g.h_flags + FL_NSCATLOG, g.d_flags + FL_NSCATLOG, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  if (num_clamped) *num_clamped = g.h_flags[FL_NSCATLOG];
  return B200_OK;
}

// ----------------------------------------------------------------------------- getvmax

// sidm.c:970-990: max |Vel| over local particles; v2 is formed in float (float products and
// sums), compared in double.
__global__ void k_vmax2(int n, const float4 *velh, float *out) {
  __shared__ float sm[256];
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 v = velh[i];
    const float v2 = fadd(fadd(fmul(v.x, v.x), fmul(v.y, v.y)), fmul(v.z, v.z));
    m = fmaxf(m, v2);
  }
  sm[threadIdx.x] = m; __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + s]); __syncthreads(); }
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

extern "C" int b200_getvmax(double *vmax) {
  if (!g.ready || g.n <= 0 || !vmax) return B200_ERR_STATE;
  const int G = 296;
  float *part = (float *)g.d_cost;
  k_vmax2<<<G, 256, 0, g.stream>>>(g.n, g.velh, part);
  count_launch();
  float h[296];
  CUDA_TRY(cudaMemcpyAsync(h, part, G * sizeof(float), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  float m = 0; for (int i = 0; i < G; i++) m = fmaxf(m, h[i]);
  *vmax = sqrt((double)m);
  return B200_OK;
}

// ----------------------------------------------------------------------------- global quantities
// compute_global_quantities_of_system(), global.c:18-135.  12 sums per particle type: mass, E_kin, E_pot,
// momentum[3], angular momentum[3], mass-weighted position[3].  One pass per type present (one in every BASELINE
// config); per-thread double accumulators over a grid-stride loop, warp shuffle + shared-memory block reduction,
// per-block partials summed in block order by the host: deterministic, no atomics.  Streaming, 48 B per particle.
constexpr int kGQ = 12, kGQBlocks = 296;
__global__ void __launch_bounds__(256) k_global_quantities(int n, int type, const float4 *posm, const float *velpred, const float *potential,
                                                           const int *ptype, double *part) {
  double a[kGQ];
  for (int k = 0; k < kGQ; k++) a[k] = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if ((ptype[i] & 7) != type) continue;
    const float4 p = posm[i];
    const float vx = velpred[3 * (size_t)i], vy = velpred[3 * (size_t)i + 1], vz = velpred[3 * (size_t)i + 2];
    const float m = p.w;
    a[0] += (double)m;                                                           // global.c:33
    a[2] += 0.5 * (double)m * (double)potential[i];                              // :35
    const float v2 = fadd(fadd(fmul(vx, vx), fmul(vy, vy)), fmul(vz, vz));       // float expression, :40-42
    a[1] += 0.5 * (double)m * (double)v2;
    a[3] += (double)fmul(m, vx); a[4] += (double)fmul(m, vy); a[5] += (double)fmul(m, vz);          // :46
    a[9] += (double)fmul(m, p.x); a[10] += (double)fmul(m, p.y); a[11] += (double)fmul(m, p.z);     // :47
    a[6] += (double)fmul(m, fadd(fmul(p.y, vz), -fmul(p.z, vy)));                // :50-55
    a[7] += (double)fmul(m, fadd(fmul(p.z, vx), -fmul(p.x, vz)));
    a[8] += (double)fmul(m, fadd(fmul(p.x, vy), -fmul(p.y, vx)));
  }
  __shared__ double sm[8][kGQ];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = 0; k < kGQ; k++) {
    double v = a[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[w][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kGQ) {
    double v = 0;
    for (int q = 0; q < 8; q++) v += sm[q][threadIdx.x];
    part[(size_t)blockIdx.x * kGQ + threadIdx.x] = v;
  }
}

extern "C" int b200_compute_global_quantities(b200_sysstate *out) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  if (!out) return B200_ERR_ARG;
  memset(out, 0, sizeof(*out));
  double *part = g.d_acc;                                   // scratch: 296 x 12 doubles
  static double h[kGQBlocks * kGQ];
  for (int t = 0; t < 5; t++) {                             // types 0..4 as global.c:24 (type 5 is not summed there either)
    if (!g.types_dirty && g.type_count[t] == 0) continue;
    k_global_quantities<<<kGQBlocks, 256, 0, g.stream>>>(g.n, t, g.posm, g.velpred, g.potential, g.ptype, part);
    count_launch();
    CUDA_TRY(cudaMemcpyAsync(h, part, sizeof(h), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    double s[kGQ];
    for (int k = 0; k < kGQ; k++) { s[k] = 0; for (int b = 0; b < kGQBlocks; b++) s[k] += h[b * kGQ + k]; }
    out->MassComp[t] = s[0]; out->EnergyKinComp[t] = s[1]; out->EnergyPotComp[t] = s[2];
    for (int j = 0; j < 3; j++) { out->MomentumComp[t][j] = s[3 + j]; out->AngMomentumComp[t][j] = s[6 + j]; out->CenterOfMassComp[t][j] = s[9 + j]; }
  }
  CUDA_TRY(cudaGetLastError());
  // totals, centre-of-mass normalisation and vector norms, global.c:71-130
  for (int i = 0; i < 5; i++) {
    out->EnergyTotComp[i] = out->EnergyKinComp[i] + out->EnergyPotComp[i] + out->EnergyIntComp[i];
    out->Mass += out->MassComp[i]; out->EnergyKin += out->EnergyKinComp[i]; out->EnergyPot += out->EnergyPotComp[i];
    out->EnergyInt += out->EnergyIntComp[i]; out->EnergyTot += out->EnergyTotComp[i];
    for (int j = 0; j < 3; j++) {
      out->Momentum[j] += out->MomentumComp[i][j]; out->AngMomentum[j] += out->AngMomentumComp[i][j];
      out->CenterOfMass[j] += out->CenterOfMassComp[i][j];
    }
  }
  for (int i = 0; i < 5; i++) for (int j = 0; j < 3; j++) if (out->MassComp[i] > 0) out->CenterOfMassComp[i][j] /= out->MassComp[i];
  for (int j = 0; j < 3; j++) if (out->Mass > 0) out->CenterOfMass[j] /= out->Mass;
  auto norm3 = [](double *v) { v[3] = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
  for (int i = 0; i < 5; i++) { norm3(out->CenterOfMassComp[i]); norm3(out->MomentumComp[i]); norm3(out->AngMomentumComp[i]); }
  norm3(out->CenterOfMass); norm3(out->Momentum); norm3(out->AngMomentum);
  return B200_OK;
}

// ----------------------------------------------------------------------------- counters / buffers

extern "C" int b200_get_counters(b200_counters *c) {
  if (!c) return B200_ERR_ARG;
  g.cnt.num_nodes = g.num_nodes; g.cnt.max_level = g.max_level;
  *c = g.cnt;
  return B200_OK;
}

extern "C" int b200_device_buffer(const char *name, void **dptr, long long *nbytes) {
  if (!g.ready || !name || !dptr || !nbytes) return B200_ERR_ARG;
  const long long n = g.n;
  struct { const char *nm; void *p; long long b; } tab[] = {
      {"posm", g.posm, n * 16}, {"velh", g.velh, n * 16}, {"accel", g.accel, n * 12}, {"dvel", g.dvel, n * 12},
      {"oldacc", g.oldacc, n * 4}, {"ngb", g.ngb, n * 4}, {"acc_raw", g.d_acc, n * 24}, {"cost", g.d_cost, n * 8},
      {"velpred", g.velpred, n * 12}, {"pos0", g.pos0, n * 12}, {"curtime", g.curtime, n * 4},
      {"gravcost", g.gravcost, n * 4}, {"left", g.left, n * 4}, {"right", g.right, n * 4}, {"maxpred", g.maxpred, n * 4}, {"potential", g.potential, n * 4}, {"ptype", g.ptype, n * 4}, {"pid", g.pid, n * 4},
      {"sidx", g.sidx, n * 4}, {"ewald", g.d_ewald, g.d_ewald ? 33LL * 33 * 33 * 16 : 0}};
  for (auto &t : tab) if (!strcmp(t.nm, name)) { *dptr = t.p; *nbytes = t.b; return B200_OK; }
  return B200_ERR_ARG;
}
