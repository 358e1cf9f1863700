// tree_build.cu - force_treebuild() on the GPU (reference: forcetree.c:90-422, 433-571).
//
// The reference inserts particles one at a time into a pointer octree and then loops over
// all particles below every node for its moments.  Here every step is data-parallel:
//   bbox reduce -> root cell -> per-particle 42-level octant keys (float-exact geometry,
//   tree_logic.h) -> radix sort (CUB) -> shared-prefix lengths -> scan -> nodes in
//   depth-first pre-order -> per-node ranges/children by binary search -> leaf-ordered
//   particle copy -> bottom-up moments by level.
// HBM traffic per particle (SURVEY.md section 8d "build bytes"): 16 B read + 16 B key write,
// 8 radix passes over (8 B key + 4 B index), 16 B gather into leaf order; per node 64 B
// record + 80 B double moments scratch.
#include <stdlib.h>
#include <cub/cub.cuh>
#include <cooperative_groups.h>
#include "ctx.cuh"
#include "build_logic.h"

namespace b200 {

// ------------------------------------------------------------------ bounding box
// only_type >= 0: bounding box of the particles of that type (one tree per type)
__global__ void k_bbox_partial(int n, const float4 *posm, const int *ptype, float *part, int *flags, int only_type) {
  __shared__ float sm[6][256];
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  const int t0 = only_type >= 0 ? only_type : (ptype[0] & 7);
  bool multi = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const bool mine = (ptype[i] & 7) == t0;
    multi |= !mine;
    if (only_type >= 0 && !mine) continue;
    const float4 p = posm[i];
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
  if (multi && only_type < 0) flags[FL_MULTITYPE] = 1;
  for (int k = 0; k < 3; k++) { sm[k][threadIdx.x] = mn[k]; sm[3 + k][threadIdx.x] = mx[k]; }
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; k++) {
        sm[k][threadIdx.x] = fminf(sm[k][threadIdx.x], sm[k][threadIdx.x + s]);
        sm[3 + k][threadIdx.x] = fmaxf(sm[3 + k][threadIdx.x], sm[3 + k][threadIdx.x + s]);
      }
    __syncthreads();
  }
  if (threadIdx.x < 6) part[6 * blockIdx.x + threadIdx.x] = sm[threadIdx.x][0];
}

__global__ void k_bbox_final(int nblk, const float *part, double *bbox, RootBox *root, float *domain) {
  // single warp; forcetree.c:179-212: extent in double from float coordinates
  const int k = threadIdx.x;
  if (k < 6) {
    float v = part[k];
    for (int b = 1; b < nblk; b++) v = (k < 3) ? fminf(v, part[6 * b + k]) : fmaxf(v, part[6 * b + k]);
    bbox[k] = (double)v;
    domain[k] = v;           // DomainMin/DomainMax of this type (forcetree.c:192-198)
  }
  __syncwarp();
  if (k == 0) {
    double mn[3] = {bbox[0], bbox[1], bbox[2]}, mx[3] = {bbox[3], bbox[4], bbox[5]};
    *root = make_root(mn, mx);
  }
}

// ------------------------------------------------------------------ keys
__global__ void k_keys(int n, const float4 *posm, const RootBox *root, uint64_t *hi, uint64_t *lo, int *iota, const int *ptype) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RootBox rb = ptype ? root[ptype[i] & 7] : root[0];    // several types: the key is the path through the particle's own tree
  const float4 p = posm[i];
  uint64_t h, l;
  make_key(p.x, p.y, p.z, rb, h, l);
  hi[i] = h; lo[i] = l; iota[i] = i;
}

__global__ void k_gather_lo(int n, const int *sidx, const uint64_t *lo, uint64_t *slo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) slo[j] = lo[sidx[j]];
}

// particles whose first 21 octants agree are ordered by the next 21 (rare: a few pairs in a
// 1e7-particle cusp); one thread per run of equal high words, insertion sort on the low word
__global__ void k_fix_ties(int n, const uint64_t *shi, uint64_t *slo, int *sidx, const unsigned char *stype) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n - 1) return;
  auto same = [&](int a, int b) { return shi[a] == shi[b] && (!stype || stype[a] == stype[b]); };   // runs never cross a type boundary
  if ((j > 0 && same(j - 1, j)) || !same(j + 1, j)) return;   // not the start of a run
  int e = j + 1;
  while (e + 1 < n && same(e + 1, j)) e++;
  for (int a = j + 1; a <= e; a++) {
    const uint64_t l = slo[a]; const int s = sidx[a];
    int b = a - 1;
    while (b >= j && (slo[b] > l || (slo[b] == l && sidx[b] > s))) { slo[b + 1] = slo[b]; sidx[b + 1] = sidx[b]; b--; }
    slo[b + 1] = l; sidx[b + 1] = s;
  }
}

// several particle types: after the key sort, a stable 3-bit sort by type puts the trees one after the other
// census of the particle types.  Warp ballots + a block histogram: one global atomic per type and block (1e7 particles of ONE
// type adding to one address took 6 ms, every step that follows a full upload)
__global__ void k_type_hist(int n, const int *ptype, int *hist) {
  __shared__ int sh[8];
  if (threadIdx.x < 8) sh[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = i < n ? (ptype[i] & 7) : -1;
  for (int k = 0; k < 8; k++) {
    const unsigned b = __ballot_sync(0xffffffffu, t == k);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&sh[k], __popc(b));
  }
  __syncthreads();
  if (threadIdx.x < 8 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void k_type_keys(int n, const int *sidx, const int *ptype, unsigned char *tkey, int *pos) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) { tkey[j] = (unsigned char)(ptype[sidx[j]] & 7); pos[j] = j; }
}
__global__ void k_type_gather(int n, const int *perm, const int *sidx_in, const uint64_t *shi_in, int *sidx_out, uint64_t *shi_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) { const int p = perm[j]; sidx_out[j] = sidx_in[p]; shi_out[j] = shi_in[p]; }
}
__global__ void k_troots(int ntrees, const int *first, const int *nodestart, int *flags) {
  const int t = threadIdx.x;
  if (t < ntrees) flags[FL_TROOT0 + t] = nodestart[first[t]];
}

// ------------------------------------------------------------------ construction kernels
__global__ void k_b1(BuildView v, int *cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n) return;
  int c = b1_common(v, i);
  if (c >= kMaxLevels) { v.flags[FL_ERR_COINCIDENT] = 1; c = kMaxLevels - 1; }
  int cp = (i > 0) ? b1_common(v, i - 1) : -1;
  if (cp >= kMaxLevels) cp = kMaxLevels - 1;
  v.clev[i] = (signed char)c;
  cnt[i] = b1_count_from(cp, c, i, v.n);
}
__global__ void k_b2(BuildView v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n) return;
  b2_body(v, i);
  if (i == 0) v.flags[FL_NUM_NODES] = v.nodestart[v.n];
  const int c = v.clev[i];
  if (c >= 0) atomicMax(&v.flags[FL_MAX_LEVEL], c);
}
__global__ void k_b3(BuildView v) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = min(v.nodestart[v.n], v.maxnodes);
  if (id < m) b3_body(v, id);
}
__global__ void k_np(BuildView v, int *np32) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = min(v.nodestart[v.n], v.maxnodes);
  if (id <= m) np32[id] = id < m ? (int)v.nnp[id] : 0;
}
__global__ void k_b4(BuildView v) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = min(v.nodestart[v.n], v.maxnodes);
  if (id < m) b4_body(v, id);
}
__global__ void k_b5(BuildView v, int level) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = min(v.nodestart[v.n], v.maxnodes);
  if (id < m && v.nlevel[id] == level) b5_body(v, id);
}
// Moments in ONE launch instead of one per level: threads start at the cells without child cells and
// climb; a cell is finished by the thread of its last-arriving child (atomic arrival counter), which reads
// the children's moments past the L1 (they were written by other SMs in this same launch).  The sums run
// over the children in octant order whatever the arrival order, so the result equals the per-level version.
__device__ __forceinline__ void b5_body_cg(const BuildView &v, int id) {
  const float4 gm = v.geom[id];
  Moments m; moments_zero(m);
  const int np = v.nnp[id], ps = v.npstart[id];
  for (int k = 0; k < np; k++) {
    const float4 p = v.leaf_posm[ps + k];
    moments_add_particle(m, p.x, p.y, p.z, p.w, gm.x, gm.y, gm.z);
  }
  int mn = v.nminidx[id];
  const int end = v.nodes[id].skip;
  for (int c = id + 1; c < end; c = v.nodes[c].skip) {     // child cells, in octant order
    const float4 cg = v.geom[c];
    Moments cm;
    const double *src = reinterpret_cast<const double *>(v.nmom + c);
    double *dst = reinterpret_cast<double *>(&cm);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Moments) / sizeof(double)); k++) dst[k] = __ldcg(src + k);
    moments_add_child(m, cm, (double)cg.x - (double)gm.x, (double)cg.y - (double)gm.y, (double)cg.z - (double)gm.z);
    const int cmn = __ldcg(v.nminidx + c);
    if (cmn < mn) mn = cmn;
  }
  v.nmom[id] = m;
  v.nminidx[id] = mn;
  NodeRec r = v.nodes[id];
  moments_finish(m, gm.x, gm.y, gm.z, gm.w, r);
  v.nodes[id] = r;
}
__global__ void k_b5_up(BuildView v) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = min(v.nodestart[v.n], v.maxnodes);
  if (id >= m || v.nnchild[id] != 0) return;
  int cur = id;
  for (;;) {
    b5_body_cg(v, cur);
    const int par = v.nparent[cur];
    if (par < 0) break;
    __threadfence();                                        // this cell's moments before the arrival count
    if (atomicAdd(&v.narrive[par], 1) + 1 < (int)v.nnchild[par]) break;
    __threadfence();
    cur = par;                                              // last child to arrive: finish the parent
  }
}
// Moments level by level in ONE cooperative launch: the cells are listed by level (6-bit radix sort of the pre-order ids), the
// grid works through the levels from the deepest up with a grid-wide barrier in between, a cell reads the finished moments of
// its child cells.  No arrival counters, no fences, no uncached loads; same sums in the same (octant) order as b5_body.
__global__ void k_level_hist(int m, const unsigned char *nlevel, int *hist) {
  __shared__ int sh[64];
  if (threadIdx.x < 64) sh[threadIdx.x] = 0;
  __syncthreads();
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < m) atomicAdd(&sh[nlevel[id] & 63], 1);
  __syncthreads();
  if (threadIdx.x < 64 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void k_level_offsets(int *hist) {          // counts [64] -> exclusive offsets [65], in place
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int at = 0;
  for (int l = 0; l < 64; l++) { const int c = hist[l]; hist[l] = at; at += c; }
  hist[64] = at;
}
template <bool REFIT>
__global__ void __launch_bounds__(256) k_b5_levels(BuildView v, const int *lev_ids, const int *lev_off) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  for (int lev = kMaxLevels; lev >= 0; lev--) {
    const int a = lev_off[lev], b = lev_off[lev + 1];
    if (b == a) continue;                                   // the same for every thread: no barrier needed
    for (int k = a + gt; k < b; k += gs) b5_body_t<REFIT>(v, lev_ids[k]);
    grid.sync();
  }
}
__global__ void k_b6(BuildView v, int level) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = min(v.nodestart[v.n], v.maxnodes);
  if (id < m && v.nlevel[id] == level) b6_body(v, id);
}

// ------------------------------------------------------------------ sibling-pair records of the packed walk
// Child cells of node a occupy the slots 2 + gbase[a] .. of the record array (two slots per 128-byte PairRec), gbase =
// exclusive scan of the child-cell counts rounded up to even; the root sits alone in pair 0.  A padding slot repeats
// its sibling's position with zero mass / oc / bmax2 / len2: it is never opened and the walk does not count it.
__global__ void k_pair_gsize(int m, const unsigned char *nnchild, int *gsize) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id <= m) gsize[id] = id < m ? ((nnchild[id] + 1) & ~1) : 0;
}
__global__ void k_pairs(int m, const NodeRec *nodes, const int *nparent, const unsigned char *nnchild, const int *gbase, PairRec *pairs) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= m) return;
  const int par = nparent[id];
  int slot = 0; bool pad = true;
  if (par >= 0) {
    int r = 0;
    for (int c = par + 1; c != id; c = nodes[c].skip) r++;       // rank among the parent's child cells (octant order)
    slot = 2 + gbase[par] + r;
    const int kp = nnchild[par];
    pad = (r == kp - 1) && (kp & 1);
  }
  const NodeRec n = nodes[id];
  const int kc = nnchild[id];
  const int cinfo = kc ? ((((2 + gbase[id]) >> 1) << 4) | kc) : 0;
  float *f = pairs[slot >> 1].f + (slot & 1);
  f[0] = n.sx; f[2] = n.sy; f[4] = n.sz; f[6] = n.mass;
  f[8] = n.oc; f[10] = n.bmax2; f[12] = __int_as_float(cinfo); f[14] = __int_as_float(n.pinfo);
  f[16] = fmul(-3.0f, n.q11); f[18] = fmul(-3.0f, n.q22); f[20] = fmul(-3.0f, n.q33); f[22] = fmul(-3.0f, n.q12);
  f[24] = fmul(-3.0f, n.q13); f[26] = fmul(-3.0f, n.q23); f[28] = fmul(-1.5f, n.p); f[30] = n.len2;
  if (pad) {
    f[1] = n.sx; f[3] = n.sy; f[5] = n.sz;
    for (int c = 3; c < 16; c++) f[2 * c + 1] = 0.0f;
  }
}

static int ensure_cub(size_t bytes) {
  if (bytes <= g.cub_tmp_bytes) return B200_OK;
  if (g.cub_tmp) cudaFree(g.cub_tmp);
  g.cub_tmp = nullptr; g.cub_tmp_bytes = 0;
  bytes = bytes + bytes / 4 + 4096;
  if (cudaMalloc(&g.cub_tmp, bytes) != cudaSuccess) return B200_ERR_ALLOC;
  g.cub_tmp_bytes = bytes;
  return B200_OK;
}

BuildView make_view() {
  BuildView v;
  v.n = g.n; v.maxnodes = g.maxnodes; v.posm = g.posm; v.shi = g.skey_hi; v.slo = g.skey_lo; v.sidx = g.sidx;
  v.clev = g.clev; v.nodestart = g.nodestart; v.root = g.d_root; v.stype = g.ntypes > 1 ? g.stype : nullptr;
  v.nodes = g.nodes; v.geom = g.geom; v.nstart = g.nstart; v.nend = g.nend; v.nparent = g.nparent; v.npstart = g.npstart;
  v.nlevel = g.nlevel; v.nnp = g.nnp; v.nnchild = g.nnchild; v.ndp = g.ndp; v.narrive = g.narrive;
  v.nminidx = g.nminidx; v.nlstart = g.nlstart; v.nmom = g.nmom;
  v.leaf_posm = g.leaf_posm; v.leaf_orig = g.leaf_orig; v.orig_leaf = g.orig_leaf; v.krank = g.krank; v.lrank = g.lrank; v.leaf_parent = g.leaf_parent;
  v.flags = g.d_flags;
  return v;
}

// ------------------------------------------------------------------ refit (option "tree_reuse")
// The reference does not rebuild its tree at every step either (gravtree.c:63-96, TreeUpdateFrequency): it lets the cells drift
// with their centre-of-mass velocity and re-moments lazily (forcetree.c:935-954, 2486-2549).  Here: between two full builds the
// topology stays (key order, cells, leaf order, level lists, search records, query groups) and the leaves and ALL moments are
// recomputed from the current predicted positions, so every cell has its exact mass, centre of mass and quadrupole; what ages is
// only which cell a particle is filed under.  Neighbour searches stay exact: they widen their cell tests by the largest coordinate
// displacement since the full build (d_pad), the sphere test uses the current positions.
__global__ void k_refit_leaves(int n, const int *leaf_orig, const float4 *posm, float4 *leaf_posm, float *padstep) {
  __shared__ float sm[256];
  float d = 0.f;
  for (int L = blockIdx.x * blockDim.x + threadIdx.x; L < n; L += gridDim.x * blockDim.x) {
    const float4 p = posm[leaf_orig[L]], o = leaf_posm[L];
    d = fmaxf(d, fmaxf(fabsf(p.x - o.x), fmaxf(fabsf(p.y - o.y), fabsf(p.z - o.z))));
    leaf_posm[L] = p;
  }
  sm[threadIdx.x] = d; __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + s]); __syncthreads(); }
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<int *>(padstep), __float_as_int(sm[0]));     // non-negative floats order like ints
}
__global__ void k_pad_accumulate(float *pad) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { pad[0] = (pad[0] + pad[1]) * 1.000001f + 1.0e-30f; pad[1] = 0.f; }   // triangle inequality, rounded up
}

static int launch_moments(const BuildView &v, int m, cudaStream_t st, bool lists_ready) {
  int *hist = g.lev_off, *lev_ids = g.sidx_tmp;             // sidx_tmp: scratch of the key sort, free between builds
  if (!lists_ready) {
    unsigned char *lkeys = (unsigned char *)g.key_tmp;
    CUDA_TRY(cudaMemsetAsync(hist, 0, 72 * sizeof(int), st));
    k_level_hist<<<cdiv(m, 256), 256, 0, st>>>(m, g.nlevel, hist);
    k_level_offsets<<<1, 32, 0, st>>>(hist);
    size_t tbl = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tbl, g.nlevel, lkeys, g.iota, lev_ids, m, 0, 6, st);
    B200_TRY(ensure_cub(tbl));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(g.cub_tmp, tbl, g.nlevel, lkeys, g.iota, lev_ids, m, 0, 6, st));
    count_launch(5);
  }
  const bool refit = v.next != nullptr;
  void *kernel = refit ? (void *)k_b5_levels<true> : (void *)k_b5_levels<false>;
  static int coop_blocks[2] = {0, 0};
  if (!coop_blocks[refit]) {
    int per_sm = 0, dev = 0, sms = 0;
    if (refit) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_b5_levels<true>, 256, 0));
    else CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_b5_levels<false>, 256, 0));
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    coop_blocks[refit] = per_sm * sms;
  }
  const int *cl = lev_ids, *co = hist;
  void *args[] = {(void *)&v, (void *)&cl, (void *)&co};
  CUDA_TRY(cudaLaunchCooperativeKernel(kernel, dim3(coop_blocks[refit]), dim3(256), args, 0, st));
  count_launch();
  return B200_OK;
}

static int tree_refit_impl() {
  const int n = g.n;
  cudaStream_t st = g.stream;
  CUDA_TRY(cudaEventRecord(g.ev0, st));
  // DomainMin/Max of the current positions (sidm.c:141-161 flags the particles near it); the root cell stays as built
  k_bbox_partial<<<296, 256, 0, st>>>(n, g.posm, g.ptype, (float *)g.d_cost, g.d_flags, -1);
  k_bbox_final<<<1, 32, 0, st>>>(296, (float *)g.d_cost, g.d_bbox, g.d_root + 7, g.d_domain);
  k_refit_leaves<<<592, 256, 0, st>>>(n, g.leaf_orig, g.posm, g.leaf_posm, g.d_pad + 1);
  k_pad_accumulate<<<1, 32, 0, st>>>(g.d_pad);
  count_launch(4);
  BuildView v = make_view();
  v.next = (float *)g.narrive;                              // per-cell extent (scratch of the build's arrival counters)
  B200_TRY(launch_moments(v, g.num_nodes, st, true));
  CUDA_TRY(cudaEventRecord(g.ev1, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  cudaEventElapsedTime(&g.cnt.ms_build, g.ev0, g.ev1);
  g.tree_valid = true; g.refits_since_build++;
  return B200_OK;
}

int tree_build_impl() {
  const int n = g.n;
  const int B = 256, G = cdiv(n, B);
  cudaStream_t st = g.stream;
  // tree reuse: a refit of the last topology instead of a build (one particle type, open boundaries, tree-order neighbour scans)
  if (g.opt_tree_reuse > 1 && g.topo_valid && g.refits_since_build + 1 < g.opt_tree_reuse && g.ntypes == 1 && !g.types_dirty &&
      !(g.par.PeriodicBoundariesOn && g.par.BoxSize > 0) && !g.par.ReferenceNgbOrder && !g.opt_walk_pairs)
    return tree_refit_impl();
  CUDA_TRY(cudaEventRecord(g.ev0, st));
  CUDA_TRY(cudaMemsetAsync(g.d_flags, 0, FL_COUNT * sizeof(int), st));

  // which particle types are present (one tree per type, forcetree.c:90-158)
  if (g.types_dirty) {
    int *hist = g.d_flags + FL_TROOT0;        // scratch: overwritten below
    CUDA_TRY(cudaMemsetAsync(hist, 0, 8 * sizeof(int), st));
    k_type_hist<<<G, B, 0, st>>>(n, g.ptype, hist);
    CUDA_TRY(cudaMemcpyAsync(g.h_flags + FL_TROOT0, hist, 6 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    g.ntypes = 0;
    for (int t = 0; t < 6; t++) { g.type_count[t] = g.h_flags[FL_TROOT0 + t]; if (g.type_count[t] > 0) g.ntypes++; }
    g.types_dirty = false;
    count_launch();
    CUDA_TRY(cudaMemsetAsync(g.d_flags, 0, FL_COUNT * sizeof(int), st));
  }
  const bool multi = g.ntypes > 1;
  if (multi && !g.stype && cudaMalloc((void **)&g.stype, (size_t)g.maxpart + 256) != cudaSuccess) return B200_ERR_ALLOC;
  // 1. bounding box -> root cell (forcetree.c:179-212), per type when there are several
  const int GB = 296;
  float *part = (float *)g.d_cost;
  if (!multi) {
    k_bbox_partial<<<GB, 256, 0, st>>>(n, g.posm, g.ptype, part, g.d_flags, -1);
    k_bbox_final<<<1, 32, 0, st>>>(GB, part, g.d_bbox, g.d_root, g.d_domain);
    count_launch(2);
  } else {
    for (int t = 0; t < 6; t++) if (g.type_count[t] > 0) {
      k_bbox_partial<<<GB, 256, 0, st>>>(n, g.posm, g.ptype, part, g.d_flags, t);
      k_bbox_final<<<1, 32, 0, st>>>(GB, part, g.d_bbox, g.d_root + t, g.d_domain + 6 * t);
      count_launch(2);
    }
  }
  // 2. keys + sort
  k_keys<<<G, B, 0, st>>>(n, g.posm, g.d_root, g.key_hi, g.key_lo, g.iota, multi ? g.ptype : nullptr);
  count_launch();
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, g.key_hi, g.skey_hi, g.iota, g.sidx, n, 0, 63, st);
  B200_TRY(ensure_cub(tb));
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(g.cub_tmp, tb, g.key_hi, g.skey_hi, g.iota, g.sidx, n, 0, 63, st));
  count_launch(9);
  if (multi) {
    // stable sort of the key order by type: tree after tree, each in its own key order
    unsigned char *tkey = g.stype, *tkey2 = (unsigned char *)g.clev;          // clev is written later (k_b1)
    int *pos = g.sidx_tmp, *perm = g.krank;                                     // krank is written later (k_b2)
    k_type_keys<<<G, B, 0, st>>>(n, g.sidx, g.ptype, tkey, pos);
    size_t tbt = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tbt, tkey, tkey2, pos, perm, n, 0, 3, st);
    B200_TRY(ensure_cub(tbt));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(g.cub_tmp, tbt, tkey, tkey2, pos, perm, n, 0, 3, st));
    k_type_gather<<<G, B, 0, st>>>(n, perm, g.sidx, g.skey_hi, g.sidx_tmp, g.key_tmp);
    CUDA_TRY(cudaMemcpyAsync(g.sidx, g.sidx_tmp, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g.skey_hi, g.key_tmp, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(g.stype, tkey2, (size_t)n, cudaMemcpyDeviceToDevice, st));
    count_launch(4);
  }
  k_gather_lo<<<G, B, 0, st>>>(n, g.sidx, g.key_lo, g.skey_lo);
  k_fix_ties<<<G, B, 0, st>>>(n, g.skey_hi, g.skey_lo, g.sidx, multi ? g.stype : nullptr);
  count_launch(2);
  // 3. prefix lengths, node counts, scan
  BuildView v = make_view();
  int *cnt = g.sidx_tmp;
  k_b1<<<G, B, 0, st>>>(v, cnt);
  size_t tb2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb2, cnt, g.nodestart, n + 1, st);
  B200_TRY(ensure_cub(tb2));
  CUDA_TRY(cudaMemsetAsync(cnt + n, 0, sizeof(int), st));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(g.cub_tmp, tb2, cnt, g.nodestart, n + 1, st));
  k_b2<<<G, B, 0, st>>>(v);
  count_launch(4);
  // tree roots (node ids) in node order
  g.ntrees = 0;
  {
    int first[6], at = 0;
    for (int t = 0; t < 6; t++) if (g.type_count[t] > 0) { first[g.ntrees] = at; g.tree_type[g.ntrees] = t; g.ntrees++; at += g.type_count[t]; }
    if (multi) {
      int *d_first = (int *)g.d_bbox;          // 8 doubles of scratch, free again by now
      CUDA_TRY(cudaMemcpyAsync(d_first, first, g.ntrees * sizeof(int), cudaMemcpyHostToDevice, st));
      k_troots<<<1, 32, 0, st>>>(g.ntrees, d_first, g.nodestart, g.d_flags);
      count_launch();
    }
  }
  // node count / depth / error flags back to the host (the only sync of the build)
  CUDA_TRY(cudaMemcpyAsync(g.h_flags, g.d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (g.h_flags[FL_MULTITYPE] && !multi) { g.types_dirty = true; return B200_ERR_TYPES; }   // stale type census
  if (g.h_flags[FL_ERR_COINCIDENT]) return B200_ERR_COINCIDENT;
  const int m = g.h_flags[FL_NUM_NODES];
  if (m >= g.maxnodes) {   // forcetree.c:233-239
    fprintf(stderr, "libsidm_b200: maximum number %d of tree-nodes reached (need %d)\n", g.maxnodes, m);
    return B200_ERR_NODES;
  }
  g.num_nodes = m; g.max_level = g.h_flags[FL_MAX_LEVEL];
  if (multi) { for (int t = 0; t < g.ntrees; t++) g.tree_root[t] = g.h_flags[FL_TROOT0 + t]; }
  else { g.ntrees = 1; g.tree_root[0] = 0; }
  g.tree_root[g.ntrees] = m;
  const int GM = cdiv(m + 1, B);
  // 4. ranges, geometry, children
  k_b3<<<GM, B, 0, st>>>(v);
  int *np32 = g.narrive;   // reuse as 32-bit copy of nnp for the scan (narrive is reset below)
  k_np<<<GM, B, 0, st>>>(v, np32);
  size_t tb3 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb3, np32, g.npstart, m + 1, st);
  B200_TRY(ensure_cub(tb3));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(g.cub_tmp, tb3, np32, g.npstart, m + 1, st));
  k_b4<<<GM, B, 0, st>>>(v);
  count_launch(5);
  // 5. moments, children before parents
  static const bool per_level = getenv("B200_MOMENTS_PER_LEVEL") != nullptr;     // the one-launch-per-level form, kept for A/B
  static const bool climb = getenv("B200_MOMENTS_CLIMB") != nullptr;             // the last-arriver climb of round 1, kept for A/B
  if (per_level) { for (int lev = g.max_level; lev >= 0; lev--) k_b5<<<GM, B, 0, st>>>(v, lev); count_launch(g.max_level + 1); }
  else if (!climb) {
    B200_TRY(launch_moments(v, m, st, false));
  } else {
    CUDA_TRY(cudaMemsetAsync(g.narrive, 0, (size_t)(m + 1) * sizeof(int), st));
    k_b5_up<<<cdiv(m + 1, 128), 128, 0, st>>>(v);
    count_launch();
  }
  // 5b. sibling-pair records for the packed walk (one tree, open boundaries)
  g.pairs_valid = false;
  if (g.opt_walk_pairs && !multi && !(g.par.PeriodicBoundariesOn && g.par.BoxSize > 0)) {
    int *gsize = g.narrive;                      // free again after the moments
    k_pair_gsize<<<GM, B, 0, st>>>(m, g.nnchild, gsize);
    size_t tb4 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb4, gsize, g.gbase, m + 1, st);
    B200_TRY(ensure_cub(tb4));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(g.cub_tmp, tb4, gsize, g.gbase, m + 1, st));
    k_pairs<<<cdiv(m, B), B, 0, st>>>(m, g.nodes, g.nparent, g.nnchild, g.gbase, g.pairs);
    count_launch(4);
    g.pairs_valid = true;
  }
  // 6. the reference's next[] chain order (only needed to scan neighbours in its order)
  if (g.par.ReferenceNgbOrder) {
    for (int lev = 0; lev <= g.max_level; lev++) k_b6<<<GM, B, 0, st>>>(v, lev);
    count_launch(g.max_level + 1);
  }
  CUDA_TRY(cudaEventRecord(g.ev1, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  cudaEventElapsedTime(&g.cnt.ms_build, g.ev0, g.ev1);
  g.tree_valid = true; g.tree_epoch++;
  g.topo_valid = true; g.refits_since_build = 0;
  CUDA_TRY(cudaMemsetAsync(g.d_pad, 0, 2 * sizeof(float), st));
  return B200_OK;
}

}  // namespace b200
using namespace b200;

extern "C" int b200_tree_build(void) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  return tree_build_impl();
}

__global__ void k_tree_dump(int m, const NodeRec *nodes, const float4 *geom, const int *nstart, const int *nend,
                            const unsigned char *nlevel, float *center, float *len, float *mass, float *s, float *Q,
                            float *oc, float *bmax2, int *count, int *level) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= m) return;
  const NodeRec r = nodes[id]; const float4 gm = geom[id];
  center[3 * id] = gm.x; center[3 * id + 1] = gm.y; center[3 * id + 2] = gm.z; len[id] = gm.w;
  mass[id] = r.mass; s[3 * id] = r.sx; s[3 * id + 1] = r.sy; s[3 * id + 2] = r.sz;
  float *q = Q + 7 * id;
  q[0] = r.q11; q[1] = r.q22; q[2] = r.q33; q[3] = r.q12; q[4] = r.q13; q[5] = r.q23; q[6] = r.p;
  oc[id] = r.oc; bmax2[id] = r.bmax2; count[id] = nend[id] - nstart[id] + 1; level[id] = nlevel[id];
}

extern "C" int b200_get_tree(int *num_nodes, float *center, float *len, float *mass, float *s, float *Q,
                             float *oc, float *bmax2, int *count, int *level) {
  if (!g.ready || !g.tree_valid) return B200_ERR_STATE;
  const int m = g.num_nodes;
  if (num_nodes) *num_nodes = m;
  if (!center && !len && !mass && !s && !Q && !oc && !bmax2 && !count && !level) return B200_OK;
  // dump into one scratch allocation, then copy out what was asked for
  float *d = nullptr;
  const size_t per = 3 + 1 + 1 + 3 + 7 + 1 + 1 + 1 + 1;
  if (cudaMalloc((void **)&d, (size_t)m * per * 4) != cudaSuccess) return B200_ERR_ALLOC;
  float *dc = d, *dl = dc + 3 * (size_t)m, *dm = dl + m, *ds = dm + m, *dq = ds + 3 * (size_t)m, *doc = dq + 7 * (size_t)m,
        *db = doc + m; int *dcnt = (int *)(db + m), *dlev = dcnt + m;
  k_tree_dump<<<cdiv(m, 256), 256, 0, g.stream>>>(m, g.nodes, g.geom, g.nstart, g.nend, g.nlevel, dc, dl, dm, ds, dq, doc, db, dcnt, dlev);
  count_launch();
  auto out = [&](void *h, const void *dv, size_t b) { if (h) cudaMemcpyAsync(h, dv, b, cudaMemcpyDeviceToHost, g.stream); };
  out(center, dc, 12 * (size_t)m); out(len, dl, 4 * (size_t)m); out(mass, dm, 4 * (size_t)m); out(s, ds, 12 * (size_t)m);
  out(Q, dq, 28 * (size_t)m); out(oc, doc, 4 * (size_t)m); out(bmax2, db, 4 * (size_t)m); out(count, dcnt, 4 * (size_t)m);
  out(level, dlev, 4 * (size_t)m);
  cudaError_t e = cudaStreamSynchronize(g.stream);
  cudaFree(d);
  if (e != cudaSuccess) { g.last_cuda = (int)e; return B200_ERR_CUDA; }
  return B200_OK;
}
