// snapshot.cu - savepositions() (io.c:16-590): GADGET snapshot file format 1 written straight
// from the device state.  The blocks the file wants (PosPred, VelPred, ID, Mass in particle-TYPE order,
// io.c:267-300) are gathered on the device - a stable 3-bit sort by type gives the order, one kernel fills the
// four blocks - and streamed to the file through two pinned staging buffers (the copy of chunk c+1 runs while
// chunk c is in fwrite), instead of packing the 124-byte AoS, downloading all of it and picking 32 bytes per
// particle out of it on the host.  Fortran-style 4-byte record markers as io.c:207-210,262,578.
#include <cub/cub.cuh>

#include <cstdio>
#include <cstring>

#include "ctx.cuh"
#include "tree_logic.h"

namespace b200 {

// struct io_header_1, allvars.h:727-746 (256 bytes)
struct SnapHeader {
  int npart[6]; double mass[6]; double time, redshift; int flag_sfr, flag_feedback; int npartTotal[6];
  int flag_cooling, num_files; double BoxSize, Omega0, OmegaLambda, HubbleParam;
  int flag_multiphase, flag_stellarage, flag_sfrhistogram; char fill[84];
};
static_assert(sizeof(SnapHeader) == 256, "io_header_1 is 256 bytes");

// hist[0..6]: types of the rows [first, first+count); hist[8..14]: types of all n rows (header npartTotal of a multi-file snapshot)
__global__ void k_snap_hist(int n, int first, int count, const int *ptype, int *hist) {
  __shared__ int sh[16];
  if (threadIdx.x < 16) sh[threadIdx.x] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = ptype[i] & 7;
    atomicAdd(&sh[8 + t], 1);
    if (i >= first && i < first + count) atomicAdd(&sh[t], 1);
  }
  __syncthreads();
  if (threadIdx.x < 16 && (threadIdx.x & 7) != 7 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void k_snap_keys(int count, int first, const int *ptype, unsigned char *key, int *iota) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) { key[i] = (unsigned char)(ptype[first + i] & 7); iota[i] = first + i; }
}
// slot j of the file order holds particle perm[j] (perm == nullptr: identity, one type)
__global__ void k_snap_gather(int nout, const int *perm, const float4 *posm, const float *velpred, const int *pid,
                              float *pos, float *vel, int *id, float *mass, int periodic, double box, int first) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nout) return;
  const int i = perm ? perm[j] : first + j;
  const float4 p = posm[i];
  float x[3] = {p.x, p.y, p.z};
  if (periodic) {                                            // io.c:275-283
    for (int k = 0; k < 3; k++) {
      while (x[k] < 0) x[k] = (float)((double)x[k] + box);
      while ((double)x[k] > box) x[k] = (float)((double)x[k] - box);
    }
  }
  for (int k = 0; k < 3; k++) { pos[3 * (size_t)j + k] = x[k]; vel[3 * (size_t)j + k] = velpred[3 * (size_t)i + k]; }
  id[j] = pid[i]; mass[j] = p.w;
}

// read_ic() + the start-up loop of init() (read_ic.c:32-481, init.c:76-100) for one format-1 file without gas:
// particle j of the file is of the type whose block range holds j; its mass is the MassTable entry of the type or
// the next entry of the mass block; PosPred = Pos, VelPred = Vel, Accel = dVel = OldAcc = Potential = 0, GravCost = 1.
struct SnapTypes { int cum[7]; int moff[6]; float mtab[6]; };
__global__ void k_snap_scatter(int n, SnapTypes T, const float *pos, const float *vel, const int *id, const float *mass, float time,
                               float4 *posm, float4 *velh, float *pos0, float *velpred, int *pid, int *ptype, float *accel, float *dvel,
                               float *curtime, float *oldacc, float *gravcost, float *left, float *right, int *ngb, float *maxpred,
                               float *potential, int dst0) {
  const int jf = blockIdx.x * blockDim.x + threadIdx.x;      // particle of the file
  if (jf >= n) return;
  int t = 0;
  while (t < 5 && jf >= T.cum[t + 1]) t++;
  const float m = T.moff[t] >= 0 ? mass[T.moff[t] + (jf - T.cum[t])] : T.mtab[t];
  const float x = pos[3 * (size_t)jf], y = pos[3 * (size_t)jf + 1], z = pos[3 * (size_t)jf + 2];
  const float vx = vel[3 * (size_t)jf], vy = vel[3 * (size_t)jf + 1], vz = vel[3 * (size_t)jf + 2];
  const int idj = id[jf];
  const int j = dst0 + jf;                                   // its place in the particle order (files one after the other)
  posm[j] = make_float4(x, y, z, m); velh[j] = make_float4(vx, vy, vz, 0.f);
  pos0[3 * (size_t)j] = x; pos0[3 * (size_t)j + 1] = y; pos0[3 * (size_t)j + 2] = z;
  velpred[3 * (size_t)j] = vx; velpred[3 * (size_t)j + 1] = vy; velpred[3 * (size_t)j + 2] = vz;
  for (int k = 0; k < 3; k++) { accel[3 * (size_t)j + k] = 0.f; dvel[3 * (size_t)j + k] = 0.f; }
  pid[j] = idj; ptype[j] = t;
  curtime[j] = time; maxpred[j] = time; oldacc[j] = 0.f; gravcost[j] = 1.f; left[j] = 0.f; right[j] = 0.f; ngb[j] = 0; potential[j] = 0.f;
}

namespace {
constexpr size_t kChunk = 32u << 20;
struct Stager {                         // two pinned staging buffers, kept between calls (page-locking 64 MB costs ~30 ms)
  char *pin[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr};
  void release() {
    for (int k = 0; k < 2; k++) { if (pin[k]) cudaFreeHost(pin[k]); if (ev[k]) cudaEventDestroy(ev[k]); pin[k] = nullptr; ev[k] = nullptr; }
  }
  int init() {
    for (int k = 0; k < 2; k++) {
      if (!pin[k] && cudaHostAlloc((void **)&pin[k], kChunk, cudaHostAllocDefault) != cudaSuccess) { pin[k] = nullptr; return B200_ERR_ALLOC; }
      if (!ev[k] && cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) { ev[k] = nullptr; return B200_ERR_CUDA; }
    }
    return B200_OK;
  }
  // device bytes -> file, double buffered
  int write(FILE *fd, const void *d_src, size_t bytes) {
    const size_t nch = (bytes + kChunk - 1) / kChunk;
    auto issue = [&](size_t c) -> cudaError_t {
      const size_t off = c * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
      cudaError_t e = cudaMemcpyAsync(pin[c & 1], (const char *)d_src + off, len, cudaMemcpyDeviceToHost, g.stream);
      return e != cudaSuccess ? e : cudaEventRecord(ev[c & 1], g.stream);
    };
    if (nch > 0) CUDA_TRY(issue(0));
    for (size_t c = 0; c < nch; c++) {
      if (c + 1 < nch) CUDA_TRY(issue(c + 1));
      CUDA_TRY(cudaEventSynchronize(ev[c & 1]));
      const size_t off = c * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
      if (fwrite(pin[c & 1], 1, len, fd) != len) return B200_ERR_IO;           // my_fwrite(), io.c:594-605
    }
    return B200_OK;
  }
};
// file bytes -> device, double buffered (the fread of chunk c+1 runs while chunk c is on the bus)
struct Loader {
  Stager &sg; explicit Loader(Stager &s) : sg(s) {}
  int read(FILE *fd, void *d_dst, size_t bytes) {
    const size_t nch = (bytes + kChunk - 1) / kChunk;
    for (size_t c = 0; c < nch; c++) {
      const size_t off = c * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
      if (c >= 2) CUDA_TRY(cudaEventSynchronize(sg.ev[c & 1]));           // the copy that last used this buffer
      if (fread(sg.pin[c & 1], 1, len, fd) != len) return B200_ERR_IO;     // my_fread(), io.c:611-622
      CUDA_TRY(cudaMemcpyAsync((char *)d_dst + off, sg.pin[c & 1], len, cudaMemcpyHostToDevice, g.stream));
      CUDA_TRY(cudaEventRecord(sg.ev[c & 1], g.stream));
    }
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return B200_OK;
  }
};
struct DevTmp {
  void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  ~DevTmp() { for (auto q : p) if (q) cudaFree(q); }
};
struct FileCloser { FILE *fd; ~FileCloser() { if (fd) fclose(fd); } };
Stager g_stager;
}  // namespace

void snapshot_release() { g_stager.release(); }      // b200_finalize

}  // namespace b200

using namespace b200;

extern "C" int b200_savepositions(const char *path, double time, const double *mass_table, double hubble_param,
                                  int *npart_out) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  return b200_savepositions_part(path, time, mass_table, hubble_param, 0, g.n, 1, npart_out);
}

// One file of a snapshot split over `num_files` files (io.c:90-103, 127-160): the rows [first, first+count) of the particle
// order in type order, header npart = this file's counts, npartTotal = the whole system's.
extern "C" int b200_savepositions_part(const char *path, double time, const double *mass_table, double hubble_param,
                                       int first, int count, int num_files, int *npart_out) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  if (!path || first < 0 || count < 0 || first + count > g.n || num_files < 1) return B200_ERR_ARG;
  const int nall = g.n, n = count, B = 256, G = cdiv(n > 0 ? n : 1, B);
  cudaStream_t st = g.stream;
  // particles per type (io.c:107-111), of this file and of the system
  int cnt[16];
  int *d_hist = (int *)g.d_bbox;                             // 8 doubles of scratch = 16 ints (free outside the tree build)
  CUDA_TRY(cudaMemsetAsync(d_hist, 0, 16 * sizeof(int), st));
  k_snap_hist<<<296, B, 0, st>>>(nall, first, count, g.ptype, d_hist);
  count_launch();
  CUDA_TRY(cudaMemcpyAsync(cnt, d_hist, 16 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (cnt[8] > 0) return B200_ERR_ARG;                       // gas blocks (u, rho, hsml of SphP) are not on this path
  long long ntot = 0, nmass = 0; int ntypes = 0;
  for (int t = 0; t < 5; t++) {                              // type 5 is not written (io.c:265)
    ntot += cnt[t]; if (cnt[t] > 0) ntypes++;
    if (!mass_table || mass_table[t] == 0) nmass += cnt[t];  // io.c:121-123
  }
  if (npart_out) for (int t = 0; t < 6; t++) npart_out[t] = t < 5 ? cnt[t] : 0;
  const bool identity = (ntypes == 1 && ntot == n) || n == 0;   // one type: the rows as they are
  DevTmp tmp;
  const int *perm = nullptr;
  if (!identity) {
    unsigned char *key, *key2; int *iota, *order; void *cubtmp;
    if (cudaMalloc(&tmp.p[0], (size_t)n) != cudaSuccess || cudaMalloc(&tmp.p[1], (size_t)n) != cudaSuccess ||
        cudaMalloc(&tmp.p[2], (size_t)n * sizeof(int)) != cudaSuccess || cudaMalloc(&tmp.p[3], (size_t)n * sizeof(int)) != cudaSuccess)
      return B200_ERR_ALLOC;
    key = (unsigned char *)tmp.p[0]; key2 = (unsigned char *)tmp.p[1]; iota = (int *)tmp.p[2]; order = (int *)tmp.p[3];
    k_snap_keys<<<G, B, 0, st>>>(n, first, g.ptype, key, iota);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key2, iota, order, n, 0, 3, st);
    if (cudaMalloc(&tmp.p[4], tb) != cudaSuccess) return B200_ERR_ALLOC;
    cubtmp = tmp.p[4];
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(cubtmp, tb, key, key2, iota, order, n, 0, 3, st));   // stable: particle order inside a type
    count_launch(4);
    perm = order;                                            // types 0..4 first, then 5.. : the first ntot entries are the file order
  }
  float *d_pos = (float *)g.d_acc, *d_vel = d_pos + 3 * (size_t)n;
  int *d_id = g.d_cost; float *d_mass = (float *)(g.d_cost + n);
  const int periodic = (g.par.PeriodicBoundariesOn && g.par.BoxSize > 0) ? 1 : 0;
  if (ntot > 0) {
    k_snap_gather<<<cdiv(ntot, B), B, 0, st>>>((int)ntot, perm, g.posm, g.velpred, g.pid, d_pos, d_vel, d_id, d_mass, periodic, g.par.BoxSize, first);
    count_launch();
  }
  CUDA_TRY(cudaGetLastError());

  SnapHeader h;
  memset(&h, 0, sizeof(h));
  for (int t = 0; t < 5; t++) { h.npart[t] = cnt[t]; h.npartTotal[t] = cnt[8 + t]; }        // io.c:140-160
  for (int t = 0; t < 6; t++) h.mass[t] = mass_table ? mass_table[t] : 0.0;
  h.time = time;
  h.redshift = g.par.ComovingIntegrationOn ? 1.0 / time - 1 : 0;                    // io.c:168-171
  h.num_files = num_files;
  h.BoxSize = g.par.BoxSize; h.Omega0 = g.par.Omega0; h.OmegaLambda = g.par.OmegaLambda; h.HubbleParam = hubble_param;

  FileCloser fc{fopen(path, "w")};
  FILE *fd = fc.fd;
  if (!fd) return B200_ERR_IO;                                // io.c:98-102 endrun(10)
  Stager &sg = g_stager;
  B200_TRY(sg.init());
  auto marker = [&](long long bytes) -> int { const int d = (int)bytes; return fwrite(&d, sizeof(d), 1, fd) == 1 ? B200_OK : B200_ERR_IO; };
  B200_TRY(marker(256));
  if (fwrite(&h, sizeof(h), 1, fd) != 1) return B200_ERR_IO;
  B200_TRY(marker(256));
  struct { const void *src; long long bytes; } blk[3] = {{d_pos, 12 * ntot}, {d_vel, 12 * ntot}, {d_id, 4 * ntot}};
  for (auto &b : blk) {
    if ((int)b.bytes > 0) B200_TRY(marker(b.bytes));         // io.c:261-262: the marker is an int4byte, written if > 0
    B200_TRY(sg.write(fd, b.src, (size_t)b.bytes));
    if ((int)b.bytes > 0) B200_TRY(marker(b.bytes));
  }
  // masses of the types without a MassTable entry, type after type (io.c:299, 371-376)
  if ((int)(4 * nmass) > 0) B200_TRY(marker(4 * nmass));
  long long off = 0;
  for (int t = 0; t < 5; t++) {
    if (!mass_table || mass_table[t] == 0) B200_TRY(sg.write(fd, d_mass + off, (size_t)cnt[t] * 4));
    off += cnt[t];
  }
  if ((int)(4 * nmass) > 0) B200_TRY(marker(4 * nmass));
  // u, rho, hsml: sizeof(float)*ntot_type[0] == 0 -> nothing written (io.c:254-262)
  fc.fd = nullptr;
  if (fclose(fd) != 0) return B200_ERR_IO;
  return B200_OK;
}

// one file of the (possibly split) snapshot into the particle order at dst0; returns the header
static int load_one_file(const char *path, int dst0, SnapHeader &h, int *n_file) {
  FileCloser fc{fopen(path, "r")};
  FILE *fd = fc.fd;
  if (!fd) return B200_ERR_IO;
  auto marker = [&](long long expect) -> int {              // SKIP of read_ic.c:35, checked
    int d = 0;
    if (fread(&d, sizeof(d), 1, fd) != 1) return B200_ERR_IO;
    return d == (int)expect ? B200_OK : B200_ERR_IO;
  };
  B200_TRY(marker(256));
  if (fread(&h, sizeof(h), 1, fd) != 1) return B200_ERR_IO;
  B200_TRY(marker(256));
  if (h.npart[0] > 0) return B200_ERR_ARG;                  // gas blocks are not on this path
  SnapTypes T;
  long long ntot = 0, nmass = 0;
  for (int t = 0; t < 6; t++) {
    if (h.npart[t] < 0) return B200_ERR_IO;
    T.cum[t] = (int)ntot; T.mtab[t] = (float)h.mass[t];
    T.moff[t] = (h.mass[t] == 0 && h.npart[t] > 0) ? (int)nmass : -1;      // read_ic.c:126,260
    if (h.mass[t] == 0) nmass += h.npart[t];
    ntot += h.npart[t];
  }
  T.cum[6] = (int)ntot;
  *n_file = (int)ntot;
  if (ntot < 0 || dst0 + ntot > g.maxpart) return B200_ERR_ARG;
  if (ntot == 0) return B200_OK;
  const int n = (int)ntot;
  float *d_pos = (float *)g.d_acc, *d_vel = d_pos + 3 * (size_t)n;
  int *d_id = g.d_cost; float *d_mass = (float *)(g.d_cost + n);
  Stager &sg = g_stager;
  B200_TRY(sg.init());
  Loader ld(sg);
  B200_TRY(marker(12 * ntot)); B200_TRY(ld.read(fd, d_pos, 12 * (size_t)ntot)); B200_TRY(marker(12 * ntot));
  B200_TRY(marker(12 * ntot)); B200_TRY(ld.read(fd, d_vel, 12 * (size_t)ntot)); B200_TRY(marker(12 * ntot));
  B200_TRY(marker(4 * ntot));  B200_TRY(ld.read(fd, d_id, 4 * (size_t)ntot));   B200_TRY(marker(4 * ntot));
  if (nmass > 0) { B200_TRY(marker(4 * nmass)); B200_TRY(ld.read(fd, d_mass, 4 * (size_t)nmass)); B200_TRY(marker(4 * nmass)); }
  k_snap_scatter<<<cdiv(n, 256), 256, 0, g.stream>>>(n, T, d_pos, d_vel, d_id, d_mass, (float)h.time, g.posm, g.velh, g.pos0, g.velpred,
                                                     g.pid, g.ptype, g.accel, g.dvel, g.curtime, g.oldacc, g.gravcost, g.left, g.right,
                                                     g.ngb, g.maxpred, g.potential, dst0);
  count_launch();
  CUDA_TRY(cudaStreamSynchronize(g.stream));                // the staging scratch is reused by the next file
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// `path` itself, or - a snapshot split over several files (read_ic.c:62-75) - path.0, path.1, ... (header num_files), the
// files' particles one after the other in the particle order (as the tasks of the reference hold them)
extern "C" int b200_load_snapshot(const char *path, double *time_out, double *mass_table_out, int *npart_out) {
  if (!g.ready) return B200_ERR_STATE;
  if (!path) return B200_ERR_ARG;
  SnapHeader h;
  int ntot = 0, nf = 0, files = 1;
  int npart[6] = {0, 0, 0, 0, 0, 0};
  FILE *probe = fopen(path, "r");
  if (probe) {
    fclose(probe);
    B200_TRY(load_one_file(path, 0, h, &nf));
    if (h.num_files > 1) return B200_ERR_ARG;               // one part of a split snapshot: name the base path instead
    ntot = nf;
    for (int t = 0; t < 6; t++) npart[t] = h.npart[t];
  } else {
    for (int k = 0; k < files; k++) {
      char name[1024];
      if (snprintf(name, sizeof(name), "%s.%d", path, k) >= (int)sizeof(name)) return B200_ERR_ARG;
      B200_TRY(load_one_file(name, ntot, h, &nf));
      if (k == 0) files = h.num_files > 1 ? h.num_files : 1;
      ntot += nf;
      for (int t = 0; t < 6; t++) npart[t] += h.npart[t];
    }
  }
  if (ntot <= 0) return B200_ERR_ARG;
  g.n = ntot;
  g.tree_valid = false; g.types_dirty = true; g.topo_valid = false;
  if (time_out) *time_out = h.time;
  if (mass_table_out) for (int t = 0; t < 6; t++) mass_table_out[t] = h.mass[t];
  if (npart_out) for (int t = 0; t < 6; t++) npart_out[t] = npart[t];
  return B200_OK;
}
