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

__global__ void k_snap_hist(int n, const int *ptype, int *hist) {
  __shared__ int sh[8];
  if (threadIdx.x < 8) sh[threadIdx.x] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&sh[ptype[i] & 7], 1);
  __syncthreads();
  if (threadIdx.x < 7 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);   // 7 words of scratch
}
__global__ void k_snap_keys(int n, const int *ptype, unsigned char *key, int *iota) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { key[i] = (unsigned char)(ptype[i] & 7); iota[i] = i; }
}
// slot j of the file order holds particle perm[j] (perm == nullptr: identity, one type)
__global__ void k_snap_gather(int nout, const int *perm, const float4 *posm, const float *velpred, const int *pid,
                              float *pos, float *vel, int *id, float *mass, int periodic, double box) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nout) return;
  const int i = perm ? perm[j] : j;
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

namespace {
constexpr size_t kChunk = 32u << 20;
struct Stager {
  char *pin[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr};
  ~Stager() { for (int k = 0; k < 2; k++) { if (pin[k]) cudaFreeHost(pin[k]); if (ev[k]) cudaEventDestroy(ev[k]); } }
  int init() {
    for (int k = 0; k < 2; k++) {
      if (cudaHostAlloc((void **)&pin[k], kChunk, cudaHostAllocDefault) != cudaSuccess) return B200_ERR_ALLOC;
      if (cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) return B200_ERR_CUDA;
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
struct DevTmp {
  void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  ~DevTmp() { for (auto q : p) if (q) cudaFree(q); }
};
struct FileCloser { FILE *fd; ~FileCloser() { if (fd) fclose(fd); } };
}  // namespace

}  // namespace b200

using namespace b200;

extern "C" int b200_savepositions(const char *path, double time, const double *mass_table, double hubble_param,
                                  int *npart_out) {
  if (!g.ready || g.n <= 0) return B200_ERR_STATE;
  if (!path) return B200_ERR_ARG;
  const int n = g.n, B = 256, G = cdiv(n, B);
  cudaStream_t st = g.stream;
  // particles per type (io.c:107-111)
  int cnt[8];
  CUDA_TRY(cudaMemsetAsync(g.d_flags + FL_TROOT0, 0, 7 * sizeof(int), st));        // 7 scratch words; type & 7 == 7 lands in none
  k_snap_hist<<<296, B, 0, st>>>(n, g.ptype, g.d_flags + FL_TROOT0);
  count_launch();
  CUDA_TRY(cudaMemcpyAsync(cnt, g.d_flags + FL_TROOT0, 7 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (cnt[0] > 0) return B200_ERR_ARG;                       // gas blocks (u, rho, hsml of SphP) are not on this path
  long long ntot = 0, nmass = 0; int ntypes = 0;
  for (int t = 0; t < 5; t++) {                              // type 5 is not written (io.c:265)
    ntot += cnt[t]; if (cnt[t] > 0) ntypes++;
    if (!mass_table || mass_table[t] == 0) nmass += cnt[t];  // io.c:121-123
  }
  if (npart_out) for (int t = 0; t < 6; t++) npart_out[t] = t < 5 ? cnt[t] : 0;
  const bool identity = (ntypes == 1 && ntot == n);
  DevTmp tmp;
  const int *perm = nullptr;
  if (!identity) {
    unsigned char *key, *key2; int *iota, *order; void *cubtmp;
    if (cudaMalloc(&tmp.p[0], (size_t)n) != cudaSuccess || cudaMalloc(&tmp.p[1], (size_t)n) != cudaSuccess ||
        cudaMalloc(&tmp.p[2], (size_t)n * sizeof(int)) != cudaSuccess || cudaMalloc(&tmp.p[3], (size_t)n * sizeof(int)) != cudaSuccess)
      return B200_ERR_ALLOC;
    key = (unsigned char *)tmp.p[0]; key2 = (unsigned char *)tmp.p[1]; iota = (int *)tmp.p[2]; order = (int *)tmp.p[3];
    k_snap_keys<<<G, B, 0, st>>>(n, g.ptype, key, iota);
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
    k_snap_gather<<<cdiv(ntot, B), B, 0, st>>>((int)ntot, perm, g.posm, g.velpred, g.pid, d_pos, d_vel, d_id, d_mass, periodic, g.par.BoxSize);
    count_launch();
  }
  CUDA_TRY(cudaGetLastError());

  SnapHeader h;
  memset(&h, 0, sizeof(h));
  for (int t = 0; t < 5; t++) { h.npart[t] = cnt[t]; h.npartTotal[t] = cnt[t]; }
  for (int t = 0; t < 6; t++) h.mass[t] = mass_table ? mass_table[t] : 0.0;
  h.time = time;
  h.redshift = g.par.ComovingIntegrationOn ? 1.0 / time - 1 : 0;                    // io.c:168-171
  h.num_files = 1;
  h.BoxSize = g.par.BoxSize; h.Omega0 = g.par.Omega0; h.OmegaLambda = g.par.OmegaLambda; h.HubbleParam = hubble_param;

  FileCloser fc{fopen(path, "w")};
  FILE *fd = fc.fd;
  if (!fd) return B200_ERR_IO;                                // io.c:98-102 endrun(10)
  Stager sg;
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
