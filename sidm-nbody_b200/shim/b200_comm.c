/* b200_comm.c - the all-gather libsidm_b200.so asks its host for when the work is sharded over several tasks
 * (b200_set_shard, include/sidm_b200.h): the C counterpart of sidm_b200/multi.py, linked with the shim.
 *
 *   -DB200_WITH_NCCL and one GPU per task: ncclAllGather over NVLink, ordered on the stream the library names
 *     (b200_current_stream()), communicator built from MPI_COMM_WORLD - task 0 broadcasts the ncclUniqueId;
 *   otherwise (no NCCL at build time, or several tasks sharing one GPU): staged through the host with MPI_Allgather.
 * Replaces the hypercube exchanges of gravtree.c:171-222 and sidm.c:204-553. */
#include <stdio.h>
#include <stdlib.h>
#include <mpi.h>
#include <cuda_runtime_api.h>
#ifdef B200_WITH_NCCL
#include <nccl.h>
#endif
#include "sidm_b200.h"

static int c_rank = 0, c_world = 1, c_nccl = 0;
static char *h_send = 0, *h_recv = 0;
static long long h_cap = 0;
#ifdef B200_WITH_NCCL
static ncclComm_t c_comm;
#endif

int b200_comm_init(int rank, int world, int use_nccl)
{
  c_rank = rank; c_world = world; c_nccl = 0;
#ifdef B200_WITH_NCCL
  if (use_nccl && world > 1) {
    ncclUniqueId id;
    if (rank == 0 && ncclGetUniqueId(&id) != ncclSuccess) return 1;
    MPI_Bcast(&id, (int)sizeof(id), MPI_BYTE, 0, MPI_COMM_WORLD);
    if (ncclCommInitRank(&c_comm, world, id, rank) != ncclSuccess) return 1;
    c_nccl = 1;
  }
#else
  (void)use_nccl;
#endif
  return 0;
}

int b200_comm_uses_nccl(void) { return c_nccl; }

/* b200_allgather_fn: send[0..bytes) of every task -> recv[task*bytes ..) on every task */
int b200_comm_allgather(long long bytes, void *user)
{
  void *send = 0, *recv = 0;
  long long cap = 0;
  cudaStream_t st = (cudaStream_t)b200_current_stream();
  (void)user;
  if (b200_shard_buffers(&send, &recv, &cap) != B200_OK || bytes > cap) return 1;
#ifdef B200_WITH_NCCL
  if (c_nccl) return ncclAllGather(send, recv, (size_t)bytes, ncclChar, c_comm, st) == ncclSuccess ? 0 : 1;
#endif
  if (bytes > h_cap) {
    if (h_send) cudaFreeHost(h_send);
    if (h_recv) cudaFreeHost(h_recv);
    h_cap = bytes + bytes / 4 + 4096;
    if (cudaMallocHost((void **)&h_send, (size_t)h_cap) != cudaSuccess || cudaMallocHost((void **)&h_recv, (size_t)h_cap * c_world) != cudaSuccess) return 1;
  }
  if (bytes > 0x7fffffffLL / c_world) return 1;                  /* int counts of MPI-1 */
  if (cudaMemcpyAsync(h_send, send, (size_t)bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
  if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
  MPI_Allgather(h_send, (int)bytes, MPI_BYTE, h_recv, (int)bytes, MPI_BYTE, MPI_COMM_WORLD);
  if (cudaMemcpyAsync(recv, h_recv, (size_t)bytes * c_world, cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}

void b200_comm_finalize(void)
{
#ifdef B200_WITH_NCCL
  if (c_nccl) ncclCommDestroy(c_comm);
#endif
  if (h_send) cudaFreeHost(h_send);
  if (h_recv) cudaFreeHost(h_recv);
  h_send = h_recv = 0; h_cap = 0; c_nccl = 0;
}
