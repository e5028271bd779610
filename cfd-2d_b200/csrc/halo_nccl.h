// halo_nccl.h -- ghost-cell exchange of one rank over NVLink 5 / NVSwitch: NCCL send/recv (default) or
// direct peer stores into the neighbours' halo slices (halo_p2p_enable).
//
// Replaces Method::exchange (reference src/methods/method.h:13-127, MPI_Send/MPI_Recv per peer in
// rank order, global.cpp:607-659): one pack kernel gathers the send lists of ALL peers into a
// staging buffer, then a single ncclGroup of ncclSend/ncclRecv; each peer's data lands contiguously
// in the halo slice [nc + recvShift[p], ...) of the destination field, so no unpack is needed.
// NCCL is dlopen()ed ("libnccl.so.2") so the single-GPU path has no NCCL dependency and a process
// that already loaded the torch-bundled NCCL shares it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/cfd2d_fvm.h"

struct HaloNccl;
HaloNccl* halo_create(const cfd2d_halo* d, int nc, int nc_ex, int device, std::string* err);
// Direct peer-store transport for halo_exchange (NVLink / NVSwitch, CUDA IPC): every rank maps its
// neighbours' copies of `fields` (each the base of its own cudaMalloc allocation, [nc_ex * rec4]
// records) and a small flag array; halo_exchange(field in fields) then is ONE kernel that gathers
// the send cells and stores them straight into the neighbours' halo slices, ordered by two flags per
// neighbour (slice free / data landed, release-acquire at system scope) -- no staging buffer, no
// ncclSend/ncclRecv rendezvous.  COLLECTIVE over the communicator.  Returns 0 and leaves the NCCL
// transport in place when any rank cannot map a neighbour (no peer access, asymmetric lists).
int halo_p2p_enable(HaloNccl* h, double4* const* fields, int nfields, cudaStream_t s);
bool halo_p2p_active(const HaloNccl* h);
void halo_destroy(HaloNccl* h);
// exchange records of rec4 double4's per cell (1: U4, 2: G8) of `field` ([nc_ex] records)
int halo_exchange(HaloNccl* h, double4* field, int rec4, cudaStream_t s, int64_t* launches);
int halo_allreduce_min(HaloNccl* h, double* v, cudaStream_t s);
// gather of byte blocks on `root` (Parallel::send/recv towards the root rank, global.cpp:607-659): every rank
// contributes nsend bytes (device), root receives counts[p] bytes of rank p at recv + sum(counts[0..p))
int halo_gather_bytes(HaloNccl* h, int root, const void* send, size_t nsend, void* recv, const size_t* counts, cudaStream_t s);
int halo_rank(const HaloNccl* h);
int halo_nranks(const HaloNccl* h);
const char* halo_error(HaloNccl* h);
