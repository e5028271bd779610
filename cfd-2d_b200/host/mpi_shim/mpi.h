/* Stand-in for <mpi.h> for building the drop-in driver (cfd2d_cuda) where no MPI is
 * installed; a site with MPI compiles the glue against its real <mpi.h> instead.
 *
 * The reference (zhrv/cfd-2d) includes "mpi.h" from src/global.h:9 and calls a
 * handful of MPI entry points from src/global.cpp:595-659.  The explicit FVM
 * path (FVM_TVD) never communicates, so to compile the reference sources where
 * they lie we only need the declarations.  Every function is a one-rank no-op.
 */
#ifndef CFD2D_HOST_MPI_SHIM_H
#define CFD2D_HOST_MPI_SHIM_H
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int unused; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_MIN 3
#define MPI_SUCCESS 0
static inline int MPI_Init(int*, char***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
/* rank / size come from the environment a process launcher sets (tools/launch_ranks.py, torchrun, mpirun's
 * OMPI_* / PMI_* equivalents): WORLD_SIZE and RANK.  Unset => one rank.  Point-to-point calls stay no-ops:
 * the only traffic of the multi-rank FVM_TVD_CUDA method (halo exchange, time-step reduction, state gather)
 * goes over NCCL inside the CUDA library, and the 128-byte ncclUniqueId is handed over through a file. */
#include <stdlib.h>
static inline int cfd2d_shim_env_int(const char* k, int dflt) { const char* v = getenv(k); return (v && *v) ? atoi(v) : dflt; }
static inline int MPI_Comm_size(MPI_Comm, int* n) { *n = cfd2d_shim_env_int("WORLD_SIZE", 1); if (*n < 1) *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = cfd2d_shim_env_int("RANK", 0); return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
static inline int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { return 0; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
    const char* a = (const char*)s; char* b = (char*)r;
    int w = (t == MPI_DOUBLE) ? 8 : 4;
    for (int i = 0; i < n * w; i++) b[i] = a[i];
    return 0;
}
#endif
