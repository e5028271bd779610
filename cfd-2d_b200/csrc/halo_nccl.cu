// halo_nccl.cu -- see halo_nccl.h
#include "halo_nccl.h"
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <vector>

// K7: halo pack -- gather the 32-byte (U4) or 64-byte (G8) records of the send lists of all peers
// into one contiguous staging buffer (Method::exchange pack loops, reference method.h:20-24)
__global__ void __launch_bounds__(256) k_pack_records(int n, int rec4, const int* __restrict__ idx,
                                                      const double4* __restrict__ src, double4* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * rec4) return;
    int s = i / rec4, k = i - s * rec4;
    const double2* q = reinterpret_cast<const double2*>(src + (size_t)idx[s] * rec4 + k);
    double2 a = __ldcg(q), b = __ldcg(q + 1);
    double2* d = reinterpret_cast<double2*>(dst + i);
    d[0] = a; d[1] = b;
}

typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
enum { NCCL_UINT8 = 1, NCCL_FLOAT64 = 8, NCCL_MIN = 3 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_uid_t*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi g_nccl;

static bool nccl_load(std::string* err) {
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { if (err) *err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) { if (err) *err = "missing NCCL symbol " name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return true;
}

struct HaloNccl {
    int rank = 0, nranks = 1, nc = 0, nc_ex = 0, device = 0;
    std::vector<int> recv_count, recv_shift, send_count, send_off;
    int total_send = 0;
    int* d_send_ind = nullptr;
    double4* d_stage = nullptr;   // total_send * 2 records (large enough for G8)
    double* d_scalar = nullptr;
    nccl_comm_t comm = nullptr;
    std::string error;
};

#define NCCL_TRY(h, call) do { int r_ = (call); if (r_ != 0) { (h)->error = std::string(#call) + ": " + g_nccl.GetErrorString(r_); return CFD2D_ENCCL; } } while (0)

HaloNccl* halo_create(const cfd2d_halo* d, int nc, int nc_ex, int device, std::string* err) {
    if (!nccl_load(err)) return nullptr;
    if (!d->nccl_unique_id) { if (err) *err = "cfd2d_halo.nccl_unique_id is NULL"; return nullptr; }
    HaloNccl* h = new HaloNccl();
    h->rank = d->rank; h->nranks = d->nranks; h->nc = nc; h->nc_ex = nc_ex; h->device = device;
    h->recv_count.assign(d->recv_count, d->recv_count + d->nranks);
    h->send_count.assign(d->send_count, d->send_count + d->nranks);
    h->recv_shift.resize(d->nranks); h->send_off.resize(d->nranks);
    int rs = 0, so = 0;
    for (int p = 0; p < d->nranks; p++) { h->recv_shift[p] = rs; rs += h->recv_count[p]; h->send_off[p] = so; so += h->send_count[p]; }
    h->total_send = so;
    if (rs != nc_ex - nc) { if (err) *err = "sum(recv_count) != nc_ex - nc"; delete h; return nullptr; }
    for (int i = 0; i < so; i++) if (d->send_ind[i] < 0 || d->send_ind[i] >= nc) { if (err) *err = "send_ind entry is not an owned cell"; delete h; return nullptr; }
    {
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_send_ind, (size_t)(so ? so : 1) * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&h->d_stage, (size_t)(so ? so : 1) * 2 * sizeof(double4));
        if (e == cudaSuccess) e = cudaMalloc(&h->d_scalar, sizeof(double));
        if (e == cudaSuccess && so) e = cudaMemcpy(h->d_send_ind, d->send_ind, (size_t)so * sizeof(int), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            if (err) *err = std::string("halo_create: ") + cudaGetErrorString(e);
            halo_destroy(h);
            return nullptr;
        }
    }
    nccl_uid_t uid;
    memcpy(&uid, d->nccl_unique_id, sizeof uid);
    int r = g_nccl.CommInitRank(&h->comm, d->nranks, uid, d->rank);
    if (r != 0) { if (err) *err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); halo_destroy(h); return nullptr; }
    return h;
}

void halo_destroy(HaloNccl* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->comm) g_nccl.CommDestroy(h->comm);
    cudaFree(h->d_send_ind); cudaFree(h->d_stage); cudaFree(h->d_scalar);
    delete h;
}

const char* halo_error(HaloNccl* h) { return h ? h->error.c_str() : ""; }

int halo_exchange(HaloNccl* h, double4* field, int rec4, cudaStream_t s, int64_t* launches) {
    if (h->total_send > 0) {
        long long n = (long long)h->total_send * rec4;
        k_pack_records<<<(int)((n + 255) / 256), 256, 0, s>>>(h->total_send, rec4, h->d_send_ind, field, h->d_stage);
        if (launches) (*launches)++;
    }
    NCCL_TRY(h, g_nccl.GroupStart());
    for (int p = 0; p < h->nranks; p++) {
        if (p == h->rank) continue;
        if (h->send_count[p] > 0)
            NCCL_TRY(h, g_nccl.Send(h->d_stage + (size_t)h->send_off[p] * rec4, (size_t)h->send_count[p] * rec4 * 4, NCCL_FLOAT64, p, h->comm, s));
        if (h->recv_count[p] > 0)
            NCCL_TRY(h, g_nccl.Recv(field + ((size_t)h->nc + h->recv_shift[p]) * rec4, (size_t)h->recv_count[p] * rec4 * 4, NCCL_FLOAT64, p, h->comm, s));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    return 0;
}

int halo_allreduce_min(HaloNccl* h, double* v, cudaStream_t s) {
    cudaMemcpyAsync(h->d_scalar, v, sizeof(double), cudaMemcpyHostToDevice, s);
    NCCL_TRY(h, g_nccl.AllReduce(h->d_scalar, h->d_scalar, 1, NCCL_FLOAT64, NCCL_MIN, h->comm, s));
    cudaMemcpyAsync(v, h->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) { h->error = "allreduce sync failed"; return CFD2D_ECUDA; }
    return 0;
}

int halo_rank(const HaloNccl* h) { return h->rank; }
int halo_nranks(const HaloNccl* h) { return h->nranks; }

int halo_gather_bytes(HaloNccl* h, int root, const void* send, size_t nsend, void* recv, const size_t* counts, cudaStream_t s) {
    if (root < 0 || root >= h->nranks) { h->error = "gather: bad root"; return CFD2D_EINVAL; }
    NCCL_TRY(h, g_nccl.GroupStart());
    if (h->rank == root) {
        size_t off = 0;
        for (int p = 0; p < h->nranks; p++) {
            if (p != root && counts[p] > 0)
                NCCL_TRY(h, g_nccl.Recv((char*)recv + off, counts[p], NCCL_UINT8, p, h->comm, s));
            off += counts[p];
        }
    } else if (nsend > 0) {
        NCCL_TRY(h, g_nccl.Send(send, nsend, NCCL_UINT8, root, h->comm, s));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    return 0;
}

extern "C" int cfd2d_nccl_get_unique_id(void* out128) {
    std::string err;
    if (!nccl_load(&err)) return CFD2D_ENCCL;
    nccl_uid_t uid;
    if (g_nccl.GetUniqueId(&uid) != 0) return CFD2D_ENCCL;
    memcpy(out128, &uid, sizeof uid);
    return 0;
}
