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
    // ---- peer-store transport (halo_p2p_enable)
    bool p2p = false;
    std::vector<double4*> p2p_fields;            // my registered fields
    std::vector<void*> p2p_mapped;               // every pointer opened with cudaIpcOpenMemHandle
    struct P2PDev* d_p2p = nullptr;              // device copy of the tables below
    unsigned long long* d_flags = nullptr;       // [2][nranks]: slice-free credit from rank r, data-landed flag from rank r
    unsigned long long* d_seq = nullptr;         // exchanges made so far (device side: graph replay safe)
    int* d_err = nullptr;                        // set by a wait that timed out
    int* d_p2p_peer_of = nullptr;                // [total_send] destination rank of send entry i
    int* d_p2p_dst = nullptr;                    // [total_send] record index in that rank's field
    int* d_p2p_peers = nullptr;                  // [npeers]
    int p2p_npeers = 0;
};

#define P2P_MAXF 4
#define P2P_MAXR 64
struct P2PDev {
    int rank, nranks, npeers, total_send;
    double4* peer_field[P2P_MAXF][P2P_MAXR];     // mapped base of field f on rank r (nullptr: not a neighbour)
    unsigned long long* peer_flags[P2P_MAXR];    // mapped flag array of rank r
    unsigned long long* my_flags;
    unsigned long long* seq;
    int* err;
    const int* peers;
    const int* peer_of;
    const int* dst;
    const int* send_ind;
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

// ---- peer-store halo exchange -----------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// wait until *flag >= q; gives up after 120 s (a peer that died): error word + trap instead of a hung GPU
__device__ __forceinline__ void p2p_wait(const unsigned long long* flag, unsigned long long q, int* err, int code) {
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(flag) < q) {
        if ((++spins & 1023u) == 0 && global_ns() - t0 > 120000000000ull) { atomicExch(err, code); __trap(); }
        __nanosleep(64);
    }
}

// Method::exchange (method.h:13-41) for ONE field as one kernel of one CTA (a few thousand 32/64-byte
// records: latency, not bandwidth).  Exchange number q, the same on every rank:
//   1. tell every neighbour that my halo slices may be overwritten (every reader of their previous
//      contents was enqueued before this kernel: halo_exchange is called behind a stream fork/join);
//   2. wait for the neighbours' same message;  3. store my send records into their slices over NVLink;
//   4. fence, then tell them the data of exchange q has landed;  5. wait for theirs.
__global__ void __launch_bounds__(256) k_halo_push(const P2PDev* __restrict__ tp, const double4* __restrict__ src, int f, int rec4) {
    const P2PDev& t = *tp;
    __shared__ unsigned long long q_s;
    const int tid = threadIdx.x;
    if (tid == 0) { q_s = *t.seq + 1; *t.seq = q_s; }
    __syncthreads();
    const unsigned long long q = q_s;
    for (int j = tid; j < t.npeers; j += blockDim.x) st_release_sys(t.peer_flags[t.peers[j]] + t.rank, q);
    for (int j = tid; j < t.npeers; j += blockDim.x) p2p_wait(t.my_flags + t.peers[j], q, t.err, 1);
    __syncthreads();
    const int n = t.total_send * rec4;
    for (int i = tid; i < n; i += blockDim.x) {
        const int s = i / rec4, k = i - s * rec4;
        const double2* a = reinterpret_cast<const double2*>(src + (size_t)t.send_ind[s] * rec4 + k);
        const double2 v0 = __ldcg(a), v1 = __ldcg(a + 1);
        double2* d = reinterpret_cast<double2*>(t.peer_field[f][t.peer_of[s]] + (size_t)t.dst[s] * rec4 + k);
        d[0] = v0; d[1] = v1;
    }
    __threadfence_system();
    __syncthreads();
    for (int j = tid; j < t.npeers; j += blockDim.x) st_release_sys(t.peer_flags[t.peers[j]] + t.nranks + t.rank, q);
    for (int j = tid; j < t.npeers; j += blockDim.x) p2p_wait(t.my_flags + t.nranks + t.peers[j], q, t.err, 2);
    __syncthreads();
}

bool halo_p2p_active(const HaloNccl* h) { return h && h->p2p; }

static void p2p_release(HaloNccl* h) {
    for (void* p : h->p2p_mapped) cudaIpcCloseMemHandle(p);
    h->p2p_mapped.clear();
    cudaFree(h->d_p2p); cudaFree(h->d_flags); cudaFree(h->d_seq); cudaFree(h->d_err);
    cudaFree(h->d_p2p_peer_of); cudaFree(h->d_p2p_dst); cudaFree(h->d_p2p_peers);
    h->d_p2p = nullptr; h->d_flags = nullptr; h->d_seq = nullptr; h->d_err = nullptr;
    h->d_p2p_peer_of = h->d_p2p_dst = h->d_p2p_peers = nullptr;
    h->p2p = false;
}

int halo_p2p_enable(HaloNccl* h, double4* const* fields, int nfields, cudaStream_t s) {
    if (h->p2p || h->nranks < 2) return 0;
    if (nfields > P2P_MAXF || h->nranks > P2P_MAXR) return 0;
    // what every neighbour needs from me: where its records land in my fields, and handles of the fields + flags
    struct Msg { long long dst_base; int ok; int pad; cudaIpcMemHandle_t field[P2P_MAXF]; cudaIpcMemHandle_t flags; };
    const int R = h->nranks;
    std::vector<int> peers;
    int ok = 1;
    for (int p = 0; p < R; p++) {
        if (p == h->rank) continue;
        const bool snd = h->send_count[p] > 0, rcv = h->recv_count[p] > 0;
        if (snd != rcv) ok = 0;                       // one-sided neighbour: keep NCCL
        if (snd || rcv) peers.push_back(p);
    }
    cudaError_t e = cudaMalloc(&h->d_flags, 2 * (size_t)R * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(h->d_flags, 0, 2 * (size_t)R * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_seq, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(h->d_seq, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_err, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->d_err, 0, sizeof(int));
    Msg mine;
    memset(&mine, 0, sizeof mine);
    if (e != cudaSuccess) ok = 0;
    for (int f = 0; f < nfields && ok; f++)
        if (cudaIpcGetMemHandle(&mine.field[f], fields[f]) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine.flags, h->d_flags) != cudaSuccess) ok = 0;
    cudaGetLastError();
    // pairwise exchange of the messages over the communicator (every rank takes part, also with ok == 0)
    Msg* d_out = nullptr; Msg* d_in = nullptr;
    if (cudaMalloc(&d_out, sizeof(Msg) * (size_t)R) != cudaSuccess || cudaMalloc(&d_in, sizeof(Msg) * (size_t)R) != cudaSuccess) {
        h->error = "halo_p2p_enable: cudaMalloc failed"; return CFD2D_ECUDA;
    }
    std::vector<Msg> out(R, mine), in(R);
    for (int p = 0; p < R; p++) { out[p].dst_base = (long long)h->nc + h->recv_shift[p]; out[p].ok = ok; }
    cudaMemcpyAsync(d_out, out.data(), sizeof(Msg) * (size_t)R, cudaMemcpyHostToDevice, s);
    NCCL_TRY(h, g_nccl.GroupStart());
    for (int p : peers) {
        NCCL_TRY(h, g_nccl.Send(d_out + p, sizeof(Msg), NCCL_UINT8, p, h->comm, s));
        NCCL_TRY(h, g_nccl.Recv(d_in + p, sizeof(Msg), NCCL_UINT8, p, h->comm, s));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    cudaMemcpyAsync(in.data(), d_in, sizeof(Msg) * (size_t)R, cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) { h->error = "halo_p2p_enable: message exchange failed"; return CFD2D_ECUDA; }
    cudaFree(d_out); cudaFree(d_in);
    // map the neighbours' memory
    P2PDev t;
    memset(&t, 0, sizeof t);
    for (int p : peers) {
        if (!in[p].ok) ok = 0;
        for (int f = 0; f < nfields && ok; f++) {
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, in[p].field[f], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
            h->p2p_mapped.push_back(q);
            t.peer_field[f][p] = (double4*)q;
        }
        if (ok) {
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, in[p].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
            else { h->p2p_mapped.push_back(q); t.peer_flags[p] = (unsigned long long*)q; }
        }
    }
    cudaGetLastError();
    // every rank must agree: min over the communicator
    {
        double v = ok ? 1.0 : 0.0;
        int rc = halo_allreduce_min(h, &v, s);
        if (rc) return rc;
        ok = v > 0.5;
    }
    if (!ok) { p2p_release(h); return 0; }
    std::vector<int> peer_of(h->total_send > 0 ? h->total_send : 1, 0), dst(h->total_send > 0 ? h->total_send : 1, 0);
    for (int p : peers)
        for (int i = 0; i < h->send_count[p]; i++) {
            peer_of[h->send_off[p] + i] = p;
            dst[h->send_off[p] + i] = (int)in[p].dst_base + i;       // my records land at rank p's nc + recvShift[me] + i
        }
    e = cudaMalloc(&h->d_p2p_peer_of, peer_of.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_p2p_dst, dst.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_p2p_peers, (peers.size() ? peers.size() : 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_p2p, sizeof(P2PDev));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_p2p_peer_of, peer_of.data(), peer_of.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_p2p_dst, dst.data(), dst.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !peers.empty()) e = cudaMemcpy(h->d_p2p_peers, peers.data(), peers.size() * sizeof(int), cudaMemcpyHostToDevice);
    t.rank = h->rank; t.nranks = R; t.npeers = (int)peers.size(); t.total_send = h->total_send;
    t.my_flags = h->d_flags; t.seq = h->d_seq; t.err = h->d_err;
    t.peers = h->d_p2p_peers; t.peer_of = h->d_p2p_peer_of; t.dst = h->d_p2p_dst; t.send_ind = h->d_send_ind;
    if (e == cudaSuccess) e = cudaMemcpy(h->d_p2p, &t, sizeof t, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { h->error = std::string("halo_p2p_enable: ") + cudaGetErrorString(e); return CFD2D_ECUDA; }
    h->p2p_fields.assign(fields, fields + nfields);
    h->p2p_npeers = (int)peers.size();
    h->p2p = true;
    return 0;
}

void halo_destroy(HaloNccl* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->p2p) {
        // nobody may free exported memory while a neighbour still has it mapped: close mine, then meet
        cudaDeviceSynchronize();
        for (void* p : h->p2p_mapped) cudaIpcCloseMemHandle(p);
        h->p2p_mapped.clear();
        double v = 1.0;
        halo_allreduce_min(h, &v, nullptr);
        p2p_release(h);
    }
    if (h->comm) g_nccl.CommDestroy(h->comm);
    cudaFree(h->d_send_ind); cudaFree(h->d_stage); cudaFree(h->d_scalar);
    delete h;
}

const char* halo_error(HaloNccl* h) { return h ? h->error.c_str() : ""; }

int halo_exchange(HaloNccl* h, double4* field, int rec4, cudaStream_t s, int64_t* launches) {
    if (h->p2p) {
        for (size_t f = 0; f < h->p2p_fields.size(); f++)
            if (h->p2p_fields[f] == field) {
                k_halo_push<<<1, 256, 0, s>>>(h->d_p2p, field, (int)f, rec4);
                if (launches) (*launches)++;
                return 0;
            }
    }
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
