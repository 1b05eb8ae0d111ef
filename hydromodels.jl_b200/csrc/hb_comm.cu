// Multi-GPU plumbing behind the C-ABI (one process per GPU): a communicator over NCCL (resolved at run time from
// libnccl.so.2 — the library itself carries no link dependency on it), symmetric peer-addressable allocations over
// CUDA IPC, and the halo exchange of the sharded dense-grid run (parallel.ShardedPanelRun / ShardedGridRun):
//
//   mode HB_HALO_NCCL : pack kernel -> ONE grouped ncclSend/ncclRecv per neighbour on the compute stream -> unpack kernel;
//   mode HB_HALO_PEER : the pack kernel stores every neighbour's block straight into that neighbour's staging buffer
//                       over NVLink (peer pointers from the IPC table), the last block to finish publishes the epoch in
//                       the neighbour's flag word; the unpack kernel of the receiving rank waits for the epoch and
//                       copies the block into its ghost cells.  No host round trip, no NCCL launch, two small kernels.
//
// What is exchanged: the bucket and river STATES of the ghost rows / columns every `halo` steps (temporal blocking,
// /root/reference/src/route.jl:387-403 moves water one cell per step, so `halo` steps can run between refreshes).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/hydrob200.h"

namespace {

// ---- the few NCCL entry points we use, resolved with dlsym ---------------------------------------------------
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
struct Nccl {
    bool ok = false;
    std::string err;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
} g_nccl;
std::mutex g_nccl_mu;
constexpr int kNcclChar = 0, kNcclDouble = 8;

extern "C" void hb_internal_set_error(const char *msg);      // hb_api.cpp: the thread-local message behind hb_last_error()

int cfail(int code, const std::string &msg) {
    hb_internal_set_error(msg.c_str());
    return code;
}

int load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.ok) return HB_OK;
    void *h = nullptr;
    if (const char *p = getenv("HB_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);      // already mapped when the host process uses NCCL (torch, MPI + NCCL)
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return cfail(HB_ERR_CUDA, std::string("libnccl.so.2 not found (set HB_NCCL_LIB): ") + dlerror());
#define HB_SYM(field, name)                                                              \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                          \
    if (!g_nccl.field) return cfail(HB_ERR_CUDA, std::string("libnccl lacks ") + name);
    HB_SYM(GetUniqueId, "ncclGetUniqueId")
    HB_SYM(CommInitRank, "ncclCommInitRank")
    HB_SYM(CommDestroy, "ncclCommDestroy")
    HB_SYM(Send, "ncclSend")
    HB_SYM(Recv, "ncclRecv")
    HB_SYM(GroupStart, "ncclGroupStart")
    HB_SYM(GroupEnd, "ncclGroupEnd")
    HB_SYM(AllGather, "ncclAllGather")
    HB_SYM(GetErrorString, "ncclGetErrorString")
#undef HB_SYM
    g_nccl.ok = true;
    return HB_OK;
}

#define HB_NCCL(expr)                                                                                           \
    do {                                                                                                        \
        int r_ = (expr);                                                                                        \
        if (r_ != 0) return cfail(HB_ERR_CUDA, std::string(#expr " failed: ") + g_nccl.GetErrorString(r_));     \
    } while (0)
#define HB_CU(expr)                                                                                             \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess) return cfail(HB_ERR_CUDA, std::string(#expr " failed: ") + cudaGetErrorString(e_)); \
    } while (0)

constexpr int kMaxPlanes = 8, kMaxEdges = 8;

struct EdgeDev {
    int s_row0, s_col0, r_row0, r_col0, nrows, ncols;
    long long buf_off;                  // element offset of this edge's block in the local send / recv buffers
    double *remote;                     // PEER mode: the neighbour's staging slot for this epoch
    unsigned long long *remote_flag;    // PEER mode: the neighbour's flag word for this edge
    const double *local_stage;          // PEER mode: my staging slot the neighbour fills
    const unsigned long long *local_flag;
};
struct HaloArgs {
    double *planes[kMaxPlanes];
    int n_planes, pitch, n_edges;
    unsigned long long epoch;
    unsigned int *done;                 // PEER mode: per-edge block counters (local)
    EdgeDev e[kMaxEdges];
};

// gather the blocks the neighbours need.  grid = (blocks per edge, n_edges).  NCCL mode: into the local send buffer; PEER
// mode: into the neighbour's staging over NVLink, then the last block of the edge publishes the epoch.
__global__ void __launch_bounds__(256) hb_halo_pack(const __grid_constant__ HaloArgs a, double *sendbuf, int peer_mode) {
    const EdgeDev &e = a.e[blockIdx.y];
    const long long per_plane = (long long)e.nrows * e.ncols, total = per_plane * a.n_planes;
    double *dst = peer_mode ? e.remote : sendbuf + e.buf_off;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int p = (int)(i / per_plane);
        const long long k = i - p * per_plane;
        const int r = (int)(k / e.ncols), c = (int)(k - (long long)r * e.ncols);
        dst[i] = a.planes[p][(long long)(e.s_row0 + r) * a.pitch + e.s_col0 + c];
    }
    if (peer_mode) {
        __threadfence_system();                                 // my stores are visible at the peer before the count moves
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned prev = atomicAdd(a.done + blockIdx.y, 1u);
            if (prev == gridDim.x - 1) {                        // last block of this edge
                a.done[blockIdx.y] = 0;
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(e.remote_flag), "l"(a.epoch) : "memory");
            }
        }
    }
}

// scatter the received blocks into the ghost cells.  PEER mode waits (per block) until the neighbour published this epoch.
__global__ void __launch_bounds__(256) hb_halo_unpack(const __grid_constant__ HaloArgs a, const double *recvbuf, int peer_mode,
                                                      unsigned int *timeout_word) {
    const EdgeDev &e = a.e[blockIdx.y];
    const double *src = peer_mode ? e.local_stage : recvbuf + e.buf_off;
    if (peer_mode) {
        if (threadIdx.x == 0) {
            unsigned long long seen = 0;
            long long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(e.local_flag) : "memory");
                if (seen >= a.epoch) break;
                if (++spins > (1ll << 26)) { atomicExch(timeout_word, 1u); break; }      // ~ seconds: the peer never arrived
                __nanosleep(100);
            }
        }
        __syncthreads();
    }
    const long long per_plane = (long long)e.nrows * e.ncols, total = per_plane * a.n_planes;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int p = (int)(i / per_plane);
        const long long k = i - p * per_plane;
        const int r = (int)(k / e.ncols), c = (int)(k - (long long)r * e.ncols);
        double v;
        if (peer_mode) asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(src + i) : "memory");   // written by another GPU: not through L1
        else v = src[i];
        a.planes[p][(long long)(e.r_row0 + r) * a.pitch + e.r_col0 + c] = v;
    }
}

}  // namespace

struct hb_comm_s {
    nccl_comm comm = nullptr;
    int rank = 0, world = 1, device = 0;
    // symmetric segments: segment k is one allocation per rank, peers[k][r] = address of rank r's copy in THIS process
    std::vector<std::vector<void *>> peers;
    std::vector<size_t> seg_bytes;
    // halo exchange state
    double *sendbuf = nullptr, *recvbuf = nullptr;
    size_t buf_cap = 0;
    int stage_seg = -1;                 // PEER mode: [2 parity][kMaxEdges slots][stage_block_bytes] + flags[kMaxEdges] (u64) at the end
    size_t stage_block = 0;
    unsigned int *done = nullptr;       // [kMaxEdges] block counters + [1] timeout word
    unsigned long long epoch = 0;
};

extern "C" {

int hb_comm_unique_id(void *id, int32_t id_bytes) {
    if (!id || id_bytes < 128) return cfail(HB_ERR_INVALID, "the unique id needs 128 bytes");
    int rc = load_nccl();
    if (rc) return rc;
    nccl_uid u;
    HB_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, 128);
    return HB_OK;
}

int hb_comm_init(const void *id, int32_t rank, int32_t world, hb_comm_t *out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return cfail(HB_ERR_INVALID, "bad communicator arguments");
    int rc = load_nccl();
    if (rc) return rc;
    hb_comm_s *c = new hb_comm_s();
    c->rank = rank;
    c->world = world;
    if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; return cfail(HB_ERR_NO_DEVICE, "no CUDA device (call hb_init first)"); }
    nccl_uid u;
    memcpy(&u, id, 128);
    int r = g_nccl.CommInitRank(&c->comm, world, u, rank);
    if (r != 0) { delete c; return cfail(HB_ERR_CUDA, std::string("ncclCommInitRank failed: ") + g_nccl.GetErrorString(r)); }
    if (cudaMalloc(&c->done, (kMaxEdges + 1) * sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(c->done, 0, (kMaxEdges + 1) * sizeof(unsigned int)) != cudaSuccess) {
        delete c;
        return cfail(HB_ERR_NOMEM, "cudaMalloc failed");
    }
    *out = c;
    return HB_OK;
}

int hb_comm_rank(hb_comm_t c, int32_t *rank, int32_t *world) {
    if (!c) return cfail(HB_ERR_INVALID, "communicator is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return HB_OK;
}

int hb_comm_free(hb_comm_t c) {
    if (!c) return HB_OK;
    cudaDeviceSynchronize();
    for (size_t k = 0; k < c->peers.size(); ++k)
        for (int r = 0; r < c->world; ++r) {
            if (!c->peers[k][r]) continue;
            if (r == c->rank) cudaFree(c->peers[k][r]);
            else cudaIpcCloseMemHandle(c->peers[k][r]);
        }
    if (c->sendbuf) cudaFree(c->sendbuf);
    if (c->recvbuf) cudaFree(c->recvbuf);
    if (c->done) cudaFree(c->done);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return HB_OK;
}

// All-gather of `bytes` per rank (objective vectors, final states): recv = [world][bytes].
int hb_comm_all_gather(hb_comm_t c, const void *send, void *recv, int64_t bytes, void *stream) {
    if (!c || !send || !recv || bytes < 0) return cfail(HB_ERR_INVALID, "bad all_gather arguments");
    HB_NCCL(g_nccl.AllGather(send, recv, (size_t)bytes, kNcclChar, c->comm, (cudaStream_t)stream));
    return HB_OK;
}

// Collective: every rank allocates `bytes` (zero-filled) and maps every other rank's allocation (CUDA IPC; NVLink peer
// access).  `*segment` identifies the allocation in hb_comm_peer_ptr.
int hb_comm_alloc(hb_comm_t c, int64_t bytes, void **local_ptr, int32_t *segment) {
    if (!c || bytes <= 0 || !local_ptr || !segment) return cfail(HB_ERR_INVALID, "bad hb_comm_alloc arguments");
    void *mine = nullptr;
    HB_CU(cudaMalloc(&mine, (size_t)bytes));
    HB_CU(cudaMemset(mine, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    HB_CU(cudaIpcGetMemHandle(&h, mine));
    char *d_all = nullptr;
    HB_CU(cudaMalloc(&d_all, sizeof(h) * (size_t)(c->world + 1)));
    HB_CU(cudaMemcpy(d_all + sizeof(h) * c->world, &h, sizeof(h), cudaMemcpyHostToDevice));
    HB_NCCL(g_nccl.AllGather(d_all + sizeof(h) * c->world, d_all, sizeof(h), kNcclChar, c->comm, (cudaStream_t)0));
    HB_CU(cudaStreamSynchronize(0));
    std::vector<cudaIpcMemHandle_t> all(c->world);
    HB_CU(cudaMemcpy(all.data(), d_all, sizeof(h) * c->world, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    std::vector<void *> table(c->world, nullptr);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) { table[r] = mine; continue; }
        cudaError_t e = cudaIpcOpenMemHandle(&table[r], all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (q != c->rank && table[q]) cudaIpcCloseMemHandle(table[q]);
            cudaFree(mine);
            return cfail(HB_ERR_CUDA, std::string("cudaIpcOpenMemHandle (peer ") + std::to_string(r) + ") failed: " + cudaGetErrorString(e));
        }
    }
    c->peers.push_back(table);
    c->seg_bytes.push_back((size_t)bytes);
    *local_ptr = mine;
    *segment = (int32_t)c->peers.size() - 1;
    return HB_OK;
}

int hb_comm_peer_ptr(hb_comm_t c, int32_t segment, int32_t peer, void **ptr) {
    if (!c || !ptr || segment < 0 || segment >= (int32_t)c->peers.size() || peer < 0 || peer >= c->world)
        return cfail(HB_ERR_INVALID, "bad hb_comm_peer_ptr arguments");
    *ptr = c->peers[segment][peer];
    return HB_OK;
}

// Collective: reserve the peer-store staging (two epochs x kMaxEdges slots of `max_block_bytes`, + flag words).
int hb_halo_enable_peer(hb_comm_t c, int64_t max_block_bytes) {
    if (!c || max_block_bytes <= 0) return cfail(HB_ERR_INVALID, "bad hb_halo_enable_peer arguments");
    if (c->stage_seg >= 0) return cfail(HB_ERR_INVALID, "peer staging is already reserved");
    const size_t blk = ((size_t)max_block_bytes + 255) / 256 * 256;
    void *p = nullptr;
    int32_t seg = -1;
    int rc = hb_comm_alloc(c, (int64_t)(2 * kMaxEdges * blk + 256), &p, &seg);
    if (rc) return rc;
    c->stage_seg = seg;
    c->stage_block = blk;
    return HB_OK;
}

int hb_halo_exchange_device(hb_comm_t c, double *const *planes, int32_t n_planes, int32_t pitch, const hb_halo_edge *edges,
                            int32_t n_edges, int32_t mode, void *stream) {
    if (!c || !planes || !edges || n_planes < 1 || n_planes > kMaxPlanes || n_edges < 0 || n_edges > kMaxEdges || pitch < 1)
        return cfail(HB_ERR_INVALID, "bad halo exchange arguments");
    if (mode != HB_HALO_NCCL && mode != HB_HALO_PEER) return cfail(HB_ERR_INVALID, "unknown halo exchange mode");
    if (mode == HB_HALO_PEER && c->stage_seg < 0) return cfail(HB_ERR_INVALID, "call hb_halo_enable_peer first");
    cudaStream_t st = (cudaStream_t)stream;
    c->epoch += 1;
    if (n_edges == 0) return HB_OK;
    HaloArgs a;
    memset(&a, 0, sizeof a);
    for (int p = 0; p < n_planes; ++p) a.planes[p] = planes[p];
    a.n_planes = n_planes;
    a.pitch = pitch;
    a.n_edges = n_edges;
    a.epoch = c->epoch;
    a.done = c->done;
    long long off = 0, biggest = 0;
    const size_t par = (size_t)(c->epoch & 1);
    for (int k = 0; k < n_edges; ++k) {
        const hb_halo_edge &h = edges[k];
        if (h.peer < 0 || h.peer >= c->world || h.peer == c->rank || h.nrows < 1 || h.ncols < 1 || h.peer_slot < 0 || h.peer_slot >= kMaxEdges)
            return cfail(HB_ERR_INVALID, "bad halo edge");
        EdgeDev &e = a.e[k];
        e.s_row0 = h.s_row0; e.s_col0 = h.s_col0; e.r_row0 = h.r_row0; e.r_col0 = h.r_col0; e.nrows = h.nrows; e.ncols = h.ncols;
        e.buf_off = off;
        const long long n = (long long)h.nrows * h.ncols * n_planes;
        off += n;
        if (n > biggest) biggest = n;
        if (mode == HB_HALO_PEER) {
            if ((size_t)n * 8 > c->stage_block) return cfail(HB_ERR_INVALID, "halo block larger than the reserved peer staging");
            char *remote = (char *)c->peers[c->stage_seg][h.peer], *local = (char *)c->peers[c->stage_seg][c->rank];
            const size_t flags_off = 2 * kMaxEdges * c->stage_block;
            e.remote = (double *)(remote + (par * kMaxEdges + h.peer_slot) * c->stage_block);
            e.remote_flag = (unsigned long long *)(remote + flags_off) + h.peer_slot;
            e.local_stage = (const double *)(local + (par * kMaxEdges + k) * c->stage_block);
            e.local_flag = (const unsigned long long *)(local + flags_off) + k;
        }
    }
    if (mode == HB_HALO_NCCL && (size_t)off * 8 > c->buf_cap) {
        cudaStreamSynchronize(st);
        if (c->sendbuf) cudaFree(c->sendbuf);
        if (c->recvbuf) cudaFree(c->recvbuf);
        c->sendbuf = c->recvbuf = nullptr;
        c->buf_cap = 0;
        HB_CU(cudaMalloc(&c->sendbuf, (size_t)off * 8));
        HB_CU(cudaMalloc(&c->recvbuf, (size_t)off * 8));
        c->buf_cap = (size_t)off * 8;
    }
    unsigned bx = (unsigned)((biggest + 256 * 8 - 1) / (256 * 8));       // ~8 elements per thread
    if (bx < 1) bx = 1;
    if (bx > 296) bx = 296;
    dim3 grid(bx, (unsigned)n_edges);
    const int peer = mode == HB_HALO_PEER;
    hb_halo_pack<<<grid, 256, 0, st>>>(a, c->sendbuf, peer);
    if (!peer) {
        HB_NCCL(g_nccl.GroupStart());
        for (int k = 0; k < n_edges; ++k) {
            const size_t n = (size_t)edges[k].nrows * edges[k].ncols * n_planes;
            HB_NCCL(g_nccl.Send(c->sendbuf + a.e[k].buf_off, n, kNcclDouble, edges[k].peer, c->comm, st));
            HB_NCCL(g_nccl.Recv(c->recvbuf + a.e[k].buf_off, n, kNcclDouble, edges[k].peer, c->comm, st));
        }
        HB_NCCL(g_nccl.GroupEnd());
    }
    hb_halo_unpack<<<grid, 256, 0, st>>>(a, c->recvbuf, peer, c->done + kMaxEdges);
    HB_CU(cudaGetLastError());
    return HB_OK;
}

// Synchronises `stream`; HB_ERR_CUDA if a peer-mode unpack gave up waiting for a neighbour's epoch.
int hb_halo_status(hb_comm_t c, void *stream) {
    if (!c) return cfail(HB_ERR_INVALID, "communicator is NULL");
    HB_CU(cudaStreamSynchronize((cudaStream_t)stream));
    unsigned int w = 0;
    HB_CU(cudaMemcpy(&w, c->done + kMaxEdges, 4, cudaMemcpyDeviceToHost));
    if (w) {
        cudaMemset(c->done + kMaxEdges, 0, 4);
        return cfail(HB_ERR_CUDA, "hb_halo_exchange: a neighbour never published its halo block (peer mode)");
    }
    return HB_OK;
}

}  // extern "C"
