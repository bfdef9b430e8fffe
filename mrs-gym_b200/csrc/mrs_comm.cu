// mrs_comm.cu -- the one exchange step of the path: the per-rollout statistics reduction across the env
// shards of one node (SURVEY.md §8b mrs_stats_allreduce, §8e), hand-written over NVLink peer memory.
//
// The payload is MRS_STATS_SLOTS counters = 64 bytes: a latency problem, not a bandwidth one.  Every rank owns
// a small device "mailbox"; the mailboxes of all ranks are mapped into every process (cudaIpc handles, or
// plain peer pointers inside one process).  One all-reduce = one 32-thread kernel per GPU:
//     lane r  : PUSH my 64 bytes into rank r's mailbox (4 x 16-byte stores through NVLink / NVSwitch),
//               fence.sys, then a release store of this call's epoch into rank r's flag word for me;
//     lane r  : spin on MY flag word for rank r (local memory: the peers wrote it) until it shows the epoch;
//     lane i<8: sum slot i over the ranks in rank order (deterministic) -> out[i].
// Nothing is ever loaded over NVLink (a peer load costs ~1 us of round trip, a store is fire-and-forget), the host
// is not involved, the epoch lives in device memory, so the kernel is capturable in a CUDA graph and replays.
// Two mailbox halves alternate by epoch parity: a rank can be at most one call ahead of the slowest peer.
// A spin that lasts longer than MRS_COMM_TIMEOUT_NS (a peer died) gives up and raises MRS_STATUS_COMM_TIMEOUT.
#include <cuda_runtime.h>
#include <new>
#include <string.h>

#include "mrs_b200.h"

namespace {

constexpr int kMaxWorld = MRS_COMM_MAX_WORLD;
constexpr unsigned long long kTimeoutNs = 4000000000ull;      // 4 s

struct Mailbox {
    unsigned long long data[2][kMaxWorld][MRS_STATS_SLOTS];   // [parity][source rank][slot]
    unsigned flag[2][kMaxWorld];                               // epoch of the newest complete push, per source
    unsigned bar[2][kMaxWorld];                                // the same for mrs_comm_barrier
    unsigned epoch, bar_epoch;                                 // calls completed by the owning rank
};

struct PeerTable {
    Mailbox* box[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// wait until *flag == want; false on timeout
__device__ __forceinline__ bool spin_until(const unsigned* flag, unsigned want) {
    if (ld_acquire_sys(flag) == want) return true;
    const unsigned long long t0 = now_ns();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 64; ++i)
            if (ld_acquire_sys(flag) == want) return true;
        if (now_ns() - t0 > kTimeoutNs) return false;
    }
}

__global__ void __launch_bounds__(32)
stats_allreduce_kernel(const PeerTable peers, int rank, int world, const unsigned long long* __restrict__ stats,
                       unsigned long long* __restrict__ out, unsigned* __restrict__ status) {
    const int lane = threadIdx.x;
    Mailbox* mine = peers.box[rank];
    const unsigned e = mine->epoch + 1u;
    const unsigned par = e & 1u;
    __syncwarp();
    bool ok = true;
    if (lane < world) {
        Mailbox* p = peers.box[lane];
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(&p->data[par][rank][0]);
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(stats);
#pragma unroll
        for (int i = 0; i < MRS_STATS_SLOTS / 2; ++i) dst[i] = src[i];
        __threadfence_system();
        st_release_sys(&p->flag[par][rank], e);
        ok = spin_until(&mine->flag[par][lane], e);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane < MRS_STATS_SLOTS) {
        unsigned long long s = 0;
        if (ok) {
            for (int r = 0; r < world; ++r) s += ld_relaxed_sys_u64(&mine->data[par][r][lane]);
        } else {
            s = stats[lane];                      // a peer is gone: report the local counters and flag it
        }
        out[lane] = s;
    }
    if (lane == 0) {
        mine->epoch = e;
        if (!ok && status) atomicOr(status, MRS_STATUS_COMM_TIMEOUT);
    }
}

__global__ void __launch_bounds__(32)
barrier_kernel(const PeerTable peers, int rank, int world, unsigned* __restrict__ status) {
    const int lane = threadIdx.x;
    Mailbox* mine = peers.box[rank];
    const unsigned e = mine->bar_epoch + 1u;
    const unsigned par = e & 1u;
    __syncwarp();
    bool ok = true;
    if (lane < world) {
        st_release_sys(&peers.box[lane]->bar[par][rank], e);
        ok = spin_until(&mine->bar[par][lane], e);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        mine->bar_epoch = e;
        if (!ok && status) atomicOr(status, MRS_STATUS_COMM_TIMEOUT);
    }
}

}  // namespace

struct MrsPeerComm {
    int rank = 0, world = 1, device = 0;
    Mailbox* mine = nullptr;
    PeerTable table = {};
    bool opened[kMaxWorld] = {};      // mapped with cudaIpcOpenMemHandle (to be closed)
    bool connected = false;
};

extern "C" {

int mrs_comm_create(int rank, int world, MrsPeerComm** out) {
    if (!out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return MRS_ERR_ARG;
    MrsPeerComm* c = new (std::nothrow) MrsPeerComm();
    if (!c) return MRS_ERR_CUDA;
    c->rank = rank;
    c->world = world;
    if (cudaGetDevice(&c->device) != cudaSuccess || cudaMalloc(&c->mine, sizeof(Mailbox)) != cudaSuccess ||
        cudaMemset(c->mine, 0, sizeof(Mailbox)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        (void)cudaGetLastError();
        if (c->mine) cudaFree(c->mine);
        delete c;
        return MRS_ERR_CUDA;
    }
    c->table.box[rank] = c->mine;
    c->connected = (world == 1);
    *out = c;
    return MRS_OK;
}

int mrs_comm_handle(const MrsPeerComm* c, unsigned char* out_handle) {
    if (!c || !out_handle) return MRS_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) <= MRS_COMM_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, c->mine) != cudaSuccess) {
        (void)cudaGetLastError();
        return MRS_ERR_CUDA;
    }
    memset(out_handle, 0, MRS_COMM_HANDLE_BYTES);
    memcpy(out_handle, &h, sizeof(h));
    return MRS_OK;
}

void* mrs_comm_mailbox(const MrsPeerComm* c) { return c ? (void*)c->mine : nullptr; }

int mrs_comm_connect(MrsPeerComm* c, const unsigned char* handles) {
    if (!c || (!handles && c->world > 1)) return MRS_ERR_ARG;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || c->table.box[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * MRS_COMM_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            (void)cudaGetLastError();
            return MRS_ERR_CUDA;
        }
        c->table.box[r] = static_cast<Mailbox*>(p);
        c->opened[r] = true;
    }
    c->connected = true;
    return MRS_OK;
}

int mrs_comm_connect_ptrs(MrsPeerComm* c, void* const* mailboxes, const int* devices) {
    if (!c || !mailboxes) return MRS_ERR_ARG;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        if (!mailboxes[r]) return MRS_ERR_ARG;
        if (devices && devices[r] != c->device) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->device, devices[r]) != cudaSuccess || !can) return MRS_ERR_UNSUPPORTED;
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                (void)cudaGetLastError();
                return MRS_ERR_CUDA;
            }
            (void)cudaGetLastError();
        }
        c->table.box[r] = static_cast<Mailbox*>(mailboxes[r]);
    }
    c->connected = true;
    return MRS_OK;
}

int mrs_comm_destroy(MrsPeerComm* c) {
    if (!c) return MRS_OK;
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->table.box[r]);
    if (c->mine) cudaFree(c->mine);
    (void)cudaGetLastError();
    delete c;
    return MRS_OK;
}

int mrs_comm_barrier(MrsPeerComm* c, unsigned int* status, void* stream) {
    if (!c || !c->connected) return MRS_ERR_ARG;
    if (c->world == 1) return MRS_OK;
    barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->table, c->rank, c->world, status);
    return cudaGetLastError() == cudaSuccess ? MRS_OK : MRS_ERR_CUDA;
}

int mrs_stats_allreduce(const MrsConfig* cfg, const MrsBuffers* bufs, MrsPeerComm* comm, unsigned long long* out,
                        void* stream) {
    (void)cfg;
    if (!bufs || !bufs->stats || !comm || !comm->connected || !out) return MRS_ERR_ARG;
    stats_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(comm->table, comm->rank, comm->world, bufs->stats, out,
                                                               bufs->status);
    return cudaGetLastError() == cudaSuccess ? MRS_OK : MRS_ERR_CUDA;
}

}  // extern "C"
