// NCCL plumbing for the z-slab decomposition: one process per GPU, halo planes
// move GPU-to-GPU over NVLink with ncclSend / ncclRecv pairs inside one group
// (SURVEY.md section 8e).  libnccl is opened at run time, so single-GPU use of
// the library has no NCCL dependency.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <utility>
#include <vector>

#include "../../include/emg3d_b200.h"
#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
enum { NCCL_UINT8 = 1, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
} g_nccl;

nccl_comm g_comm = nullptr;
int g_nranks = 1, g_rank = 0;
char g_msg[512];

bool load_nccl() {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return false;
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.Send && g_nccl.Recv &&
           g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.AllReduce;
}

}  // namespace

// provided by api.cu
extern "C" int emg3d_b200_internal_fail(const char* msg);
extern "C" void* emg3d_b200_internal_stream(void);

#define NCK(call)                                                                        \
    do {                                                                                 \
        int _r = (call);                                                                 \
        if (_r != 0) {                                                                   \
            snprintf(g_msg, sizeof g_msg, "%s: %s", #call,                               \
                     g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "NCCL error");  \
            return emg3d_b200_internal_fail(g_msg);                                      \
        }                                                                                \
    } while (0)

extern "C" {

int emg3d_b200_comm_unique_id(void* out128) {
    if (!load_nccl()) return emg3d_b200_internal_fail("libnccl.so.2 could not be loaded");
    nccl_uid id;
    NCK(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return 0;
}

int emg3d_b200_comm_init(const void* unique_id128, int nranks, int rank) {
    if (!load_nccl()) return emg3d_b200_internal_fail("libnccl.so.2 could not be loaded");
    if (!emg3d_b200_internal_stream()) return emg3d_b200_internal_fail("emg3d_b200_init() has not been called");
    if (g_comm) return emg3d_b200_internal_fail("comm_init: communicator already exists");
    nccl_uid id;
    memcpy(&id, unique_id128, sizeof id);
    NCK(g_nccl.CommInitRank(&g_comm, nranks, id, rank));
    g_nranks = nranks;
    g_rank = rank;
    return 0;
}

int emg3d_b200_comm_size(int* nranks, int* rank) {
    *nranks = g_comm ? g_nranks : 1;
    *rank = g_comm ? g_rank : 0;
    return 0;
}

int emg3d_b200_comm_destroy(void) {
    emg3d_b200_p2p_shutdown();
    if (g_comm) {
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_nranks = 1;
    g_rank = 0;
    return 0;
}

// n transfers in one NCCL group on the library stream.  is_send[i] != 0: send
// nbytes[i] from ptrs[i] to peers[i]; else receive into ptrs[i] from peers[i].
int emg3d_b200_comm_sendrecv(int n, void* const* ptrs, const size_t* nbytes, const int* peers,
                             const int* is_send) {
    if (!g_comm) return emg3d_b200_internal_fail("comm_sendrecv: no communicator (comm_init)");
    cudaStream_t st = (cudaStream_t)emg3d_b200_internal_stream();
    NCK(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
        if (nbytes[i] == 0) continue;
        if (is_send[i]) NCK(g_nccl.Send(ptrs[i], nbytes[i], NCCL_UINT8, peers[i], g_comm, st));
        else NCK(g_nccl.Recv(ptrs[i], nbytes[i], NCCL_UINT8, peers[i], g_comm, st));
    }
    NCK(g_nccl.GroupEnd());
    return 0;
}

// in-place sum of n doubles (device memory) over all ranks
int emg3d_b200_comm_allreduce_sum(double* dev, int n) {
    if (!g_comm) return 0;     // single rank: nothing to do
    cudaStream_t st = (cudaStream_t)emg3d_b200_internal_stream();
    NCK(g_nccl.AllReduce(dev, dev, (size_t)n, NCCL_FLOAT64, NCCL_SUM, g_comm, st));
    return 0;
}


}  // extern "C"

// ================================================================================
// Halo exchange over peer memory (NVLink, CUDA IPC): ONE kernel per exchange.
//
// Every rank maps the field arrays and a small flag block of its z-neighbours
// (rank - 1 = "lower", rank + 1 = "upper") into its own address space.  An
// exchange is then a single launch of `halo_pull_kernel`:
//   1. tell both neighbours "my planes are final"          (ready flag, remote store)
//   2. wait for the neighbours' ready flags                 (local poll)
//   3. pull their boundary planes into my halo planes       (remote 16-byte loads)
//   4. last block: tell the neighbours "I have read yours"  (done flag), then wait
//      for their done flags, so that the next kernel in the stream may overwrite
//      my boundary planes.
// PUSH variant (r2, same kernel, `push` = 1): step 1 means "my halo planes may be overwritten",
// step 3 WRITES my boundary planes into the neighbours' halo planes (posted remote stores do not
// wait for a round trip per 16 bytes like remote loads do), step 4 means "my planes have arrived
// at yours" and the final wait is for the neighbours' planes to have arrived here.
// Flags carry a sequence number that lives in device memory and is advanced by the
// kernel itself, so the launch is replayable from a CUDA graph.  Latency per
// exchange: one launch + two NVLink flag round trips, instead of an NCCL group
// of 6-12 send/recv operations (measured r1: 18-61 us per exchange with NCCL).
// ================================================================================
namespace {

typedef unsigned long long u64;

struct P2pSeg {
    const char* src;     // peer memory
    char* dst;           // local halo
    u64 nbytes;
};
constexpr int P2P_MAX_SEG = 8;
struct P2pArgs {
    P2pSeg seg[P2P_MAX_SEG];
    int nseg;
    u64* my_flags;        // [0] ready(lower) [1] ready(upper) [2] done(lower) [3] done(upper)
    u64* peer_flags[2];   // flag blocks of the lower / upper neighbour (null: none)
    u64* seq;             // device: number of the next exchange (starts at 1)
    unsigned* counter;    // device: blocks that finished copying
    int* status;          // device: 1 after a spin timed out
    int push;             // 1: src is local, dst is peer memory
};

__device__ __forceinline__ void st_release_sys(u64* p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_relaxed_sys(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// poll with relaxed loads, acquire once the value is there
__device__ __forceinline__ void spin_until(const u64* p, u64 want, int* status) {
    const long long t0 = clock64();
    while (ld_relaxed_sys(p) < want) {
        if (clock64() - t0 > 20000000000ll) {   // ~10 s: the neighbour is gone
            atomicExch(status, 1);
            break;
        }
    }
    (void)ld_acquire_sys(p);
}

__global__ void __launch_bounds__(256) halo_pull_kernel(P2pArgs a) {
    const u64 seq = *a.seq;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            // my planes are final: everything before this kernel in the stream is done
            if (a.peer_flags[0]) st_release_sys(a.peer_flags[0] + 1, seq);   // I am its upper
            if (a.peer_flags[1]) st_release_sys(a.peer_flags[1] + 0, seq);   // I am its lower
        }
        if (a.peer_flags[0]) spin_until(a.my_flags + 0, seq, a.status);
        if (a.peer_flags[1]) spin_until(a.my_flags + 1, seq, a.status);
    }
    __syncthreads();
    const u64 tid = blockIdx.x * (u64)blockDim.x + threadIdx.x, nth = gridDim.x * (u64)blockDim.x;
    for (int k = 0; k < a.nseg; ++k) {
        const P2pSeg sg = a.seg[k];
        if (((reinterpret_cast<u64>(sg.src) | reinterpret_cast<u64>(sg.dst) | sg.nbytes) & 15) == 0) {
            const int4* src = reinterpret_cast<const int4*>(sg.src);
            int4* dst = reinterpret_cast<int4*>(sg.dst);
            const u64 n = sg.nbytes >> 4;
            u64 i = tid;
            for (; i + 3 * nth < n; i += 4 * nth) {     // four loads in flight per thread
                const int4 v0 = __ldcg(src + i), v1 = __ldcg(src + i + nth),
                           v2 = __ldcg(src + i + 2 * nth), v3 = __ldcg(src + i + 3 * nth);
                dst[i] = v0; dst[i + nth] = v1; dst[i + 2 * nth] = v2; dst[i + 3 * nth] = v3;
            }
            for (; i < n; i += nth) dst[i] = __ldcg(src + i);
        } else {
            const double* src = reinterpret_cast<const double*>(sg.src);
            double* dst = reinterpret_cast<double*>(sg.dst);
            const u64 n = sg.nbytes >> 3;
            for (u64 i = tid; i < n; i += nth) dst[i] = __ldcg(src + i);
        }
    }
    __threadfence_system();                               // (push: the stores went to another GPU)
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.counter, 1u);
        if (prev == gridDim.x - 1) {                     // last block: everything is copied
            __threadfence_system();
            *a.counter = 0;
            *a.seq = seq + 1;
            if (a.peer_flags[0]) st_release_sys(a.peer_flags[0] + 3, seq);
            if (a.peer_flags[1]) st_release_sys(a.peer_flags[1] + 2, seq);
            if (a.peer_flags[0]) spin_until(a.my_flags + 2, seq, a.status);
            if (a.peer_flags[1]) spin_until(a.my_flags + 3, seq, a.status);
        }
    }
}

struct IpcBlob {                 // what a rank publishes about one device pointer
    cudaIpcMemHandle_t handle;   // handle of the allocation that contains it
    u64 offset;                  // pointer - allocation base
};

struct P2pSlot {
    char* local;
    char* peer[2];               // mapped pointer of the lower / upper neighbour's array
};

struct P2p {
    bool on = false;
    u64* flags = nullptr;        // 4 flags + seq (5) + counter/status words
    u64* peer_flags[2] = {nullptr, nullptr};
    std::vector<P2pSlot> slots;
    std::vector<std::pair<IpcBlob, char*>> opened;   // (handle, mapped base)
    size_t n_keep = 0;                               // the first n_keep mappings: the flag blocks
    IpcBlob* stage_dev = nullptr;                    // nranks blobs (device, for ncclAllGather)
} g_p2p;

typedef int (*cuMemGetAddressRange_t)(u64*, size_t*, u64);
cuMemGetAddressRange_t g_addr_range = nullptr;

int make_blob(const void* p, IpcBlob* out) {
    if (!g_addr_range) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
            return 1;
        g_addr_range = (cuMemGetAddressRange_t)fn;
    }
    u64 base = 0;
    size_t size = 0;
    if (g_addr_range(&base, &size, (u64)p) != 0) return 1;
    memset(out, 0, sizeof *out);
    if (cudaIpcGetMemHandle(&out->handle, (void*)base) != cudaSuccess) return 1;
    out->offset = (u64)p - base;
    return 0;
}

char* open_blob(const IpcBlob& b) {
    for (auto& o : g_p2p.opened)
        if (memcmp(&o.first.handle, &b.handle, sizeof b.handle) == 0) return o.second + b.offset;
    void* base = nullptr;
    if (cudaIpcOpenMemHandle(&base, b.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    g_p2p.opened.push_back({b, (char*)base});
    return (char*)base + b.offset;
}

// all ranks publish one pointer; returns the neighbours' mapped pointers
int publish(const void* p, char* peer_out[2]) {
    cudaStream_t st = (cudaStream_t)emg3d_b200_internal_stream();
    std::vector<IpcBlob> all(g_nranks);
    IpcBlob mine;
    int bad = make_blob(p, &mine);
    if (bad) memset(&mine, 0xff, sizeof mine);          // poisoned: everybody falls back
    if (cudaMemcpyAsync(g_p2p.stage_dev + g_rank, &mine, sizeof mine, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return 1;
    NCK(g_nccl.AllGather(g_p2p.stage_dev + g_rank, g_p2p.stage_dev, sizeof(IpcBlob), NCCL_UINT8, g_comm, st));
    if (cudaMemcpyAsync(all.data(), g_p2p.stage_dev, sizeof(IpcBlob) * g_nranks, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return 1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    for (int r = 0; r < g_nranks; ++r)
        if (all[r].offset == ~0ull) return 1;           // somebody could not export
    peer_out[0] = peer_out[1] = nullptr;
    int fail = 0;
    if (g_rank > 0 && !(peer_out[0] = open_blob(all[g_rank - 1]))) fail = 1;
    if (g_rank < g_nranks - 1 && !(peer_out[1] = open_blob(all[g_rank + 1]))) fail = 1;
    // agree on the outcome (a rank that cannot map its neighbour makes everybody fall back)
    double* flag_dev = reinterpret_cast<double*>(g_p2p.stage_dev);
    double f = fail;
    cudaMemcpyAsync(flag_dev, &f, sizeof f, cudaMemcpyHostToDevice, st);
    NCK(g_nccl.AllReduce(flag_dev, flag_dev, 1, NCCL_FLOAT64, NCCL_SUM, g_comm, st));
    cudaMemcpyAsync(&f, flag_dev, sizeof f, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    return f != 0.0;
}

}  // namespace

extern "C" {

// Collective.  Sets *enabled = 1 if every rank could map its neighbours' memory.
int emg3d_b200_p2p_init(int* enabled) {
    *enabled = 0;
    if (!g_comm) return emg3d_b200_internal_fail("p2p_init: no communicator (comm_init)");
    if (!g_nccl.AllGather) return 0;
    if (g_p2p.on) { *enabled = 1; return 0; }
    cudaStream_t st = (cudaStream_t)emg3d_b200_internal_stream();
    if (!g_p2p.flags) {
        if (cudaMalloc(&g_p2p.flags, 64 * sizeof(u64)) != cudaSuccess) return 0;
        if (cudaMalloc(&g_p2p.stage_dev, sizeof(IpcBlob) * (g_nranks + 1)) != cudaSuccess) return 0;
        std::vector<u64> init(64, 0);
        init[4] = 1;                                      // seq
        cudaMemcpyAsync(g_p2p.flags, init.data(), 64 * sizeof(u64), cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
    }
    char* peers[2];
    if (publish(g_p2p.flags, peers)) { cudaGetLastError(); return 0; }
    g_p2p.peer_flags[0] = (u64*)peers[0];
    g_p2p.peer_flags[1] = (u64*)peers[1];
    g_p2p.n_keep = g_p2p.opened.size();
    g_p2p.on = true;
    *enabled = 1;
    return 0;
}

// Forget all registered arrays and unmap the neighbours' (everything but the flag blocks).
// Call on every rank BEFORE the registered arrays are freed, and synchronise the ranks
// before the next registration.
int emg3d_b200_p2p_release(void) {
    if (!g_p2p.on) return 0;
    cudaStreamSynchronize((cudaStream_t)emg3d_b200_internal_stream());
    for (size_t i = g_p2p.n_keep; i < g_p2p.opened.size(); ++i) cudaIpcCloseMemHandle(g_p2p.opened[i].second);
    g_p2p.opened.resize(g_p2p.n_keep);
    g_p2p.slots.clear();
    return 0;
}

// Collective: every rank registers the array it will exchange halos of (same order
// on all ranks).  *slot < 0 if a neighbour's array could not be mapped.
int emg3d_b200_p2p_register(void* dev_ptr, int* slot) {
    *slot = -1;
    if (!g_p2p.on) return emg3d_b200_internal_fail("p2p_register: p2p_init has not succeeded");
    P2pSlot s;
    s.local = (char*)dev_ptr;
    if (publish(dev_ptr, s.peer)) { cudaGetLastError(); return 0; }
    g_p2p.slots.push_back(s);
    *slot = (int)g_p2p.slots.size() - 1;
    return 0;
}

// Pull n segments: nbytes[i] from byte offset peer_off[i] of the array the neighbour
// (from_upper[i] ? rank + 1 : rank - 1) registered in the same slot, to byte offset
// my_off[i] of mine.  One kernel launch on the library stream; see above.
int emg3d_b200_p2p_exchange(int slot, int n, const size_t* my_off, const size_t* peer_off,
                            const size_t* nbytes, const int* from_upper, int push) {
    if (!g_p2p.on || slot < 0 || slot >= (int)g_p2p.slots.size())
        return emg3d_b200_internal_fail("p2p_exchange: unknown slot");
    if (n > P2P_MAX_SEG) return emg3d_b200_internal_fail("p2p_exchange: too many segments");
    const P2pSlot& s = g_p2p.slots[slot];
    P2pArgs a;
    memset(&a, 0, sizeof a);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        const int q = from_upper[i] ? 1 : 0;
        if (!s.peer[q]) return emg3d_b200_internal_fail("p2p_exchange: no such neighbour");
        if (push) {
            a.seg[i].src = s.local + my_off[i];
            a.seg[i].dst = s.peer[q] + peer_off[i];
        } else {
            a.seg[i].src = s.peer[q] + peer_off[i];
            a.seg[i].dst = s.local + my_off[i];
        }
        a.seg[i].nbytes = nbytes[i];
        total += nbytes[i];
    }
    a.nseg = n;
    a.push = push ? 1 : 0;
    a.my_flags = g_p2p.flags;
    a.peer_flags[0] = g_p2p.peer_flags[0];
    a.peer_flags[1] = g_p2p.peer_flags[1];
    a.seq = g_p2p.flags + 4;
    a.counter = reinterpret_cast<unsigned*>(g_p2p.flags + 5);
    a.status = reinterpret_cast<int*>(g_p2p.flags + 6);
    // enough blocks to keep the links busy, few enough to stay one wave (every block
    // polls a flag before copying)
    int blocks = (int)((total / 16 + 1023) / 1024);
    if (blocks < 1) blocks = 1;
    if (blocks > 264) blocks = 264;
    halo_pull_kernel<<<blocks, 256, 0, (cudaStream_t)emg3d_b200_internal_stream()>>>(a);
    if (cudaGetLastError() != cudaSuccess) return emg3d_b200_internal_fail("p2p_exchange: launch failed");
    return 0;
}

// 0: fine; 1: a wait for a neighbour timed out (results are invalid)
int emg3d_b200_p2p_status(int* status) {
    *status = 0;
    if (!g_p2p.on) return 0;
    cudaStream_t st = (cudaStream_t)emg3d_b200_internal_stream();
    if (cudaMemcpyAsync(status, g_p2p.flags + 6, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return emg3d_b200_internal_fail("p2p_status: copy failed");
    return 0;
}

int emg3d_b200_p2p_shutdown(void) {
    if (g_p2p.flags && emg3d_b200_internal_stream()) {
        // back to the initial state (all flags 0, sequence number 1), so that a later
        // communicator in this process starts in step with its new neighbours
        cudaStream_t st = (cudaStream_t)emg3d_b200_internal_stream();
        u64 init[8] = {0, 0, 0, 0, 1, 0, 0, 0};
        cudaStreamSynchronize(st);
        cudaMemcpyAsync(g_p2p.flags, init, sizeof init, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
    }
    for (auto& o : g_p2p.opened) cudaIpcCloseMemHandle(o.second);
    g_p2p.opened.clear();
    g_p2p.slots.clear();
    g_p2p.peer_flags[0] = g_p2p.peer_flags[1] = nullptr;
    g_p2p.on = false;
    return 0;
}

}  // extern "C"
