// NCCL plumbing for the z-slab decomposition: one process per GPU, halo planes
// move GPU-to-GPU over NVLink with ncclSend / ncclRecv pairs inside one group
// (SURVEY.md section 8e).  libnccl is opened at run time, so single-GPU use of
// the library has no NCCL dependency.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

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
