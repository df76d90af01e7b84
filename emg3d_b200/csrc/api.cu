// C ABI of the emg3d_b200 library (see include/emg3d_b200.h).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/emg3d_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace emg;

namespace emg {
long long g_launch_count = 0;
}

static thread_local std::string g_err;
static cudaStream_t g_stream = nullptr;
static int g_device = -1;
static double* g_dot_scratch = nullptr;   // partials for dot products
static double* g_dot_out = nullptr;       // 2 doubles
static int64_t g_dot_scratch_n = 0;

static int fail(const char* where, cudaError_t e) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    g_err = buf;
    return (int)e ? (int)e : -1;
}
static int fail_msg(const char* msg) {
    g_err = msg;
    return -1;
}
#define CK(call)                                          \
    do {                                                  \
        cudaError_t _e = (call);                          \
        if (_e != cudaSuccess) return fail(#call, _e);    \
    } while (0)
#define CK_LAUNCH(name)                                   \
    do {                                                  \
        cudaError_t _e = cudaPeekAtLastError();           \
        if (_e != cudaSuccess) {                          \
            cudaGetLastError();                           \
            return fail(name, _e);                        \
        }                                                 \
    } while (0)
#define NEED_INIT()                                                              \
    do {                                                                         \
        if (!g_stream) return fail_msg("emg3d_b200_init() has not been called"); \
    } while (0)

struct emg3d_b200_level {
    Dims d;
    double* h[3];        // device
    double* rh[3];       // device
    int cplx;            // -1 until a model is set
    const void* eta[3];
    const double* zeta;
    void* fac[3];        // cached line factorisations (device), per direction
    void* fac2[3];       // cached data of the segment-parallel line kernels (gs_line_seg.cu)
    void* chain_in[3];   // lines cut by multi-GPU slabs: factors entering from the lower rank (owned)
    int chained[3];      // the direction's lines continue on other ranks (emg3d_b200_level_line_chain)
    void* diag;          // cached diagonal of A per edge (device), point smoother
    double* scratch;     // residual-norm partials
    double* norm2;       // device scalar
    // link to the parent (fine) level
    Dims fine;
    int linked;
    int cflag[3];
    double* w[9];        // device weights
    int* lo[3];
    double* fr[3];
    int is_window;       // z-window of another level: h / rh are borrowed
};

// Every live level, so that a failed allocation can reclaim cached line
// factorisations anywhere in the hierarchy (they are recomputed on demand).
static std::vector<emg3d_b200_level*> g_levels;
// levels at or below this many cells may be baked into CUDA graphs of the host
// driver (solver._GRAPH_MAX_CELLS): their cached arrays must keep their addresses
static const int64_t EVICT_MIN_CELLS = (int64_t)160 * 160 * 160 + 1;

// cudaMalloc that, when out of memory, frees cached factorisations -- largest
// first, never those of `keep` in direction `keep_dir` -- and retries.
static cudaError_t malloc_evicting(void** p, size_t nbytes, const emg3d_b200_level* keep, int keep_dir) {
    cudaError_t e = cudaMalloc(p, nbytes);
    while (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        emg3d_b200_level* best = nullptr;
        int best_dir = -1;
        size_t best_bytes = 0;
        for (emg3d_b200_level* lv : g_levels) {
            if (n_cells(lv->d) < EVICT_MIN_CELLS) continue;
            const size_t el = lv->cplx == 0 ? sizeof(double) : sizeof(cplx);
            for (int a = 0; a < 3; ++a) {
                if (lv == keep && a == keep_dir) continue;
                if (lv->fac[a]) {
                    const size_t b = (size_t)line_factor_elems(lv->d, a) * el;
                    if (b > best_bytes) { best = lv; best_dir = a; best_bytes = b; }
                }
                if (lv->fac2[a]) {
                    const size_t b = (size_t)line_seg_elems(lv->d, a) * el;
                    if (b > best_bytes) { best = lv; best_dir = a + 3; best_bytes = b; }
                }
            }
        }
        if (!best) return cudaErrorMemoryAllocation;
        cudaStreamSynchronize(g_stream);
        void** victim = best_dir < 3 ? &best->fac[best_dir] : &best->fac2[best_dir - 3];
        cudaFree(*victim);
        *victim = nullptr;
        e = cudaMalloc(p, nbytes);
    }
    return e;
}

template <typename T>
static Model<T> model_of(const emg3d_b200_level* lv) {
    Model<T> m;
    m.d = lv->d;
    for (int a = 0; a < 3; ++a) {
        m.eta[a] = (const T*)lv->eta[a];
        m.h[a] = lv->h[a];
        m.rh[a] = lv->rh[a];
    }
    m.zeta = lv->zeta;
    // the point smoother addresses the diagonal relative to the x-component of the
    // (windowed) field view, so the window offset of that component is applied here
    m.diag = lv->diag ? (const T*)lv->diag + (int64_t)lv->d.n[0] * (lv->d.n[1] + 1) * lv->d.zoff : nullptr;
    return m;
}

extern "C" {

// hooks for comm.cu (not part of the public header)
int emg3d_b200_internal_fail(const char* msg) { return fail_msg(msg); }
void* emg3d_b200_internal_stream(void) { return (void*)g_stream; }

int emg3d_b200_abi_version(void) { return 1; }

const char* emg3d_b200_last_error(void) { return g_err.c_str(); }

int emg3d_b200_device_count(int* count) {
    CK(cudaGetDeviceCount(count));
    return 0;
}

int emg3d_b200_init(int device) {
    if (g_stream && device == g_device) return 0;
    if (g_stream) return fail_msg("emg3d_b200_init: already initialised on another device");
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_device = device;
    g_dot_scratch_n = dot_scratch_doubles((int64_t)1 << 40);
    CK(cudaMalloc(&g_dot_scratch, sizeof(double) * g_dot_scratch_n));
    CK(cudaMalloc(&g_dot_out, sizeof(double) * 2));
    return 0;
}

int emg3d_b200_device_name(char* buf, int buflen) {
    NEED_INIT();
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, g_device));
    snprintf(buf, buflen, "%s (sm_%d%d, %d SMs)", p.name, p.major, p.minor, p.multiProcessorCount);
    return 0;
}

int emg3d_b200_mem_info(size_t* free_bytes, size_t* total_bytes) {
    NEED_INIT();
    CK(cudaMemGetInfo(free_bytes, total_bytes));
    return 0;
}

int emg3d_b200_sync(void) {
    NEED_INIT();
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int emg3d_b200_launch_count(long long* count) {
    *count = emg::g_launch_count;
    return 0;
}

// ---- memory ------------------------------------------------------------------
int emg3d_b200_malloc(void** dptr, size_t nbytes) {
    NEED_INIT();
    *dptr = nullptr;
    {
        cudaError_t me = malloc_evicting(dptr, nbytes ? nbytes : 16, nullptr, -1);
        if (me != cudaSuccess) return fail("cudaMalloc", me);
    }
    return 0;
}
int emg3d_b200_free(void* dptr) {
    if (!dptr) return 0;
    CK(cudaFree(dptr));
    return 0;
}
// Short-lived work arrays (interpolation tables, spline coefficients, sampled values): stream-
// ordered allocations from the device's memory pool.  cudaMalloc / cudaFree synchronise the
// device and cost milliseconds next to a 10 ms cycle (measured: 32 ms per cudaFree with the
// field arrays of a 256^3 solve resident); the pool keeps its memory between calls.
int emg3d_b200_malloc_scratch(void** dptr, size_t nbytes) {
    NEED_INIT();
    static bool pool_set = false;
    if (!pool_set) {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        pool_set = true;
    }
    *dptr = nullptr;
    cudaError_t me = cudaMallocAsync(dptr, nbytes ? nbytes : 16, g_stream);
    if (me != cudaSuccess) {
        // out of memory: hand the pool's idle blocks back and let the evicting allocator make room
        // (cached line factorisations are recomputed on demand), then try once more
        cudaGetLastError();
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            cudaStreamSynchronize(g_stream);
            cudaMemPoolTrimTo(pool, 0);
        }
        void* probe = nullptr;
        if (malloc_evicting(&probe, nbytes ? nbytes : 16, nullptr, -1) == cudaSuccess) cudaFree(probe);
        cudaGetLastError();
        me = cudaMallocAsync(dptr, nbytes ? nbytes : 16, g_stream);
        if (me != cudaSuccess) return fail("cudaMallocAsync", me);
    }
    return 0;
}
int emg3d_b200_free_scratch(void* dptr) {
    if (!dptr) return 0;
    CK(cudaFreeAsync(dptr, g_stream));
    return 0;
}
int emg3d_b200_memset(void* dptr, int byte, size_t nbytes) {
    NEED_INIT();
    CK(cudaMemsetAsync(dptr, byte, nbytes, g_stream));
    return 0;
}
int emg3d_b200_h2d(void* dst, const void* src, size_t nbytes) {
    NEED_INIT();
    CK(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}
int emg3d_b200_d2h(void* dst, const void* src, size_t nbytes) {
    NEED_INIT();
    CK(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}
int emg3d_b200_d2d(void* dst, const void* src, size_t nbytes) {
    NEED_INIT();
    CK(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, g_stream));
    return 0;
}
// ---- sparse upload ---------------------------------------------------------------
// Source fields of dipoles and wires are constant (zero, or -0.0 after the scaling
// by -s mu_0) except on a handful of edges (emg3d/fields.py:386-519), yet they
// arrive as dense host arrays of n_edges values (0.8 GB at 256^3).  The host array
// is compared against its background BIT PATTERN by a few threads at memory
// bandwidth; if only a small fraction differs, the device array is filled with the
// background and the (index, value) pairs are scattered into it, so that only
// those cross PCIe.  The device array is bit-identical to a plain copy.
extern "C++" {
namespace {
template <typename V>
__global__ void scatter_kernel(V* __restrict__ dst, const long long* __restrict__ idx,
                               const V* __restrict__ val, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = val[i];
}
template <typename V>
__global__ void fill_kernel(V* __restrict__ dst, V v, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) dst[i] = v;
}
struct SparseStage {
    void* host = nullptr;     // pinned: [idx (8 B) x cap | values (elsize) x cap]
    void* dev = nullptr;
    size_t cap = 0, elsize = 0;
} g_stage;

// indices of the elements (WPE 64-bit words each) that differ from bg; gives up
// (returns false) as soon as more than cap were found
template <int WPE>
bool scan_sparse(const uint64_t* w, size_t n, const uint64_t* bg, size_t cap, int nt,
                 std::vector<std::vector<long long>>& found) {
    std::atomic<int> dense(0);
    auto scan = [&](int t) {
        const size_t e0 = n * t / nt, e1 = n * (t + 1) / nt;
        std::vector<long long>& out = found[t];
        const size_t B = 64 / WPE;                    // elements per 512-byte block
        for (size_t e = e0; e < e1;) {
            const size_t eb = e + B < e1 ? e + B : e1;
            uint64_t any = 0;
            for (size_t i = e; i < eb; ++i)
                for (int k = 0; k < WPE; ++k) any |= w[i * WPE + k] ^ bg[k];
            if (any) {
                for (size_t i = e; i < eb; ++i) {
                    uint64_t a = 0;
                    for (int k = 0; k < WPE; ++k) a |= w[i * WPE + k] ^ bg[k];
                    if (a) out.push_back((long long)i);
                }
                if (out.size() > cap) { dense.store(1); return; }
            }
            e = eb;
            if ((e & 0xfffff) < B && dense.load(std::memory_order_relaxed)) return;
        }
    };
    if (nt == 1) scan(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(scan, t);
        for (auto& x : th) x.join();
    }
    size_t cnt = 0;
    for (auto& v : found) cnt += v.size();
    return !dense.load() && cnt <= cap;
}
}  // namespace
}  // extern "C++"

int emg3d_b200_h2d_sparse(void* dst, const void* src, size_t n, int elsize, int* used_sparse) {
    NEED_INIT();
    if (elsize != 8 && elsize != 16) return fail_msg("h2d_sparse: elsize must be 8 or 16");
    if (used_sparse) *used_sparse = 0;
    if (n == 0) return 0;
    const size_t cap = (n / 64 + 1024) & ~(size_t)1;   // beyond ~1.5 % a plain copy is as good
    const int wpe = elsize / 8;                        // 64-bit words per element
    const uint64_t* w = (const uint64_t*)src;
    // background = the most frequent pattern among five probes
    uint64_t bg[2] = {0, 0};
    {
        const size_t cand[5] = {0, n / 4, n / 2, n - 1 - n / 4, n - 1};
        int pick = 0, best = -1;
        for (int i = 0; i < 5; ++i) {
            int votes = 0;
            for (int j = 0; j < 5; ++j)
                votes += memcmp(w + cand[i] * wpe, w + cand[j] * wpe, elsize) == 0;
            if (votes > best) { best = votes; pick = i; }
        }
        memcpy(bg, w + cand[pick] * wpe, elsize);
    }
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)(hw ? hw : 8);
    if (nt > 32) nt = 32;
    if (n < (size_t)1 << 20) nt = 1;
    std::vector<std::vector<long long>> found(nt);
    const bool sparse = wpe == 2 ? scan_sparse<2>(w, n, bg, cap, nt, found)
                                 : scan_sparse<1>(w, n, bg, cap, nt, found);
    size_t cnt = 0;
    for (auto& v : found) cnt += v.size();
    if (getenv("EMG3D_B200_DEBUG"))
        fprintf(stderr, "h2d_sparse: n=%zu elsize=%d threads=%d differing=%zu cap=%zu sparse=%d\n", n,
                elsize, nt, cnt, cap, (int)sparse);
    if (!sparse) return emg3d_b200_h2d(dst, src, n * (size_t)elsize);
    emg3d_b200_sync();                                // staging buffer free again
    if (g_stage.cap < cap || g_stage.elsize != (size_t)elsize) {
        if (g_stage.host) cudaFreeHost(g_stage.host);
        if (g_stage.dev) cudaFree(g_stage.dev);
        g_stage.host = g_stage.dev = nullptr;
        g_stage.cap = 0;
        CK(cudaMallocHost(&g_stage.host, cap * (8 + (size_t)elsize)));
        CK(cudaMalloc(&g_stage.dev, cap * (8 + (size_t)elsize)));
        g_stage.cap = cap;
        g_stage.elsize = elsize;
    }
    long long* hidx = (long long*)g_stage.host;
    char* hval = (char*)g_stage.host + 8 * g_stage.cap;
    size_t k = 0;
    for (auto& v : found)
        for (long long i : v) {
            hidx[k] = i;
            memcpy(hval + k * elsize, (const char*)src + (size_t)i * elsize, elsize);
            ++k;
        }
    const int bs = 256;
    ++emg::g_launch_count;
    if (elsize == 16) {
        double2 v;
        memcpy(&v, bg, 16);
        fill_kernel<double2><<<148 * 8, bs, 0, g_stream>>>((double2*)dst, v, (long long)n);
    } else {
        double v;
        memcpy(&v, bg, 8);
        fill_kernel<double><<<148 * 8, bs, 0, g_stream>>>((double*)dst, v, (long long)n);
    }
    CK_LAUNCH("fill");
    if (cnt) {
        CK(cudaMemcpyAsync(g_stage.dev, hidx, 8 * cnt, cudaMemcpyHostToDevice, g_stream));
        char* dval = (char*)g_stage.dev + 8 * g_stage.cap;
        CK(cudaMemcpyAsync(dval, hval, cnt * (size_t)elsize, cudaMemcpyHostToDevice, g_stream));
        const unsigned gs = (unsigned)((cnt + bs - 1) / bs);
        ++emg::g_launch_count;
        if (elsize == 16)
            scatter_kernel<double2><<<gs, bs, 0, g_stream>>>((double2*)dst, (const long long*)g_stage.dev,
                                                             (const double2*)dval, (long long)cnt);
        else
            scatter_kernel<double><<<gs, bs, 0, g_stream>>>((double*)dst, (const long long*)g_stage.dev,
                                                            (const double*)dval, (long long)cnt);
        CK_LAUNCH("scatter");
    }
    CK(cudaStreamSynchronize(g_stream));              // the staging buffer is reused
    if (used_sparse) *used_sparse = 1;
    return 0;
}

// A source field given as what it is: a constant background and `count` (index, value) pairs
// (emg3d_b200/fields.py: SourceField) -- no dense host array is ever built or scanned.
int emg3d_b200_fill_scatter(void* dst, size_t n, int elsize, const void* fill_value, const long long* idx,
                            const void* val, size_t count) {
    NEED_INIT();
    if (elsize != 8 && elsize != 16) return fail_msg("fill_scatter: elsize must be 8 or 16");
    if (n == 0) return 0;
    for (size_t k = 0; k < count; ++k)
        if (idx[k] < 0 || (size_t)idx[k] >= n) return fail_msg("fill_scatter: index out of range");
    const int bs = 256;
    ++emg::g_launch_count;
    if (elsize == 16) {
        double2 v;
        memcpy(&v, fill_value, 16);
        fill_kernel<double2><<<148 * 8, bs, 0, g_stream>>>((double2*)dst, v, (long long)n);
    } else {
        double v;
        memcpy(&v, fill_value, 8);
        fill_kernel<double><<<148 * 8, bs, 0, g_stream>>>((double*)dst, v, (long long)n);
    }
    CK_LAUNCH("fill");
    if (count) {
        void* tmp = nullptr;
        CK(cudaMalloc(&tmp, count * (8 + (size_t)elsize)));
        char* dval = (char*)tmp + 8 * count;
        CK(cudaMemcpyAsync(tmp, idx, 8 * count, cudaMemcpyHostToDevice, g_stream));
        CK(cudaMemcpyAsync(dval, val, count * (size_t)elsize, cudaMemcpyHostToDevice, g_stream));
        const unsigned gs = (unsigned)((count + bs - 1) / bs);
        ++emg::g_launch_count;
        if (elsize == 16)
            scatter_kernel<double2><<<gs, bs, 0, g_stream>>>((double2*)dst, (const long long*)tmp, (const double2*)dval,
                                                             (long long)count);
        else
            scatter_kernel<double><<<gs, bs, 0, g_stream>>>((double*)dst, (const long long*)tmp, (const double*)dval,
                                                            (long long)count);
        CK_LAUNCH("scatter");
        CK(cudaStreamSynchronize(g_stream));
        cudaFree(tmp);
    }
    return 0;
}

int emg3d_b200_host_alloc(void** hptr, size_t nbytes) {
    CK(cudaMallocHost(hptr, nbytes ? nbytes : 16));
    return 0;
}
int emg3d_b200_host_free(void* hptr) {
    if (!hptr) return 0;
    CK(cudaFreeHost(hptr));
    return 0;
}

// ---- events / graphs ---------------------------------------------------------
int emg3d_b200_event_create(void** ev) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    *ev = (void*)e;
    return 0;
}
int emg3d_b200_event_record(void* ev) {
    NEED_INIT();
    CK(cudaEventRecord((cudaEvent_t)ev, g_stream));
    return 0;
}
int emg3d_b200_event_elapsed_ms(void* a, void* b, float* ms) {
    CK(cudaEventSynchronize((cudaEvent_t)b));
    CK(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
    return 0;
}
int emg3d_b200_event_destroy(void* ev) {
    CK(cudaEventDestroy((cudaEvent_t)ev));
    return 0;
}
// A captured graph remembers how many kernel launches it holds, so that the
// launch counter keeps counting kernels when a graph is replayed.
struct GraphHandle {
    cudaGraphExec_t exec;
    long long launches;
};
static long long g_capture_start = 0;

int emg3d_b200_graph_begin(void) {
    NEED_INIT();
    g_capture_start = emg::g_launch_count;
    CK(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
    return 0;
}
int emg3d_b200_graph_end(void** graph_exec) {
    NEED_INIT();
    cudaGraph_t g;
    CK(cudaStreamEndCapture(g_stream, &g));
    cudaGraphExec_t ge;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail("cudaGraphInstantiate", e);
    GraphHandle* h = new GraphHandle{ge, emg::g_launch_count - g_capture_start};
    emg::g_launch_count = g_capture_start;            // nothing ran during the capture
    *graph_exec = (void*)h;
    return 0;
}
int emg3d_b200_graph_launch(void* graph_exec) {
    NEED_INIT();
    GraphHandle* h = (GraphHandle*)graph_exec;
    CK(cudaGraphLaunch(h->exec, g_stream));
    emg::g_launch_count += h->launches;
    return 0;
}
int emg3d_b200_graph_destroy(void* graph_exec) {
    GraphHandle* h = (GraphHandle*)graph_exec;
    if (!h) return 0;
    cudaGraphExecDestroy(h->exec);
    delete h;
    return 0;
}

// ---- levels ------------------------------------------------------------------
int emg3d_b200_level_create(emg3d_b200_level** out, int nx, int ny, int nz, const double* hx,
                            const double* hy, const double* hz) {
    NEED_INIT();
    if (nx < 1 || ny < 1 || nz < 1) return fail_msg("level_create: need at least 1 cell per axis");
    emg3d_b200_level* lv = new emg3d_b200_level();
    memset(lv, 0, sizeof *lv);
    lv->d.n[0] = nx; lv->d.n[1] = ny; lv->d.n[2] = nz;
    lv->cplx = -1;
    const double* hh[3] = {hx, hy, hz};
    for (int a = 0; a < 3; ++a) {
        const int n = lv->d.n[a];
        std::vector<double> r(n);
        for (int i = 0; i < n; ++i) r[i] = 1.0 / hh[a][i];
        CK(cudaMalloc(&lv->h[a], sizeof(double) * n));
        CK(cudaMalloc(&lv->rh[a], sizeof(double) * n));
        CK(cudaMemcpyAsync(lv->h[a], hh[a], sizeof(double) * n, cudaMemcpyHostToDevice, g_stream));
        CK(cudaMemcpyAsync(lv->rh[a], r.data(), sizeof(double) * n, cudaMemcpyHostToDevice, g_stream));
        CK(cudaStreamSynchronize(g_stream));
    }
    CK(cudaMalloc(&lv->scratch, sizeof(double) * residual_scratch_doubles(lv->d)));
    CK(cudaMalloc(&lv->norm2, sizeof(double) * 2));
    g_levels.push_back(lv);
    *out = lv;
    return 0;
}

int emg3d_b200_level_drop_factors(emg3d_b200_level* lv) {
    if (lv->diag) cudaFree(lv->diag);
    lv->diag = nullptr;
    for (int a = 0; a < 3; ++a) {
        if (lv->fac[a]) cudaFree(lv->fac[a]);
        lv->fac[a] = nullptr;
        if (lv->fac2[a]) cudaFree(lv->fac2[a]);
        lv->fac2[a] = nullptr;
    }
    return 0;
}

int emg3d_b200_level_destroy(emg3d_b200_level* lv) {
    if (!lv) return 0;
    for (size_t i = 0; i < g_levels.size(); ++i)
        if (g_levels[i] == lv) { g_levels.erase(g_levels.begin() + i); break; }
    emg3d_b200_level_drop_factors(lv);
    for (int a = 0; a < 3; ++a) {
        if (lv->chain_in[a]) cudaFree(lv->chain_in[a]);
        if (!lv->is_window) {
            cudaFree(lv->h[a]);
            cudaFree(lv->rh[a]);
        }
        cudaFree(lv->lo[a]);
        cudaFree(lv->fr[a]);
    }
    for (int k = 0; k < 9; ++k) cudaFree(lv->w[k]);
    cudaFree(lv->scratch);
    cudaFree(lv->norm2);
    delete lv;
    return 0;
}

int emg3d_b200_level_set_model(emg3d_b200_level* lv, int cplx, const void* eta_x, const void* eta_y,
                               const void* eta_z, const double* zeta) {
    NEED_INIT();
    emg3d_b200_level_drop_factors(lv);
    lv->cplx = cplx ? 1 : 0;
    lv->eta[0] = eta_x; lv->eta[1] = eta_y; lv->eta[2] = eta_z;
    lv->zeta = zeta;
    return 0;
}

int emg3d_b200_level_window(emg3d_b200_level** out, const emg3d_b200_level* parent, int z0, int nz) {
    NEED_INIT();
    if (parent->is_window) return fail_msg("level_window: parent is a window itself");
    if (parent->cplx < 0) return fail_msg("level_window: parent has no model");
    if (z0 < 0 || nz < 1 || z0 + nz > parent->d.n[2]) return fail_msg("level_window: range outside the parent");
    emg3d_b200_level* lv = new emg3d_b200_level();
    memset(lv, 0, sizeof *lv);
    lv->is_window = 1;
    lv->d = parent->d;
    lv->d.n[2] = nz;
    lv->d.zoff = z0;
    lv->d.nzf = parent->d.n[2];
    for (int a = 0; a < 3; ++a) {
        lv->h[a] = parent->h[a] + (a == 2 ? z0 : 0);
        lv->rh[a] = parent->rh[a] + (a == 2 ? z0 : 0);
    }
    lv->cplx = parent->cplx;
    const size_t el = parent->cplx ? sizeof(cplx) : sizeof(double);
    const size_t coff = (size_t)parent->d.n[0] * parent->d.n[1] * z0;
    for (int a = 0; a < 3; ++a) lv->eta[a] = (const char*)parent->eta[a] + coff * el;
    lv->zeta = parent->zeta + coff;
    CK(cudaMalloc(&lv->scratch, sizeof(double) * residual_scratch_doubles(lv->d)));
    CK(cudaMalloc(&lv->norm2, sizeof(double) * 2));
    g_levels.push_back(lv);
    *out = lv;
    return 0;
}

int emg3d_b200_level_set_owned(emg3d_b200_level* lv, int plane0, int plane1) {
    if (plane0 < 0 || plane1 < plane0 || plane1 > lv->d.n[2] + 1)
        return fail_msg("level_set_owned: plane range outside the level");
    lv->d.own0 = plane0;
    lv->d.own1 = plane1;
    return 0;
}

int emg3d_b200_level_set_zflip(emg3d_b200_level* lv, int flip) {
    lv->d.zflip = flip ? 1 : 0;
    return 0;
}

int emg3d_b200_point_schedule_kind(const emg3d_b200_level* lv, int* kind) {
    const int64_t nint = (int64_t)(lv->d.n[0] - 1) * (lv->d.n[1] - 1) * (lv->d.n[2] - 1);
    *kind = nint <= SMALL_GRID_NODES ? 0 : nint > TILE_MIN_NODES ? 2 : 1;
    return 0;
}

int emg3d_b200_level_factor_bytes(const emg3d_b200_level* lv, int ldir, size_t* nbytes) {
    if (ldir < 1 || ldir > 3) return fail_msg("level_factor_bytes: ldir must be 1, 2 or 3");
    const size_t el = lv->cplx == 0 ? sizeof(double) : sizeof(cplx);
    *nbytes = (size_t)line_factor_elems(lv->d, ldir - 1) * el;
    return 0;
}

int emg3d_b200_level_link(emg3d_b200_level* c, const emg3d_b200_level* f, const int* cflag,
                          const double* const* weights, const int* const* lo,
                          const double* const* frac) {
    NEED_INIT();
    for (int a = 0; a < 3; ++a) {
        const int expect = cflag[a] ? f->d.n[a] / 2 : f->d.n[a];
        if (cflag[a] && (f->d.n[a] % 2)) return fail_msg("level_link: odd cell count cannot be coarsened");
        if (c->d.n[a] != expect) return fail_msg("level_link: coarse shape does not match fine shape and cflag");
    }
    c->fine = f->d;
    for (int a = 0; a < 3; ++a) {
        c->cflag[a] = cflag[a] ? 1 : 0;
        const int ncn = c->d.n[a] + 1, nfn = f->d.n[a] + 1;
        for (int k = 0; k < 3; ++k) {
            cudaFree(c->w[3 * a + k]);
            c->w[3 * a + k] = nullptr;
            if (cflag[a]) {
                if (!weights || !weights[3 * a + k]) return fail_msg("level_link: missing restriction weights");
                CK(cudaMalloc(&c->w[3 * a + k], sizeof(double) * ncn));
                CK(cudaMemcpyAsync(c->w[3 * a + k], weights[3 * a + k], sizeof(double) * ncn,
                                   cudaMemcpyHostToDevice, g_stream));
            }
        }
        cudaFree(c->lo[a]);
        cudaFree(c->fr[a]);
        CK(cudaMalloc(&c->lo[a], sizeof(int) * nfn));
        CK(cudaMalloc(&c->fr[a], sizeof(double) * nfn));
        CK(cudaMemcpyAsync(c->lo[a], lo[a], sizeof(int) * nfn, cudaMemcpyHostToDevice, g_stream));
        CK(cudaMemcpyAsync(c->fr[a], frac[a], sizeof(double) * nfn, cudaMemcpyHostToDevice, g_stream));
    }
    CK(cudaStreamSynchronize(g_stream));
    c->linked = 1;
    return 0;
}

// ---- kernels -------------------------------------------------------------------
#define NEED_MODEL(lv)                                                            \
    do {                                                                          \
        NEED_INIT();                                                              \
        if ((lv)->cplx < 0) return fail_msg("level has no model (level_set_model)"); \
    } while (0)

static int residual_impl(emg3d_b200_level* lv, const void* s, const void* e, void* r, double* norm2_dev,
                         int apply_only) {
    NEED_MODEL(lv);
    if (lv->cplx)
        launch_residual<cplx>(model_of<cplx>(lv), (const cplx*)s, (const cplx*)e, (cplx*)r, norm2_dev,
                              lv->scratch, apply_only, g_stream);
    else
        launch_residual<double>(model_of<double>(lv), (const double*)s, (const double*)e, (double*)r,
                                norm2_dev, lv->scratch, apply_only, g_stream);
    CK_LAUNCH("residual");
    return 0;
}

int emg3d_b200_residual(emg3d_b200_level* lv, const void* s, const void* e, void* r, double* norm2_dev) {
    return residual_impl(lv, s, e, r, norm2_dev, 0);
}

int emg3d_b200_apply(emg3d_b200_level* lv, const void* e, void* out) {
    return residual_impl(lv, nullptr, e, out, nullptr, 1);
}

int emg3d_b200_amat_x(emg3d_b200_level* lv, void* r, const void* e) {
    return emg3d_b200_residual(lv, r, e, r, nullptr);
}

int emg3d_b200_residual_norm(emg3d_b200_level* lv, const void* s, const void* e, void* r, double* norm_host) {
    int rc = emg3d_b200_residual(lv, s, e, r, lv->norm2);
    if (rc) return rc;
    double v = 0.0;
    CK(cudaMemcpyAsync(&v, lv->norm2, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    *norm_host = sqrt(v);
    return 0;
}

int emg3d_b200_gauss_seidel(emg3d_b200_level* lv, void* e, const void* s, int nu, int ldir, int order) {
    NEED_MODEL(lv);
    if (ldir < 0 || ldir > 3) return fail_msg("gauss_seidel: ldir must be 0..3");
    // bits 8+ of `order` carry the number of sweeps already done (sweep direction phase)
    if ((order & 0xff) != ORDER_LEX && (order & 0xff) != ORDER_COLOR)
        return fail_msg("gauss_seidel: unknown order");
    if (nu <= 0) return 0;
    if (ldir == 0) {
        if (!lv->diag) {
            const size_t el = lv->cplx ? sizeof(cplx) : sizeof(double);
            CK(cudaMalloc(&lv->diag, (size_t)n_edges(lv->d) * el));
            if (lv->cplx) launch_edge_diag<cplx>(model_of<cplx>(lv), (cplx*)lv->diag, g_stream);
            else launch_edge_diag<double>(model_of<double>(lv), (double*)lv->diag, g_stream);
            CK_LAUNCH("edge_diag");
        }
        if (lv->cplx)
            launch_gs_point<cplx>(model_of<cplx>(lv), (cplx*)e, (const cplx*)s, nu, order, g_stream);
        else
            launch_gs_point<double>(model_of<double>(lv), (double*)e, (const double*)s, nu, order, g_stream);
        CK_LAUNCH("gauss_seidel");
        return 0;
    }
    const int dir = ldir - 1;
    const size_t el = lv->cplx ? sizeof(cplx) : sizeof(double);
    // long lines in multicolour order: segment-parallel kernels with their own cached data
    const bool seg = (order & 0xff) == ORDER_COLOR && !lv->chained[dir] && line_seg_elems(lv->d, dir) > 0;
    if (seg && !lv->fac2[dir]) {
        const size_t nbytes = (size_t)line_seg_elems(lv->d, dir) * el;
        cudaError_t me = malloc_evicting(&lv->fac2[dir], nbytes, lv, dir);
        if (me != cudaSuccess) { lv->fac2[dir] = nullptr; return fail("cudaMalloc (line factorisation)", me); }
        if (lv->cplx)
            launch_line_seg_factor<cplx>(model_of<cplx>(lv), dir, (cplx*)lv->fac2[dir], g_stream);
        else
            launch_line_seg_factor<double>(model_of<double>(lv), dir, (double*)lv->fac2[dir], g_stream);
        CK_LAUNCH("line_seg_factor");
    }
    if (!seg && !lv->fac[dir]) {
        size_t nbytes;
        emg3d_b200_level_factor_bytes(lv, ldir, &nbytes);
        if (nbytes == 0) return 0;
        {
            // out of memory: reclaim cached factorisations elsewhere (recomputed when needed)
            cudaError_t me = malloc_evicting(&lv->fac[dir], nbytes, lv, dir);
            if (me != cudaSuccess) { lv->fac[dir] = nullptr; return fail("cudaMalloc (line factorisation)", me); }
        }
        // (an evicted factorisation of a chained direction is rebuilt from the kept incoming factors)
        if (lv->cplx)
            launch_line_factor<cplx>(model_of<cplx>(lv), dir, (cplx*)lv->fac[dir], (const cplx*)lv->chain_in[dir],
                                     nullptr, g_stream);
        else
            launch_line_factor<double>(model_of<double>(lv), dir, (double*)lv->fac[dir],
                                       (const double*)lv->chain_in[dir], nullptr, g_stream);
        CK_LAUNCH("line_factor");
    }
    const void* f2 = seg ? lv->fac2[dir] : nullptr;
    if (lv->cplx)
        launch_gs_line<cplx>(model_of<cplx>(lv), dir, (const cplx*)lv->fac[dir], (const cplx*)f2, (cplx*)e,
                             (const cplx*)s, nu, order, g_stream);
    else
        launch_gs_line<double>(model_of<double>(lv), dir, (const double*)lv->fac[dir], (const double*)f2,
                               (double*)e, (const double*)s, nu, order, g_stream);
    CK_LAUNCH("gauss_seidel_line");
    return 0;
}

// Lines of direction ldir that continue on the neighbouring ranks of a multi-GPU slab
// decomposition: (re)factorise them as pieces of the global lines.  chain_in (device, or NULL
// on the first rank): [line slot][10] factors of the last block of the lower rank's pieces;
// chain_out (device, or NULL): receives those of this rank's last blocks.  *n_elems: elements
// (of the level's dtype) of either buffer.  The incoming factors are kept (copied) so that an
// evicted factorisation can be rebuilt.
int emg3d_b200_level_line_chain(emg3d_b200_level* lv, int ldir, const void* chain_in, void* chain_out,
                                size_t* n_elems) {
    NEED_MODEL(lv);
    if (ldir < 1 || ldir > 3) return fail_msg("level_line_chain: ldir must be 1, 2 or 3");
    const int dir = ldir - 1;
    const size_t el = lv->cplx ? sizeof(cplx) : sizeof(double);
    const size_t ne = (size_t)line_chain_elems(lv->d, dir);
    if (n_elems) *n_elems = ne;
    if (!chain_in && !chain_out) return 0;                  // size query
    lv->chained[dir] = 1;
    if (chain_in) {
        if (!lv->chain_in[dir]) CK(cudaMalloc(&lv->chain_in[dir], ne * el));
        CK(cudaMemcpyAsync(lv->chain_in[dir], chain_in, ne * el, cudaMemcpyDeviceToDevice, g_stream));
    }
    if (!lv->fac[dir]) {
        size_t nbytes;
        emg3d_b200_level_factor_bytes(lv, ldir, &nbytes);
        if (nbytes == 0) return 0;
        cudaError_t me = malloc_evicting(&lv->fac[dir], nbytes, lv, dir);
        if (me != cudaSuccess) { lv->fac[dir] = nullptr; return fail("cudaMalloc (line factorisation)", me); }
    }
    if (lv->cplx)
        launch_line_factor<cplx>(model_of<cplx>(lv), dir, (cplx*)lv->fac[dir], (const cplx*)lv->chain_in[dir],
                                 (cplx*)chain_out, g_stream);
    else
        launch_line_factor<double>(model_of<double>(lv), dir, (double*)lv->fac[dir],
                                   (const double*)lv->chain_in[dir], (double*)chain_out, g_stream);
    CK_LAUNCH("line_factor (chained)");
    return 0;
}

int emg3d_b200_point_tile_schedule(int* variant) {
    *variant = point_tile_schedule();
    return 0;
}

int emg3d_b200_line_seg_mask(int mask, int* previous) {
    const int prev = line_seg_mask(mask);
    if (previous) *previous = prev;
    return 0;
}

int emg3d_b200_point_tile_shape(int* txyz) {
    point_tile_shape(txyz);
    return 0;
}

int emg3d_b200_restrict(emg3d_b200_level* c, const void* r_fine, void* s_coarse) {
    NEED_INIT();
    if (!c->linked) return fail_msg("restrict: coarse level is not linked to a fine level");
    if (c->cplx < 0) return fail_msg("restrict: coarse level has no model (dtype unknown)");
    const double* wl[3] = {c->w[0], c->w[3], c->w[6]};
    const double* w0[3] = {c->w[1], c->w[4], c->w[7]};
    const double* wr[3] = {c->w[2], c->w[5], c->w[8]};
    if (c->cplx)
        launch_restrict<cplx>(c->fine, c->cflag, (const cplx*)r_fine, (cplx*)s_coarse, wl, w0, wr, g_stream);
    else
        launch_restrict<double>(c->fine, c->cflag, (const double*)r_fine, (double*)s_coarse, wl, w0, wr, g_stream);
    CK_LAUNCH("restrict");
    return 0;
}

int emg3d_b200_prolong(emg3d_b200_level* c, void* e_fine, const void* e_coarse) {
    NEED_INIT();
    if (!c->linked) return fail_msg("prolong: coarse level is not linked to a fine level");
    if (c->cplx < 0) return fail_msg("prolong: coarse level has no model (dtype unknown)");
    if (c->cplx)
        launch_prolong<cplx>(c->fine, c->cflag, (cplx*)e_fine, (const cplx*)e_coarse, c->lo, c->fr, g_stream);
    else
        launch_prolong<double>(c->fine, c->cflag, (double*)e_fine, (const double*)e_coarse, c->lo, c->fr, g_stream);
    CK_LAUNCH("prolong");
    return 0;
}

int emg3d_b200_restrict_cells(emg3d_b200_level* c, int is_cplx, const void* p_fine, void* p_coarse) {
    NEED_INIT();
    if (!c->linked) return fail_msg("restrict_cells: coarse level is not linked to a fine level");
    if (is_cplx)
        launch_restrict_cells<cplx>(c->fine, c->cflag, (const cplx*)p_fine, (cplx*)p_coarse, g_stream);
    else
        launch_restrict_cells<double>(c->fine, c->cflag, (const double*)p_fine, (double*)p_coarse, g_stream);
    CK_LAUNCH("restrict_cells");
    return 0;
}

int emg3d_b200_volume_model(emg3d_b200_level* lv, int is_cplx, double c_re, double c_im, double s_re,
                            double s_im, int map_code, const double* prop_x, const double* prop_y,
                            const double* prop_z, const double* mu_r, const double* eps_r, void* eta_x,
                            void* eta_y, void* eta_z, double* zeta) {
    NEED_INIT();
    if (map_code < 0 || map_code > 5) return fail_msg("volume_model: unknown property mapping");
    if (!prop_x || !eta_x || !zeta) return fail_msg("volume_model: prop_x, eta_x and zeta are required");
    const double eps0 = 1.0;   // the caller passes s * eps_0 (its own CODATA value) as (s_re, s_im)
    if (is_cplx)
        launch_volume_model<cplx>(lv->d, lv->h[0], lv->h[1], lv->h[2], c_re, c_im, s_re, s_im, eps0, map_code,
                                  prop_x, prop_y, prop_z, mu_r, eps_r, (cplx*)eta_x, (cplx*)eta_y,
                                  (cplx*)eta_z, zeta, g_stream);
    else
        launch_volume_model<double>(lv->d, lv->h[0], lv->h[1], lv->h[2], c_re, c_im, s_re, s_im, eps0,
                                    map_code, prop_x, prop_y, prop_z, mu_r, eps_r, (double*)eta_x,
                                    (double*)eta_y, (double*)eta_z, zeta, g_stream);
    CK_LAUNCH("volume_model");
    return 0;
}

int emg3d_b200_pec_zero(emg3d_b200_level* lv, void* e) {
    NEED_MODEL(lv);
    if (lv->cplx) launch_pec_zero<cplx>(lv->d, (cplx*)e, g_stream);
    else launch_pec_zero<double>(lv->d, (double*)e, g_stream);
    CK_LAUNCH("pec_zero");
    return 0;
}

// ---- vector helpers --------------------------------------------------------------
int emg3d_b200_dot(int is_cplx, long long n, const void* x, const void* y, int conj_x, double* dot2_dev) {
    NEED_INIT();
    if (is_cplx) launch_dot<cplx>(n, (const cplx*)x, (const cplx*)y, conj_x, dot2_dev, g_dot_scratch, g_stream);
    else launch_dot<double>(n, (const double*)x, (const double*)y, 0, dot2_dev, g_dot_scratch, g_stream);
    CK_LAUNCH("dot");
    return 0;
}

int emg3d_b200_dot_host(int is_cplx, long long n, const void* x, const void* y, int conj_x, double* dot2_host) {
    int rc = emg3d_b200_dot(is_cplx, n, x, y, conj_x, g_dot_out);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dot2_host, g_dot_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    if (!is_cplx) dot2_host[1] = 0.0;
    return 0;
}

int emg3d_b200_axpby(int is_cplx, long long n, double a_re, double a_im, const void* x, double b_re,
                     double b_im, void* y) {
    NEED_INIT();
    if (is_cplx) launch_axpby<cplx>(n, make_c(a_re, a_im), (const cplx*)x, make_c(b_re, b_im), (cplx*)y, g_stream);
    else launch_axpby<double>(n, a_re, (const double*)x, b_re, (double*)y, g_stream);
    CK_LAUNCH("axpby");
    return 0;
}

// ---- host-array entry points -------------------------------------------------------
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { return cudaMalloc(&p, n ? n : 16) == cudaSuccess ? 0 : -1; }
};

struct HostCall {
    emg3d_b200_level* lv = nullptr;
    DevBuf eta[3], zeta;
    ~HostCall() { emg3d_b200_level_destroy(lv); }
    int setup(int is_cplx, int nx, int ny, int nz, const void* ex_, const void* ey_, const void* ez_,
              const double* zt, const double* hx, const double* hy, const double* hz) {
        int rc = emg3d_b200_level_create(&lv, nx, ny, nz, hx, hy, hz);
        if (rc) return rc;
        const size_t el = is_cplx ? sizeof(cplx) : sizeof(double);
        const size_t nc = (size_t)nx * ny * nz;
        const void* src[3] = {ex_, ey_, ez_};
        const void* dev[3];
        for (int a = 0; a < 3; ++a) {
            int same = -1;
            for (int b = 0; b < a; ++b) if (src[b] == src[a]) same = b;
            if (same >= 0) { dev[a] = dev[same]; continue; }
            if (eta[a].alloc(nc * el)) return fail_msg("host call: out of device memory");
            rc = emg3d_b200_h2d(eta[a].p, src[a], nc * el);
            if (rc) return rc;
            dev[a] = eta[a].p;
        }
        if (zeta.alloc(nc * sizeof(double))) return fail_msg("host call: out of device memory");
        rc = emg3d_b200_h2d(zeta.p, zt, nc * sizeof(double));
        if (rc) return rc;
        return emg3d_b200_level_set_model(lv, is_cplx, dev[0], dev[1], dev[2], (const double*)zeta.p);
    }
};

int field_io(bool up, int is_cplx, const Dims& d, void* dev, void* x, void* y, void* z) {
    const size_t el = is_cplx ? sizeof(cplx) : sizeof(double);
    void* host[3] = {x, y, z};
    for (int c = 0; c < 3; ++c) {
        char* dp = (char*)dev + comp_offset(d, c) * el;
        const size_t nb = comp_size(d, c) * el;
        int rc = up ? emg3d_b200_h2d(dp, host[c], nb) : emg3d_b200_d2h(host[c], dp, nb);
        if (rc) return rc;
    }
    return 0;
}
}  // namespace

int emg3d_b200_host_amat_x(int is_cplx, int nx, int ny, int nz, void* rx, void* ry, void* rz, const void* ex,
                           const void* ey, const void* ez, const void* eta_x, const void* eta_y,
                           const void* eta_z, const double* zeta, const double* hx, const double* hy,
                           const double* hz) {
    NEED_INIT();
    HostCall hc;
    int rc = hc.setup(is_cplx, nx, ny, nz, eta_x, eta_y, eta_z, zeta, hx, hy, hz);
    if (rc) return rc;
    const size_t el = is_cplx ? sizeof(cplx) : sizeof(double);
    const size_t nb = (size_t)n_edges(hc.lv->d) * el;
    DevBuf r, e;
    if (r.alloc(nb) || e.alloc(nb)) return fail_msg("host_amat_x: out of device memory");
    if ((rc = field_io(true, is_cplx, hc.lv->d, r.p, rx, ry, rz))) return rc;
    if ((rc = field_io(true, is_cplx, hc.lv->d, e.p, (void*)ex, (void*)ey, (void*)ez))) return rc;
    if ((rc = emg3d_b200_amat_x(hc.lv, r.p, e.p))) return rc;
    return field_io(false, is_cplx, hc.lv->d, r.p, rx, ry, rz);
}

// H = zeta_avg / (s mu_0) * curl E / dual measure on the level's grid (device pointers);
// scale = 1 / (s mu_0) as (re, im), im ignored for real (Laplace-domain) fields
int emg3d_b200_magnetic_field(emg3d_b200_level* lv, const void* e, void* hfield, double scale_re,
                              double scale_im) {
    NEED_MODEL(lv);
    if (lv->is_window) return fail_msg("magnetic_field: not defined on a z-window");
    if (lv->cplx)
        launch_edge_curl<cplx, double>(lv->d, (const cplx*)e, (cplx*)hfield, lv->h[0], lv->h[1], lv->h[2],
                                       lv->zeta, make_c(scale_re, scale_im), g_stream);
    else
        launch_edge_curl<double, double>(lv->d, (const double*)e, (double*)hfield, lv->h[0], lv->h[1],
                                         lv->h[2], lv->zeta, scale_re, g_stream);
    CK_LAUNCH("edge_curl");
    return 0;
}

// ---- interpolation next to the solve (interp.cu; emg3d/maps.py, emg3d/fields.py:522-615) -------
int emg3d_b200_volume_average(const double* values, int nx, int ny, int nz, double* out, int mx, int my,
                              int mz, const double* const* w, const int* const* iin, const int* const* start,
                              const double* const* hnew, int log_scale, int add) {
    NEED_INIT();
    if (nx < 1 || ny < 1 || nz < 1 || mx < 1 || my < 1 || mz < 1) return fail_msg("volume_average: empty grid");
    launch_volume_average(values, nx, ny, out, mx, my, mz, w, iin, start, hnew, log_scale, add, g_stream);
    CK_LAUNCH("volume_average");
    return 0;
}

int emg3d_b200_edges_to_vol_averages(int is_cplx, int nx, int ny, int nz, const void* field, const double* hx,
                                     const double* hy, const double* hz, void* out) {
    NEED_INIT();
    if (nx < 1 || ny < 1 || nz < 1) return fail_msg("edges_to_vol_averages: need at least 1 cell per axis");
    Dims d = {};
    d.n[0] = nx; d.n[1] = ny; d.n[2] = nz;
    if (is_cplx) launch_edges_to_vol<cplx>(d, (const cplx*)field, hx, hy, hz, (cplx*)out, g_stream);
    else launch_edges_to_vol<double>(d, (const double*)field, hx, hy, hz, (double*)out, g_stream);
    CK_LAUNCH("edges_to_vol_averages");
    return 0;
}

int emg3d_b200_gradient_field(int nx, int ny, int nz, const void* efield, const void* bfield, double smu0_re,
                              double smu0_im, const double* hx, const double* hy, const double* hz, double* out) {
    NEED_INIT();
    if (nx < 1 || ny < 1 || nz < 1) return fail_msg("gradient_field: need at least 1 cell per axis");
    Dims d = {};
    d.n[0] = nx; d.n[1] = ny; d.n[2] = nz;
    launch_gradient_field(d, (const cplx*)efield, (const cplx*)bfield, make_c(smu0_re, smu0_im), hx, hy, hz, out,
                          g_stream);
    CK_LAUNCH("gradient_field");
    return 0;
}

int emg3d_b200_spline_filter3(int is_cplx, int n0, int n1, int n2, void* data, int reflect) {
    NEED_INIT();
    if (is_cplx) launch_spline_filter<cplx>((cplx*)data, n0, n1, n2, reflect, g_stream);
    else launch_spline_filter<double>((double*)data, n0, n1, n2, reflect, g_stream);
    CK_LAUNCH("spline_filter3");
    return 0;
}

int emg3d_b200_copy_box3(int is_cplx, int n0, int n1, int n2, const void* src, const int* lo, const int* m,
                         void* dst) {
    NEED_INIT();
    for (int a = 0; a < 3; ++a) {
        const int n = a == 0 ? n0 : a == 1 ? n1 : n2;
        if (lo[a] < 0 || m[a] < 1 || lo[a] + m[a] > n) return fail_msg("copy_box3: box outside the array");
    }
    if (is_cplx) launch_copy_box<cplx>((const cplx*)src, n0, n1, lo, m, (cplx*)dst, g_stream);
    else launch_copy_box<double>((const double*)src, n0, n1, lo, m, (double*)dst, g_stream);
    CK_LAUNCH("copy_box3");
    return 0;
}

int emg3d_b200_pad_edge3(int is_cplx, int n0, int n1, int n2, const void* src, int npad, void* dst) {
    NEED_INIT();
    if (is_cplx) launch_pad_edge<cplx>((const cplx*)src, n0, n1, n2, npad, (cplx*)dst, g_stream);
    else launch_pad_edge<double>((const double*)src, n0, n1, n2, npad, (double*)dst, g_stream);
    CK_LAUNCH("pad_edge3");
    return 0;
}

int emg3d_b200_interp_points(int is_cplx, int method, int n0, int n1, int n2, const void* data, int npad, int mode,
                             double fill_re, double fill_im, const double* cx, const double* cy, const double* cz,
                             long long npts, int tensor, int m0, int m1, double scale_re, double scale_im,
                             int accumulate, void* out) {
    NEED_INIT();
    if (method != 1 && method != 3) return fail_msg("interp_points: method must be 1 (linear) or 3 (cubic)");
    if (npts <= 0) return 0;
    if (is_cplx) {
        const cplx fill = make_c(fill_re, fill_im), scale = make_c(scale_re, scale_im);
        if (method == 3)
            launch_spline_eval<cplx>((const cplx*)data, n0, n1, n2, npad, mode, fill, cx, cy, cz, npts, tensor, m0, m1,
                                     scale, accumulate, (cplx*)out, g_stream);
        else
            launch_linear_eval<cplx>((const cplx*)data, n0, n1, n2, fill, cx, cy, cz, npts, tensor, m0, m1, scale,
                                     accumulate, (cplx*)out, g_stream);
    } else {
        if (method == 3)
            launch_spline_eval<double>((const double*)data, n0, n1, n2, npad, mode, fill_re, cx, cy, cz, npts, tensor,
                                       m0, m1, scale_re, accumulate, (double*)out, g_stream);
        else
            launch_linear_eval<double>((const double*)data, n0, n1, n2, fill_re, cx, cy, cz, npts, tensor, m0, m1,
                                       scale_re, accumulate, (double*)out, g_stream);
    }
    CK_LAUNCH("interp_points");
    return 0;
}

int emg3d_b200_host_edge_curl_factor(int is_cplx, int nx, int ny, int nz, void* mx, void* my, void* mz,
                                     const void* ex, const void* ey, const void* ez, const double* hx,
                                     const double* hy, const double* hz, const void* zeta) {
    NEED_INIT();
    if (nx < 1 || ny < 1 || nz < 1) return fail_msg("host_edge_curl_factor: need at least 1 cell per axis");
    Dims d = {};
    d.n[0] = nx; d.n[1] = ny; d.n[2] = nz;
    const size_t el = is_cplx ? sizeof(cplx) : sizeof(double);
    const size_t nc = (size_t)nx * ny * nz;
    DevBuf e, hf, z, h[3];
    const double* hh[3] = {hx, hy, hz};
    int rc;
    if (e.alloc((size_t)n_edges(d) * el) || hf.alloc((size_t)n_faces(d) * el) || z.alloc(nc * el))
        return fail_msg("host_edge_curl_factor: out of device memory");
    for (int a = 0; a < 3; ++a) {
        if (h[a].alloc(sizeof(double) * d.n[a])) return fail_msg("host_edge_curl_factor: out of device memory");
        if ((rc = emg3d_b200_h2d(h[a].p, hh[a], sizeof(double) * d.n[a]))) return rc;
    }
    if ((rc = field_io(true, is_cplx, d, e.p, (void*)ex, (void*)ey, (void*)ez))) return rc;
    if ((rc = emg3d_b200_h2d(z.p, zeta, nc * el))) return rc;
    if (is_cplx)
        launch_edge_curl<cplx, cplx>(d, (const cplx*)e.p, (cplx*)hf.p, (const double*)h[0].p,
                                     (const double*)h[1].p, (const double*)h[2].p, (const cplx*)z.p,
                                     make_c(1.0, 0.0), g_stream);
    else
        launch_edge_curl<double, double>(d, (const double*)e.p, (double*)hf.p, (const double*)h[0].p,
                                         (const double*)h[1].p, (const double*)h[2].p, (const double*)z.p,
                                         1.0, g_stream);
    CK_LAUNCH("edge_curl");
    // faces: hx (nx+1, ny, nz), hy (nx, ny+1, nz), hz (nx, ny, nz+1)
    const size_t nf[3] = {(size_t)(nx + 1) * ny * nz, (size_t)nx * (ny + 1) * nz, (size_t)nx * ny * (nz + 1)};
    void* out[3] = {mx, my, mz};
    size_t off = 0;
    for (int c = 0; c < 3; ++c) {
        if ((rc = emg3d_b200_d2h(out[c], (char*)hf.p + off * el, nf[c] * el))) return rc;
        off += nf[c];
    }
    return 0;
}

int emg3d_b200_host_solve(int is_cplx, int n, void* amat, void* bvec) {
    NEED_INIT();
    if (n < 1) return fail_msg("host_solve: n must be positive");
    const size_t el = is_cplx ? sizeof(cplx) : sizeof(double);
    DevBuf a, b;
    if (a.alloc(6 * (size_t)n * el) || b.alloc((size_t)n * el)) return fail_msg("host_solve: out of device memory");
    int rc;
    if ((rc = emg3d_b200_h2d(a.p, amat, 6 * (size_t)n * el))) return rc;
    if ((rc = emg3d_b200_h2d(b.p, bvec, (size_t)n * el))) return rc;
    if (is_cplx) launch_band_solve<cplx>(n, (cplx*)a.p, (cplx*)b.p, g_stream);
    else launch_band_solve<double>(n, (double*)a.p, (double*)b.p, g_stream);
    CK_LAUNCH("band_solve");
    if ((rc = emg3d_b200_d2h(amat, a.p, 6 * (size_t)n * el))) return rc;
    return emg3d_b200_d2h(bvec, b.p, (size_t)n * el);
}

// core.restrict (core.py:1620-1621; call site solver.py:937-938) on host arrays: (nx, ny, nz) are the
// FINE cell counts, weights[3 a + {0, 1, 2}] = wl, w0, wr of axis a (core.restrict_weights), NULL where
// sc_dir leaves the axis alone.  Two bare levels (unit widths: the restriction only uses the weights)
// are linked for the call.
int emg3d_b200_host_restrict(int is_cplx, int nx, int ny, int nz, int sc_dir, void* crx, void* cry, void* crz,
                             const void* rx, const void* ry, const void* rz, const double* const* weights) {
    NEED_INIT();
    static const int flags[7][3] = {{1, 1, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (sc_dir < 0 || sc_dir > 6) return fail_msg("host_restrict: sc_dir must be 0 .. 6");
    if (!weights) return fail_msg("host_restrict: weights missing");
    const int* cf = flags[sc_dir];
    const int nf[3] = {nx, ny, nz};
    int nc[3], nmax = 1;
    for (int a = 0; a < 3; ++a) {
        if (nf[a] < 1) return fail_msg("host_restrict: need at least 1 cell per axis");
        if (cf[a] && (nf[a] % 2 || nf[a] < 2))
            return fail_msg("host_restrict: a coarsened axis needs an even number of cells");
        if (cf[a] && !(weights[3 * a] && weights[3 * a + 1] && weights[3 * a + 2]))
            return fail_msg("host_restrict: weights of a coarsened axis missing");
        nc[a] = cf[a] ? nf[a] / 2 : nf[a];
        if (nf[a] > nmax) nmax = nf[a];
    }
    struct Owned {
        emg3d_b200_level* p = nullptr;
        ~Owned() { emg3d_b200_level_destroy(p); }
    } fine, coarse;                                       // (coarse is destroyed first)
    std::vector<double> ones((size_t)nmax, 1.0), frac0((size_t)nmax + 1, 0.0);
    std::vector<int> lo0((size_t)nmax + 1, 0);            // prolongation tables: unused here
    int rc;
    if ((rc = emg3d_b200_level_create(&fine.p, nx, ny, nz, ones.data(), ones.data(), ones.data()))) return rc;
    if ((rc = emg3d_b200_level_create(&coarse.p, nc[0], nc[1], nc[2], ones.data(), ones.data(), ones.data())))
        return rc;
    const int* lo[3] = {lo0.data(), lo0.data(), lo0.data()};
    const double* fr[3] = {frac0.data(), frac0.data(), frac0.data()};
    if ((rc = emg3d_b200_level_link(coarse.p, fine.p, cf, weights, lo, fr))) return rc;
    if ((rc = emg3d_b200_level_set_model(coarse.p, is_cplx, nullptr, nullptr, nullptr, nullptr))) return rc;
    const size_t el = is_cplx ? sizeof(cplx) : sizeof(double);
    DevBuf r, c;
    if (r.alloc((size_t)n_edges(fine.p->d) * el) || c.alloc((size_t)n_edges(coarse.p->d) * el))
        return fail_msg("host_restrict: out of device memory");
    if ((rc = field_io(true, is_cplx, fine.p->d, r.p, (void*)rx, (void*)ry, (void*)rz))) return rc;
    if ((rc = emg3d_b200_restrict(coarse.p, r.p, c.p))) return rc;
    return field_io(false, is_cplx, coarse.p->d, c.p, crx, cry, crz);
}

int emg3d_b200_host_gauss_seidel(int is_cplx, int ldir, int order, int nx, int ny, int nz, void* ex, void* ey,
                                 void* ez, const void* sx, const void* sy, const void* sz,
                                 const void* eta_x, const void* eta_y, const void* eta_z,
                                 const double* zeta, const double* hx, const double* hy, const double* hz,
                                 int nu) {
    NEED_INIT();
    HostCall hc;
    int rc = hc.setup(is_cplx, nx, ny, nz, eta_x, eta_y, eta_z, zeta, hx, hy, hz);
    if (rc) return rc;
    const size_t el = is_cplx ? sizeof(cplx) : sizeof(double);
    const size_t nb = (size_t)n_edges(hc.lv->d) * el;
    DevBuf s, e;
    if (s.alloc(nb) || e.alloc(nb)) return fail_msg("host_gauss_seidel: out of device memory");
    if ((rc = field_io(true, is_cplx, hc.lv->d, e.p, ex, ey, ez))) return rc;
    if ((rc = field_io(true, is_cplx, hc.lv->d, s.p, (void*)sx, (void*)sy, (void*)sz))) return rc;
    if ((rc = emg3d_b200_gauss_seidel(hc.lv, e.p, s.p, nu, ldir, order))) return rc;
    return field_io(false, is_cplx, hc.lv->d, e.p, ex, ey, ez);
}

}  // extern "C"
