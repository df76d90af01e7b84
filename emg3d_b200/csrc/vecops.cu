// Small vector kernels that keep the multigrid cycle and the Krylov wrapper
// resident on the device: PEC zeroing (emg3d/solver.py:350-355), dot products
// and norms (scipy.linalg.norm / numpy.vdot at solver.py:312, 1066 and inside
// scipy's bicgstab), and y = a x + b y.
#include "common.cuh"
#include "kernels.h"

namespace emg {

// zero all tangential boundary edges
template <typename T>
__global__ void __launch_bounds__(256) pec_kernel(Dims d, T* e) {
    FieldView<T> E(e, d);
    int f[3];
    f[0] = blockIdx.x * blockDim.x + threadIdx.x;
    f[1] = blockIdx.y * blockDim.y + threadIdx.y;
    f[2] = blockIdx.z * blockDim.z + threadIdx.z;
    if (f[0] > d.n[0] || f[1] > d.n[1] || f[2] > d.n[2]) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int u = (c + 1) % 3, v = (c + 2) % 3;
        if (f[c] >= d.n[c]) continue;
        if (f[u] == 0 || f[u] == d.n[u] || f[v] == 0 || f[v] == d.n[v]) E.p[c][E.idx(c, f)] = zero_<T>();
    }
}

template <typename T>
void launch_pec_zero(const Dims& d, T* e, cudaStream_t st) {
    dim3 b(32, 4, 2);
    dim3 g((d.n[0] + 1 + b.x - 1) / b.x, (d.n[1] + 1 + b.y - 1) / b.y, (d.n[2] + 1 + b.z - 1) / b.z);
    ++g_launch_count; pec_kernel<T><<<g, b, 0, st>>>(d, e);
}

constexpr int DOT_THREADS = 256;
constexpr int DOT_MAX_BLOCKS = 148 * 8;

__device__ __forceinline__ void dot_acc(double x, double y, int, double& re, double& im) { re += x * y; (void)im; }
__device__ __forceinline__ void dot_acc(cplx x, cplx y, int cj, double& re, double& im) {
    // conj(x) * y if cj else x * y
    const double xi = cj ? -x.im : x.im;
    re += x.re * y.re - xi * y.im;
    im += x.re * y.im + xi * y.re;
}

template <typename T>
__global__ void __launch_bounds__(DOT_THREADS)
dot_kernel(int64_t n, const T* __restrict__ x, const T* __restrict__ y, int cj, double* __restrict__ part) {
    double re = 0.0, im = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dot_acc(ldg(x + i), ldg(y + i), cj, re, im);
    __shared__ double r0[DOT_THREADS / 32], r1[DOT_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_down_sync(0xffffffffu, re, o);
        im += __shfl_down_sync(0xffffffffu, im, o);
    }
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = re; r1[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < DOT_THREADS / 32; ++w) { a += r0[w]; b += r1[w]; }
        part[2 * blockIdx.x] = a;
        part[2 * blockIdx.x + 1] = b;
    }
}

__global__ void __launch_bounds__(1024) dot_final_kernel(const double* __restrict__ part, int nb,
                                                        double* __restrict__ out2) {
    __shared__ double r0[32], r1[32];
    double re = 0.0, im = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) { re += part[2 * i]; im += part[2 * i + 1]; }
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_down_sync(0xffffffffu, re, o);
        im += __shfl_down_sync(0xffffffffu, im, o);
    }
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = re; r1[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 32; ++w) { a += r0[w]; b += r1[w]; }
        out2[0] = a;
        out2[1] = b;
    }
}

static int dot_blocks(int64_t n) {
    int64_t nb = (n + DOT_THREADS - 1) / DOT_THREADS;
    if (nb > DOT_MAX_BLOCKS) nb = DOT_MAX_BLOCKS;
    if (nb < 1) nb = 1;
    return (int)nb;
}

int64_t dot_scratch_doubles(int64_t n) { return 2 * (int64_t)dot_blocks(n); }

template <typename T>
void launch_dot(int64_t n, const T* x, const T* y, int conj_x, double* out2, double* scratch,
                cudaStream_t st) {
    const int nb = dot_blocks(n);
    ++g_launch_count; dot_kernel<T><<<nb, DOT_THREADS, 0, st>>>(n, x, y, conj_x, scratch);
    ++g_launch_count; dot_final_kernel<<<1, 1024, 0, st>>>(scratch, nb, out2);
}

template <typename T>
__global__ void __launch_bounds__(256) axpby_kernel(int64_t n, T a, const T* __restrict__ x, T b, T* y,
                                                   int bzero) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const T ax = a * ldg(x + i);
        y[i] = bzero ? ax : ax + b * y[i];
    }
}

template <typename T>
void launch_axpby(int64_t n, T a, const T* x, T b, T* y, cudaStream_t st) {
    int64_t nb = (n + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    if (nb < 1) nb = 1;
    const int bzero = abs2(b) == 0.0;
    ++g_launch_count; axpby_kernel<T><<<(int)nb, 256, 0, st>>>(n, a, x, b, y, bzero);
}

#define INST(T)                                                                                    \
    template void launch_pec_zero<T>(const Dims&, T*, cudaStream_t);                               \
    template void launch_dot<T>(int64_t, const T*, const T*, int, double*, double*, cudaStream_t); \
    template void launch_axpby<T>(int64_t, T, const T*, T, T*, cudaStream_t);
INST(double)
INST(cplx)

}  // namespace emg
