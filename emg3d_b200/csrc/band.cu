// Banded complex-symmetric LDL^T solve, half-bandwidth 5, no pivoting: the
// interface of emg3d/core.py:1481-1616 `solve` (a[p + 5 q] = A(p, q),
// 0 <= p - q <= 5; a is overwritten by L and 1/D, b by the solution).
//
// The smoothers do not use this routine (they work on 6x6 node blocks and on
// block-tridiagonal line factors, see gs_point.cu / gs_line.cu); it exists so
// that the reference's stand-alone `solve` has a device counterpart with the
// same contract.  Column-oriented (right-looking) elimination: one column is
// finished before the next is touched, 15 independent updates per column, which
// one warp carries out with lane (r, c) owning A(j+r, j+c).
#include "common.cuh"
#include "kernels.h"

namespace emg {

template <typename T>
__global__ void __launch_bounds__(32) band_solve_kernel(int n, T* __restrict__ a, T* __restrict__ b) {
    const int lane = threadIdx.x;
    // lanes 0..14 <-> (r, c), 1 <= c <= r <= 5 : update A(j+r, j+c) -= L(r) L(c) D
    int r = 0, c = 0;
    {
        int k = 0;
        for (int rr = 1; rr <= 5; ++rr)
            for (int cc = 1; cc <= rr; ++cc) {
                if (k == lane) { r = rr; c = cc; }
                ++k;
            }
    }
    for (int j = 0; j < n; ++j) {
        const T d = a[6 * j];
        const T dinv = rcp(d);
        __syncwarp();
        if (lane < 15 && j + r < n) {
            const T lr = a[(j + r) + 5 * j] * dinv;
            const T lc = a[(j + c) + 5 * j];          // still unscaled: L(c) D
            a[(j + r) + 5 * (j + c)] -= lr * lc;
        }
        __syncwarp();
        if (lane >= 1 && lane <= 5 && j + lane < n) a[(j + lane) + 5 * j] *= dinv;
        if (lane == 0) a[6 * j] = dinv;
        __syncwarp();
    }
    if (lane == 0) {
        for (int j = 1; j < n; ++j) {
            T acc = zero_<T>();
            for (int k = max(0, j - 5); k < j; ++k) acc += a[j + 5 * k] * b[k];
            b[j] -= acc;
        }
        for (int j = 0; j < n; ++j) b[j] = b[j] * a[6 * j];
        for (int j = n - 2; j >= 0; --j) {
            T acc = zero_<T>();
            for (int k = j + 1; k < min(n, j + 6); ++k) acc += a[k + 5 * j] * b[k];
            b[j] -= acc;
        }
    }
}

template <typename T>
void launch_band_solve(int n, T* a, T* b, cudaStream_t st) {
    ++g_launch_count; band_solve_kernel<T><<<1, 32, 0, st>>>(n, a, b);
}
template void launch_band_solve<double>(int, double*, double*, cudaStream_t);
template void launch_band_solve<cplx>(int, cplx*, cplx*, cudaStream_t);

}  // namespace emg
