// Residual r = s - A e of the 3-D staggered-grid EM diffusion operator
// (what emg3d/core.py:57-206 `amat_x` computes, fused with the copy of
// emg3d/solver.py:1059 and the norm of solver.py:1066).
//
// Bandwidth-bound streaming stencil: one thread per node (ix, iy, iz) of the
// (nx+1, ny+1, nz+1) node grid updates the three edges that start at that node.
// Algorithmic traffic per cell (complex, triaxial): s 48 + e 48 + eta 48 +
// zeta 8 + r 48 = 200 B (152 B when only the norm is wanted).
#include "common.cuh"
#include "kernels.h"

namespace emg {

// zeta-weighted curl through one face ("flux"); see header comment of
// gs_point.cu for the face convention.  All indices are valid by construction.
template <typename T>
struct Flux {
    const Model<T>& m;
    FieldView<const T> e;
    int64_t cs1, cs2;   // cell strides
    __device__ Flux(const Model<T>& m_, const T* eb)
        : m(m_), e(eb, m_.d), cs1(m_.d.n[0]), cs2((int64_t)m_.d.n[0] * m_.d.n[1]) {}
    __device__ __forceinline__ double z(int i, int j, int k) const {
        return ldg(m.zeta + i + cs1 * j + cs2 * k);
    }
    __device__ __forceinline__ T E(int c, int i, int j, int k) const {
        return ldg(e.p[c] + e.idx(c, i, j, k));
    }
    // face normal z at cells (i, j), node plane k
    __device__ __forceinline__ T fz(int i, int j, int k) const {
        double M = z(i, j, k - 1) + z(i, j, k);
        T c = ldg(m.rh[0] + i) * (E(1, i + 1, j, k) - E(1, i, j, k)) -
              ldg(m.rh[1] + j) * (E(0, i, j + 1, k) - E(0, i, j, k));
        return M * c;
    }
    // face normal y at cells (i, k), node line j
    __device__ __forceinline__ T fy(int i, int j, int k) const {
        double M = z(i, j - 1, k) + z(i, j, k);
        T c = ldg(m.rh[2] + k) * (E(0, i, j, k + 1) - E(0, i, j, k)) -
              ldg(m.rh[0] + i) * (E(2, i + 1, j, k) - E(2, i, j, k));
        return M * c;
    }
    // face normal x at cells (j, k), node i
    __device__ __forceinline__ T fx(int i, int j, int k) const {
        double M = z(i - 1, j, k) + z(i, j, k);
        T c = ldg(m.rh[1] + j) * (E(2, i, j + 1, k) - E(2, i, j, k)) -
              ldg(m.rh[2] + k) * (E(1, i, j, k + 1) - E(1, i, j, k));
        return M * c;
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
residual_kernel(Model<T> m, const T* s, const T* __restrict__ e,
                T* r, double* __restrict__ partial, int apply_only) {
    const int nx = m.d.n[0], ny = m.d.n[1], nz = m.d.n[2];
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    const int iz = blockIdx.z * blockDim.z + threadIdx.z;
    double acc = 0.0;
    if (ix <= nx && iy <= ny && iz <= nz) {
        Flux<T> F(m, e);
        FieldView<const T> S(s, m.d);
        FieldView<T> R(r, m.d);
        const bool inner = ix < nx && iy < ny && iz < nz;   // edges the reference touches
        const int64_t cs1 = nx, cs2 = (int64_t)nx * ny;
        const int ixm = max(ix - 1, 0), iym = max(iy - 1, 0), izm = max(iz - 1, 0);
        T fzpp = zero_<T>(), fypp = zero_<T>(), fxpp = zero_<T>();
        if (inner) {
            if (iz > 0) fzpp = F.fz(ix, iy, iz);
            if (iy > 0) fypp = F.fy(ix, iy, iz);
            if (ix > 0) fxpp = F.fx(ix, iy, iz);
        }
        // x-edge (ix, iy, iz)
        if (ix < nx) {
            int64_t id = S.idx(0, ix, iy, iz);
            T val = apply_only ? zero_<T>() : ldg(S.p[0] + id);
            if (inner) {
                T cc = zero_<T>();
                if (iy > 0 && iz > 0) {
                    cc = ldg(m.rh[1] + iy) * fzpp - ldg(m.rh[1] + iy - 1) * F.fz(ix, iy - 1, iz) -
                         ldg(m.rh[2] + iz) * fypp + ldg(m.rh[2] + iz - 1) * F.fy(ix, iy, iz - 1);
                }
                const T* et = m.eta[0];
                T st = ldg(et + ix + cs1 * iym + cs2 * izm) + ldg(et + ix + cs1 * iym + cs2 * iz) +
                       ldg(et + ix + cs1 * iy + cs2 * izm) + ldg(et + ix + cs1 * iy + cs2 * iz);
                val -= 0.5 * cc - 0.25 * (st * F.E(0, ix, iy, iz));
            }
            if (apply_only) val = -val;
            if (r) R.p[0][id] = val;
            acc += abs2(val);
        }
        // y-edge
        if (iy < ny) {
            int64_t id = S.idx(1, ix, iy, iz);
            T val = apply_only ? zero_<T>() : ldg(S.p[1] + id);
            if (inner) {
                T cc = zero_<T>();
                if (ix > 0 && iz > 0) {
                    cc = ldg(m.rh[2] + iz) * fxpp - ldg(m.rh[2] + iz - 1) * F.fx(ix, iy, iz - 1) -
                         ldg(m.rh[0] + ix) * fzpp + ldg(m.rh[0] + ix - 1) * F.fz(ix - 1, iy, iz);
                }
                const T* et = m.eta[1];
                T st = ldg(et + ixm + cs1 * iy + cs2 * izm) + ldg(et + ix + cs1 * iy + cs2 * izm) +
                       ldg(et + ixm + cs1 * iy + cs2 * iz) + ldg(et + ix + cs1 * iy + cs2 * iz);
                val -= 0.5 * cc - 0.25 * (st * F.E(1, ix, iy, iz));
            }
            if (apply_only) val = -val;
            if (r) R.p[1][id] = val;
            acc += abs2(val);
        }
        // z-edge
        if (iz < nz) {
            int64_t id = S.idx(2, ix, iy, iz);
            T val = apply_only ? zero_<T>() : ldg(S.p[2] + id);
            if (inner) {
                T cc = zero_<T>();
                if (ix > 0 && iy > 0) {
                    cc = ldg(m.rh[0] + ix) * fypp - ldg(m.rh[0] + ix - 1) * F.fy(ix - 1, iy, iz) -
                         ldg(m.rh[1] + iy) * fxpp + ldg(m.rh[1] + iy - 1) * F.fx(ix, iy - 1, iz);
                }
                const T* et = m.eta[2];
                T st = ldg(et + ixm + cs1 * iym + cs2 * iz) + ldg(et + ix + cs1 * iym + cs2 * iz) +
                       ldg(et + ixm + cs1 * iy + cs2 * iz) + ldg(et + ix + cs1 * iy + cs2 * iz);
                val -= 0.5 * cc - 0.25 * (st * F.E(2, ix, iy, iz));
            }
            if (apply_only) val = -val;
            if (r) R.p[2][id] = val;
            acc += abs2(val);
        }
    }
    if (partial) {
        // deterministic block reduction -> one partial per block
        __shared__ double red[8];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
            for (int w = 0; w < nw; ++w) t += red[w];
            partial[blockIdx.x + gridDim.x * (blockIdx.y + (int64_t)gridDim.y * blockIdx.z)] = t;
        }
    }
}

// single-block deterministic sum of the per-block partials
__global__ void __launch_bounds__(1024) sum_partials_kernel(const double* __restrict__ p, int64_t n,
                                                           double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += p[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) out[0] = acc;
    }
}

template <typename T>
void launch_residual(const Model<T>& m, const T* s, const T* e, T* r, double* norm2_out,
                     double* scratch, int apply_only, cudaStream_t st) {
    dim3 b(32, 4, 2);
    dim3 g((m.d.n[0] + 1 + b.x - 1) / b.x, (m.d.n[1] + 1 + b.y - 1) / b.y,
           (m.d.n[2] + 1 + b.z - 1) / b.z);
    ++g_launch_count; residual_kernel<T><<<g, b, 0, st>>>(m, s, e, r, norm2_out ? scratch : nullptr, apply_only);
    if (norm2_out) {
        int64_t nb = (int64_t)g.x * g.y * g.z;
        ++g_launch_count; sum_partials_kernel<<<1, 1024, 0, st>>>(scratch, nb, norm2_out);
    }
}

int64_t residual_scratch_doubles(const Dims& d) {
    dim3 b(32, 4, 2);
    return (int64_t)((d.n[0] + 1 + b.x - 1) / b.x) * ((d.n[1] + 1 + b.y - 1) / b.y) *
           ((d.n[2] + 1 + b.z - 1) / b.z);
}

template void launch_residual<double>(const Model<double>&, const double*, const double*, double*,
                                      double*, double*, int, cudaStream_t);
template void launch_residual<cplx>(const Model<cplx>&, const cplx*, const cplx*, cplx*, double*,
                                    double*, int, cudaStream_t);

}  // namespace emg
