// Magnetic field from the electric field by Faraday's law (what
// emg3d/fields.py:941-1009 `_edge_curl_factor` computes for
// fields.py:617-659 `get_magnetic_field`): the curl of the edge field lives on
// the faces; it is scaled by a per-cell factor averaged over the two cells that
// share the face and by the inverse dual-cell measure.  Same stencil family as
// the first curl of amat_x; one thread per cell, x fastest, 3 x 4 edge loads and
// 3 face stores per cell (streaming, HBM-bound: 48 B of E + 8..16 B of factor in,
// 48 B of H out per cell).
//
// Layout of the result (a magnetic Field, fields.py:116 with meshes.py:110-112):
// [hx | hy | hz], hx (nx+1, ny, nz), hy (nx, ny+1, nz), hz (nx, ny, nz+1), x
// fastest.  Faces on the low boundary (index 0) and the high boundary are not
// written by the reference (fields.py:1002-1007 skip index 0; the loops stop at
// n - 1): they keep the zeros the caller (here: launch_edge_curl) puts there.
#include "common.cuh"
#include "kernels.h"

namespace emg {

__device__ __forceinline__ cplx mulz(cplx a, cplx b) { return a * b; }
__device__ __forceinline__ cplx mulz(cplx a, double b) { return a * b; }
__device__ __forceinline__ double mulz(double a, double b) { return a * b; }

// Z: type of the per-cell factor (double: zeta of the level, scaled by `scale`;
// T: a caller-provided array as in the reference's kernel signature)
template <typename T, typename Z>
__global__ void __launch_bounds__(256)
edge_curl_kernel(Dims d, const T* __restrict__ e, T* __restrict__ hf, const double* __restrict__ hx,
                 const double* __restrict__ hy, const double* __restrict__ hz,
                 const Z* __restrict__ zeta, T scale) {
    const int nx = d.n[0], ny = d.n[1], nz = d.n[2];
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    const int iz = blockIdx.z * blockDim.z + threadIdx.z;
    if (ix >= nx || iy >= ny || iz >= nz) return;
    FieldView<const T> E(e, d);
    const int ixm = max(ix - 1, 0), iym = max(iy - 1, 0), izm = max(iz - 1, 0);
    const double wx = ldg(hx + ix), wy = ldg(hy + iy), wz = ldg(hz + iz);
    const T ex0 = ldg(E.p[0] + E.idx(0, ix, iy, iz)), ey0 = ldg(E.p[1] + E.idx(1, ix, iy, iz)),
            ez0 = ldg(E.p[2] + E.idx(2, ix, iy, iz));
    // curl E through the three faces that start at the cell's low corner
    const T fx = (1.0 / wy) * (ldg(E.p[2] + E.idx(2, ix, iy + 1, iz)) - ez0) -
                 (1.0 / wz) * (ldg(E.p[1] + E.idx(1, ix, iy, iz + 1)) - ey0);
    const T fy = (1.0 / wz) * (ldg(E.p[0] + E.idx(0, ix, iy, iz + 1)) - ex0) -
                 (1.0 / wx) * (ldg(E.p[2] + E.idx(2, ix + 1, iy, iz)) - ez0);
    const T fz = (1.0 / wx) * (ldg(E.p[1] + E.idx(1, ix + 1, iy, iz)) - ey0) -
                 (1.0 / wy) * (ldg(E.p[0] + E.idx(0, ix, iy + 1, iz)) - ex0);
    const int64_t cs1 = nx, cs2 = (int64_t)nx * ny;
    const int64_t c = ix + cs1 * iy + cs2 * iz;
    const Z z0 = ldg(zeta + c);
    // face arrays: hx (nx+1, ny, nz), hy (nx, ny+1, nz), hz (nx, ny, nz+1)
    T* const mx = hf;
    T* const my = mx + (int64_t)(nx + 1) * ny * nz;
    T* const mz = my + (int64_t)nx * (ny + 1) * nz;
    if (ix != 0) {
        const Z zs = ldg(zeta + ixm + cs1 * iy + cs2 * iz) + z0;
        const double dm = (ldg(hx + ixm) + wx) * wy * wz;
        mx[ix + (int64_t)(nx + 1) * (iy + (int64_t)ny * iz)] = mulz(mulz(fx, zs) * (1.0 / dm), scale);
    }
    if (iy != 0) {
        const Z zs = ldg(zeta + ix + cs1 * iym + cs2 * iz) + z0;
        const double dm = wx * (ldg(hy + iym) + wy) * wz;
        my[ix + (int64_t)nx * (iy + (int64_t)(ny + 1) * iz)] = mulz(mulz(fy, zs) * (1.0 / dm), scale);
    }
    if (iz != 0) {
        const Z zs = ldg(zeta + ix + cs1 * iy + cs2 * izm) + z0;
        const double dm = wx * wy * (ldg(hz + izm) + wz);
        mz[ix + (int64_t)nx * (iy + (int64_t)ny * iz)] = mulz(mulz(fz, zs) * (1.0 / dm), scale);
    }
}

int64_t n_faces(const Dims& d) {
    const int64_t nx = d.n[0], ny = d.n[1], nz = d.n[2];
    return (nx + 1) * ny * nz + nx * (ny + 1) * nz + nx * ny * (nz + 1);
}

template <typename T, typename Z>
void launch_edge_curl(const Dims& d, const T* e, T* hf, const double* hx, const double* hy,
                      const double* hz, const Z* zeta, T scale, cudaStream_t st) {
    cudaMemsetAsync(hf, 0, sizeof(T) * n_faces(d), st);
    dim3 b(32, 4, 2);
    dim3 g((d.n[0] + b.x - 1) / b.x, (d.n[1] + b.y - 1) / b.y, (d.n[2] + b.z - 1) / b.z);
    ++g_launch_count; edge_curl_kernel<T, Z><<<g, b, 0, st>>>(d, e, hf, hx, hy, hz, zeta, scale);
}

template void launch_edge_curl<cplx, double>(const Dims&, const cplx*, cplx*, const double*, const double*,
                                             const double*, const double*, cplx, cudaStream_t);
template void launch_edge_curl<cplx, cplx>(const Dims&, const cplx*, cplx*, const double*, const double*,
                                           const double*, const cplx*, cplx, cudaStream_t);
template void launch_edge_curl<double, double>(const Dims&, const double*, double*, const double*,
                                               const double*, const double*, const double*, double,
                                               cudaStream_t);

}  // namespace emg
