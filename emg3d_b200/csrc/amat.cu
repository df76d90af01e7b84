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
        const bool own_xy = owns_plane(m.d, iz), own_z = owns_layer(m.d, iz);
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
            if (own_xy) acc += abs2(val);
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
            if (own_xy) acc += abs2(val);
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
            if (own_z) acc += abs2(val);
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

// ---- plane-streaming variant for large grids ----------------------------------
// Each flux (zeta-weighted curl through a face) enters the residual of four
// edges; the simple kernel above recomputes it for each of them (9 fluxes, ~80
// loads per node).  Here a block of 32 x 8 threads owns an (x, y) tile and
// marches through KZ node planes: every thread computes the three fluxes at its
// own node once (18 loads) and obtains the neighbours'
//   -z : from its own registers (previous plane),
//   -x : by warp shuffle (a warp is one x-row of the tile),
//   -y : through a double-buffered shared-memory row exchange,
// recomputing only on the low tile faces.  DRAM traffic stays at the algorithmic
// 200 B per cell; L1 traffic drops by more than half.
__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ cplx shfl_up1(cplx v) {
    return make_c(__shfl_up_sync(0xffffffffu, v.re, 1), __shfl_up_sync(0xffffffffu, v.im, 1));
}

#ifndef EMG_RZ_BY
#define EMG_RZ_BY 8
#endif
#ifndef EMG_RZ_MINB
#define EMG_RZ_MINB 4
#endif
#ifndef EMG_RZ_KZ
#define EMG_RZ_KZ 16
#endif
constexpr int RZ_BX = 32, RZ_BY = EMG_RZ_BY;

template <typename T>
__global__ void __launch_bounds__(RZ_BX * RZ_BY, EMG_RZ_MINB)
residual_zmarch_kernel(Model<T> m, const T* s, const T* __restrict__ e, T* r,
                       double* __restrict__ partial, int apply_only, int kz) {
    const int nx = m.d.n[0], ny = m.d.n[1], nz = m.d.n[2];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * RZ_BX, y0 = blockIdx.y * RZ_BY;
    const int ix = x0 + tx, iy = y0 + ty;
    const int k0 = blockIdx.z * kz, k1 = min(k0 + kz, nz + 1);
    const bool node = ix <= nx && iy <= ny;
    const bool cellxy = ix < nx && iy < ny;
    __shared__ T sFz[2][RZ_BY + 1][RZ_BX];
    __shared__ T sFx[2][RZ_BY + 1][RZ_BX];

    Flux<T> F(m, e);
    FieldView<const T> S(s, m.d);
    FieldView<T> R(r, m.d);
    const int64_t cs1 = nx, cs2 = (int64_t)nx * ny;
    const int ixm = max(ix - 1, 0), iym = max(iy - 1, 0);
    const double rhx = ix < nx ? ldg(m.rh[0] + ix) : 0.0, rhxm = ldg(m.rh[0] + ixm);
    const double rhy = iy < ny ? ldg(m.rh[1] + iy) : 0.0, rhym = ldg(m.rh[1] + iym);

    // fluxes of the plane below the first one (needed by x- and y-edges)
    T fx_dn = zero_<T>(), fy_dn = zero_<T>();
    if (cellxy && k0 > 0) {
        if (ix > 0) fx_dn = F.fx(ix, iy, k0 - 1);
        if (iy > 0) fy_dn = F.fy(ix, iy, k0 - 1);
    }
    double acc = 0.0;
    for (int k = k0; k < k1; ++k) {
        const int buf = (k - k0) & 1;
        const bool cell = cellxy && k < nz;          // the edges the reference touches
        const int km = max(k - 1, 0);
        // own fluxes
        T fx = zero_<T>(), fy = zero_<T>(), fz = zero_<T>();
        if (cell) {
            if (ix > 0) fx = F.fx(ix, iy, k);
            if (iy > 0) fy = F.fy(ix, iy, k);
            if (k > 0) fz = F.fz(ix, iy, k);
        }
        // -x neighbours by shuffle; the first lane of the row recomputes
        T fz_xm = shfl_up1(fz), fy_xm = shfl_up1(fy);
        if (tx == 0) {
            fz_xm = zero_<T>();
            fy_xm = zero_<T>();
            if (cell && ix > 0) {
                if (k > 0) fz_xm = F.fz(ix - 1, iy, k);
                if (iy > 0) fy_xm = F.fy(ix - 1, iy, k);
            }
        }
        // -y neighbours through shared memory; row 0 holds the tile's lower halo
        sFz[buf][ty + 1][tx] = fz;
        sFx[buf][ty + 1][tx] = fx;
        if (ty == 0) {
            T hz = zero_<T>(), hx = zero_<T>();
            if (cell && iy > 0) {
                if (k > 0) hz = F.fz(ix, iy - 1, k);
                if (ix > 0) hx = F.fx(ix, iy - 1, k);
            }
            sFz[buf][0][tx] = hz;
            sFx[buf][0][tx] = hx;
        }
        __syncthreads();
        const T fz_ym = sFz[buf][ty][tx], fx_ym = sFx[buf][ty][tx];

        if (node) {
            const double rhz = k < nz ? ldg(m.rh[2] + k) : 0.0, rhzm = ldg(m.rh[2] + km);
            const bool own_xy = owns_plane(m.d, k), own_z = owns_layer(m.d, k);
            // x-edge (ix, iy, k)
            if (ix < nx) {
                const int64_t id = S.idx(0, ix, iy, k);
                T val = apply_only ? zero_<T>() : ldg(S.p[0] + id);
                if (cell) {
                    T cc = zero_<T>();
                    if (iy > 0 && k > 0) cc = rhy * fz - rhym * fz_ym - rhz * fy + rhzm * fy_dn;
                    const T* et = m.eta[0];
                    const T st = ldg(et + ix + cs1 * iym + cs2 * km) + ldg(et + ix + cs1 * iym + cs2 * k) +
                                 ldg(et + ix + cs1 * iy + cs2 * km) + ldg(et + ix + cs1 * iy + cs2 * k);
                    val -= 0.5 * cc - 0.25 * (st * F.E(0, ix, iy, k));
                }
                if (apply_only) val = -val;
                if (r) R.p[0][id] = val;
                if (own_xy) acc += abs2(val);
            }
            // y-edge
            if (iy < ny) {
                const int64_t id = S.idx(1, ix, iy, k);
                T val = apply_only ? zero_<T>() : ldg(S.p[1] + id);
                if (cell) {
                    T cc = zero_<T>();
                    if (ix > 0 && k > 0) cc = rhz * fx - rhzm * fx_dn - rhx * fz + rhxm * fz_xm;
                    const T* et = m.eta[1];
                    const T st = ldg(et + ixm + cs1 * iy + cs2 * km) + ldg(et + ix + cs1 * iy + cs2 * km) +
                                 ldg(et + ixm + cs1 * iy + cs2 * k) + ldg(et + ix + cs1 * iy + cs2 * k);
                    val -= 0.5 * cc - 0.25 * (st * F.E(1, ix, iy, k));
                }
                if (apply_only) val = -val;
                if (r) R.p[1][id] = val;
                if (own_xy) acc += abs2(val);
            }
            // z-edge
            if (k < nz) {
                const int64_t id = S.idx(2, ix, iy, k);
                T val = apply_only ? zero_<T>() : ldg(S.p[2] + id);
                if (cell) {
                    T cc = zero_<T>();
                    if (ix > 0 && iy > 0) cc = rhx * fy - rhxm * fy_xm - rhy * fx + rhym * fx_ym;
                    const T* et = m.eta[2];
                    const T st = ldg(et + ixm + cs1 * iym + cs2 * k) + ldg(et + ix + cs1 * iym + cs2 * k) +
                                 ldg(et + ixm + cs1 * iy + cs2 * k) + ldg(et + ix + cs1 * iy + cs2 * k);
                    val -= 0.5 * cc - 0.25 * (st * F.E(2, ix, iy, k));
                }
                if (apply_only) val = -val;
                if (r) R.p[2][id] = val;
                if (own_z) acc += abs2(val);
            }
        }
        fx_dn = fx;
        fy_dn = fy;
    }
    if (partial) {
        __shared__ double red[RZ_BX * RZ_BY / 32];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        const int tid = tx + RZ_BX * ty;
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < RZ_BX * RZ_BY / 32; ++w) t += red[w];
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

constexpr int64_t ZMARCH_MIN_CELLS = 100 * 100 * 100;   // below: too few blocks to march
constexpr int ZMARCH_KZ = EMG_RZ_KZ;

template <typename T>
void launch_residual(const Model<T>& m, const T* s, const T* e, T* r, double* norm2_out,
                     double* scratch, int apply_only, cudaStream_t st) {
    dim3 b(32, 4, 2);
    dim3 g((m.d.n[0] + 1 + b.x - 1) / b.x, (m.d.n[1] + 1 + b.y - 1) / b.y,
           (m.d.n[2] + 1 + b.z - 1) / b.z);
    if (n_cells(m.d) >= ZMARCH_MIN_CELLS) {
        b = dim3(RZ_BX, RZ_BY, 1);
        g = dim3((m.d.n[0] + RZ_BX) / RZ_BX, (m.d.n[1] + RZ_BY) / RZ_BY,
                 (m.d.n[2] + ZMARCH_KZ) / ZMARCH_KZ);
        ++g_launch_count; residual_zmarch_kernel<T><<<g, b, 0, st>>>(m, s, e, r, norm2_out ? scratch : nullptr, apply_only, ZMARCH_KZ);
    } else {
        ++g_launch_count; residual_kernel<T><<<g, b, 0, st>>>(m, s, e, r, norm2_out ? scratch : nullptr, apply_only);
    }
    if (norm2_out) {
        int64_t nb = (int64_t)g.x * g.y * g.z;
        ++g_launch_count; sum_partials_kernel<<<1, 1024, 0, st>>>(scratch, nb, norm2_out);
    }
}

int64_t residual_scratch_doubles(const Dims& d) {
    dim3 b(32, 4, 2);     // the simple kernel has the finer grid of the two variants
    return (int64_t)((d.n[0] + 1 + b.x - 1) / b.x) * ((d.n[1] + 1 + b.y - 1) / b.y) *
           ((d.n[2] + 1 + b.z - 1) / b.z);
}

template void launch_residual<double>(const Model<double>&, const double*, const double*, double*,
                                      double*, double*, int, cudaStream_t);
template void launch_residual<cplx>(const Model<cplx>&, const cplx*, const cplx*, cplx*, double*,
                                    double*, int, cudaStream_t);

}  // namespace emg
