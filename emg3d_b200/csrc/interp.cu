// Interpolation kernels on either side of a solve:
//
//   volume_average_kernel      emg3d/maps.py:556-617 `interp_volume_average` (+ the log10 /
//                              10** wrapping of maps.interpolate(method='volume', log=True),
//                              maps.py:306-369): model properties from one tensor grid to another
//   edges_to_vol_kernel        maps.py:668-720 `interp_edges_to_vol_averages`, optionally fused
//                              with Re(b * s mu0 * e) of Simulation.gradient
//                              (emg3d/simulations.py:1028-1046)
//   spline_filter_kernel /     cubic-spline interpolation of fields: maps.py:500-553
//   spline_eval_kernel         `interp_spline_3d` = scipy.ndimage.map_coordinates(order=3); used by
//                              get_receiver (emg3d/fields.py:522-615) and
//                              Field.interpolate_to_grid (fields.py:303-346)
//   linear_eval_kernel         scipy RegularGridInterpolator(method='linear') (maps.py:355-362)
//
// All are streaming / gather kernels on device-resident data (the field stays on the device
// after a solve: sampling receivers there avoids the D2H copy of the whole field).  The
// per-axis bookkeeping (merged-node weights, index coordinates) is O(n) per axis and prepared
// by the host side (emg3d_b200/maps.py).
#include "common.cuh"
#include "kernels.h"

#include <type_traits>

namespace emg {

// ---- volume averaging ---------------------------------------------------------------------
// Per axis: the merged segments (weights w, input cell i_in) sorted by output cell, start[o] ..
// start[o + 1] = the segments of output cell o.  One thread per output cell; the sums run z
// outermost, x innermost, like the reference's loops.
struct AxisCsr {
    const double* w;
    const int* iin;
    const int* start;
};

__global__ void __launch_bounds__(256)
volume_average_kernel(const double* __restrict__ values, int nx, int ny, double* __restrict__ out, int mx,
                      int my, int mz, AxisCsr ax, AxisCsr ay, AxisCsr az, const double* __restrict__ hx,
                      const double* __restrict__ hy, const double* __restrict__ hz, int log_scale, int add) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)mx * my * mz) return;
    const int ox = (int)(t % mx), oy = (int)((t / mx) % my), oz = (int)(t / ((int64_t)mx * my));
    double acc = add ? out[t] : 0.0;
    for (int sz = ldg(az.start + oz); sz < ldg(az.start + oz + 1); ++sz) {
        const double wz = ldg(az.w + sz);
        const int64_t kz = (int64_t)ldg(az.iin + sz) * ny;
        for (int sy = ldg(ay.start + oy); sy < ldg(ay.start + oy + 1); ++sy) {
            const double wzy = wz * ldg(ay.w + sy);
            const int64_t row = (kz + ldg(ay.iin + sy)) * nx;
            for (int sx = ldg(ax.start + ox); sx < ldg(ax.start + ox + 1); ++sx) {
                double v = ldg(values + row + ldg(ax.iin + sx));
                if (log_scale) v = log10(v);
                acc += wzy * ldg(ax.w + sx) * v;
            }
        }
    }
    acc /= ldg(hx + ox) * ldg(hy + oy) * ldg(hz + oz);
    out[t] = log_scale ? pow(10.0, acc) : acc;
}

void launch_volume_average(const double* values, int nx, int ny, double* out, int mx, int my, int mz,
                           const double* const* w, const int* const* iin, const int* const* start,
                           const double* const* hnew, int log_scale, int add, cudaStream_t st) {
    const int64_t n = (int64_t)mx * my * mz;
    AxisCsr a[3];
    for (int k = 0; k < 3; ++k) { a[k].w = w[k]; a[k].iin = iin[k]; a[k].start = start[k]; }
    ++g_launch_count;
    volume_average_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(values, nx, ny, out, mx, my, mz, a[0], a[1],
                                                                       a[2], hnew[0], hnew[1], hnew[2], log_scale, add);
}

// ---- edges -> volume-weighted cell averages -------------------------------------------------
// Cell (ix, iy, iz) collects the four edges of each direction around it, each weighted
// volume / 4; an edge row on the boundary adds to the same cell twice (the reference clamps
// both of its target cells into the grid, maps.py:697-719).
// GRAD: the edge value is Re(b * smu0 * e) of two complex fields (Simulation.gradient).
template <typename T, bool GRAD>
__global__ void __launch_bounds__(256)
edges_to_vol_kernel(Dims d, const T* __restrict__ e, const cplx* __restrict__ b, cplx smu0,
                    const double* __restrict__ hx, const double* __restrict__ hy,
                    const double* __restrict__ hz, typename std::conditional<GRAD, double, T>::type* out) {
    using O = typename std::conditional<GRAD, double, T>::type;
    const int nx = d.n[0], ny = d.n[1], nz = d.n[2];
    const int64_t nc = (int64_t)nx * ny * nz;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nc) return;
    const int c[3] = {(int)(t % nx), (int)((t / nx) % ny), (int)(t / ((int64_t)nx * ny))};
    const double vol4 = ldg(hx + c[0]) * ldg(hy + c[1]) * ldg(hz + c[2]) * 0.25;
    FieldView<const T> E(e, d);
    FieldView<const cplx> B(b, d);
#pragma unroll
    for (int comp = 0; comp < 3; ++comp) {
        const int u = (comp + 1) % 3, v = (comp + 2) % 3;
        O acc = zero_<O>();
#pragma unroll
        for (int du = 0; du < 2; ++du)
#pragma unroll
            for (int dv = 0; dv < 2; ++dv) {
                int q[3];
                q[comp] = c[comp]; q[u] = c[u] + du; q[v] = c[v] + dv;
                // boundary rows count twice
                double wgt = 1.0;
                if ((du == 0 && c[u] == 0) || (du == 1 && c[u] == d.n[u] - 1)) wgt *= 2.0;
                if ((dv == 0 && c[v] == 0) || (dv == 1 && c[v] == d.n[v] - 1)) wgt *= 2.0;
                // (a doubled row replaces the row the clamped cell index would have taken from
                // outside: with one cell along an axis both rows are doubled)
                const int64_t id = E.idx(comp, q);
                if constexpr (GRAD) {
                    const cplx ev = ldg(reinterpret_cast<const cplx*>(E.p[comp]) + id);
                    const cplx bv = ldg(B.p[comp] + id);
                    const cplx pr = bv * smu0 * ev;
                    acc += wgt * (vol4 * pr.re);
                } else {
                    acc += wgt * (vol4 * ldg(E.p[comp] + id));
                }
            }
        out[comp * nc + t] = acc;
    }
}

template <typename T>
void launch_edges_to_vol(const Dims& d, const T* e, const double* hx, const double* hy, const double* hz, T* out,
                         cudaStream_t st) {
    const int64_t nc = n_cells(d);
    ++g_launch_count;
    edges_to_vol_kernel<T, false><<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(d, e, nullptr, make_c(0, 0), hx, hy, hz, out);
}
void launch_gradient_field(const Dims& d, const cplx* e, const cplx* b, cplx smu0, const double* hx,
                           const double* hy, const double* hz, double* out, cudaStream_t st) {
    const int64_t nc = n_cells(d);
    ++g_launch_count;
    edges_to_vol_kernel<cplx, true><<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(d, e, b, smu0, hx, hy, hz, out);
}
template void launch_edges_to_vol<double>(const Dims&, const double*, const double*, const double*, const double*,
                                          double*, cudaStream_t);
template void launch_edges_to_vol<cplx>(const Dims&, const cplx*, const double*, const double*, const double*, cplx*,
                                        cudaStream_t);

// ---- cubic B-spline prefilter (scipy.ndimage.spline_filter, order 3) -------------------------
// One thread per line of the (n0, n1, n2) array (x fastest) along `axis`, in place: gain 6, causal
// and anti-causal recursion with pole sqrt(3) - 2; boundary rule `reflect` = 0: mirror (whole-sample
// symmetric; modes 'constant' / 'mirror' of SciPy), 1: reflect (half-sample; mode 'nearest' after
// its 12-sample edge padding).  The start-up sums are truncated where pole^i < 1e-18 relative.
template <typename T>
__global__ void __launch_bounds__(128)
spline_filter_kernel(T* __restrict__ data, int n0, int n1, int n2, int axis, int reflect) {
    const int dims[3] = {n0, n1, n2};
    const int64_t strides[3] = {1, n0, (int64_t)n0 * n1};
    const int n = dims[axis];
    const int a = (axis + 1) % 3, b = (axis + 2) % 3;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)dims[a] * dims[b] || n < 2) return;
    // (thread index runs over the lower of the two other axes first: coalesced for axis != 0)
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    T* c = data + (t % dims[lo]) * strides[lo] + (t / dims[lo]) * strides[hi];
    const int64_t s = strides[axis];
    const double z = -0.26794919243112270647;        // sqrt(3) - 2
    const double gain = (1.0 - z) * (1.0 - 1.0 / z);
    const int horizon = min(n, 48);                  // |z|^48 = 4e-28
    // causal initialisation
    T c0;
    if (!reflect) {
        const double zn1 = n - 1 < 1000 ? pow(z, (double)(n - 1)) : 0.0;
        c0 = gain * c[0] + zn1 * (gain * c[(n - 1) * s]);
        double zi = z;
        for (int i = 1; i < n - 1 && i < horizon; ++i) {
            c0 += zi * (gain * c[i * s]);
            zi *= z;
        }
        if (zn1 != 0.0) {                            // short lines: the far end still matters
            zi = z;
            for (int i = 1; i < n - 1 && i < horizon; ++i) {
                c0 += (zi * zn1) * (gain * c[(n - 1 - i) * s]);
                zi *= z;
            }
        }
        c0 = c0 * (1.0 / (1.0 - zn1 * zn1));
    } else {
        const double zn = n < 1000 ? pow(z, (double)n) : 0.0;
        T sum = gain * c[0] + zn * (gain * c[(n - 1) * s]);
        double zi = z;
        for (int i = 1; i < n && i < horizon; ++i) {
            sum += zi * (gain * c[i * s]);
            zi *= z;
        }
        if (zn != 0.0) {
            zi = z;
            for (int i = 1; i < n && i < horizon; ++i) {
                sum += (zi * zn) * (gain * c[(n - 1 - i) * s]);
                zi *= z;
            }
        }
        c0 = gain * c[0] + sum * (z / (1.0 - zn * zn));
    }
    // causal pass
    T prev = c0;
    c[0] = c0;
    for (int i = 1; i < n; ++i) {
        prev = gain * c[i * s] + z * prev;
        c[i * s] = prev;
    }
    // anti-causal initialisation and pass
    T last;
    if (!reflect) last = (z / (z * z - 1.0)) * (c[(n - 1) * s] + z * c[(n - 2) * s]);
    else last = (z / (z - 1.0)) * c[(n - 1) * s];
    c[(n - 1) * s] = last;
    for (int i = n - 2; i >= 0; --i) {
        last = z * (last - c[i * s]);
        c[i * s] = last;
    }
}

template <typename T>
void launch_spline_filter(T* data, int n0, int n1, int n2, int reflect, cudaStream_t st) {
    const int dims[3] = {n0, n1, n2};
    for (int axis = 0; axis < 3; ++axis) {
        const int64_t lines = (int64_t)dims[(axis + 1) % 3] * dims[(axis + 2) % 3];
        ++g_launch_count;
        spline_filter_kernel<T><<<(unsigned)((lines + 127) / 128), 128, 0, st>>>(data, n0, n1, n2, axis, reflect);
    }
}
template void launch_spline_filter<double>(double*, int, int, int, int, cudaStream_t);
template void launch_spline_filter<cplx>(cplx*, int, int, int, int, cudaStream_t);

// replicate the edge samples npad times on every side (np.pad(mode='edge')): dst is
// (n0 + 2 npad, n1 + 2 npad, n2 + 2 npad)
template <typename T>
__global__ void __launch_bounds__(256)
pad_edge_kernel(const T* __restrict__ src, int n0, int n1, int n2, int npad, T* __restrict__ dst) {
    const int m0 = n0 + 2 * npad, m1 = n1 + 2 * npad, m2 = n2 + 2 * npad;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)m0 * m1 * m2) return;
    const int i = min(max((int)(t % m0) - npad, 0), n0 - 1);
    const int j = min(max((int)((t / m0) % m1) - npad, 0), n1 - 1);
    const int k = min(max((int)(t / ((int64_t)m0 * m1)) - npad, 0), n2 - 1);
    dst[t] = src[i + (int64_t)n0 * (j + (int64_t)n1 * k)];
}
template <typename T>
void launch_pad_edge(const T* src, int n0, int n1, int n2, int npad, T* dst, cudaStream_t st) {
    const int64_t n = (int64_t)(n0 + 2 * npad) * (n1 + 2 * npad) * (n2 + 2 * npad);
    ++g_launch_count;
    pad_edge_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, n0, n1, n2, npad, dst);
}
template void launch_pad_edge<double>(const double*, int, int, int, int, double*, cudaStream_t);
template void launch_pad_edge<cplx>(const cplx*, int, int, int, int, cplx*, cudaStream_t);

// sub-box [lo, lo + m) of an (n0, n1, n2) array (x fastest) into a compact (m0, m1, m2) array
template <typename T>
__global__ void __launch_bounds__(256)
copy_box_kernel(const T* __restrict__ src, int n0, int n1, int l0, int l1, int l2, int m0, int m1, int m2,
                T* __restrict__ dst) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)m0 * m1 * m2) return;
    const int i = (int)(t % m0), j = (int)((t / m0) % m1), k = (int)(t / ((int64_t)m0 * m1));
    dst[t] = ldg(src + (l0 + i) + (int64_t)n0 * ((l1 + j) + (int64_t)n1 * (l2 + k)));
}
template <typename T>
void launch_copy_box(const T* src, int n0, int n1, const int* lo, const int* m, T* dst, cudaStream_t st) {
    const int64_t n = (int64_t)m[0] * m[1] * m[2];
    ++g_launch_count;
    copy_box_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, n0, n1, lo[0], lo[1], lo[2], m[0], m[1], m[2], dst);
}
template void launch_copy_box<double>(const double*, int, int, const int*, const int*, double*, cudaStream_t);
template void launch_copy_box<cplx>(const cplx*, int, int, const int*, const int*, cplx*, cudaStream_t);

// ---- evaluation at points ------------------------------------------------------------------
// Points: scattered (tensor = 0: point p has coordinates cx[p], cy[p], cz[p]) or a tensor grid
// (tensor = 1: p = a + m0 (b + m1 c) has cx[a], cy[b], cz[c]).  Coordinates are in index units of
// the UNPADDED array.  out[p] = scale * value (+ out[p] if accumulate).
struct Points {
    const double *cx, *cy, *cz;
    int64_t npts;
    int tensor, m0, m1;
    __device__ __forceinline__ void get(int64_t p, double c[3]) const {
        if (tensor) {
            c[0] = ldg(cx + p % m0); c[1] = ldg(cy + (p / m0) % m1); c[2] = ldg(cz + p / ((int64_t)m0 * m1));
        } else {
            c[0] = ldg(cx + p); c[1] = ldg(cy + p); c[2] = ldg(cz + p);
        }
    }
};

__device__ __forceinline__ void bspline3_weights(double x, double w[4]) {
    w[1] = (x * x * (x - 2.0) * 3.0 + 4.0) / 6.0;
    const double zc = 1.0 - x;
    w[2] = (zc * zc * (zc - 2.0) * 3.0 + 4.0) / 6.0;
    w[0] = zc * zc * zc / 6.0;
    w[3] = 1.0 - w[0] - w[1] - w[2];
}

// mode 0: 'constant' (outside the data: cval; spline support mirrored), 1: 'nearest' (coef is
// the prefiltered PADDED array, npad samples per side; coordinates clamped to it)
template <typename T>
__global__ void __launch_bounds__(128)
spline_eval_kernel(const T* __restrict__ coef, int n0, int n1, int n2, int npad, int mode, T cval, Points pts,
                   T scale, int accumulate, T* __restrict__ out) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= pts.npts) return;
    double c[3];
    pts.get(p, c);
    const int n[3] = {n0, n1, n2};
    const int m[3] = {n0 + 2 * npad, n1 + 2 * npad, n2 + 2 * npad};
    bool constant = false;
    int idx[3][4];
    double w[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double cc = c[d];
        if (mode == 0 && (!(cc >= 0.0) || cc > n[d] - 1)) constant = true;   // (NaN coordinates too)
        cc += npad;
        if (mode == 1) cc = fmin(fmax(cc, 0.0), (double)(m[d] - 1));
        if (constant) cc = 0.0;
        const double f = floor(cc);
        bspline3_weights(cc - f, w[d]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int ii = (int)f - 1 + k;
            if (mode == 0) {
                if (ii < 0) ii = -ii;
                if (ii > m[d] - 1) ii = 2 * (m[d] - 1) - ii;
                ii = min(max(ii, 0), m[d] - 1);                            // (arrays of one sample)
            } else {
                ii = min(max(ii, 0), m[d] - 1);
            }
            idx[d][k] = ii;
        }
    }
    T val = cval;
    if (!constant) {
        val = zero_<T>();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            T si = zero_<T>();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                T sj = zero_<T>();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    sj += w[2][k] * ldg(coef + idx[0][i] + (int64_t)m[0] * (idx[1][j] + (int64_t)m[1] * idx[2][k]));
                si += w[1][j] * sj;
            }
            val += w[0][i] * si;
        }
    }
    val = scale * val;
    out[p] = accumulate ? out[p] + val : val;
}

template <typename T>
void launch_spline_eval(const T* coef, int n0, int n1, int n2, int npad, int mode, T cval, const double* cx,
                        const double* cy, const double* cz, int64_t npts, int tensor, int m0, int m1, T scale,
                        int accumulate, T* out, cudaStream_t st) {
    Points pts{cx, cy, cz, npts, tensor, m0, m1};
    ++g_launch_count;
    spline_eval_kernel<T><<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(coef, n0, n1, n2, npad, mode, cval, pts,
                                                                          scale, accumulate, out);
}
template void launch_spline_eval<double>(const double*, int, int, int, int, int, double, const double*, const double*,
                                         const double*, int64_t, int, int, int, double, int, double*, cudaStream_t);
template void launch_spline_eval<cplx>(const cplx*, int, int, int, int, int, cplx, const double*, const double*,
                                       const double*, int64_t, int, int, int, cplx, int, cplx*, cudaStream_t);

// Linear interpolation (RegularGridInterpolator): the coordinate of a point along an axis is
// i + t with the lower grid index i (0 .. n - 2) and the normalised distance t (outside [0, 1]:
// extrapolation); NaN marks a point outside the grid that gets `fill`.
template <typename T>
__global__ void __launch_bounds__(128)
linear_eval_kernel(const T* __restrict__ data, int n0, int n1, int n2, T fill, Points pts, T scale,
                   int accumulate, T* __restrict__ out) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= pts.npts) return;
    double c[3];
    pts.get(p, c);
    const int n[3] = {n0, n1, n2};
    bool outside = false;
    int i0[3];
    double t[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (!(c[d] == c[d])) { outside = true; c[d] = 0.0; }
        const int hi = max(n[d] - 2, 0);
        i0[d] = min(max((int)floor(c[d]), 0), hi);
        t[d] = c[d] - i0[d];
    }
    T val = fill;
    if (!outside) {
        val = zero_<T>();
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const double wgt = (a ? t[0] : 1.0 - t[0]) * (b ? t[1] : 1.0 - t[1]) * (k ? t[2] : 1.0 - t[2]);
                    const int ia = min(i0[0] + a, n0 - 1), ib = min(i0[1] + b, n1 - 1), ik = min(i0[2] + k, n2 - 1);
                    val += wgt * ldg(data + ia + (int64_t)n0 * (ib + (int64_t)n1 * ik));
                }
    }
    val = scale * val;
    out[p] = accumulate ? out[p] + val : val;
}

template <typename T>
void launch_linear_eval(const T* data, int n0, int n1, int n2, T fill, const double* cx, const double* cy,
                        const double* cz, int64_t npts, int tensor, int m0, int m1, T scale, int accumulate, T* out,
                        cudaStream_t st) {
    Points pts{cx, cy, cz, npts, tensor, m0, m1};
    ++g_launch_count;
    linear_eval_kernel<T><<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(data, n0, n1, n2, fill, pts, scale, accumulate,
                                                                          out);
}
template void launch_linear_eval<double>(const double*, int, int, int, double, const double*, const double*,
                                         const double*, int64_t, int, int, int, double, int, double*, cudaStream_t);
template void launch_linear_eval<cplx>(const cplx*, int, int, int, cplx, const double*, const double*, const double*,
                                       int64_t, int, int, int, cplx, int, cplx*, cudaStream_t);

}  // namespace emg
