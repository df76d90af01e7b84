// Grid-transfer operators of the multigrid cycle.
//
//  restrict       full-weighting restriction of the edge residual
//                 (emg3d/core.py:1620-2001, all seven sc_dir patterns from one
//                 kernel: an axis is either coarsened or not)
//  prolong        e_fine += P e_coarse on interior edges: constant along the
//                 edge, linear between coarse nodes transversally
//                 (emg3d/solver.py:947-1019, 1385-1478)
//  restrict_cells coarse eta / zeta = sum over the 8/4/2 fine cells
//                 (emg3d/solver.py:1667-1718)
//
// All three are pure streaming kernels, one thread per output (coarse edge,
// fine edge, coarse cell), x fastest so that loads and stores coalesce.
#include "common.cuh"
#include "kernels.h"

namespace emg {

struct Axis3 {
    int v[3];
};
struct WPtrs {
    const double* wl[3];
    const double* w0[3];
    const double* wr[3];
};

template <typename T>
__global__ void __launch_bounds__(256)
restrict_kernel(Dims fine, Dims coarse, Axis3 cf, const T* __restrict__ r, T* __restrict__ cr,
                WPtrs w) {
    FieldView<const T> R(r, fine);
    FieldView<T> C(cr, coarse);
    int ci[3];
    ci[0] = blockIdx.x * blockDim.x + threadIdx.x;
    ci[1] = blockIdx.y * blockDim.y + threadIdx.y;
    ci[2] = blockIdx.z * blockDim.z + threadIdx.z;
    if (ci[0] > coarse.n[0] || ci[1] > coarse.n[1] || ci[2] > coarse.n[2]) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (ci[c] >= coarse.n[c]) continue;          // no coarse edge of this component here
        const int u = (c + 1) % 3, v = (c + 2) % 3;
        T acc = zero_<T>();
        const int nu_ = cf.v[u] ? 3 : 1, nv_ = cf.v[v] ? 3 : 1;
        for (int a = 0; a < nu_; ++a) {
            const int ju = cf.v[u] ? a - 1 : 0;
            const double wu = !cf.v[u] ? 1.0
                              : ju < 0 ? ldg(w.wl[u] + ci[u])
                              : ju == 0 ? ldg(w.w0[u] + ci[u]) : ldg(w.wr[u] + ci[u]);
            int fu = cf.v[u] ? 2 * ci[u] + ju : ci[u];
            fu = min(max(fu, 0), fine.n[u]);            // clamp into the array (core.py:1676-1689)
            for (int b = 0; b < nv_; ++b) {
                const int jv = cf.v[v] ? b - 1 : 0;
                const double wv = !cf.v[v] ? 1.0
                                  : jv < 0 ? ldg(w.wl[v] + ci[v])
                                  : jv == 0 ? ldg(w.w0[v] + ci[v]) : ldg(w.wr[v] + ci[v]);
                int fv = cf.v[v] ? 2 * ci[v] + jv : ci[v];
                fv = min(max(fv, 0), fine.n[v]);
                int f[3];
                f[u] = fu; f[v] = fv;
                f[c] = cf.v[c] ? 2 * ci[c] : ci[c];
                T val = ldg(R.p[c] + R.idx(c, f));
                if (cf.v[c]) {
                    f[c] += 1;
                    val += ldg(R.p[c] + R.idx(c, f));
                }
                acc += (wu * wv) * val;
            }
        }
        C.p[c][C.idx(c, ci)] = acc;
    }
}

template <typename T>
void launch_restrict(const Dims& fine, const int* cflag, const T* r, T* cr, const double* const* wl,
                     const double* const* w0, const double* const* wr, cudaStream_t st) {
    Dims coarse = {};
    Axis3 cf;
    WPtrs w;
    for (int a = 0; a < 3; ++a) {
        cf.v[a] = cflag[a];
        coarse.n[a] = cflag[a] ? fine.n[a] / 2 : fine.n[a];
        w.wl[a] = wl[a]; w.w0[a] = w0[a]; w.wr[a] = wr[a];
    }
    dim3 b(32, 4, 2);
    dim3 g((coarse.n[0] + 1 + b.x - 1) / b.x, (coarse.n[1] + 1 + b.y - 1) / b.y,
           (coarse.n[2] + 1 + b.z - 1) / b.z);
    ++g_launch_count; restrict_kernel<T><<<g, b, 0, st>>>(fine, coarse, cf, r, cr, w);
}

struct IPtrs {
    const int* lo[3];
    const double* fr[3];
};

template <typename T>
__global__ void __launch_bounds__(256)
prolong_kernel(Dims fine, Dims coarse, Axis3 cf, T* __restrict__ e, const T* __restrict__ ce,
               IPtrs ip) {
    FieldView<T> E(e, fine);
    FieldView<const T> C(ce, coarse);
    int f[3];
    f[0] = blockIdx.x * blockDim.x + threadIdx.x;
    f[1] = blockIdx.y * blockDim.y + threadIdx.y;
    f[2] = blockIdx.z * blockDim.z + threadIdx.z;
    if (f[0] > fine.n[0] || f[1] > fine.n[1] || f[2] > fine.n[2]) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int u = (c + 1) % 3, v = (c + 2) % 3;
        if (f[c] >= fine.n[c]) continue;
        // interior edges only: PEC boundary values stay untouched (solver.py:984-1019)
        if (f[u] < 1 || f[u] > fine.n[u] - 1 || f[v] < 1 || f[v] > fine.n[v] - 1) continue;
        const int iu = ldg(ip.lo[u] + f[u]), iv = ldg(ip.lo[v] + f[v]);
        const double tu = ldg(ip.fr[u] + f[u]), tv = ldg(ip.fr[v] + f[v]);
        int q[3];
        q[c] = cf.v[c] ? f[c] >> 1 : f[c];
        T acc;
        q[u] = iu; q[v] = iv;
        acc = ((1.0 - tu) * (1.0 - tv)) * ldg(C.p[c] + C.idx(c, q));
        q[v] = iv + 1;
        acc += ((1.0 - tu) * tv) * ldg(C.p[c] + C.idx(c, q));
        q[u] = iu + 1; q[v] = iv;
        acc += (tu * (1.0 - tv)) * ldg(C.p[c] + C.idx(c, q));
        q[v] = iv + 1;
        acc += (tu * tv) * ldg(C.p[c] + C.idx(c, q));
        const int64_t id = E.idx(c, f);
        E.p[c][id] = E.p[c][id] + acc;
    }
}

template <typename T>
void launch_prolong(const Dims& fine, const int* cflag, T* e, const T* ce, const int* const* lo,
                    const double* const* frac, cudaStream_t st) {
    Dims coarse = {};
    Axis3 cf;
    IPtrs ip;
    for (int a = 0; a < 3; ++a) {
        cf.v[a] = cflag[a];
        coarse.n[a] = cflag[a] ? fine.n[a] / 2 : fine.n[a];
        ip.lo[a] = lo[a]; ip.fr[a] = frac[a];
    }
    dim3 b(32, 4, 2);
    dim3 g((fine.n[0] + 1 + b.x - 1) / b.x, (fine.n[1] + 1 + b.y - 1) / b.y,
           (fine.n[2] + 1 + b.z - 1) / b.z);
    ++g_launch_count; prolong_kernel<T><<<g, b, 0, st>>>(fine, coarse, cf, e, ce, ip);
}

template <typename T>
__global__ void __launch_bounds__(256)
restrict_cells_kernel(Dims fine, Dims coarse, Axis3 cf, const T* __restrict__ p, T* __restrict__ cp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= coarse.n[0] || j >= coarse.n[1] || k >= coarse.n[2]) return;
    const int64_t s1 = fine.n[0], s2 = (int64_t)fine.n[0] * fine.n[1];
    const int fi = cf.v[0] ? 2 * i : i, fj = cf.v[1] ? 2 * j : j, fk = cf.v[2] ? 2 * k : k;
    // Summation order of the reference (solver.py:1686-1716): pairs along the
    // first coarsened axis, pairs accumulated with the remaining axes ordered
    // y outer / z inner -- so coarse coefficients are bitwise reproducible.
    const int pa = cf.v[0] ? 0 : cf.v[1] ? 1 : 2;
    const int64_t st[3] = {1, s1, s2};
    const int64_t base = fi + s1 * fj + s2 * fk;
    T acc = zero_<T>();
    bool first = true;
    for (int d1 = 0; d1 <= ((pa != 1 && cf.v[1]) ? 1 : 0); ++d1)
        for (int d2 = 0; d2 <= ((pa != 2 && cf.v[2]) ? 1 : 0); ++d2) {
            const int64_t o = base + s1 * d1 + s2 * d2;
            const T pair = ldg(p + o) + ldg(p + o + st[pa]);
            acc = first ? pair : acc + pair;
            first = false;
        }
    cp[i + (int64_t)coarse.n[0] * (j + (int64_t)coarse.n[1] * k)] = acc;
}

template <typename T>
void launch_restrict_cells(const Dims& fine, const int* cflag, const T* p, T* cp, cudaStream_t st) {
    Dims coarse = {};
    Axis3 cf;
    for (int a = 0; a < 3; ++a) {
        cf.v[a] = cflag[a];
        coarse.n[a] = cflag[a] ? fine.n[a] / 2 : fine.n[a];
    }
    dim3 b(32, 4, 2);
    dim3 g((coarse.n[0] + b.x - 1) / b.x, (coarse.n[1] + b.y - 1) / b.y, (coarse.n[2] + b.z - 1) / b.z);
    ++g_launch_count; restrict_cells_kernel<T><<<g, b, 0, st>>>(fine, coarse, cf, p, cp);
}

#define INST(T)                                                                                     \
    template void launch_restrict<T>(const Dims&, const int*, const T*, T*, const double* const*,  \
                                     const double* const*, const double* const*, cudaStream_t);    \
    template void launch_prolong<T>(const Dims&, const int*, T*, const T*, const int* const*,      \
                                    const double* const*, cudaStream_t);                           \
    template void launch_restrict_cells<T>(const Dims&, const int*, const T*, T*, cudaStream_t);
INST(double)
INST(cplx)

}  // namespace emg
