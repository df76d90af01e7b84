// Line-relaxation Gauss-Seidel smoothers along x, y or z (what
// emg3d/core.py:506-783 `gauss_seidel_x`, 786-1068 `_y`, 1071-1348 `_z` compute):
// all edges touching the interior nodes of one grid line, plus the line's own
// edges, are solved for simultaneously.
//
// Unknowns of a line of N cells: the line edges L_i (cell i, between nodes i and
// i+1), i = 0 .. N-1, and the four transverse edges T_m = [p-, p+, q-, q+] at the
// interior nodes m = 1 .. N-1, (p, q) = (y,z) / (x,z) / (x,y) for x- / y- / z-lines;
// T_0 and T_N lie on the line's end planes and are fixed data (zero on a PEC
// boundary, halo values on a multi-GPU z-window).  The reference orders them in
// 5-unknown blocks [L_i, T_{i+1}] (core.py:775-783) and factorises a banded matrix
// per line and sweep.  Here the line edges are eliminated first: L_i couples only to
// T_i and T_{i+1}, with  c_i = -f_i,
//     L_i = (bL_i - f_i . (T_i - T_{i+1})) / dL_i ,
// which leaves a block-tridiagonal system of 4x4 blocks for the T_m alone,
//     E_m T_{m-1} + D_m T_m + E_{m+1} T_{m+1} = r_m ,
//     D_m = C_m - f_{m-1} f_{m-1}^T / dL_{m-1} - f_m f_m^T / dL_m ,
//     E_m = diag(d_{m-1}) + f_{m-1} f_{m-1}^T / dL_{m-1}      (symmetric; d, f real),
//     r_m = bT_m + f_{m-1} bL_{m-1} / dL_{m-1} - f_m bL_m / dL_m ,
// solved by block elimination:  S_1 = D_1,  S_m = D_m - E_m S_{m-1}^{-1} E_m.
// Same equations, same solution (to rounding) as the reference's block solve.
//
// B200 design: the matrix depends on (grid, model, s) only, not
// on E, so X_m = S_m^{-1} (10 numbers, symmetric) and 1/dL (1 number) are computed
// ONCE per level and direction (`line_factor`) and streamed from HBM by every
// sweep -- 11 numbers per cell instead of the 15 of a 5x5 block LDL^T, and E_m is
// rebuilt from zeta and the widths on the fly (the kernels are bound by the bytes
// they move, with the fp64 pipe below 10 %).  The explicit inverse is stored rather
// than LDL^T factors: applying it is a matrix-vector product without a dependent
// chain (measured 8-17 % faster with one warp per scheduler) and as accurate on the
// golden cases (the inverse is needed by the block recurrence anyway).  A sweep is
// one forward pass
//     g_m = X_m (r_m - E_m g_{m-1}),  g_0 = T_0,
// and one backward pass
//     T_m = g_m - X_m E_{m+1} T_{m+1},   L_m = (bL_m - f_m . (T_m - T_{m+1})) / dL_m,
// one thread per line, g_m and bL_i stored in place in E between the passes.
// Factor layout [group of 32 lines][block][entry][lane] makes the factor stream one
// contiguous chunk per warp and block.
//
// Orderings: `lex` runs hyperplanes t = tp + 2 tq of the reference's line
// order (p fastest, then q; core.py:601-624, 886-917, 1166-1197) and is
// sequentially equivalent to it; `color` runs the four parity classes
// (tp&1, tq&1), which are conflict-free.
#include "common.cuh"
#include "kernels.h"
#include "line_common.cuh"

namespace emg {

// ---- one line sweep: forward and backward block substitution ----------------
//
// What one thread streams per cell of its line:
//   forward  : 11 factor entries, 5 sources, 4 parallel line edges of the next
//              cell, 8 outer transverse edges of the end faces, 4 zeta;
//              writes g_m (4) and bL (1)
//   backward : 11 factor entries, g_m (4), bL (1), 4 zeta; writes T_m (4), L (1)
// PH = 0: the whole sweep; 1: forward pass only, 2: backward pass only (g_m and bL live in E
// between the two) -- the pieces of a z-line cut by multi-GPU z-slabs run the forward pass rank
// after rank upwards (g of the lower piece's last node arrives in the halo plane, where the
// pass reads its start value g_0 = T_0) and the backward pass rank after rank downwards.
template <typename T, int D, int PH = 0>
__device__ void sweep_line(const Line<T, D>& ln, const LineAddr<T, D>& a) {
    using A = Ax<D>;
    const Model<T>& m = ln.m;
    const int N = ln.N;

    double zc[2][2], zn[2][2], gs[4];
    CellCoef cc, cn;
    T rl_c = zero_<T>(), bl_c = zero_<T>();
    if constexpr (PH != 2) {
    // ---------------- forward ----------------
    ln.load_zeta(0, zc);
    ln.side_g(zc, gs);
    cell_coef<T, D>(ln, gs, ldg(m.rh[A::d]), cc);
    T eo_c[4], eo_n[4];                  // outer parallel neighbours of L in the current / next cell
#pragma unroll
    for (int k = 0; k < 4; ++k) eo_c[k] = a.ed[a.oLn[k]];
    rl_c = ldg(a.fac + (int64_t)(N - 1) * FAC_BS);           // 1 / dL_0
    bl_c = line_rhs<T, D>(a, 0, cc, eo_c);
    a.ed[a.oL] = bl_c;
    // g_0 = T_0: the (fixed) transverse edges on the line's start plane: zero on a PEC
    // boundary, halo data of the neighbouring slab on a multi-GPU z-window
    T g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = *a.t_ptr(k, 0);

    for (int i = 0; i < N - 1; ++i) {                        // node mn = i + 1
        const int mn = i + 1;
        const T* fp = a.fac + (int64_t)i * FAC_BS;
        T X[10];
#pragma unroll
        for (int e = 0; e < 10; ++e) X[e] = ldg(fp + e * FAC_ES);
        const T rl_n = ldg(fp + 10 * FAC_ES);
        ln.load_zeta(mn, zn);
        double gn[4];
        ln.side_g(zn, gn);
        cell_coef<T, D>(ln, gn, ldg(m.rh[A::d] + mn), cn);
#pragma unroll
        for (int k = 0; k < 4; ++k) eo_n[k] = a.ed[a.oLn[k] + a.sd * mn];
        const T bl_n = line_rhs<T, D>(a, mn, cn, eo_n);
        // bT_m: sources, side faces of cell i (+) and cell i+1 (-), end faces at node m
        T r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = ldg(a.ts_ptr(k, mn)) + cc.gra[k] * eo_c[k] - cn.gra[k] * eo_n[k];
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) {
                const double gf = 0.5 * (zc[jp][jq] + zn[jp][jq]);
                const double ap = ln.al_p(jq), aq = ln.al_q(jp);
                const T out = ap * a.epo(jp, jq, mn) + aq * a.eqo(jp, jq, mn);
                r[jp] += (gf * ap) * out;
                r[2 + jq] += (gf * aq) * out;
            }
        // r_m = bT_m + f_c bL_c / dL_c - f_n bL_n / dL_n;   v = r_m - E_m g_{m-1}
        const T sc = bl_c * rl_c, sn = bl_n * rl_n;
        T eg[4], v[4];
        apply_E<T>(cc.d, cc.f, rl_c, g, eg);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = r[k] + cc.f[k] * sc - cn.f[k] * sn - eg[k];
        symv4<T>(X, v, g);                                   // g_m = S_m^{-1} v
#pragma unroll
        for (int k = 0; k < 4; ++k) *a.t_ptr(k, mn) = g[k];
        a.ed[a.oL + a.sd * mn] = bl_n;
        // shift
        cc = cn;
        rl_c = rl_n;
        bl_c = bl_n;
#pragma unroll
        for (int k = 0; k < 4; ++k) eo_c[k] = eo_n[k];
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) zc[jp][jq] = zn[jp][jq];
    }
    }
    if constexpr (PH != 1) {
    // ---------------- backward ----------------
    // cc / rl_c / bl_c now belong to the last cell N-1 (PH = 2: reloaded below); T_N is fixed data
    T tn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tn[k] = *a.t_ptr(k, N);
    for (int i = N - 2; i >= 0; --i) {                       // node mn = i + 1, cell mn
        const int mn = i + 1;
        const T* fp = a.fac + (int64_t)i * FAC_BS;
        T X[10];
#pragma unroll
        for (int e = 0; e < 10; ++e) X[e] = ldg(fp + e * FAC_ES);
        if (PH == 2 || i < N - 2) {                          // (the last cell's data are at hand)
            rl_c = ldg(fp + 10 * FAC_ES);
            bl_c = a.ed[a.oL + a.sd * mn];
            ln.load_zeta(mn, zc);
            ln.side_g(zc, gs);
            cell_coef<T, D>(ln, gs, ldg(m.rh[A::d] + mn), cc);
        }
        T gm[4], w[4], xw[4], tm[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) gm[k] = *a.t_ptr(k, mn);
        apply_E<T>(cc.d, cc.f, rl_c, tn, w);                 // E_{m+1} T_{m+1}
        symv4<T>(X, w, xw);
        T fd = zero_<T>();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            tm[k] = gm[k] - xw[k];
            fd += cc.f[k] * (tm[k] - tn[k]);
        }
        a.ed[a.oL + a.sd * mn] = rl_c * (bl_c - fd);         // L_m
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            *a.t_ptr(k, mn) = tm[k];
            tn[k] = tm[k];
        }
    }
    // L_0 = (bL_0 - f_0 . (T_0 - T_1)) / dL_0
    if (PH == 2 || N > 1) {
        rl_c = ldg(a.fac + (int64_t)(N - 1) * FAC_BS);
        bl_c = a.ed[a.oL];
        ln.load_zeta(0, zc);
        ln.side_g(zc, gs);
        cell_coef<T, D>(ln, gs, ldg(m.rh[A::d]), cc);
    }
    T fd = zero_<T>();
#pragma unroll
    for (int k = 0; k < 4; ++k) fd += cc.f[k] * (*a.t_ptr(k, 0) - tn[k]);
    a.ed[a.oL] = rl_c * (bl_c - fd);
    }
}

template <typename T, int D, int PH = 0>
__device__ __forceinline__ void sweep_line_direct(const Model<T>& m, int tp, int tq, const T* fac,
                                                  const LineSlots& ls, const FieldView<T>& E,
                                                  const FieldView<const T>& S) {
    Line<T, D> ln(m, tp, tq);
    LineAddr<T, D> a(E, S, fac + ls.base(ls.slot(tp, tq), ln.N), tp, tq);
    sweep_line<T, D, PH>(ln, a);
}


// ---- kernels ---------------------------------------------------------------
#ifndef EMG_LINE_THREADS
#define EMG_LINE_THREADS 64       // lines (threads) per block of the thread-per-line sweep kernel
#endif

// xin / xout: [slot][10] factors handed across a z-slab cut (see factor_line), or null
template <typename T, int D>
__global__ void __launch_bounds__(64)
line_factor_kernel(Model<T> m, T* fac, LineSlots ls, int c, const T* xin, T* xout) {
    int tp, tq;
    if (!class_line(ls, c, blockIdx.x * blockDim.x + threadIdx.x, tp, tq)) return;
    const int64_t slot = ls.slot(tp, tq);
    factor_line<T, D, 0>(m, tp, tq, fac + ls.base(slot, m.d.n[Ax<D>::d]), xin ? xin + slot * 10 : nullptr,
                         xout ? xout + slot * 10 : nullptr);
}

// lines t0 .. t1 - 1 of colour class c (the whole class, or one batch of it)
template <typename T, int D, int PH>
__global__ void __launch_bounds__(EMG_LINE_THREADS)
gs_line_color_kernel(Model<T> m, const T* fac, LineSlots ls, T* e, const T* s, int c, int t0, int t1) {
    int tp, tq;
    const int t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t1 || !class_line(ls, c, t, tp, tq)) return;
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    sweep_line_direct<T, D, PH>(m, tp, tq, fac, ls, E, S);
}

template <typename T, int D>
__global__ void __launch_bounds__(128)
gs_line_front_kernel(Model<T> m, const T* fac, LineSlots ls, T* e, const T* s, int t) {
    using A = Ax<D>;
    const int tq = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (tq >= m.d.n[A::q]) return;
    const int tp = t - 2 * tq;
    if (tp < 1 || tp >= m.d.n[A::p]) return;
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    sweep_line_direct<T, D>(m, tp, tq, fac, ls, E, S);
}

template <typename T, int D>
__global__ void __launch_bounds__(256)
gs_line_small_kernel(Model<T> m, const T* fac, LineSlots ls, T* e, const T* s, int nu, int order) {
    using A = Ax<D>;
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    const int npi = m.d.n[A::p] - 1, nqi = m.d.n[A::q] - 1;
    bool back = (order >> 8) & 1;   // bits 8+ of `order`: sweeps already done (phase)
    order &= 0xff;
    for (int sw = 0; sw < nu; ++sw) {
        back = !back;
        if (order == ORDER_LEX) {
            const int tmin = 3, tmax = npi + 2 * nqi;
            for (int tt = tmin; tt <= tmax; ++tt) {
                const int t = back ? tmax + tmin - tt : tt;
                for (int b = threadIdx.x; b < nqi; b += blockDim.x) {
                    const int tq = 1 + b, tp = t - 2 * tq;
                    if (tp >= 1 && tp <= npi) sweep_line_direct<T, D>(m, tp, tq, fac, ls, E, S);
                }
                __syncthreads();
            }
        } else {
            for (int cc = 0; cc < 4; ++cc) {
                const int c = back ? 3 - cc : cc;
                if (sw > 0 && cc == 0) continue;   // idempotent repeat, see gs_dir
                for (int t = threadIdx.x; t < ls.cnt[c]; t += blockDim.x) {
                    int tp, tq;
                    class_line(ls, c, t, tp, tq);
                    sweep_line_direct<T, D>(m, tp, tq, fac, ls, E, S);
                }
                __syncthreads();
            }
        }
    }
}

int64_t line_factor_elems(const Dims& d, int dir) {
    const int p = dir == 0 ? 1 : 0, q = dir == 2 ? 1 : 2;
    if (d.n[p] < 2 || d.n[q] < 2) return 0;
    LineSlots ls(d.n[p] - 1, d.n[q] - 1);
    return (int64_t)FAC_BS * d.n[dir] * (ls.nl / 32);
}

int64_t line_chain_elems(const Dims& d, int dir) {
    const int p = dir == 0 ? 1 : 0, q = dir == 2 ? 1 : 2;
    if (d.n[p] < 2 || d.n[q] < 2) return 0;
    LineSlots ls(d.n[p] - 1, d.n[q] - 1);
    return ls.nl * 10;
}

template <typename T, int D>
static void factor_dir(const Model<T>& m, T* fac, const T* xin, T* xout, cudaStream_t st) {
    using A = Ax<D>;
    const int npi = m.d.n[A::p] - 1, nqi = m.d.n[A::q] - 1;
    if (npi < 1 || nqi < 1) return;
    LineSlots ls(npi, nqi);
    for (int c = 0; c < 4; ++c) {
        if (ls.cnt[c] == 0) continue;
        ++g_launch_count; line_factor_kernel<T, D><<<(ls.cnt[c] + 63) / 64, 64, 0, st>>>(m, fac, ls, c, xin, xout);
    }
}

template <typename T>
void launch_line_factor(const Model<T>& m, int dir, T* fac, const T* xin, T* xout, cudaStream_t st) {
    if (dir == 0) factor_dir<T, 0>(m, fac, xin, xout, st);
    else if (dir == 1) factor_dir<T, 1>(m, fac, xin, xout, st);
    else factor_dir<T, 2>(m, fac, xin, xout, st);
}

template <typename T, int D>
static void gs_dir(const Model<T>& m, const T* fac, const T* fac2, T* e, const T* s, int nu, int order,
                   cudaStream_t st) {
    using A = Ax<D>;
    const int npi = m.d.n[A::p] - 1, nqi = m.d.n[A::q] - 1;
    if (npi < 1 || nqi < 1) return;
    LineSlots ls(npi, nqi);
    if (!fac2 && !((order >> 18) & 0x1fff) && (int64_t)npi * nqi <= 1024) {
        int threads = 32;
        const int want = order == ORDER_LEX ? nqi : (npi * nqi + 3) / 4;
        while (threads < 256 && threads < want) threads <<= 1;
        ++g_launch_count; gs_line_small_kernel<T, D><<<1, threads, 0, st>>>(m, fac, ls, e, s, nu, order);
        return;
    }
    bool back = (order >> 8) & 1;   // bit 8 of `order`: sweeps already done (phase)
    // bits 16-17: z-half of a multicolour sweep (see gs_point.cu); x- and y-lines are coloured
    // by (p, z) parity, class index cp + 2 cz
    const int zsel = D == 2 ? 0 : (order >> 16) & 3;
    // bits 18-19: 1 = forward pass only, 2 = backward pass only; bits 20-22: 1 + the one colour
    // class to run (0: all) -- the pieces of z-lines cut by z-slabs (see sweep_line)
    const int phase = (order >> 18) & 3, csel = (order >> 20) & 7;
    // bits 23-26: batch index, bits 27-30: number of batches - 1 (the lines of the selected class
    // are cut into equal batches: the pipeline over the ranks of emg3d_b200/parallel.py)
    const int batch = (order >> 23) & 15, nbatch = ((order >> 27) & 15) + 1;
    order &= 0xff;
    for (int sw = 0; sw < nu; ++sw) {
        back = !back;
        if (order == ORDER_LEX) {
            const int tmin = 3, tmax = npi + 2 * nqi;
            dim3 b(64);
            dim3 g((nqi + b.x - 1) / b.x);
            for (int tt = tmin; tt <= tmax; ++tt) {
                const int t = back ? tmax + tmin - tt : tt;
                ++g_launch_count; gs_line_front_kernel<T, D><<<g, b, 0, st>>>(m, fac, ls, e, s, t);
            }
        } else {
            for (int cc = 0; cc < 4; ++cc) {
                const int cg = back ? 3 - cc : cc;              // class by global z-parity
                const int c = D == 2 ? cg : cg ^ ((m.d.zflip & 1) << 1);   // local class
                if (ls.cnt[c] == 0) continue;
                // Consecutive sweeps run the colours in opposite order, so the first
                // colour of a sweep is the last colour of the previous one.  Lines of one
                // colour do not interact and nothing changed in between: solving them
                // again reproduces the same values (block relaxation is idempotent), so
                // that launch is skipped -- 7 instead of 8 colour launches for nu = 2.
                if (sw > 0 && cc == 0) continue;
                if (zsel && (cg >> 1) != zsel - 1) continue;
                if (csel && cg != csel - 1) continue;
                if (fac2) {          // segment-parallel kernel (gs_line_seg.cu), same colour sequence
                    launch_gs_line_seg_color<T>(m, D, fac2, e, s, c, st);
                    continue;
                }
                const int threads = EMG_LINE_THREADS;
                const int t0 = (int)((int64_t)ls.cnt[c] * batch / nbatch);
                const int t1 = (int)((int64_t)ls.cnt[c] * (batch + 1) / nbatch);
                if (t1 <= t0) continue;
                const dim3 grid((t1 - t0 + threads - 1) / threads);
                ++g_launch_count;
                if (phase == 1) gs_line_color_kernel<T, D, 1><<<grid, threads, 0, st>>>(m, fac, ls, e, s, c, t0, t1);
                else if (phase == 2) gs_line_color_kernel<T, D, 2><<<grid, threads, 0, st>>>(m, fac, ls, e, s, c, t0, t1);
                else gs_line_color_kernel<T, D, 0><<<grid, threads, 0, st>>>(m, fac, ls, e, s, c, t0, t1);
            }
        }
    }
}

template <typename T>
void launch_gs_line(const Model<T>& m, int dir, const T* fac, const T* fac2, T* e, const T* s, int nu,
                    int order, cudaStream_t st) {
    if (dir == 0) gs_dir<T, 0>(m, fac, fac2, e, s, nu, order, st);
    else if (dir == 1) gs_dir<T, 1>(m, fac, fac2, e, s, nu, order, st);
    else gs_dir<T, 2>(m, fac, fac2, e, s, nu, order, st);
}

template void launch_line_factor<double>(const Model<double>&, int, double*, const double*, double*, cudaStream_t);
template void launch_line_factor<cplx>(const Model<cplx>&, int, cplx*, const cplx*, cplx*, cudaStream_t);
template void launch_gs_line<double>(const Model<double>&, int, const double*, const double*, double*,
                                     const double*, int, int, cudaStream_t);
template void launch_gs_line<cplx>(const Model<cplx>&, int, const cplx*, const cplx*, cplx*, const cplx*, int,
                                   int, cudaStream_t);

}  // namespace emg
