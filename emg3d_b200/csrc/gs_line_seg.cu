// Segment-parallel line smoother: one WARP per line, every lane owns a segment of
// 8 consecutive blocks of the line's block-tridiagonal system (same equations as
// gs_line.cu, i.e. what emg3d/core.py:506-783 `gauss_seidel_x` -- and 786-1348 for
// y / z -- solve per line; multicolour order only).
//
// Why: the one-thread-per-line kernel of gs_line.cu streams the cached factors from
// HBM twice per sweep (forward and backward substitution), parks the forward
// intermediate in E (HBM) and runs 16 k threads per colour launch (r1: 0.93 kB moved
// per cell-sweep against 184 B algorithmic, 5 % occupancy).  Here
//   * the factors of a line (X_m = S_m^{-1}, 1/dL_m: 11 numbers per cell) come from
//     HBM ONCE per sweep, as one TMA bulk copy (cp.async.bulk, mbarrier completion)
//     into shared memory, and are reused from there by all passes;
//   * the intermediate vectors never leave the chip (registers);
//   * field data are read and written with coalesced accesses in a cell-parallel
//     layout and transposed to / from the segment layout through shared memory;
//   * the line recurrences are cut into 32 independent segments by precomputed JUMP
//     matrices (products of the recurrence matrices over a segment; they depend on
//     the model only and are cached with the factors: 2 x 16 numbers per 8 cells).
//
// The block elimination of gs_line.cu is
//     forward   g_m = X_m (r_m - E_m g_{m-1}) = p_m + M_m g_{m-1},   M_m = -X_m E_m,  g_0 = T_0
//     backward  T_m = g_m - X_m E_{m+1} T_{m+1} = g_m + N_m T_{m+1}, N_m = -X_m E_{m+1}
// With a segment q = blocks a..b:
//     g_b = g~_b + Phi_q g_{a-1},  Phi_q = M_b ... M_a   (g~: the recurrence started from 0)
//     T_a = T~_a + Psi_q T_{b+1},  Psi_q = N_a ... N_b   (T~: the recurrence started from 0)
// so that each direction is: local pass from zero (all segments in parallel), a short
// scan over the segment ends with the jump matrices (4x4 matrix-vector products),
// and a local correction pass that propagates the incoming value through the segment.
// Same solution as the sequential elimination up to rounding (sums are associated
// differently); the parity tests of the line smoothers apply unchanged.
#include "common.cuh"
#include "kernels.h"
#include "line_common.cuh"

#include <stdlib.h>

namespace emg {

constexpr int SEG_K = 8;                       // blocks per lane
#ifndef SEG_PF_W
#define SEG_PF_W 0                             // L2 prefetch of the line's cached data at the start (measured: -8 %)
#endif
#ifndef SEG_PF_ROWS
#define SEG_PF_ROWS 1                          // L2 prefetch of the field rows of an x-line
#endif

// Layout of one line's chunk of the cached data (elements of T), QP = lanes per line:
//   [ W: (j, e, q) at (j * FAC_NE + e) * QP + q  | header: 1/dL_0 + 7 pad | Phi: (q, r, c) | Psi: (q, r, c) ]
// block i (node m = i + 1) = 8 q + j; entries e = 0..9: X_m (packed lower triangle), e = 10: 1/dL_m.
template <int QP>
struct SegL {
    static constexpr int HDR = SEG_K * FAC_NE * QP;
    static constexpr int W = HDR + 8;           // what the bulk copy brings on chip
    static constexpr int PHI = W;
    static constexpr int PSI = W + 16 * QP;
    static constexpr int LINE = W + 32 * QP;
    static constexpr int ST = 9 * QP;           // padded length of a staged component
    static constexpr int ZR = 9 * QP + 2;       // padded length of a staged zeta row
};
__device__ __forceinline__ constexpr int pad8(int i) { return i + (i >> 3); }

// shared memory of one warp (bytes): factor / staging region, zeta rows, scan buffer, the lines'
// end data (T_0, T_N, bL_0), barrier
template <typename T>
constexpr int seg_warp_smem() {
    return 2832 * (int)sizeof(T) + 4 * 292 * (int)sizeof(double) + (128 + 20) * (int)sizeof(T) + 16;
}

__device__ __forceinline__ double shfl_t(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ cplx shfl_t(cplx v, int src) {
    return make_c(__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src));
}
// value of the lane below within groups of `width` lanes (the lowest lane keeps its own)
__device__ __forceinline__ double shfl_up_t(double v, int width) { return __shfl_up_sync(0xffffffffu, v, 1, width); }
__device__ __forceinline__ cplx shfl_up_t(cplx v, int width) {
    return make_c(__shfl_up_sync(0xffffffffu, v.re, 1, width), __shfl_up_sync(0xffffffffu, v.im, 1, width));
}

// acc += a * b with every product fused into the accumulation (4 DFMA per complex
// multiply-add instead of 2 DMUL + 2 DFMA + 2 DADD)
__device__ __forceinline__ void fma_t(double& acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void fma_t(cplx& acc, cplx a, cplx b) {
    acc.re = fma(a.re, b.re, acc.re);
    acc.im = fma(a.re, b.im, acc.im);
    acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.im, b.re, acc.im);
}
__device__ __forceinline__ void fma_t(cplx& acc, double a, cplx b) {
    acc.re = fma(a, b.re, acc.re);
    acc.im = fma(a, b.im, acc.im);
}
template <typename T>
__device__ __forceinline__ void seg_symv4(const T f[10], const T v[4], T out[4]) {
    // out = S^{-1} v by substitution with the LDL^T factors (line_common.cuh: ldl4_factor), every
    // product fused into its accumulation
    T y0 = v[0], y1 = v[1], y2 = v[2], y3 = v[3];
    fma_t(y1, -f[tri(1, 0)], y0);
    fma_t(y2, -f[tri(2, 0)], y0);
    fma_t(y3, -f[tri(3, 0)], y0);
    fma_t(y2, -f[tri(2, 1)], y1);
    fma_t(y3, -f[tri(3, 1)], y1);
    fma_t(y3, -f[tri(3, 2)], y2);
    y0 = f[tri(0, 0)] * y0;
    y1 = f[tri(1, 1)] * y1;
    y2 = f[tri(2, 2)] * y2;
    y3 = f[tri(3, 3)] * y3;
    fma_t(y2, -f[tri(3, 2)], y3);
    fma_t(y1, -f[tri(3, 1)], y3);
    fma_t(y0, -f[tri(3, 0)], y3);
    fma_t(y1, -f[tri(2, 1)], y2);
    fma_t(y0, -f[tri(2, 0)], y2);
    fma_t(y0, -f[tri(1, 0)], y1);
    out[0] = y0; out[1] = y1; out[2] = y2; out[3] = y3;
}
// E v with E = diag(d) + rl f f^T
template <typename T>
__device__ __forceinline__ void seg_apply_E(const double d[4], const double f[4], T rl, const T v[4], T out[4]) {
    T f0 = f[0] * v[0];
    T f1 = f[2] * v[2];
    fma_t(f0, f[1], v[1]);
    fma_t(f1, f[3], v[3]);
    const T a = rl * (f0 + f1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        out[k] = d[k] * v[k];
        fma_t(out[k], f[k], a);
    }
}

// coefficients of line cell `cell` from the staged zeta rows: f (L <-> T) and d (T <-> T)
template <typename T, int D, int QP>
__device__ __forceinline__ void seg_cell(const Line<T, D>& ln, const double* __restrict__ zs, int cell,
                                         double f[4], double d[4]) {
    using A = Ax<D>;
    const int pc = pad8(cell);
    const double z00 = zs[pc], z01 = zs[SegL<QP>::ZR + pc], z10 = zs[2 * SegL<QP>::ZR + pc],
                 z11 = zs[3 * SegL<QP>::ZR + pc];
    double gs[4];
    gs[0] = 0.5 * (z00 + z01);
    gs[1] = 0.5 * (z10 + z11);
    gs[2] = 0.5 * (z00 + z10);
    gs[3] = 0.5 * (z01 + z11);
    const double rd = ldg(ln.m.rh[A::d] + min(cell, ln.N - 1));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double as = ln.a_side(k);
        f[k] = -(gs[k] * rd * as);
        d[k] = -(gs[k] * rd * rd);
    }
}

// Sequential scan over the Qn segment ends of a line: sb[qq] <- sb[qq] + J_qq * (previous result),
// forward (qq = 0 .. Qn-1) or backward (qq = Qn-1 .. 0); G = the value entering the first step.
// The jump matrices come in batches of QP / 4 by one coalesced load: lane q = 4 u + r holds row r
// of matrix `batch + u` and the matching entry of sb, and computes that row of step u itself, so
// a step is 4 fused complex multiply-adds and one broadcast of the new 4-vector (no memory).
template <typename T, int QP, bool FWD>
__device__ __forceinline__ void seg_scan(const T* __restrict__ jump, T* __restrict__ sb, T G[4], int Qn,
                                         int q, int sub) {
    constexpr int MB = QP / 4;                   // matrices per batch
    const int nbatch = (Qn + MB - 1) / MB;
    const int mu = q >> 2;                       // my matrix within a batch
    auto fetch = [&](int bb, T ph[4], T& v) {
        const int b = FWD ? bb * MB : (nbatch - 1 - bb) * MB;
        const T* p = jump + (b + mu) * 16 + (q & 3) * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) ph[c] = ldg(p + c);
        v = sb[(b + mu) * 4 + (q & 3)];
    };
    T ph[4], v;
    fetch(0, ph, v);
    for (int bb = 0; bb < nbatch; ++bb) {
        const int b = FWD ? bb * MB : (nbatch - 1 - bb) * MB;
        T ph2[4], v2;                            // next batch: in flight during this one
        if (bb + 1 < nbatch) fetch(bb + 1, ph2, v2);
        T res = v;
#pragma unroll
        for (int uu = 0; uu < MB; ++uu) {
            const int u = FWD ? uu : MB - 1 - uu;
            if (b + u < Qn) {                    // uniform over the warp
                T s0 = ph[0] * G[0], s1 = ph[2] * G[2];
                fma_t(s0, ph[1], G[1]);
                fma_t(s1, ph[3], G[3]);
                const T s_ = v + (s0 + s1);
                if (mu == u) res = s_;
#pragma unroll
                for (int k = 0; k < 4; ++k) G[k] = shfl_t(s_, sub * QP + 4 * u + k);
            }
        }
        if (b + mu < Qn) sb[(b + mu) * 4 + (q & 3)] = res;
        if (bb + 1 < nbatch) {
#pragma unroll
            for (int c = 0; c < 4; ++c) ph[c] = ph2[c];
            v = v2;
        }
    }
}

// 1/dL of line cell `cell` = 8 q + jc (jc = 0..8) from the staged factors
template <typename T, int QP>
__device__ __forceinline__ T seg_rl(const T* __restrict__ wA, int q, int jc) {
    if (jc > 0) return wA[((jc - 1) * FAC_NE + 10) * QP + q];
    return q == 0 ? wA[SegL<QP>::HDR] : wA[(7 * FAC_NE + 10) * QP + q - 1];
}

// One multicolour relaxation of the line (tp, tq), executed by the QP lanes lane % QP of a
// warp (32 / QP lines per warp).  `store` = false: everything but the final stores (the
// padding line of a partially filled warp).
template <typename T, int D, int QP>
__device__ void seg_line_sweep(const Model<T>& m, const T* __restrict__ fac2, const LineSlots& ls,
                               const FieldView<T>& E, const FieldView<const T>& S, int tp, int tq,
                               bool store, T* __restrict__ smA, double* __restrict__ smZ,
                               T* __restrict__ smS, T* __restrict__ smT, uint64_t* bar, unsigned& parity) {
    using A = Ax<D>;
    using SL = SegL<QP>;
    constexpr int LPW = 32 / QP;
    const int lane = threadIdx.x & 31;
    const int q = lane % QP, sub = lane / QP;
    T* const wA = smA + sub * SL::W;
    double* const zs = smZ + sub * 4 * SL::ZR;
    T* const sb = smS + sub * QP * 4;
    T* const se = smT + sub * 10;                // [0..3] T_0, [4..7] T_N, [8] bL_0 (fixed data of the line)

    const Line<T, D> ln(m, tp, tq);
    const int N = ln.N, NB = N - 1;              // cells, blocks
    const int Qn = (NB + SEG_K - 1) / SEG_K;     // segments in use
    const T* const chunk = fac2 + ls.slot(tp, tq) * (int64_t)SL::LINE;
    // the factors are needed after the right-hand sides: start their way towards the L2 now
    if (SEG_PF_W && q == 0) bulk_prefetch_l2(chunk, (unsigned)(SL::LINE * sizeof(T)));

    // ---- right-hand sides, cell-parallel (coalesced): block i = q + QP t ----------------
    {
    double cz[4];                                // carry: zeta of the cell below lane 0's block
    T cu[4];                                     //        and its gra * (outer line edges)
    {
    const LineAddr<T, D> a(E, S, nullptr, tp, tq);
    if (SEG_PF_ROWS && D == 0 && sizeof(T) == 16) {
        // x-lines: every array row the line touches is one contiguous run; 21 bulk prefetches,
        // one per lane, bring them into the L2 while the first loads are in flight
        const void* row = nullptr;
        int cnt = N + 1;
        if (q == 8) { row = a.sdp + a.oL; cnt = N; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {            // (static indices: no local-memory arrays)
            if (q == k) { row = a.ed + a.oLn[k]; cnt = N; }
            if (q == 4 + k) row = a.ts_ptr(k, 0);
            if (q == 9 + k) row = a.ep + a.oP[k >> 1][2 * (k & 1)];
            if (q == 13 + k) row = a.eq + a.oQ[k & 1][2 * (k >> 1)];
        }
        if (row) bulk_prefetch_l2(row, (unsigned)(cnt * sizeof(T)));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (q == k) {
            se[k] = *a.t_ptr(k, 0);
            se[4 + k] = *a.t_ptr(k, N);
        }
    // Every lane loads the data of the cell ABOVE its block (cell i + 1) and of the block's node;
    // the cell below is the neighbouring lane's cell above (shuffle), lane 0 takes it from lane
    // QP - 1 of the previous round (carry), the very first one is line cell 0.
    {
        double z0[2][2], g0[4];
        ln.load_zeta(0, z0);
        ln.side_g(z0, g0);
        CellCoef c0;
        cell_coef<T, D>(ln, g0, ldg(m.rh[A::d]), c0);
        T e0[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            e0[k] = a.ed[a.oLn[k]];
            cu[k] = c0.gra[k] * e0[k];
        }
        cz[0] = z0[0][0]; cz[1] = z0[0][1]; cz[2] = z0[1][0]; cz[3] = z0[1][1];
        if (q == 0) {
            se[8] = line_rhs<T, D>(a, 0, c0, e0);
#pragma unroll
            for (int k = 0; k < 4; ++k) zs[k * SL::ZR] = cz[k];
        }
    }
    }
#pragma unroll 2
    for (int t = 0; t < SEG_K; ++t) {
        // The ~26 row offsets of the line are recomputed every round (a few integer multiply-adds
        // each): hoisted out of the loop they do not fit the register file next to the 22 loads
        // in flight, and were spilled and re-read from local memory (= L2) every round.
        int tpl = tp, tql = tq;
        asm volatile("" : "+r"(tpl), "+r"(tql));
        const LineAddr<T, D> a(E, S, nullptr, tpl, tql);
        const int i = q + QP * t;
        const bool valid = i < NB;
        const int mn = valid ? i + 1 : NB;       // (clamped: loads stay inside the line)
        // ---- loads: cell mn, node mn
        double zn[4];
        {
            double z[2][2];
            ln.load_zeta(mn, z);
            zn[0] = z[0][0]; zn[1] = z[0][1]; zn[2] = z[1][0]; zn[3] = z[1][1];
        }
        const double rdn = ldg(m.rh[A::d] + mn);
        T eo_n[4], st[4], op[4], oq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) eo_n[k] = a.ed[a.oLn[k] + a.sd * mn];
        const T sl = ldg(a.sdp + a.oL + a.sd * mn);
#pragma unroll
        for (int k = 0; k < 4; ++k) st[k] = ldg(a.ts_ptr(k, mn));
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) {
                op[jp * 2 + jq] = a.epo(jp, jq, mn);
                oq[jp * 2 + jq] = a.eqo(jp, jq, mn);
            }
        // ---- the cell above
        T un[4], bln = sl;
        {
            double gn[4];
            gn[0] = 0.5 * (zn[0] + zn[1]);
            gn[1] = 0.5 * (zn[2] + zn[3]);
            gn[2] = 0.5 * (zn[0] + zn[2]);
            gn[3] = 0.5 * (zn[1] + zn[3]);
            CellCoef cn;
            cell_coef<T, D>(ln, gn, rdn, cn);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                un[k] = cn.gra[k] * eo_n[k];
                bln += cn.gaa[k] * eo_n[k];
            }
        }
        if (!valid) {
            bln = zero_<T>();
#pragma unroll
            for (int k = 0; k < 4; ++k) { un[k] = zero_<T>(); zn[k] = 0.0; st[k] = zero_<T>(); }
        }
        // ---- the cell below: neighbouring lane / carry
        double zc[4];
        T uc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            zc[k] = __shfl_up_sync(0xffffffffu, zn[k], 1, QP);
            uc[k] = shfl_up_t(un[k], QP);
            if (q == 0) { zc[k] = cz[k]; uc[k] = cu[k]; }
            cz[k] = __shfl_sync(0xffffffffu, zn[k], sub * QP + QP - 1);
            cu[k] = shfl_t(un[k], sub * QP + QP - 1);
        }
        // ---- right-hand side of the block: sources, side faces of both cells, end faces
        T r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = valid ? st[k] + uc[k] - un[k] : zero_<T>();
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) {
                const double gf = valid ? 0.5 * (zc[jp * 2 + jq] + zn[jp * 2 + jq]) : 0.0;
                const double ap = ln.al_p(jq), aq = ln.al_q(jp);
                const T out = ap * op[jp * 2 + jq] + aq * oq[jp * 2 + jq];
                r[jp] += (gf * ap) * out;
                r[2 + jq] += (gf * aq) * out;
            }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            zs[k * SL::ZR + pad8(i + 1)] = zn[k];
            wA[k * SL::ST + pad8(i)] = r[k];
        }
        wA[4 * SL::ST + pad8(i)] = bln;
    }
    }
    __syncwarp();

    // ---- to the segment layout: lane q owns blocks 8 q .. 8 q + 7 ---------------------
    T g[SEG_K][4], bl[SEG_K];
#pragma unroll
    for (int j = 0; j < SEG_K; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) g[j][k] = wA[k * SL::ST + 9 * q + j];
        bl[j] = wA[4 * SL::ST + 9 * q + j];
    }
    const T blm1 = q == 0 ? se[8] : wA[4 * SL::ST + 9 * (q - 1) + 7];  // bL of line cell 8 q
    __syncwarp();

    // ---- factors: one bulk copy per line into the region the staging just left ---------
    fence_proxy_async_smem();
    {
        const unsigned bytes = (unsigned)(SL::W * sizeof(T));
        if (lane == 0) mbar_expect_tx(bar, bytes * LPW);
        __syncwarp();
        if (q == 0) bulk_g2s(wA, chunk, bytes, bar);
        mbar_wait(bar, parity);
        parity ^= 1u;
    }

    // ---- pass 1: forward recurrence from zero; bl[j] <- bL_m / dL_m ---------------------
    {
        double fc[4], dc[4];
        seg_cell<T, D, QP>(ln, zs, SEG_K * q, fc, dc);
        T sc = blm1 * seg_rl<T, QP>(wA, q, 0);
        T eg[4];                                 // E_m g_{m-1}
#pragma unroll
        for (int k = 0; k < 4; ++k) eg[k] = zero_<T>();
#pragma unroll
        for (int j = 0; j < SEG_K; ++j) {
            double fn[4], dn[4];
            seg_cell<T, D, QP>(ln, zs, SEG_K * q + j + 1, fn, dn);
            const T rln = wA[(j * FAC_NE + 10) * QP + q];
            const T sn = bl[j] * rln;
            T v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = g[j][k] + fc[k] * sc - fn[k] * sn - eg[k];
            {
                T X[10];
#pragma unroll
                for (int e = 0; e < 10; ++e) X[e] = wA[(j * FAC_NE + e) * QP + q];
                seg_symv4<T>(X, v, g[j]);
            }
            bl[j] = sn;
            sc = sn;
            if (SEG_K * q + j < NB) {
#pragma unroll
                for (int k = 0; k < 4; ++k) sb[q * 4 + k] = g[j][k];
            }
            if (j < SEG_K - 1) seg_apply_E<T>(dn, fn, rln, g[j], eg);
#pragma unroll
            for (int k = 0; k < 4; ++k) fc[k] = fn[k];
        }
    }
    __syncwarp();

    // ---- forward scan over the segment ends ---------------------------------------------
    // G_q = g~_end(q) + Phi_q G_{q-1}, sequential over the segments.  The jump matrices come in
    // batches of QP / 4: one coalesced load (lane q holds row q % 4 of matrix batch + q / 4);
    // lanes q < 4 own one row of the running 4x4 matrix-vector product and fetch their row of
    // the current matrix by shuffle, so a step never waits for memory.
    T G[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) G[k] = se[k];                    // g_0 = T_0 (fixed data)
    seg_scan<T, QP, true>(chunk + SL::PHI, sb, G, Qn, q, sub);
    __syncwarp();

    // ---- pass 2: propagate the incoming g through the segment ---------------------------
    {
        T dl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dl[k] = q == 0 ? se[k] : sb[(q - 1) * 4 + k];
#pragma unroll
        for (int j = 0; j < SEG_K; ++j) {
            double fc[4], dc[4];
            seg_cell<T, D, QP>(ln, zs, SEG_K * q + j, fc, dc);
            const T rlc = seg_rl<T, QP>(wA, q, j);
            T X[10];
#pragma unroll
            for (int e = 0; e < 10; ++e) X[e] = wA[(j * FAC_NE + e) * QP + q];
            T eg[4], xw[4];
            seg_apply_E<T>(dc, fc, rlc, dl, eg);
            seg_symv4<T>(X, eg, xw);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                dl[k] = -xw[k];
                g[j][k] += dl[k];
            }
        }
    }
    __syncwarp();

    // ---- pass 3: backward recurrence from zero ------------------------------------------
    {
        T tt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) tt[k] = zero_<T>();
#pragma unroll
        for (int j = SEG_K - 1; j >= 0; --j) {
            if (j < SEG_K - 1) {
                double fn[4], dn[4];
                seg_cell<T, D, QP>(ln, zs, SEG_K * q + j + 1, fn, dn);
                const T rln = wA[(j * FAC_NE + 10) * QP + q];
                T X[10];
#pragma unroll
                for (int e = 0; e < 10; ++e) X[e] = wA[(j * FAC_NE + e) * QP + q];
                T eg[4], xw[4];
                seg_apply_E<T>(dn, fn, rln, tt, eg);
                seg_symv4<T>(X, eg, xw);
#pragma unroll
                for (int k = 0; k < 4; ++k) g[j][k] -= xw[k];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) tt[k] = g[j][k];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sb[q * 4 + k] = g[0][k];
    }
    __syncwarp();

    // ---- backward scan: H_q = T~_first(q) + Psi_q H_{q+1} -------------------------------
#pragma unroll
    for (int k = 0; k < 4; ++k) G[k] = se[4 + k];                // T_N (fixed data)
    seg_scan<T, QP, false>(chunk + SL::PSI, sb, G, Qn, q, sub);
    __syncwarp();

    // ---- pass 4: propagate the incoming T, line edges -----------------------------------
    T l0 = zero_<T>();
    {
        T ep[4], tnx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ep[k] = q == Qn - 1 ? se[4 + k] : (q < Qn - 1 ? sb[(q + 1) * 4 + k] : zero_<T>());
            tnx[k] = ep[k];
        }
#pragma unroll
        for (int j = SEG_K - 1; j >= 0; --j) {
            if (SEG_K * q + j < NB) {
                double fn[4], dn[4];
                seg_cell<T, D, QP>(ln, zs, SEG_K * q + j + 1, fn, dn);
                const T rln = wA[(j * FAC_NE + 10) * QP + q];
                T X[10];
#pragma unroll
                for (int e = 0; e < 10; ++e) X[e] = wA[(j * FAC_NE + e) * QP + q];
                T eg[4], xw[4];
                seg_apply_E<T>(dn, fn, rln, ep, eg);
                seg_symv4<T>(X, eg, xw);
                T fd = zero_<T>();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ep[k] = -xw[k];
                    g[j][k] += ep[k];
                    fd += fn[k] * (g[j][k] - tnx[k]);
                    tnx[k] = g[j][k];
                }
                bl[j] = bl[j] - rln * fd;                        // L_m
            }
        }
        if (q == 0) {                                            // L_0 = (bL_0 - f_0 . (T_0 - T_1)) / dL_0
            double fc[4], dc[4];
            seg_cell<T, D, QP>(ln, zs, 0, fc, dc);
            T fd = zero_<T>();
#pragma unroll
            for (int k = 0; k < 4; ++k) fd += fc[k] * (se[k] - tnx[k]);
            l0 = wA[SL::HDR] * (se[8] - fd);
        }
    }
    __syncwarp();

    // ---- back to the cell-parallel layout, coalesced stores ------------------------------
#pragma unroll
    for (int j = 0; j < SEG_K; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) wA[k * SL::ST + 9 * q + j] = g[j][k];
        wA[4 * SL::ST + 9 * q + j] = bl[j];
    }
    __syncwarp();
    if (store) {
        const LineAddr<T, D> a(E, S, nullptr, tp, tq);
#pragma unroll 1
        for (int t = 0; t < SEG_K; ++t) {
            const int i = q + QP * t;
            if (i < NB) {
#pragma unroll
                for (int k = 0; k < 4; ++k) *a.t_ptr(k, i + 1) = wA[k * SL::ST + pad8(i)];
                a.ed[a.oL + a.sd * (i + 1)] = wA[4 * SL::ST + pad8(i)];
            }
        }
        if (q == 0) a.ed[a.oL] = l0;
    }
    __syncwarp();
}

template <typename T>
__device__ __forceinline__ void seg_carve(unsigned char* base, T*& smA, double*& smZ, T*& smS, T*& smT,
                                          uint64_t*& bar) {
    smA = reinterpret_cast<T*>(base);
    smZ = reinterpret_cast<double*>(base + 2832 * sizeof(T));
    smS = reinterpret_cast<T*>(base + 2832 * sizeof(T) + 4 * 292 * sizeof(double));
    smT = smS + 128;
    bar = reinterpret_cast<uint64_t*>(base + 2832 * sizeof(T) + 4 * 292 * sizeof(double) + (128 + 20) * sizeof(T));
}

// one warp per block; lines t = blockIdx.x * LPW + (lane / QP) of parity class c
template <typename T, int D, int QP>
__global__ void __launch_bounds__(32)
gs_line_seg_color_kernel(Model<T> m, const T* fac2, LineSlots ls, T* e, const T* s, int c) {
    extern __shared__ __align__(128) unsigned char seg_smem[];
    constexpr int LPW = 32 / QP;
    T *smA, *smS, *smT;
    double* smZ;
    uint64_t* bar;
    seg_carve<T>(seg_smem, smA, smZ, smS, smT, bar);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    __syncwarp();
    unsigned parity = 0;
    const int sub = (threadIdx.x & 31) / QP;
    int tp, tq;
    bool store = class_line(ls, c, blockIdx.x * LPW + sub, tp, tq);
    if (!store) class_line(ls, c, blockIdx.x * LPW, tp, tq);     // padding: redo line 0, no stores
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    seg_line_sweep<T, D, QP>(m, fac2, ls, E, S, tp, tq, store, smA, smZ, smS, smT, bar, parity);
}

// ---- cached data -------------------------------------------------------------------------
template <typename T, int D, int QP>
__global__ void __launch_bounds__(64)
line_factor_seg_kernel(Model<T> m, T* fac2, LineSlots ls, int c) {
    int tp, tq;
    if (!class_line(ls, c, blockIdx.x * blockDim.x + threadIdx.x, tp, tq)) return;
    factor_line<T, D, QP>(m, tp, tq, fac2 + ls.slot(tp, tq) * (int64_t)SegL<QP>::LINE);
}

// jump matrices: thread (line, segment q); Phi_q = M_b ... M_a, Psi_q = N_a ... N_b
template <typename T, int D, int QP>
__global__ void __launch_bounds__(128)
line_jump_kernel(Model<T> m, T* fac2, LineSlots ls, int c) {
    using A = Ax<D>;
    using SL = SegL<QP>;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = gt % QP;
    int tp, tq;
    if (!class_line(ls, c, gt / QP, tp, tq)) return;
    const Line<T, D> ln(m, tp, tq);
    const int NB = ln.N - 1;
    if (SEG_K * q >= NB) return;
    T* const chunk = fac2 + ls.slot(tp, tq) * (int64_t)SL::LINE;
    const int nv = min(SEG_K, NB - SEG_K * q);

    auto coef = [&](int cell, double f[4], double d[4]) {
        double z[2][2], gs[4];
        ln.load_zeta(cell, z);
        ln.side_g(z, gs);
        CellCoef cc;
        cell_coef<T, D>(ln, gs, ldg(m.rh[A::d] + cell), cc);
#pragma unroll
        for (int k = 0; k < 4; ++k) { f[k] = cc.f[k]; d[k] = cc.d[k]; }
    };
    auto rl_of = [&](int cell) -> T {                        // 1/dL of a line cell
        if (cell == 0) return chunk[SL::HDR];
        const int b = cell - 1;
        return chunk[((b & 7) * FAC_NE + 10) * QP + (b >> 3)];
    };

    T P[4][4];                                               // columns (Phi), then rows (Psi)
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                P[u][k] = zero_<T>();
                if (u == k) add_real(P[u][k], 1.0);
            }
        for (int j = 0; j < nv; ++j) {
            const int i = SEG_K * q + j;
            T X[10];
#pragma unroll
            for (int e = 0; e < 10; ++e) X[e] = chunk[(j * FAC_NE + e) * QP + q];
            double f[4], d[4];
            const int cell = pass == 0 ? i : i + 1;          // E_m: cell m-1 = i;  E_{m+1}: cell m = i+1
            coef(cell, f, d);
            const T rl = rl_of(cell);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                T t1[4], t2[4];
                if (pass == 0) {                             // column u <- -X (E column)
                    apply_E<T>(d, f, rl, P[u], t1);
                    symv4<T>(X, t1, t2);
                } else {                                     // row u <- -E (X row^T)
                    symv4<T>(X, P[u], t1);
                    apply_E<T>(d, f, rl, t1, t2);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) P[u][k] = -t2[k];
            }
        }
        T* out = chunk + (pass == 0 ? SL::PHI : SL::PSI) + q * 16;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) out[r * 4 + cc] = pass == 0 ? P[cc][r] : P[r][cc];
    }
}

// ---- host side ---------------------------------------------------------------------------
// lanes per line of the segment kernels for lines of `n` cells along `dir`; 0 = not used
static int g_seg_mask = -1;
int line_seg_mask(int mask) {
    if (g_seg_mask < 0) {
        const char* v = getenv("EMG3D_B200_LINE_SEG");       // bit a: direction a (default: x)
        g_seg_mask = v ? (atoi(v) & 7) : 1;
    }
    const int prev = g_seg_mask;
    if (mask >= 0) g_seg_mask = mask & 7;
    return prev;
}

int line_seg_qp(const Dims& d, int dir) {
    const int mask = line_seg_mask(-1);
    if (!((mask >> dir) & 1)) return 0;
    const int p = dir == 0 ? 1 : 0, q = dir == 2 ? 1 : 2;
    if (d.n[p] < 2 || d.n[q] < 2) return 0;
    const int nb = d.n[dir] - 1;
    if (nb > 256 || nb <= 64) return 0;
    return nb > 128 ? 32 : 16;
}

int64_t line_seg_elems(const Dims& d, int dir) {
    const int qp = line_seg_qp(d, dir);
    if (!qp) return 0;
    const int p = dir == 0 ? 1 : 0, q = dir == 2 ? 1 : 2;
    LineSlots ls(d.n[p] - 1, d.n[q] - 1);
    return ls.nl * (int64_t)(qp == 32 ? SegL<32>::LINE : SegL<16>::LINE);
}

template <typename T, int D, int QP>
static void seg_factor_dir(const Model<T>& m, T* fac2, cudaStream_t st) {
    using A = Ax<D>;
    LineSlots ls(m.d.n[A::p] - 1, m.d.n[A::q] - 1);
    cudaMemsetAsync(fac2, 0, sizeof(T) * ls.nl * (size_t)SegL<QP>::LINE, st);
    for (int c = 0; c < 4; ++c) {
        if (ls.cnt[c] == 0) continue;
        ++g_launch_count; line_factor_seg_kernel<T, D, QP><<<(ls.cnt[c] + 63) / 64, 64, 0, st>>>(m, fac2, ls, c);
        const int64_t nt = (int64_t)ls.cnt[c] * QP;
        ++g_launch_count; line_jump_kernel<T, D, QP><<<(unsigned)((nt + 127) / 128), 128, 0, st>>>(m, fac2, ls, c);
    }
}

template <typename T>
void launch_line_seg_factor(const Model<T>& m, int dir, T* fac2, cudaStream_t st) {
    const int qp = line_seg_qp(m.d, dir);
#define EMG_SEGF(DD)                                                    \
    if (qp == 32) seg_factor_dir<T, DD, 32>(m, fac2, st);               \
    else seg_factor_dir<T, DD, 16>(m, fac2, st)
    if (dir == 0) { EMG_SEGF(0); }
    else if (dir == 1) { EMG_SEGF(1); }
    else { EMG_SEGF(2); }
#undef EMG_SEGF
}

template <typename T, int D, int QP>
static void seg_color(const Model<T>& m, const T* fac2, T* e, const T* s, int c, cudaStream_t st) {
    using A = Ax<D>;
    LineSlots ls(m.d.n[A::p] - 1, m.d.n[A::q] - 1);
    if (ls.cnt[c] == 0) return;
    constexpr int LPW = 32 / QP;
    constexpr int smem = seg_warp_smem<T>();
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gs_line_seg_color_kernel<T, D, QP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        // four blocks per SM need the whole shared-memory carve-out
        cudaFuncSetAttribute(gs_line_seg_color_kernel<T, D, QP>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    ++g_launch_count;
    gs_line_seg_color_kernel<T, D, QP><<<(ls.cnt[c] + LPW - 1) / LPW, 32, smem, st>>>(m, fac2, ls, e, s, c);
}

// one colour class of one multicolour sweep
template <typename T>
void launch_gs_line_seg_color(const Model<T>& m, int dir, const T* fac2, T* e, const T* s, int c,
                              cudaStream_t st) {
    const int qp = line_seg_qp(m.d, dir);
#define EMG_SEGC(DD)                                                    \
    if (qp == 32) seg_color<T, DD, 32>(m, fac2, e, s, c, st);           \
    else seg_color<T, DD, 16>(m, fac2, e, s, c, st)
    if (dir == 0) { EMG_SEGC(0); }
    else if (dir == 1) { EMG_SEGC(1); }
    else { EMG_SEGC(2); }
#undef EMG_SEGC
}

template void launch_line_seg_factor<double>(const Model<double>&, int, double*, cudaStream_t);
template void launch_line_seg_factor<cplx>(const Model<cplx>&, int, cplx*, cudaStream_t);
template void launch_gs_line_seg_color<double>(const Model<double>&, int, const double*, double*,
                                               const double*, int, cudaStream_t);
template void launch_gs_line_seg_color<cplx>(const Model<cplx>&, int, const cplx*, cplx*, const cplx*, int,
                                             cudaStream_t);

}  // namespace emg
