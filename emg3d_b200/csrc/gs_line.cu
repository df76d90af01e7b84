// Line-relaxation Gauss-Seidel smoothers along x, y or z (what
// emg3d/core.py:506-783 `gauss_seidel_x`, 786-1068 `_y`, 1071-1348 `_z` compute):
// all edges touching the interior nodes of one grid line, plus the line's own
// edges, are solved for simultaneously.
//
// Unknown order per line, as in the reference (core.py:775-783, 1060-1068,
// 1340-1348): block i = [ L_i, T_{i+1} ],  L_i = line edge between nodes i and
// i+1, T_m = the four transverse edges at interior node m in slot order
// [p-, p+, q-, q+], (p, q) = (y,z) / (x,z) / (x,y) for x- / y- / z-lines; the
// last block holds L_{N-1} only.  The matrix is block tridiagonal,
//     S_0 = M_0,  S_i = M_i - F_i (S_{i-1}^{-1}) F_i^T,
// with a *real* sparse coupling F_i: row 0 = (0, f_0..f_3), rows 1..4 =
// diag(d_0..d_3).  Only the trailing 4x4 block of S_{i-1}^{-1} enters.
//
// B200 design: the matrix depends on (grid, model, s) only, not
// on E, so the block factors L_i D_i L_i^T of S_i (15 numbers per block) are
// computed ONCE per level and direction (`line_factor`) and streamed from HBM
// by every sweep; the reference refactors every line in every sweep
// (core.py:769-772).  A sweep is then one forward and one backward block
// substitution per line, one thread per line, with the intermediate vector
// stored in place in E.  Factor layout [block][entry][line] makes the factor
// stream perfectly coalesced across the threads of a warp.
//
// Orderings: `lex` runs hyperplanes t = tp + 2 tq of the reference's line
// order (p fastest, then q; core.py:601-624, 886-917, 1166-1197) and is
// sequentially equivalent to it; `color` runs the four parity classes
// (tp&1, tq&1), which are conflict-free.
#include "common.cuh"
#include "kernels.h"

namespace emg {

template <int D> struct Ax {
    static constexpr int d = D;
    static constexpr int p = (D == 0) ? 1 : 0;
    static constexpr int q = (D == 2) ? 1 : 2;
};

// colour-major slot of line (tp, tq) in the factor array
struct LineSlots {
    int na[2], off[4];
    int64_t nl;
    __host__ __device__ LineSlots() {}
    __host__ __device__ LineSlots(int npi, int nqi) {
        na[0] = (npi + 1) / 2; na[1] = npi / 2;
        const int nb0 = (nqi + 1) / 2, nb1 = nqi / 2;
        off[0] = 0;
        off[1] = na[0] * nb0;
        off[2] = off[1] + na[1] * nb0;
        off[3] = off[2] + na[0] * nb1;
        nl = (int64_t)npi * nqi;
    }
    __host__ __device__ __forceinline__ int64_t slot(int tp, int tq) const {
        const int cp = (tp - 1) & 1, cq = (tq - 1) & 1;
        return off[cp + 2 * cq] + ((tp - 1) >> 1) + (int64_t)na[cp] * ((tq - 1) >> 1);
    }
};

// ---- per-line geometry shared by factor and solve ----------------------------
template <typename T, int D>
struct Line {
    using A = Ax<D>;
    const Model<T>& m;
    int N, tp, tq;
    double rp[2], rq[2];
    int64_t cstr;          // cell stride along the line
    int64_t cbase[2][2];   // cell offset of column (jp, jq) at line cell 0

    __device__ Line(const Model<T>& m_, int tp_, int tq_) : m(m_), tp(tp_), tq(tq_) {
        N = m.d.n[A::d];
        rp[0] = ldg(m.rh[A::p] + tp - 1); rp[1] = ldg(m.rh[A::p] + tp);
        rq[0] = ldg(m.rh[A::q] + tq - 1); rq[1] = ldg(m.rh[A::q] + tq);
        const int64_t cs[3] = {1, m.d.n[0], (int64_t)m.d.n[0] * m.d.n[1]};
        cstr = cs[A::d];
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq)
                cbase[jp][jq] = cs[A::p] * (tp - 1 + jp) + cs[A::q] * (tq - 1 + jq);
    }
    // stencil entry of the line edge in its side face k (k = p-, p+, q-, q+)
    __device__ __forceinline__ double a_side(int k) const {
        return k == 0 ? -rp[0] : k == 1 ? rp[1] : k == 2 ? -rq[0] : rq[1];
    }
    __device__ __forceinline__ void load_zeta(int i, double z[2][2]) const {
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) z[jp][jq] = ldg(m.zeta + cbase[jp][jq] + cstr * i);
    }
    // 1/2 (zeta + zeta) of the four side faces of line cell i
    __device__ __forceinline__ void side_g(const double z[2][2], double g[4]) const {
        g[0] = 0.5 * (z[0][0] + z[0][1]);
        g[1] = 0.5 * (z[1][0] + z[1][1]);
        g[2] = 0.5 * (z[0][0] + z[1][0]);
        g[3] = 0.5 * (z[0][1] + z[1][1]);
    }
    // stencil entries of the local p-/q-edge in the end face of quadrant (jp, jq)
    __device__ __forceinline__ double al_p(int jq) const { return jq == 0 ? -rq[0] : rq[1]; }
    __device__ __forceinline__ double al_q(int jp) const { return jp == 0 ? rp[0] : -rp[1]; }
};

// 5x5 symmetric block stored as s[r][c], r >= c.
// In-place LDL^T: s[r][c] (r>c) <- L(r,c), dinv[r] <- 1/D(r).
template <typename T>
__device__ __forceinline__ void ldlt5(T s[5][5], T dinv[5]) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        T v[5];
        T dj = s[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) {
            v[k] = s[j][k] * s[k][k];
            dj -= s[j][k] * v[k];
        }
        s[j][j] = dj;
        const T r = rcp(dj);
        dinv[j] = r;
#pragma unroll
        for (int i = j + 1; i < 5; ++i) {
            T t = s[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) t -= s[i][k] * v[k];
            s[i][j] = t * r;
        }
    }
}

// ---- factorisation: one thread per line ------------------------------------
template <typename T, int D>
__device__ void factor_line(const Model<T>& m, int tp, int tq, T* __restrict__ fac,
                            const LineSlots& ls) {
    using A = Ax<D>;
    Line<T, D> ln(m, tp, tq);
    const int N = ln.N;
    const int64_t slot = ls.slot(tp, tq), nl = ls.nl;

    double zc[2][2], zn[2][2];
    ln.load_zeta(0, zc);
    T X[4][4];                       // trailing 4x4 of S_{i-1}^{-1} (lower part used)
    // eta sums carried from the previous line cell for the transverse diagonals
    T etp_c[2], etq_c[2];            // sum over jq (resp. jp) of eta_p / eta_q at line cell i
    {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            etp_c[j] = ldg(m.eta[A::p] + ln.cbase[j][0]) + ldg(m.eta[A::p] + ln.cbase[j][1]);
            etq_c[j] = ldg(m.eta[A::q] + ln.cbase[0][j]) + ldg(m.eta[A::q] + ln.cbase[1][j]);
        }
    }
    for (int i = 0; i < N; ++i) {
        const double rd = ldg(m.rh[A::d] + i);
        double gs[4];
        ln.side_g(zc, gs);
        T S[5][5];
        // line edge diagonal
        {
            T st = zero_<T>();
#pragma unroll
            for (int jp = 0; jp < 2; ++jp)
#pragma unroll
                for (int jq = 0; jq < 2; ++jq)
                    st += ldg(m.eta[A::d] + ln.cbase[jp][jq] + ln.cstr * i);
            S[0][0] = -0.25 * st;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc += gs[k] * ln.a_side(k) * ln.a_side(k);
            add_real(S[0][0], acc);
        }
        double f[4], dk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double c = gs[k] * ln.a_side(k) * rd;
            f[k] = -c;                   // L_i <-> T_{i,k}
            dk[k] = -gs[k] * rd * rd;    // T_{i+1,k} <-> T_{i,k}
            S[1 + k][0] = zero_<T>();
            add_real(S[1 + k][0], c);    // L_i <-> T_{i+1,k}
        }
        const bool last = (i == N - 1);
        if (!last) {
            ln.load_zeta(i + 1, zn);
            const double rdn = ldg(m.rh[A::d] + i + 1);
            double gn[4];
            ln.side_g(zn, gn);
            T etp_n[2], etq_n[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int64_t o = ln.cstr * (i + 1);
                etp_n[j] = ldg(m.eta[A::p] + ln.cbase[j][0] + o) + ldg(m.eta[A::p] + ln.cbase[j][1] + o);
                etq_n[j] = ldg(m.eta[A::q] + ln.cbase[0][j] + o) + ldg(m.eta[A::q] + ln.cbase[1][j] + o);
            }
            // transverse diagonals: eta, side faces of L_i and L_{i+1}
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const T et = k < 2 ? etp_c[k] + etp_n[k] : etq_c[k - 2] + etq_n[k - 2];
                S[1 + k][1 + k] = -0.25 * et;
                add_real(S[1 + k][1 + k], gs[k] * rd * rd + gn[k] * rdn * rdn);
            }
            S[2][1] = zero_<T>();
            S[4][3] = zero_<T>();
            S[3][1] = S[3][2] = S[4][1] = S[4][2] = zero_<T>();
            // end faces at node i+1
#pragma unroll
            for (int jp = 0; jp < 2; ++jp)
#pragma unroll
                for (int jq = 0; jq < 2; ++jq) {
                    const double g = 0.5 * (zc[jp][jq] + zn[jp][jq]);
                    const double ap = ln.al_p(jq), aq = ln.al_q(jp);
                    add_real(S[1 + jp][1 + jp], g * ap * ap);
                    add_real(S[3 + jq][3 + jq], g * aq * aq);
                    add_real(S[3 + jq][1 + jp], g * ap * aq);
                }
#pragma unroll
            for (int j = 0; j < 2; ++j) { etp_c[j] = etp_n[j]; etq_c[j] = etq_n[j]; }
        }
        // Schur update with the previous block
        if (i > 0) {
            T xf[4];                     // X f
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                T t = zero_<T>();
#pragma unroll
                for (int c = 0; c < 4; ++c) t += f[c] * (r >= c ? X[r][c] : X[c][r]);
                xf[r] = t;
            }
            T z00 = zero_<T>();
#pragma unroll
            for (int r = 0; r < 4; ++r) z00 += f[r] * xf[r];
            S[0][0] -= z00;
            if (!last) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    S[1 + r][0] -= dk[r] * xf[r];
#pragma unroll
                    for (int c = 0; c <= r; ++c) S[1 + r][1 + c] -= (dk[r] * dk[c]) * X[r][c];
                }
            }
        }
        T* out = fac + ((int64_t)i * 15) * nl + slot;
        if (last) {
            out[0] = rcp(S[0][0]);
            break;
        }
        T dinv[5];
        ldlt5<T>(S, dinv);
        {
            int e = 0;
#pragma unroll
            for (int r = 1; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < r; ++c) out[(int64_t)(e++) * nl] = S[r][c];
#pragma unroll
            for (int r = 0; r < 5; ++r) out[(int64_t)(10 + r) * nl] = dinv[r];
        }
        // X = (L4 D4 L4^T)^-1 with L4 = L[1:,1:], D4 = D[1:]
        {
            T Li[4][4];                  // inverse of unit lower L4 (strict lower part)
            Li[1][0] = -S[2][1];
            Li[2][1] = -S[3][2];
            Li[3][2] = -S[4][3];
            Li[2][0] = -S[3][1] - S[3][2] * Li[1][0];
            Li[3][1] = -S[4][2] - S[4][3] * Li[2][1];
            Li[3][0] = -S[4][1] - S[4][2] * Li[1][0] - S[4][3] * Li[2][0];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) {
                    // sum_{k >= r} Li[k][r] dinv[k+1] Li[k][c], Li[k][k] = 1
                    T t = (r == c) ? dinv[1 + r] : dinv[1 + r] * Li[r][c];
#pragma unroll
                    for (int k = r + 1; k < 4; ++k) t += Li[k][r] * dinv[1 + k] * Li[k][c];
                    X[r][c] = t;
                }
        }
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) zc[jp][jq] = zn[jp][jq];
    }
}

// ---- one line sweep: forward and backward block substitution ----------------
template <typename T>
__device__ __forceinline__ void solve5(const T* __restrict__ fp, int64_t nl, T y[5], bool first_zero) {
    // fp points at entry 0 of this block for this line; entries strided by nl
    T L[10], dinv[5];
#pragma unroll
    for (int e = 0; e < 10; ++e) L[e] = ldg(fp + (int64_t)e * nl);
#pragma unroll
    for (int r = 0; r < 5; ++r) dinv[r] = ldg(fp + (int64_t)(10 + r) * nl);
    // L index: (1,0)=0 (2,0)=1 (2,1)=2 (3,0)=3 (3,1)=4 (3,2)=5 (4,0)=6 (4,1)=7 (4,2)=8 (4,3)=9
    if (!first_zero) {
        y[1] -= L[0] * y[0];
        y[2] -= L[1] * y[0];
        y[3] -= L[3] * y[0];
        y[4] -= L[6] * y[0];
    }
    y[2] -= L[2] * y[1];
    y[3] -= L[4] * y[1];
    y[4] -= L[7] * y[1];
    y[3] -= L[5] * y[2];
    y[4] -= L[8] * y[2];
    y[4] -= L[9] * y[3];
#pragma unroll
    for (int r = 0; r < 5; ++r) y[r] = y[r] * dinv[r];
    y[3] -= L[9] * y[4];
    y[2] -= L[8] * y[4] + L[5] * y[3];
    y[1] -= L[7] * y[4] + L[4] * y[3] + L[2] * y[2];
    y[0] -= L[6] * y[4] + L[3] * y[3] + L[1] * y[2] + L[0] * y[1];
}

template <typename T, int D>
__device__ void sweep_line(const Model<T>& m, int tp, int tq, const T* __restrict__ fac,
                           const LineSlots& ls, const FieldView<T>& E, const FieldView<const T>& S) {
    using A = Ax<D>;
    Line<T, D> ln(m, tp, tq);
    const int N = ln.N;
    const int64_t slot = ls.slot(tp, tq), nl = ls.nl;

    // element strides along the line in the three component arrays
    const int64_t sd = D == 0 ? 1 : D == 1 ? E.s1[A::d] : E.s2[A::d];
    const int64_t sp = D == 0 ? 1 : D == 1 ? E.s1[A::p] : E.s2[A::p];
    const int64_t sq = D == 0 ? 1 : D == 1 ? E.s1[A::q] : E.s2[A::q];
    int pos[3];
    // line edge L_0 and its four parallel neighbours
    pos[A::d] = 0; pos[A::p] = tp; pos[A::q] = tq;
    const int64_t oL = E.idx(A::d, pos);
    int64_t oLn[4];
    pos[A::p] = tp - 1; oLn[0] = E.idx(A::d, pos);
    pos[A::p] = tp + 1; oLn[1] = E.idx(A::d, pos);
    pos[A::p] = tp; pos[A::q] = tq - 1; oLn[2] = E.idx(A::d, pos);
    pos[A::q] = tq + 1; oLn[3] = E.idx(A::d, pos);
    // transverse edges at line node 0: p-edges (jp) with q-node tq-1, tq, tq+1
    int64_t oP[2][3], oQ[2][3];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            pos[A::d] = 0; pos[A::p] = tp - 1 + j; pos[A::q] = tq - 1 + o;
            oP[j][o] = E.idx(A::p, pos);     // p-edge in p-cell j at q-node tq-1+o
            pos[A::p] = tp - 1 + o; pos[A::q] = tq - 1 + j;
            oQ[j][o] = E.idx(A::q, pos);     // q-edge in q-cell j at p-node tp-1+o
        }
    T* ed = E.p[A::d];
    T* ep = E.p[A::p];
    T* eq = E.p[A::q];
    const T* sdp = S.p[A::d];
    const T* spp = S.p[A::p];
    const T* sqp = S.p[A::q];

    // ---------------- forward ----------------
    double zc[2][2], zn[2][2], gs[4], gn[4];
    ln.load_zeta(0, zc);
    ln.side_g(zc, gs);
    double rd = ldg(m.rh[A::d]);
    T eo[4];                             // parallel neighbours of L_i
#pragma unroll
    for (int k = 0; k < 4; ++k) eo[k] = ed[oLn[k]];
    T wT[4];                             // transverse part of previous block's solution
#pragma unroll
    for (int k = 0; k < 4; ++k) wT[k] = zero_<T>();

    for (int i = 0; i < N; ++i) {
        const bool last = (i == N - 1);
        T y[5];
        // line edge
        {
            T acc = ldg(sdp + oL + sd * i);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double a = ln.a_side(k);
                acc += (gs[k] * a * a) * eo[k];
            }
            y[0] = acc;
        }
        double f[4], dk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[k] = -gs[k] * ln.a_side(k) * rd;
            dk[k] = -gs[k] * rd * rd;
        }
        if (i > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) y[0] -= f[k] * wT[k];
        }
        const T* fp = fac + ((int64_t)i * 15) * nl + slot;
        if (last) {
            ed[oL + sd * i] = y[0] * ldg(fp);
            break;
        }
        // transverse edges at node m = i+1
        const int mnode = i + 1;
        ln.load_zeta(i + 1, zn);
        ln.side_g(zn, gn);
        const double rdn = ldg(m.rh[A::d] + i + 1);
        T en[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) en[k] = ed[oLn[k] + sd * (i + 1)];
        y[1] = ldg(spp + oP[0][1] + sp * mnode);
        y[2] = ldg(spp + oP[1][1] + sp * mnode);
        y[3] = ldg(sqp + oQ[0][1] + sq * mnode);
        y[4] = ldg(sqp + oQ[1][1] + sq * mnode);
        // side faces of L_i (+) and L_{i+1} (-)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double a = ln.a_side(k);
            y[1 + k] += (gs[k] * rd * a) * eo[k] - (gn[k] * rdn * a) * en[k];
        }
        // end faces at node m
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) {
                const double g = 0.5 * (zc[jp][jq] + zn[jp][jq]);
                const double ap = ln.al_p(jq), aq = ln.al_q(jp);
                const T epo = ep[oP[jp][jq == 0 ? 0 : 2] + sp * mnode];
                const T eqo = eq[oQ[jq][jp == 0 ? 0 : 2] + sq * mnode];
                const T out = ap * epo + aq * eqo;
                y[1 + jp] += (g * ap) * out;
                y[3 + jq] += (g * aq) * out;
            }
        if (i > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) y[1 + k] -= dk[k] * wT[k];
        }
        solve5<T>(fp, nl, y, false);
        ed[oL + sd * i] = y[0];
        ep[oP[0][1] + sp * mnode] = y[1];
        ep[oP[1][1] + sp * mnode] = y[2];
        eq[oQ[0][1] + sq * mnode] = y[3];
        eq[oQ[1][1] + sq * mnode] = y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { wT[k] = y[1 + k]; eo[k] = en[k]; gs[k] = gn[k]; }
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) zc[jp][jq] = zn[jp][jq];
        rd = rdn;
    }

    // ---------------- backward ----------------
    // x_i = w_i - S_i^{-1} F_{i+1}^T x_{i+1};  gs / rd now belong to line cell N-1
    T xL = ed[oL + sd * (N - 1)];
    T xT[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xT[k] = zero_<T>();
    for (int i = N - 2; i >= 0; --i) {
        T v[5];
        v[0] = zero_<T>();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double fk = -gs[k] * ln.a_side(k) * rd;
            const double dkk = -gs[k] * rd * rd;
            v[1 + k] = fk * xL + dkk * xT[k];
        }
        const T* fp = fac + ((int64_t)i * 15) * nl + slot;
        solve5<T>(fp, nl, v, true);
        const int mnode = i + 1;
        xL = ed[oL + sd * i] - v[0];
        xT[0] = ep[oP[0][1] + sp * mnode] - v[1];
        xT[1] = ep[oP[1][1] + sp * mnode] - v[2];
        xT[2] = eq[oQ[0][1] + sq * mnode] - v[3];
        xT[3] = eq[oQ[1][1] + sq * mnode] - v[4];
        ed[oL + sd * i] = xL;
        ep[oP[0][1] + sp * mnode] = xT[0];
        ep[oP[1][1] + sp * mnode] = xT[1];
        eq[oQ[0][1] + sq * mnode] = xT[2];
        eq[oQ[1][1] + sq * mnode] = xT[3];
        if (i > 0) {
            ln.load_zeta(i, zc);
            ln.side_g(zc, gs);
            rd = ldg(m.rh[A::d] + i);
        }
    }
}

// ---- kernels ---------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(128)
line_factor_kernel(Model<T> m, T* fac, LineSlots ls, int npi, int nqi) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y * blockDim.y + threadIdx.y;
    if (a >= npi || b >= nqi) return;
    factor_line<T, D>(m, 1 + a, 1 + b, fac, ls);
}

template <typename T, int D>
__global__ void __launch_bounds__(128)
gs_line_color_kernel(Model<T> m, const T* fac, LineSlots ls, T* e, const T* s, int fp, int fq,
                     int cp, int cq) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y * blockDim.y + threadIdx.y;
    if (a >= cp || b >= cq) return;
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    sweep_line<T, D>(m, fp + 2 * a, fq + 2 * b, fac, ls, E, S);
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
    sweep_line<T, D>(m, tp, tq, fac, ls, E, S);
}

template <typename T, int D>
__global__ void __launch_bounds__(256)
gs_line_small_kernel(Model<T> m, const T* fac, LineSlots ls, T* e, const T* s, int nu, int order) {
    using A = Ax<D>;
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    const int npi = m.d.n[A::p] - 1, nqi = m.d.n[A::q] - 1;
    bool back = false;
    for (int sw = 0; sw < nu; ++sw) {
        back = !back;
        if (order == ORDER_LEX) {
            const int tmin = 3, tmax = npi + 2 * nqi;
            for (int tt = tmin; tt <= tmax; ++tt) {
                const int t = back ? tmax + tmin - tt : tt;
                for (int b = threadIdx.x; b < nqi; b += blockDim.x) {
                    const int tq = 1 + b, tp = t - 2 * tq;
                    if (tp >= 1 && tp <= npi) sweep_line<T, D>(m, tp, tq, fac, ls, E, S);
                }
                __syncthreads();
            }
        } else {
            for (int cc = 0; cc < 4; ++cc) {
                const int c = back ? 3 - cc : cc;
                for (int l = threadIdx.x; l < npi * nqi; l += blockDim.x) {
                    const int tp = 1 + l % npi, tq = 1 + l / npi;
                    if (((tp - 1) & 1) == (c & 1) && ((tq - 1) & 1) == (c >> 1))
                        sweep_line<T, D>(m, tp, tq, fac, ls, E, S);
                }
                __syncthreads();
            }
        }
    }
}

int64_t line_factor_elems(const Dims& d, int dir) {
    const int p = dir == 0 ? 1 : 0, q = dir == 2 ? 1 : 2;
    return (int64_t)15 * d.n[dir] * (d.n[p] - 1) * (d.n[q] - 1);
}

template <typename T, int D>
static void factor_dir(const Model<T>& m, T* fac, cudaStream_t st) {
    using A = Ax<D>;
    const int npi = m.d.n[A::p] - 1, nqi = m.d.n[A::q] - 1;
    if (npi < 1 || nqi < 1) return;
    LineSlots ls(npi, nqi);
    dim3 b(32, 4);
    dim3 g((npi + b.x - 1) / b.x, (nqi + b.y - 1) / b.y);
    ++g_launch_count; line_factor_kernel<T, D><<<g, b, 0, st>>>(m, fac, ls, npi, nqi);
}

template <typename T>
void launch_line_factor(const Model<T>& m, int dir, T* fac, cudaStream_t st) {
    if (dir == 0) factor_dir<T, 0>(m, fac, st);
    else if (dir == 1) factor_dir<T, 1>(m, fac, st);
    else factor_dir<T, 2>(m, fac, st);
}

template <typename T, int D>
static void gs_dir(const Model<T>& m, const T* fac, T* e, const T* s, int nu, int order,
                   cudaStream_t st) {
    using A = Ax<D>;
    const int npi = m.d.n[A::p] - 1, nqi = m.d.n[A::q] - 1;
    if (npi < 1 || nqi < 1) return;
    LineSlots ls(npi, nqi);
    if ((int64_t)npi * nqi <= 1024) {
        int threads = 32;
        const int want = order == ORDER_LEX ? nqi : npi * nqi;
        while (threads < 256 && threads < want) threads <<= 1;
        ++g_launch_count; gs_line_small_kernel<T, D><<<1, threads, 0, st>>>(m, fac, ls, e, s, nu, order);
        return;
    }
    bool back = false;
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
                const int c = back ? 3 - cc : cc;
                const int fp = 1 + (c & 1), fq = 1 + (c >> 1);
                const int cp = (npi - (fp - 1) + 1) / 2, cq = (nqi - (fq - 1) + 1) / 2;
                if (cp <= 0 || cq <= 0) continue;
                dim3 b(32, 2);
                dim3 g((cp + b.x - 1) / b.x, (cq + b.y - 1) / b.y);
                ++g_launch_count; gs_line_color_kernel<T, D><<<g, b, 0, st>>>(m, fac, ls, e, s, fp, fq, cp, cq);
            }
        }
    }
}

template <typename T>
void launch_gs_line(const Model<T>& m, int dir, const T* fac, T* e, const T* s, int nu, int order,
                    cudaStream_t st) {
    if (dir == 0) gs_dir<T, 0>(m, fac, e, s, nu, order, st);
    else if (dir == 1) gs_dir<T, 1>(m, fac, e, s, nu, order, st);
    else gs_dir<T, 2>(m, fac, e, s, nu, order, st);
}

template void launch_line_factor<double>(const Model<double>&, int, double*, cudaStream_t);
template void launch_line_factor<cplx>(const Model<cplx>&, int, cplx*, cudaStream_t);
template void launch_gs_line<double>(const Model<double>&, int, const double*, double*, const double*,
                                     int, int, cudaStream_t);
template void launch_gs_line<cplx>(const Model<cplx>&, int, const cplx*, cplx*, const cplx*, int, int,
                                   cudaStream_t);

}  // namespace emg
