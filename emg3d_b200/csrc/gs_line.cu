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

namespace emg {

template <int D> struct Ax {
    static constexpr int d = D;
    static constexpr int p = (D == 0) ? 1 : 0;
    static constexpr int q = (D == 2) ? 1 : 2;
};

// Factor layout.  Lines are numbered colour-major ("slot"): the four parity
// classes one after the other, each padded to a multiple of 32 lines, and
// within a class p fastest.  A warp of the colour kernel therefore owns one
// aligned group of 32 consecutive slots, and the factors are stored
//     [group][block i][entry e][lane]          (32 lanes, 11 entries, N blocks)
// block i < N-1: entries 0..9 = X_{i+1} = S_{i+1}^{-1} (symmetric, lower triangle
// row-wise), entry 10 = 1/dL_{i+1}; block N-1: entry 0 = 1/dL_0.
constexpr int FAC_NE = 11;          // entries per block
constexpr int FAC_ES = 32;          // stride between entries of one block
constexpr int FAC_BS = FAC_NE * 32; // stride between blocks of one line

struct LineSlots {
    int na[2], nb[2], off[4], cnt[4];
    int64_t nl;                     // padded number of slots
    __host__ __device__ LineSlots() {}
    __host__ __device__ LineSlots(int npi, int nqi) {
        na[0] = (npi + 1) / 2; na[1] = npi / 2;
        nb[0] = (nqi + 1) / 2; nb[1] = nqi / 2;
        int o = 0;
        for (int c = 0; c < 4; ++c) {
            off[c] = o;
            cnt[c] = na[c & 1] * nb[c >> 1];
            o += (cnt[c] + 31) / 32 * 32;
        }
        nl = o;
    }
    __host__ __device__ __forceinline__ int64_t slot(int tp, int tq) const {
        const int cp = (tp - 1) & 1, cq = (tq - 1) & 1;
        return off[cp + 2 * cq] + ((tp - 1) >> 1) + (int64_t)na[cp] * ((tq - 1) >> 1);
    }
    // offset of (block 0, entry 0) of a slot for lines of N blocks
    __host__ __device__ __forceinline__ int64_t base(int64_t slot, int N) const {
        return (slot >> 5) * ((int64_t)N * FAC_BS) + (slot & 31);
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

// ---- small dense helpers (everything unrolled, registers only) -----------------
// symmetric 4x4 in packed lower-triangular storage: (r, c), r >= c, at r (r+1)/2 + c
__device__ __forceinline__ constexpr int tri(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }

// x <- x^{-1} for a complex-symmetric (not Hermitian) 4x4 in packed lower-triangular
// storage, through its LDL^T without pivoting.
template <typename T>
__device__ __forceinline__ void inv4sym(T x[10]) {
    T l[4][4], dinv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        T v[4];
        T dj = x[tri(j, j)];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < j) {
                v[k] = l[j][k] * x[tri(k, k)];      // L(j,k) D(k); x(k,k) holds D(k)
                dj -= l[j][k] * v[k];
            }
        x[tri(j, j)] = dj;
        dinv[j] = rcp(dj);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i > j) {
                T t = x[tri(i, j)];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < j) t -= l[i][k] * v[k];
                l[i][j] = t * dinv[j];
            }
    }
    // inverse of the unit lower factor (strict lower part)
    T li[4][4];
    li[1][0] = -l[1][0];
    li[2][1] = -l[2][1];
    li[3][2] = -l[3][2];
    li[2][0] = -l[2][0] - l[2][1] * li[1][0];
    li[3][1] = -l[3][1] - l[3][2] * li[2][1];
    li[3][0] = -l[3][0] - l[3][1] * li[1][0] - l[3][2] * li[2][0];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            // sum_{k >= r} li[k][r] dinv[k] li[k][c], li[k][k] = 1
            T t = (r == c) ? dinv[r] : dinv[r] * li[r][c];
#pragma unroll
            for (int k = r + 1; k < 4; ++k) t += li[k][r] * dinv[k] * li[k][c];
            x[tri(r, c)] = t;
        }
}

template <typename T>
__device__ __forceinline__ void symv4(const T x[10], const T v[4], T out[4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        T t = x[tri(r, 0)] * v[0];
#pragma unroll
        for (int c = 1; c < 4; ++c) t += x[tri(r, c)] * v[c];
        out[r] = t;
    }
}

// E v with E = diag(d) + rl f f^T  (d, f real; rl = 1/dL complex)
template <typename T>
__device__ __forceinline__ void apply_E(const double d[4], const double f[4], T rl, const T v[4], T out[4]) {
    T fv = f[0] * v[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) fv += f[k] * v[k];
    const T a = rl * fv;
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = d[k] * v[k] + f[k] * a;
}

// geometry of one line cell: couplings of its line edge
struct CellCoef {
    double f[4];      // L <-> T at the cell's lower node (upper node: -f)
    double d[4];      // T(lower node) <-> T(upper node) through the cell's side faces
    double gaa[4];    // gs * a_side^2: L <-> outer parallel line edges (and part of dL)
    double gra[4];    // gs * rd * a_side: T <-> outer parallel line edges
    double grr[4];    // gs * rd^2: contribution to the transverse diagonals
};
template <typename T, int D>
__device__ __forceinline__ void cell_coef(const Line<T, D>& ln, const double gs[4], double rd, CellCoef& c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double as = ln.a_side(k);
        c.gaa[k] = gs[k] * as * as;
        c.gra[k] = gs[k] * rd * as;
        c.grr[k] = gs[k] * rd * rd;
        c.f[k] = -c.gra[k];
        c.d[k] = -c.grr[k];
    }
}

// diagonal entry of the line edge of line cell i
template <typename T, int D>
__device__ __forceinline__ T line_diag(const Line<T, D>& ln, int i, const CellCoef& c) {
    using A = Ax<D>;
    T st = zero_<T>();
#pragma unroll
    for (int jp = 0; jp < 2; ++jp)
#pragma unroll
        for (int jq = 0; jq < 2; ++jq) st += ldg(ln.m.eta[A::d] + ln.cbase[jp][jq] + ln.cstr * i);
    T dl = -0.25 * st;
    add_real(dl, c.gaa[0] + c.gaa[1] + c.gaa[2] + c.gaa[3]);
    return dl;
}

// ---- factorisation: one thread per line ------------------------------------
template <typename T, int D>
__device__ void factor_line(const Model<T>& m, int tp, int tq, T* __restrict__ fac,
                            const LineSlots& ls) {
    using A = Ax<D>;
    Line<T, D> ln(m, tp, tq);
    const int N = ln.N;
    T* const fbase = fac + ls.base(ls.slot(tp, tq), N);

    double zc[2][2], zn[2][2], gs[4];
    ln.load_zeta(0, zc);
    ln.side_g(zc, gs);
    CellCoef cc, cn;
    cell_coef<T, D>(ln, gs, ldg(m.rh[A::d]), cc);
    T rl_c = rcp(line_diag<T, D>(ln, 0, cc));
    fbase[(int64_t)(N - 1) * FAC_BS] = rl_c;                 // 1 / dL_0
    // eta sums carried from the previous line cell for the transverse diagonals
    T etp_c[2], etq_c[2];            // sum over jq (resp. jp) of eta_p / eta_q at line cell i
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        etp_c[j] = ldg(m.eta[A::p] + ln.cbase[j][0]) + ldg(m.eta[A::p] + ln.cbase[j][1]);
        etq_c[j] = ldg(m.eta[A::q] + ln.cbase[0][j]) + ldg(m.eta[A::q] + ln.cbase[1][j]);
    }
    T X[10];                         // X_{m-1}
    for (int i = 0; i < N - 1; ++i) {                        // node m = i + 1
        ln.load_zeta(i + 1, zn);
        double gn[4];
        ln.side_g(zn, gn);
        cell_coef<T, D>(ln, gn, ldg(m.rh[A::d] + i + 1), cn);
        const T rl_n = rcp(line_diag<T, D>(ln, i + 1, cn));
        T etp_n[2], etq_n[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int64_t o = ln.cstr * (i + 1);
            etp_n[j] = ldg(m.eta[A::p] + ln.cbase[j][0] + o) + ldg(m.eta[A::p] + ln.cbase[j][1] + o);
            etq_n[j] = ldg(m.eta[A::q] + ln.cbase[0][j] + o) + ldg(m.eta[A::q] + ln.cbase[1][j] + o);
        }
        // C_m: transverse block at node m (eta, side faces of both cells, end faces)
        T S[10];
#pragma unroll
        for (int e = 0; e < 10; ++e) S[e] = zero_<T>();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const T et = k < 2 ? etp_c[k] + etp_n[k] : etq_c[k - 2] + etq_n[k - 2];
            S[tri(k, k)] = -0.25 * et;
            add_real(S[tri(k, k)], cc.grr[k] + cn.grr[k]);
        }
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) {
                const double g = 0.5 * (zc[jp][jq] + zn[jp][jq]);
                const double ap = ln.al_p(jq), aq = ln.al_q(jp);
                add_real(S[tri(jp, jp)], g * ap * ap);
                add_real(S[tri(2 + jq, 2 + jq)], g * aq * aq);
                add_real(S[tri(2 + jq, jp)], g * ap * aq);
            }
        // D_m = C_m - f_c f_c^T / dL_c - f_n f_n^T / dL_n
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c)
                S[tri(r, c)] -= (cc.f[r] * cc.f[c]) * rl_c + (cn.f[r] * cn.f[c]) * rl_n;
        // S_m = D_m - E_m X_{m-1} E_m,  E_m = diag(d_c) + rl_c f_c f_c^T
        if (i > 0) {
            T fr[4], u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) fr[k] = cc.f[k] * rl_c;          // rl_c f_c  (complex)
            T fc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { fc[k] = zero_<T>(); add_real(fc[k], cc.f[k]); }
            symv4<T>(X, fc, u);                                          // u = X f_c
            T beta = cc.f[0] * u[0];
#pragma unroll
            for (int k = 1; k < 4; ++k) beta += cc.f[k] * u[k];          // f_c^T X f_c
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c)
                    S[tri(r, c)] -= (cc.d[r] * cc.d[c]) * X[tri(r, c)] + (cc.d[r] * u[r]) * fr[c] +
                                    fr[r] * (u[c] * cc.d[c]) + (fr[r] * fr[c]) * beta;
        }
        inv4sym<T>(S);
        T* out = fbase + (int64_t)i * FAC_BS;
#pragma unroll
        for (int e = 0; e < 10; ++e) {
            out[e * FAC_ES] = S[e];               // the inverse (see file header)
            X[e] = S[e];
        }
        out[10 * FAC_ES] = rl_n;
        // shift: cell i+1 becomes the current cell
        cc = cn;
        rl_c = rl_n;
#pragma unroll
        for (int j = 0; j < 2; ++j) { etp_c[j] = etp_n[j]; etq_c[j] = etq_n[j]; }
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) zc[jp][jq] = zn[jp][jq];
    }
}

// ---- addresses of everything a line touches -----------------------------------
template <typename T, int D>
struct LineAddr {
    using A = Ax<D>;
    T *ed, *ep, *eq;
    const T *sdp, *spp, *sqp;
    const T* fac;             // (block 0, entry 0) of this line
    int64_t sd, sp, sq;       // element strides along the line
    int64_t oL, oLn[4], oP[2][3], oQ[2][3];

    __device__ LineAddr(const FieldView<T>& E, const FieldView<const T>& S, const T* fac_,
                        int tp, int tq) {
        ed = E.p[A::d]; ep = E.p[A::p]; eq = E.p[A::q];
        sdp = S.p[A::d]; spp = S.p[A::p]; sqp = S.p[A::q];
        fac = fac_;
        sd = D == 0 ? 1 : D == 1 ? E.s1[A::d] : E.s2[A::d];
        sp = D == 0 ? 1 : D == 1 ? E.s1[A::p] : E.s2[A::p];
        sq = D == 0 ? 1 : D == 1 ? E.s1[A::q] : E.s2[A::q];
        int pos[3];
        pos[A::d] = 0; pos[A::p] = tp; pos[A::q] = tq;
        oL = E.idx(A::d, pos);
        pos[A::p] = tp - 1; oLn[0] = E.idx(A::d, pos);
        pos[A::p] = tp + 1; oLn[1] = E.idx(A::d, pos);
        pos[A::p] = tp; pos[A::q] = tq - 1; oLn[2] = E.idx(A::d, pos);
        pos[A::q] = tq + 1; oLn[3] = E.idx(A::d, pos);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                pos[A::d] = 0; pos[A::p] = tp - 1 + j; pos[A::q] = tq - 1 + o;
                oP[j][o] = E.idx(A::p, pos);     // p-edge in p-cell j at q-node tq-1+o
                pos[A::p] = tp - 1 + o; pos[A::q] = tq - 1 + j;
                oQ[j][o] = E.idx(A::q, pos);     // q-edge in q-cell j at p-node tp-1+o
            }
    }
    // the four transverse edges [p-, p+, q-, q+] of this line at node m
    __device__ __forceinline__ T* t_ptr(int k, int m) const {
        return k == 0 ? ep + oP[0][1] + sp * m : k == 1 ? ep + oP[1][1] + sp * m
             : k == 2 ? eq + oQ[0][1] + sq * m : eq + oQ[1][1] + sq * m;
    }
    __device__ __forceinline__ const T* ts_ptr(int k, int m) const {      // their sources
        return k == 0 ? spp + oP[0][1] + sp * m : k == 1 ? spp + oP[1][1] + sp * m
             : k == 2 ? sqp + oQ[0][1] + sq * m : sqp + oQ[1][1] + sq * m;
    }
    // outer transverse edges of the end faces at node m: quadrant (jp, jq)
    __device__ __forceinline__ T epo(int jp, int jq, int m) const { return ep[oP[jp][2 * jq] + sp * m]; }
    __device__ __forceinline__ T eqo(int jp, int jq, int m) const { return eq[oQ[jq][2 * jp] + sq * m]; }
};

// right-hand side of the line edge of cell i: source + outer parallel line edges
template <typename T, int D>
__device__ __forceinline__ T line_rhs(const LineAddr<T, D>& a, int i, const CellCoef& c, const T eo[4]) {
    T acc = ldg(a.sdp + a.oL + a.sd * i);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += c.gaa[k] * eo[k];
    return acc;
}

// ---- one line sweep: forward and backward block substitution ----------------
//
// What one thread streams per cell of its line:
//   forward  : 11 factor entries, 5 sources, 4 parallel line edges of the next
//              cell, 8 outer transverse edges of the end faces, 4 zeta;
//              writes g_m (4) and bL (1)
//   backward : 11 factor entries, g_m (4), bL (1), 4 zeta; writes T_m (4), L (1)
template <typename T, int D>
__device__ void sweep_line(const Line<T, D>& ln, const LineAddr<T, D>& a) {
    using A = Ax<D>;
    const Model<T>& m = ln.m;
    const int N = ln.N;

    // ---------------- forward ----------------
    double zc[2][2], zn[2][2], gs[4];
    ln.load_zeta(0, zc);
    ln.side_g(zc, gs);
    CellCoef cc, cn;
    cell_coef<T, D>(ln, gs, ldg(m.rh[A::d]), cc);
    T eo_c[4], eo_n[4];                  // outer parallel neighbours of L in the current / next cell
#pragma unroll
    for (int k = 0; k < 4; ++k) eo_c[k] = a.ed[a.oLn[k]];
    T rl_c = ldg(a.fac + (int64_t)(N - 1) * FAC_BS);         // 1 / dL_0
    T bl_c = line_rhs<T, D>(a, 0, cc, eo_c);
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

    // ---------------- backward ----------------
    // cc / rl_c / bl_c now belong to the last cell N-1; T_N is fixed data
    T tn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tn[k] = *a.t_ptr(k, N);
    for (int i = N - 2; i >= 0; --i) {                       // node mn = i + 1, cell mn
        const int mn = i + 1;
        const T* fp = a.fac + (int64_t)i * FAC_BS;
        T X[10];
#pragma unroll
        for (int e = 0; e < 10; ++e) X[e] = ldg(fp + e * FAC_ES);
        if (i < N - 2) {                                     // (the last cell's data are at hand)
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
    if (N > 1) {
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

template <typename T, int D>
__device__ __forceinline__ void sweep_line_direct(const Model<T>& m, int tp, int tq, const T* fac,
                                                  const LineSlots& ls, const FieldView<T>& E,
                                                  const FieldView<const T>& S) {
    Line<T, D> ln(m, tp, tq);
    LineAddr<T, D> a(E, S, fac + ls.base(ls.slot(tp, tq), ln.N), tp, tq);
    sweep_line<T, D>(ln, a);
}


// ---- kernels ---------------------------------------------------------------
// thread t of parity class c  <->  slot off[c] + t  <->  line (1 + cp + 2 a, 1 + cq + 2 b)
__device__ __forceinline__ bool class_line(const LineSlots& ls, int c, int t, int& tp, int& tq) {
    if (t >= ls.cnt[c]) return false;
    const int cp = c & 1, cq = c >> 1;
    tp = 1 + cp + 2 * (t % ls.na[cp]);
    tq = 1 + cq + 2 * (t / ls.na[cp]);
    return true;
}

template <typename T, int D>
__global__ void __launch_bounds__(64)
line_factor_kernel(Model<T> m, T* fac, LineSlots ls, int c) {
    int tp, tq;
    if (!class_line(ls, c, blockIdx.x * blockDim.x + threadIdx.x, tp, tq)) return;
    factor_line<T, D>(m, tp, tq, fac, ls);
}

template <typename T, int D>
__global__ void __launch_bounds__(64)
gs_line_color_kernel(Model<T> m, const T* fac, LineSlots ls, T* e, const T* s, int c) {
    int tp, tq;
    if (!class_line(ls, c, blockIdx.x * blockDim.x + threadIdx.x, tp, tq)) return;
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    sweep_line_direct<T, D>(m, tp, tq, fac, ls, E, S);
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

template <typename T, int D>
static void factor_dir(const Model<T>& m, T* fac, cudaStream_t st) {
    using A = Ax<D>;
    const int npi = m.d.n[A::p] - 1, nqi = m.d.n[A::q] - 1;
    if (npi < 1 || nqi < 1) return;
    LineSlots ls(npi, nqi);
    for (int c = 0; c < 4; ++c) {
        if (ls.cnt[c] == 0) continue;
        ++g_launch_count; line_factor_kernel<T, D><<<(ls.cnt[c] + 63) / 64, 64, 0, st>>>(m, fac, ls, c);
    }
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
        const int want = order == ORDER_LEX ? nqi : (npi * nqi + 3) / 4;
        while (threads < 256 && threads < want) threads <<= 1;
        ++g_launch_count; gs_line_small_kernel<T, D><<<1, threads, 0, st>>>(m, fac, ls, e, s, nu, order);
        return;
    }
    bool back = (order >> 8) & 1;   // bits 8+ of `order`: sweeps already done (phase)
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
                const int c = back ? 3 - cc : cc;
                if (ls.cnt[c] == 0) continue;
                // Consecutive sweeps run the colours in opposite order, so the first
                // colour of a sweep is the last colour of the previous one.  Lines of one
                // colour do not interact and nothing changed in between: solving them
                // again reproduces the same values (block relaxation is idempotent), so
                // that launch is skipped -- 7 instead of 8 colour launches for nu = 2.
                if (sw > 0 && cc == 0) continue;
                const int threads = 64;
                ++g_launch_count; gs_line_color_kernel<T, D><<<(ls.cnt[c] + threads - 1) / threads, threads, 0, st>>>(m, fac, ls, e, s, c);
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
