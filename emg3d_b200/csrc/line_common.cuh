// Geometry, addressing and small dense helpers shared by the line smoothers
// (gs_line.cu: one thread per line; gs_line_seg.cu: one warp per line).
#pragma once
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

// LDL^T (no pivoting) of a complex-symmetric (not Hermitian) 4x4 in packed lower-triangular
// storage, in place: the diagonal entries become 1 / D(j), the strict lower triangle the unit
// factor L.  The blocks are applied by SUBSTITUTION with these factors (backward stable).  An
// explicit inverse assembled from them (r1) is not: where the conductivity is tiny (air, 1e8 Ohm m)
// the blocks are conditioned ~1e10 (the gradient of the node's hat function is almost in the null
// space), the Schur recurrence S_m = D_m - E_m S_{m-1}^{-1} E_m then lost all digits after two
// blocks and line relaxation diverged on the marine model from 64^3 cells on.
template <typename T>
__device__ __forceinline__ void ldl4_factor(T x[10]) {
    T l[4][4], dinv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        T v[4];
        T dj = x[tri(j, j)];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < j) {
                v[k] = l[j][k] * x[tri(k, k)];      // L(j,k) D(k); x(k,k) still holds D(k)
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
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        x[tri(r, r)] = dinv[r];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < r) x[tri(r, c)] = l[r][c];
    }
}

// out = S^{-1} v with the factors of ldl4_factor: L y = v, z = D^{-1} y, L^T out = z
template <typename T>
__device__ __forceinline__ void symv4(const T f[10], const T v[4], T out[4]) {
    T y0 = v[0];
    T y1 = v[1] - f[tri(1, 0)] * y0;
    T y2 = v[2] - f[tri(2, 0)] * y0 - f[tri(2, 1)] * y1;
    T y3 = v[3] - f[tri(3, 0)] * y0 - f[tri(3, 1)] * y1 - f[tri(3, 2)] * y2;
    y0 = f[tri(0, 0)] * y0;
    y1 = f[tri(1, 1)] * y1;
    y2 = f[tri(2, 2)] * y2;
    y3 = f[tri(3, 3)] * y3;
    out[3] = y3;
    out[2] = y2 - f[tri(3, 2)] * y3;
    out[1] = y1 - f[tri(2, 1)] * out[2] - f[tri(3, 1)] * y3;
    out[0] = y0 - f[tri(1, 0)] * out[1] - f[tri(2, 0)] * out[2] - f[tri(3, 0)] * y3;
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

// ---- factorisation: one thread per line ------------------------------------
// QP = 0: layout of the one-thread-per-line kernels, `fbase` = (block 0, entry 0) of the
// line, entries FAC_ES apart, blocks FAC_BS apart, 1/dL_0 in block N-1.
// QP > 0: segment layout of gs_line_seg.cu, `fbase` = the line's chunk: block i = 8 q + j
// at ((j * FAC_NE + e) * QP + q), 1/dL_0 in the header behind the 8 * FAC_NE * QP entries.
// Lines cut by multi-GPU z-slabs (z-lines only): `xin` = the factors of the block BELOW the
// line's first one (the last block of the lower rank's piece, 10 entries) continue the global
// recurrence S_m = D_m - E_m S_{m-1}^{-1} E_m across the cut; `xout` receives the factors of
// the last block.
template <typename T, int D, int QP>
__device__ void factor_line(const Model<T>& m, int tp, int tq, T* __restrict__ fbase,
                            const T* __restrict__ xin = nullptr, T* __restrict__ xout = nullptr) {
    using A = Ax<D>;
    Line<T, D> ln(m, tp, tq);
    const int N = ln.N;

    double zc[2][2], zn[2][2], gs[4];
    ln.load_zeta(0, zc);
    ln.side_g(zc, gs);
    CellCoef cc, cn;
    cell_coef<T, D>(ln, gs, ldg(m.rh[A::d]), cc);
    T rl_c = rcp(line_diag<T, D>(ln, 0, cc));
    if (QP == 0) fbase[(int64_t)(N - 1) * FAC_BS] = rl_c;    // 1 / dL_0
    else fbase[8 * FAC_NE * QP] = rl_c;
    // eta sums carried from the previous line cell for the transverse diagonals
    T etp_c[2], etq_c[2];            // sum over jq (resp. jp) of eta_p / eta_q at line cell i
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        etp_c[j] = ldg(m.eta[A::p] + ln.cbase[j][0]) + ldg(m.eta[A::p] + ln.cbase[j][1]);
        etq_c[j] = ldg(m.eta[A::q] + ln.cbase[0][j]) + ldg(m.eta[A::q] + ln.cbase[1][j]);
    }
    T X[10];                         // factors of S_{m-1}
    if (xin) {
#pragma unroll
        for (int e = 0; e < 10; ++e) X[e] = xin[e];
    }
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
        // S_m = D_m - E_m S_{m-1}^{-1} E_m,  E_m = diag(d_c) + rl_c f_c f_c^T: column by column,
        // Z e_c = S_{m-1}^{-1} (E_m e_c) by substitution with the previous block's factors
        if (i > 0 || xin) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                T ecol[4], z[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    ecol[r] = (cc.f[c] * cc.f[r]) * rl_c;
                    if (r == c) add_real(ecol[r], cc.d[c]);
                }
                symv4<T>(X, ecol, z);
                T fz = cc.f[0] * z[0];
#pragma unroll
                for (int k = 1; k < 4; ++k) fz += cc.f[k] * z[k];
                fz *= rl_c;
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (r >= c) S[tri(r, c)] -= cc.d[r] * z[r] + cc.f[r] * fz;
            }
        }
        ldl4_factor<T>(S);
        T* out = QP == 0 ? fbase + (int64_t)i * FAC_BS : fbase + ((i & 7) * FAC_NE) * QP + (i >> 3);
        constexpr int es = QP == 0 ? FAC_ES : QP;
#pragma unroll
        for (int e = 0; e < 10; ++e) {
            out[e * es] = S[e];                   // the LDL^T factors (see ldl4_factor)
            X[e] = S[e];
        }
        out[10 * es] = rl_n;
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
    if (xout && (N > 1 || xin)) {
#pragma unroll
        for (int e = 0; e < 10; ++e) xout[e] = X[e];
    }
}


// thread t of parity class c  <->  slot off[c] + t  <->  line (1 + cp + 2 a, 1 + cq + 2 b)
__device__ __forceinline__ bool class_line(const LineSlots& ls, int c, int t, int& tp, int& tq) {
    if (t >= ls.cnt[c]) return false;
    const int cp = c & 1, cq = c >> 1;
    tp = 1 + cp + 2 * (t % ls.na[cp]);
    tq = 1 + cq + 2 * (t / ls.na[cp]);
    return true;
}

}  // namespace emg
