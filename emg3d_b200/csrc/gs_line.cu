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

constexpr int LINE_STAGES = 4;   // factor blocks in flight per line (cp.async ring)

template <int D> struct Ax {
    static constexpr int d = D;
    static constexpr int p = (D == 0) ? 1 : 0;
    static constexpr int q = (D == 2) ? 1 : 2;
};

// Factor layout.  Lines are numbered colour-major ("slot"): the four parity
// classes one after the other, each padded to a multiple of 32 lines, and
// within a class p fastest.  A warp of the colour kernel therefore owns one
// aligned group of 32 consecutive slots, and the factors are stored
//     [group][block i][entry e][lane]          (32 lanes, 15 entries, N blocks)
// so that a warp reads ONE contiguous 15*32*sizeof(T) chunk per block and walks
// through HBM sequentially from block to block.
constexpr int FAC_ES = 32;          // stride between entries of one block
constexpr int FAC_BS = 15 * 32;     // stride between blocks of one line

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
    T* const fbase = fac + ls.base(ls.slot(tp, tq), N);

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
        T* out = fbase + (int64_t)i * FAC_BS;
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
                for (int c = 0; c < r; ++c) out[(e++) * FAC_ES] = S[r][c];
#pragma unroll
            for (int r = 0; r < 5; ++r) out[(10 + r) * FAC_ES] = dinv[r];
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
//
// What one thread streams per block of its line:
//   forward  : 15 factor entries, 5 sources, 4 parallel line edges of the next
//              cell, 8 outer transverse edges of the end faces, 4 zeta
//   backward : 15 factor entries, the 5 intermediate values it stored, 4 zeta
// One thread per line leaves one warp per scheduler (16 k lines per colour at
// 256^3), so nothing hides memory latency unless the loads of later blocks are
// in flight while the current block is being solved.  `Staged` does that with
// cp.async (LDGSTS) into a shared-memory ring [stage][word][thread], STAGES
// blocks ahead, without touching the register file; `Direct` loads on demand
// (used by the small-grid and hyperplane kernels).
constexpr int FWD_WORDS = 32;   // T-words per forward stage: 15 + 5 + 4 + 8
constexpr int BWD_WORDS = 20;   // 15 + 5

// addresses of everything a line touches
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
    // source pointers of the forward words 15.. of block j (node m = j+1, cell j+1)
    __device__ __forceinline__ const T* fwd_src(int w, int j) const {
        const int m = j + 1;
        switch (w) {
            case 15: return sdp + oL + sd * j;
            case 16: return spp + oP[0][1] + sp * m;
            case 17: return spp + oP[1][1] + sp * m;
            case 18: return sqp + oQ[0][1] + sq * m;
            case 19: return sqp + oQ[1][1] + sq * m;
            case 20: case 21: case 22: case 23: return ed + oLn[w - 20] + sd * m;
            case 24: return ep + oP[0][0] + sp * m;   // (jp, jq) = (0, 0)
            case 25: return ep + oP[0][2] + sp * m;   // (0, 1)
            case 26: return ep + oP[1][0] + sp * m;   // (1, 0)
            case 27: return ep + oP[1][2] + sp * m;   // (1, 1)
            case 28: return eq + oQ[0][0] + sq * m;   // (0, 0)
            case 29: return eq + oQ[1][0] + sq * m;   // (0, 1): q-cell 1, p-node tp-1
            case 30: return eq + oQ[0][2] + sq * m;   // (1, 0)
            default: return eq + oQ[1][2] + sq * m;   // (1, 1)
        }
    }
    __device__ __forceinline__ const T* bwd_src(int w, int i) const {
        const int m = i + 1;
        switch (w) {
            case 15: return ed + oL + sd * i;
            case 16: return ep + oP[0][1] + sp * m;
            case 17: return ep + oP[1][1] + sp * m;
            case 18: return eq + oQ[0][1] + sq * m;
            default: return eq + oQ[1][1] + sq * m;
        }
    }
};

// Forward words: f15 = w[0..14]; s5 = w[15..19]; en = w[20..23];
// epo(jp,jq) = w[24 + 2 jp + jq]; eqo(jp,jq) = w[28 + 2 jp + jq].
template <typename T, int NW>
struct RegWords {            // words held in registers
    T v[NW];
    __device__ __forceinline__ T operator[](int e) const { return v[e]; }
};
template <typename T>
struct SmemWords {           // words read from the shared-memory ring on use
    const T* src;
    int nt;
    __device__ __forceinline__ T operator[](int e) const { return src[e * nt]; }
};

template <typename T, int D>
struct Direct {
    using FwdView = RegWords<T, FWD_WORDS>;
    using BwdView = RegWords<T, BWD_WORDS>;
    const LineAddr<T, D>& a;
    const Line<T, D>& ln;
    int N;
    __device__ Direct(const LineAddr<T, D>& a_, const Line<T, D>& ln_) : a(a_), ln(ln_), N(ln_.N) {}
    __device__ __forceinline__ void fwd_start() {}
    __device__ __forceinline__ void bwd_start() {}
    __device__ __forceinline__ FwdView fwd_get(int j, double zn[2][2]) {
        FwdView w;
        const T* fp = a.fac + (int64_t)j * FAC_BS;
#pragma unroll
        for (int e = 0; e < 15; ++e) w.v[e] = ldg(fp + e * FAC_ES);
        w.v[15] = ldg(a.fwd_src(15, j));
        if (j < N - 1) {
#pragma unroll
            for (int e = 16; e < 20; ++e) w.v[e] = ldg(a.fwd_src(e, j));
#pragma unroll
            for (int e = 20; e < FWD_WORDS; ++e) w.v[e] = *a.fwd_src(e, j);
            ln.load_zeta(j + 1, zn);
        }
        return w;
    }
    __device__ __forceinline__ BwdView bwd_get(int, int i, double zc[2][2]) {
        BwdView w;
        const T* fp = a.fac + (int64_t)i * FAC_BS;
#pragma unroll
        for (int e = 0; e < 15; ++e) w.v[e] = ldg(fp + e * FAC_ES);
#pragma unroll
        for (int e = 15; e < BWD_WORDS; ++e) w.v[e] = *a.bwd_src(e, i);
        if (i > 0) ln.load_zeta(i, zc);
        return w;
    }
};

template <typename T, int D, int STAGES>
struct Staged {
    using FwdView = SmemWords<T>;
    using BwdView = SmemWords<T>;
    const LineAddr<T, D>& a;
    const Line<T, D>& ln;
    int N, nt;
    T* smT;          // this thread's column: word w of stage s at smT[(s * FWD_WORDS + w) * nt]
    double* smZ;     // zeta ring: smZ[(s * 4 + c) * nt]
    __device__ Staged(const LineAddr<T, D>& a_, const Line<T, D>& ln_, T* smT_, double* smZ_, int nt_)
        : a(a_), ln(ln_), N(ln_.N), nt(nt_), smT(smT_), smZ(smZ_) {}

    __device__ __forceinline__ void fwd_issue(int j) {
        if (j < N) {
            T* dst = smT + (int64_t)(j % STAGES) * FWD_WORDS * nt;
            const T* fp = a.fac + (int64_t)j * FAC_BS;
#pragma unroll
            for (int e = 0; e < 15; ++e) cp_async(dst + e * nt, fp + e * FAC_ES);
            cp_async(dst + 15 * nt, a.fwd_src(15, j));
            if (j < N - 1) {
#pragma unroll
                for (int e = 16; e < FWD_WORDS; ++e) cp_async(dst + e * nt, a.fwd_src(e, j));
                double* dz = smZ + (int64_t)(j % STAGES) * 4 * nt;
#pragma unroll
                for (int jp = 0; jp < 2; ++jp)
#pragma unroll
                    for (int jq = 0; jq < 2; ++jq)
                        cp_async(dz + (2 * jp + jq) * nt,
                                 ln.m.zeta + ln.cbase[jp][jq] + ln.cstr * (j + 1));
            }
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void fwd_start() {
#pragma unroll
        for (int k = 0; k < STAGES - 1; ++k) fwd_issue(k);
    }
    // The stage overwritten by the new copies is the one consumed in the previous
    // step; its words were all read (into registers) before this call.
    __device__ __forceinline__ FwdView fwd_get(int j, double zn[2][2]) {
        fwd_issue(j + STAGES - 1);
        cp_async_wait<STAGES - 1>();
        const double* sz = smZ + (int64_t)(j % STAGES) * 4 * nt;
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) zn[jp][jq] = sz[(2 * jp + jq) * nt];
        return FwdView{smT + (int64_t)(j % STAGES) * FWD_WORDS * nt, nt};
    }
    // backward sequence: k-th step handles block i = N - 2 - k
    __device__ __forceinline__ void bwd_issue(int k) {
        const int i = N - 2 - k;
        if (i >= 0) {
            T* dst = smT + (int64_t)(k % STAGES) * FWD_WORDS * nt;
            const T* fp = a.fac + (int64_t)i * FAC_BS;
#pragma unroll
            for (int e = 0; e < 15; ++e) cp_async(dst + e * nt, fp + e * FAC_ES);
#pragma unroll
            for (int e = 15; e < BWD_WORDS; ++e) cp_async(dst + e * nt, a.bwd_src(e, i));
            if (i > 0) {
                double* dz = smZ + (int64_t)(k % STAGES) * 4 * nt;
#pragma unroll
                for (int jp = 0; jp < 2; ++jp)
#pragma unroll
                    for (int jq = 0; jq < 2; ++jq)
                        cp_async(dz + (2 * jp + jq) * nt, ln.m.zeta + ln.cbase[jp][jq] + ln.cstr * i);
            }
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void bwd_start() {
        // our own stores of the forward pass must be visible to the async copies
        __threadfence_block();
#pragma unroll
        for (int k = 0; k < STAGES - 1; ++k) bwd_issue(k);
    }
    __device__ __forceinline__ BwdView bwd_get(int k, int, double zc[2][2]) {
        bwd_issue(k + STAGES - 1);
        cp_async_wait<STAGES - 1>();
        const double* sz = smZ + (int64_t)(k % STAGES) * 4 * nt;
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) zc[jp][jq] = sz[(2 * jp + jq) * nt];
        return BwdView{smT + (int64_t)(k % STAGES) * FWD_WORDS * nt, nt};
    }
};

// Factor stream through TMA: the factors of one block of the 32 lines of a warp are
// ONE contiguous chunk of 15 * 32 words (layout above), so a single elected lane
// fetches it with one cp.async.bulk (UBLKCP) into a per-warp shared-memory ring,
// BULK_STAGES blocks ahead of the block being solved, completion on an mbarrier.
// The fetch sequence runs through the forward pass (blocks 0 .. N-1) and straight on
// into the backward pass (blocks N-2 .. 0), so the ring never drains between the
// passes.  No registers and no LSU instructions are spent on the factor stream; the
// strided E / S / zeta accesses stay ordinary loads through the L1.
#ifndef EMG_LINE_BULK
#define EMG_LINE_BULK 0
#endif
#ifndef EMG_LINE_BULK_STAGES
#define EMG_LINE_BULK_STAGES 4
#endif
constexpr int BULK_STAGES = EMG_LINE_BULK_STAGES;

template <typename T, int D>
struct BulkFac {
    using FwdView = RegWords<T, FWD_WORDS>;
    using BwdView = RegWords<T, BWD_WORDS>;
    static constexpr unsigned BYTES = FAC_BS * sizeof(T);
    const LineAddr<T, D>& a;
    const Line<T, D>& ln;
    int N, lane, q;
    unsigned mask;
    T* ring;              // this warp's ring: [stage][entry][lane]
    uint64_t* bars;       // this warp's mbarriers: [stage]
    const T* gsrc;        // (block 0, entry 0, lane 0) of the warp's 32-line group
    __device__ BulkFac(const LineAddr<T, D>& a_, const Line<T, D>& ln_, T* ring_, uint64_t* bars_,
                       unsigned mask_)
        : a(a_), ln(ln_), N(ln_.N), lane(threadIdx.x & 31), q(0), mask(mask_), ring(ring_), bars(bars_),
          gsrc(a_.fac - (threadIdx.x & 31)) {}

    // fetch number qq of the sequence: forward blocks 0 .. N-1, then backward N-2 .. 0
    __device__ __forceinline__ void issue(int qq) {
        if (qq > 2 * N - 2) return;
        const int blk = qq < N ? qq : 2 * N - 2 - qq;
        const int st = qq % BULK_STAGES;
        mbar_expect_tx(bars + st, BYTES);
        bulk_g2s(ring + st * FAC_BS, gsrc + (int64_t)blk * FAC_BS, BYTES, bars + st);
    }
    __device__ __forceinline__ void fwd_start() {
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < BULK_STAGES; ++k) issue(k);
        }
    }
    __device__ __forceinline__ void bwd_start() {}
    template <class V>
    __device__ __forceinline__ void factors(V& w) {
        const int st = q % BULK_STAGES;
        mbar_wait(bars + st, (q / BULK_STAGES) & 1);
        const T* p = ring + st * FAC_BS + lane;
#pragma unroll
        for (int e = 0; e < 15; ++e) w.v[e] = p[e * FAC_ES];
        __syncwarp(mask);                       // every lane has read the stage
        if (lane == 0) {
            fence_proxy_async_smem();
            issue(q + BULK_STAGES);
        }
        ++q;
    }
    __device__ __forceinline__ FwdView fwd_get(int j, double zn[2][2]) {
        FwdView w;
        factors(w);
        w.v[15] = ldg(a.fwd_src(15, j));
        if (j < N - 1) {
#pragma unroll
            for (int e = 16; e < 20; ++e) w.v[e] = ldg(a.fwd_src(e, j));
#pragma unroll
            for (int e = 20; e < FWD_WORDS; ++e) w.v[e] = *a.fwd_src(e, j);
            ln.load_zeta(j + 1, zn);
        }
        return w;
    }
    __device__ __forceinline__ BwdView bwd_get(int, int i, double zc[2][2]) {
        BwdView w;
        factors(w);
#pragma unroll
        for (int e = 15; e < BWD_WORDS; ++e) w.v[e] = *a.bwd_src(e, i);
        if (i > 0) ln.load_zeta(i, zc);
        return w;
    }
};

template <typename T, class W>
__device__ __forceinline__ void solve5(const W& f, T y[5], bool first_zero) {
    T L[10], dinv[5];
#pragma unroll
    for (int e = 0; e < 10; ++e) L[e] = f[e];
#pragma unroll
    for (int e = 0; e < 5; ++e) dinv[e] = f[10 + e];
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

template <typename T, int D, class Loader>
__device__ void sweep_line(const Line<T, D>& ln, const LineAddr<T, D>& a, Loader& ld) {
    using A = Ax<D>;
    const Model<T>& m = ln.m;
    const int N = ln.N;

    // ---------------- forward ----------------
    ld.fwd_start();
    double zc[2][2], zn[2][2], gs[4], gn[4];
    ln.load_zeta(0, zc);
    ln.side_g(zc, gs);
    double rd = ldg(m.rh[A::d]);
    T eo[4];                             // parallel neighbours of L_i
#pragma unroll
    for (int k = 0; k < 4; ++k) eo[k] = a.ed[a.oLn[k]];
    // transverse part of the previous block's solution; for the first block the
    // (fixed) transverse edges on the line's start plane: zero on a PEC boundary,
    // halo data of the neighbouring slab on a multi-GPU z-window
    T wT[4];
    wT[0] = a.ep[a.oP[0][1]];
    wT[1] = a.ep[a.oP[1][1]];
    wT[2] = a.eq[a.oQ[0][1]];
    wT[3] = a.eq[a.oQ[1][1]];

    for (int i = 0; i < N; ++i) {
        const bool last = (i == N - 1);
        const typename Loader::FwdView w = ld.fwd_get(i, zn);
        T y[5];
        // line edge
        {
            T acc = w[15];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double as = ln.a_side(k);
                acc += (gs[k] * as * as) * eo[k];
            }
            y[0] = acc;
        }
        double f[4], dk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[k] = -gs[k] * ln.a_side(k) * rd;
            dk[k] = -gs[k] * rd * rd;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) y[0] -= f[k] * wT[k];
        if (last) {
            // coupling to the fixed transverse edges on the line's end plane
            const T tN[4] = {a.ep[a.oP[0][1] + a.sp * N], a.ep[a.oP[1][1] + a.sp * N],
                             a.eq[a.oQ[0][1] + a.sq * N], a.eq[a.oQ[1][1] + a.sq * N]};
#pragma unroll
            for (int k = 0; k < 4; ++k) y[0] += f[k] * tN[k];
            a.ed[a.oL + a.sd * i] = y[0] * w[0];
            break;
        }
        // transverse edges at node m = i+1
        const int mnode = i + 1;
        ln.side_g(zn, gn);
        const double rdn = ldg(m.rh[A::d] + i + 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) y[1 + k] = w[16 + k];
        // side faces of L_i (+) and L_{i+1} (-)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double as = ln.a_side(k);
            y[1 + k] += (gs[k] * rd * as) * eo[k] - (gn[k] * rdn * as) * w[20 + k];
        }
        // end faces at node m
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) {
                const double g = 0.5 * (zc[jp][jq] + zn[jp][jq]);
                const double ap = ln.al_p(jq), aq = ln.al_q(jp);
                const T out = ap * w[24 + 2 * jp + jq] + aq * w[28 + 2 * jp + jq];
                y[1 + jp] += (g * ap) * out;
                y[3 + jq] += (g * aq) * out;
            }
#pragma unroll
        for (int k = 0; k < 4; ++k) y[1 + k] -= dk[k] * wT[k];
        if (i == N - 2) {
            // last interior node: its neighbours on the end plane are fixed data
            const T tN[4] = {a.ep[a.oP[0][1] + a.sp * N], a.ep[a.oP[1][1] + a.sp * N],
                             a.eq[a.oQ[0][1] + a.sq * N], a.eq[a.oQ[1][1] + a.sq * N]};
#pragma unroll
            for (int k = 0; k < 4; ++k) y[1 + k] += (gn[k] * rdn * rdn) * tN[k];
        }
        solve5<T>(w, y, false);
        a.ed[a.oL + a.sd * i] = y[0];
        a.ep[a.oP[0][1] + a.sp * mnode] = y[1];
        a.ep[a.oP[1][1] + a.sp * mnode] = y[2];
        a.eq[a.oQ[0][1] + a.sq * mnode] = y[3];
        a.eq[a.oQ[1][1] + a.sq * mnode] = y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { wT[k] = y[1 + k]; eo[k] = w[20 + k]; gs[k] = gn[k]; }
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) zc[jp][jq] = zn[jp][jq];
        rd = rdn;
    }

    // ---------------- backward ----------------
    // x_i = w_i - S_i^{-1} F_{i+1}^T x_{i+1};  gs / rd now belong to line cell N-1
    T xL = a.ed[a.oL + a.sd * (N - 1)];
    ld.bwd_start();
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
        const typename Loader::BwdView w = ld.bwd_get(N - 2 - i, i, zc);
        solve5<T>(w, v, true);
        const int mnode = i + 1;
        xL = w[15] - v[0];
#pragma unroll
        for (int k = 0; k < 4; ++k) xT[k] = w[16 + k] - v[1 + k];
        a.ed[a.oL + a.sd * i] = xL;
        a.ep[a.oP[0][1] + a.sp * mnode] = xT[0];
        a.ep[a.oP[1][1] + a.sp * mnode] = xT[1];
        a.eq[a.oQ[0][1] + a.sq * mnode] = xT[2];
        a.eq[a.oQ[1][1] + a.sq * mnode] = xT[3];
        if (i > 0) {
            ln.side_g(zc, gs);
            rd = ldg(m.rh[A::d] + i);
        }
    }
}

template <typename T, int D>
__device__ __forceinline__ void sweep_line_direct(const Model<T>& m, int tp, int tq, const T* fac,
                                                  const LineSlots& ls, const FieldView<T>& E,
                                                  const FieldView<const T>& S) {
    Line<T, D> ln(m, tp, tq);
    LineAddr<T, D> a(E, S, fac + ls.base(ls.slot(tp, tq), ln.N), tp, tq);
    Direct<T, D> ld(a, ln);
    sweep_line<T, D>(ln, a, ld);
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

#ifndef EMG_LINE_STAGED
#define EMG_LINE_STAGED 0
#endif

template <typename T, int D>
__global__ void __launch_bounds__(64)
gs_line_color_kernel(Model<T> m, const T* fac, LineSlots ls, T* e, const T* s, int c) {
    int tp, tq;
    const bool valid = class_line(ls, c, blockIdx.x * blockDim.x + threadIdx.x, tp, tq);
#if EMG_LINE_BULK && !EMG_LINE_STAGED
    const unsigned mask = __ballot_sync(0xffffffffu, valid);
#endif
    if (!valid) return;
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    Line<T, D> ln(m, tp, tq);
    LineAddr<T, D> ad(E, S, fac + ls.base(ls.slot(tp, tq), ln.N), tp, tq);
#if EMG_LINE_BULK && !EMG_LINE_STAGED
    extern __shared__ __align__(128) unsigned char ring_raw[];
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    T* ring = reinterpret_cast<T*>(ring_raw) + (size_t)warp * BULK_STAGES * FAC_BS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<T*>(ring_raw) +
                                                 (size_t)nwarps * BULK_STAGES * FAC_BS) + warp * BULK_STAGES;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < BULK_STAGES; ++k) mbar_init(bars + k, 1);
        mbar_init_fence();
    }
    __syncwarp(mask);
    BulkFac<T, D> ld(ad, ln, ring, bars, mask);
#elif EMG_LINE_STAGED
    extern __shared__ __align__(16) unsigned char ring_raw[];
    const int nt = blockDim.x;
    T* smT = reinterpret_cast<T*>(ring_raw) + threadIdx.x;
    double* smZ = reinterpret_cast<double*>(reinterpret_cast<T*>(ring_raw) +
                                            (size_t)LINE_STAGES * FWD_WORDS * nt) + threadIdx.x;
    Staged<T, D, LINE_STAGES> ld(ad, ln, smT, smZ, nt);
#else
    Direct<T, D> ld(ad, ln);
#endif
    sweep_line<T, D>(ln, ad, ld);
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
                const int threads = EMG_LINE_STAGED ? 32 : 64;
                size_t smem = 0;
#if EMG_LINE_BULK && !EMG_LINE_STAGED
                smem = (size_t)(threads / 32) * BULK_STAGES * (FAC_BS * sizeof(T) + sizeof(uint64_t));
                static bool bulk_attr_set = false;  // per template instance
                if (!bulk_attr_set) {
                    cudaFuncSetAttribute(gs_line_color_kernel<T, D>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    bulk_attr_set = true;
                }
#endif
#if EMG_LINE_STAGED
                smem = (size_t)LINE_STAGES * threads * (FWD_WORDS * sizeof(T) + 4 * sizeof(double));
                static bool attr_set = false;       // per template instance
                if (!attr_set) {
                    cudaFuncSetAttribute(gs_line_color_kernel<T, D>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    attr_set = true;
                }
#endif
                ++g_launch_count; gs_line_color_kernel<T, D><<<(ls.cnt[c] + threads - 1) / threads, threads, smem, st>>>(m, fac, ls, e, s, c);
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
