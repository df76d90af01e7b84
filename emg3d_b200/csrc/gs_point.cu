// Point-block Gauss-Seidel smoother: for every interior node, solve
// simultaneously for the six edges meeting at it (what emg3d/core.py:210-503
// `gauss_seidel` computes).
//
// Local system of node (ix, iy, iz), unknown order as in the reference
// [ex(ix-1), ex(ix), ey(iy-1), ey(iy), ez(iz-1), ez(iz)]:
//   - the 12 faces containing the node each hold two local and two outer
//     edges; a face in the plane of axes (p, q) at cells (cp, cq) has the curl
//     stencil  e_p(q-node cq): +1/h_q, e_p(cq+1): -1/h_q,
//              e_q(p-node cp+1): +1/h_p, e_q(cp): -1/h_p
//     (only products of stencil entries matter, so one global sign is free);
//   - matrix  += 1/2 M_f c_loc c_loc^T,  rhs -= 1/2 M_f c_loc (c_out . e_out);
//   - diagonal -= 1/4 (sum of eta over the 4 cells around the edge); rhs += s.
// The 6x6 complex-symmetric system is solved in registers: the two x-edges are
// eliminated analytically, the remaining 4x4 Schur complement by LDL^T.
//
// Orderings: `lex` reproduces the reference's lexicographic
// sweep exactly by running hyperplanes t = ix + 2 iy + 3 iz one after the
// other (nodes on one hyperplane never share or neighbour an edge, and every
// dependency of the sequential sweep points to a smaller t); `color` runs the
// 8 parity classes (ix&1, iy&1, iz&1), which are conflict-free as well.
#include "common.cuh"
#include "kernels.h"

#include <stdio.h>
#include <stdlib.h>

namespace emg {

// Local system of one node in structured form.  Unknown order as in the
// reference: X = (ex-, ex+), T = (ey-, ey+, ez-, ez+).  The two x-edges do not
// couple (core.py:416-418), and neither do ey-/ey+ nor ez-/ez+, and all
// off-diagonal entries are real:
//
//        | dX   B  |      dX  : 2 complex diagonal entries
//    A = |         |      B   : 2 x 4 real (x-edge <-> transverse edge)
//        | B^T  C  |      C   : 4 complex diagonal entries dT + 4 real entries
//                               cyz[jz][jy] (ez <-> ey)
//
// Eliminating X first is the reference's own elimination order (0, 1, 2, ...)
// with the structural zeros skipped: S = C - B^T dX^-1 B is a dense 4x4
// complex-symmetric matrix that is then factorised by LDL^T.  Compared with a
// dense 6x6 LDL^T this needs about 40 % of the flops and half the registers.
template <typename T>
struct NodeSys {
    T dX[2], dT[4], bX[2], bT[4];
    double B[2][4], cyz[2][2];
};

// Contribution of the face in plane (P, Q), quadrant (SP, SQ) of the node.
// All indices are template parameters so that everything stays in registers.
// Addresses relative to the node: pe[c] points at element (ix, iy, iz) of component c, an
// element at offset (dx, dy, dz) is pe[c] + dx + dy s1[c] + dz s2[c] with COMPILE-TIME offsets in
// {-1, 0, 1}: the x-offset becomes an immediate of the load, the few distinct (dy, dz) rows are
// shared sub-expressions -- instead of a 64-bit index polynomial per access (r2: 850 instructions
// per node, ~ 300 of them integer / address arithmetic; the kernel is latency-, i.e. instruction-
// bound at 12 warps per SM).
template <typename P_>
struct NodePtr {
    P_* pe[3];
    int64_t s1[3], s2[3];
    template <int C, int DX, int DY, int DZ>
    __device__ __forceinline__ P_* at() const { return pe[C] + DX + DY * s1[C] + DZ * s2[C]; }
};

template <typename T>
struct Faces {
    template <int P, int Q, int SP, int SQ>
    static __device__ __forceinline__ void quad(const NodePtr<T>& E,
                                                const double (&z)[2][2][2], const double (&rh)[3][2],
                                                NodeSys<T>& n) {
        constexpr int W = 3 - P - Q;
        // cells of this face: P-index 1-SP, Q-index 1-SQ, both W-indices
        constexpr int i0x = P == 0 ? 1 - SP : Q == 0 ? 1 - SQ : 0;
        constexpr int i0y = P == 1 ? 1 - SP : Q == 1 ? 1 - SQ : 0;
        constexpr int i0z = P == 2 ? 1 - SP : Q == 2 ? 1 - SQ : 0;
        constexpr int i1x = W == 0 ? 1 : i0x, i1y = W == 1 ? 1 : i0y, i1z = W == 2 ? 1 : i0z;
        const double g = 0.5 * (z[i0x][i0y][i0z] + z[i1x][i1y][i1z]);
        const double rp = rh[P][1 - SP], rq = rh[Q][1 - SQ];
        const double al_p = SQ ? -rq : rq;   // stencil entry of the local P-edge
        const double al_q = SP ? rp : -rp;   // stencil entry of the local Q-edge
        // outer P-edge: same P-cell, Q-node moved away from the node
        constexpr int dq = SQ ? -1 : 1, dp = SP ? -1 : 1;
        constexpr int qx = (P == 0 ? -SP : 0) + (Q == 0 ? dq : 0), qy = (P == 1 ? -SP : 0) + (Q == 1 ? dq : 0),
                      qz = (P == 2 ? -SP : 0) + (Q == 2 ? dq : 0);
        const T ep = *E.template at<P, qx, qy, qz>();
        constexpr int rx = (Q == 0 ? -SQ : 0) + (P == 0 ? dp : 0), ry = (Q == 1 ? -SQ : 0) + (P == 1 ? dp : 0),
                      rz = (Q == 2 ? -SQ : 0) + (P == 2 ? dp : 0);
        const T eq = *E.template at<Q, rx, ry, rz>();
        const T out = al_p * ep + al_q * eq;   // = -(c_out . e_out)
        constexpr int kq = 2 * (Q - 1) + (1 - SQ);      // transverse index of the Q-edge
        // (diagonal entries come precombined from m.diag, see edge_diag_kernel)
        if (P == 0) {
            constexpr int j = 1 - SP;                   // which x-edge
            n.B[j][kq] += g * al_p * al_q;
            n.bX[j] += (g * al_p) * out;
        } else {
            constexpr int kp = 1 - SP;                  // P == 1: a y-edge
            n.cyz[1 - SQ][1 - SP] += g * al_p * al_q;
            n.bT[kp] += (g * al_p) * out;
        }
        n.bT[kq] += (g * al_q) * out;
    }
    template <int P, int Q>
    static __device__ __forceinline__ void plane(const NodePtr<T>& E,
                                                 const double (&z)[2][2][2], const double (&rh)[3][2],
                                                 NodeSys<T>& n) {
        quad<P, Q, 0, 0>(E, z, rh, n);
        quad<P, Q, 0, 1>(E, z, rh, n);
        quad<P, Q, 1, 0>(E, z, rh, n);
        quad<P, Q, 1, 1>(E, z, rh, n);
    }
};

#define SS(r, c) s4[((r) * ((r) + 1)) / 2 + (c)]

template <typename T>
__device__ __forceinline__ void node_update(const Model<T>& m, const FieldView<T>& E,
                                            const FieldView<const T>& S, int ix, int iy, int iz) {
    const int nd[3] = {ix, iy, iz};
    const int64_t cs[3] = {1, m.d.n[0], (int64_t)m.d.n[0] * m.d.n[1]};
    const int64_t c0 = (ix - 1) + cs[1] * (iy - 1) + cs[2] * (iz - 1);  // low corner cell

    // reciprocal widths of the two cells along each axis: rh[a][0] = 1/h_a[nd-1]
    double rh[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        rh[a][0] = ldg(m.rh[a] + nd[a] - 1);
        rh[a][1] = ldg(m.rh[a] + nd[a]);
    }
    // zeta of the 8 cells around the node, z[i][j][k] <-> cell (ix-1+i, iy-1+j, iz-1+k)
    double z[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) z[i][j][k] = ldg(m.zeta + c0 + i + cs[1] * j + cs[2] * k);

    NodeSys<T> n;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) n.B[j][k] = 0.0;
    n.cyz[0][0] = n.cyz[0][1] = n.cyz[1][0] = n.cyz[1][1] = 0.0;

    // pointers at the node's own position in every component (E, source, diagonal share the
    // field layout); the edge below the node along its own axis is one own-axis stride back
    NodePtr<T> pe;
    const T* ps[3];
    const T* pd[3];
    int64_t own[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int64_t id = E.idx(c, ix, iy, iz);
        pe.pe[c] = E.p[c] + id;
        pe.s1[c] = E.s1[c];
        pe.s2[c] = E.s2[c];
        ps[c] = S.p[c] + id;
        pd[c] = m.diag + (S.p[c] - S.p[0]) + id;
        own[c] = c == 0 ? 1 : c == 1 ? E.s1[1] : E.s2[2];
    }
    // diagonal of A at the six local edges (precombined per level: four face terms
    // minus 1/4 of the eta sum, see edge_diag_kernel); rhs: source
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            const T src = sg ? ldg(ps[c]) : ldg(ps[c] - own[c]);
            const T dg = sg ? ldg(pd[c]) : ldg(pd[c] - own[c]);
            if (c == 0) {
                n.dX[sg] = dg;
                n.bX[sg] = src;
            } else {
                n.dT[2 * (c - 1) + sg] = dg;
                n.bT[2 * (c - 1) + sg] = src;
            }
        }
    }

    // the three coordinate planes (p, q), four quadrants each (compile-time indices)
    Faces<T>::template plane<0, 1>(pe, z, rh, n);
    Faces<T>::template plane<0, 2>(pe, z, rh, n);
    Faces<T>::template plane<1, 2>(pe, z, rh, n);

    // eliminate the x-edges: S = C - B^T dX^-1 B,  bT' = bT - B^T dX^-1 bX
    T rX[2], tX[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        rX[j] = rcp(n.dX[j]);
        tX[j] = rX[j] * n.bX[j];
    }
    T s4[10];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (l <= k) {
                T v = (n.B[0][k] * n.B[0][l]) * rX[0] + (n.B[1][k] * n.B[1][l]) * rX[1];
                T cc = zero_<T>();
                if (k == l) cc = n.dT[k];
                else if (k >= 2 && l < 2) add_real(cc, n.cyz[k - 2][l]);
                SS(k, l) = cc - v;
            }
        }
        n.bT[k] -= n.B[0][k] * tX[0] + n.B[1][k] * tX[1];
    }

    // 4x4 LDL^T without pivoting, then forward / diagonal / backward substitution.
    // Full-range loops with compile-time guards unroll completely (registers only).
    T dinv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        T v[4];
        T dj = SS(j, j);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < j) {
                v[k] = SS(j, k) * SS(k, k);        // L(j,k) D(k); SS(k, k) holds D(k)
                dj -= SS(j, k) * v[k];
            }
        }
        SS(j, j) = dj;
        const T r = rcp(dj);
        dinv[j] = r;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i > j) {
                T t = SS(i, j);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < j) t -= SS(i, k) * v[k];
                SS(i, j) = t * r;
            }
        }
    }
#pragma unroll
    for (int j = 1; j < 4; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < j) n.bT[j] -= SS(j, k) * n.bT[k];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) n.bT[j] = n.bT[j] * dinv[j];
#pragma unroll
    for (int j = 2; j >= 0; --j) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k > j) n.bT[j] -= SS(k, j) * n.bT[k];
    }
    // back-substitute the x-edges
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        T acc = n.bX[j];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc -= n.B[j][k] * n.bT[k];
        n.bX[j] = rX[j] * acc;
    }

#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            T* dst = sg ? pe.pe[c] : pe.pe[c] - own[c];
            *dst = c == 0 ? n.bX[sg] : n.bT[2 * (c - 1) + sg];
        }
    }
}

#undef SS

// ---- precombined diagonal ------------------------------------------------------
// diag(edge) = sum over the 4 faces around the edge of 1/2 M_f / h^2  -  1/4 sum
// of eta over the 4 cells around it: the diagonal of A.  It depends on the grid
// and the model only, so it is computed once per level; the point smoother then
// loads 6 numbers per node instead of 24 eta values (and skips 12 updates).
// Only interior edges (the unknowns) are filled; the rest is zero.
template <typename T>
__global__ void __launch_bounds__(256) edge_diag_kernel(Model<T> m, T* __restrict__ diag) {
    int nd[3];
    nd[0] = blockIdx.x * blockDim.x + threadIdx.x;
    nd[1] = blockIdx.y * blockDim.y + threadIdx.y;
    nd[2] = blockIdx.z * blockDim.z + threadIdx.z;
    if (nd[0] > m.d.n[0] || nd[1] > m.d.n[1] || nd[2] > m.d.n[2]) return;
    FieldView<T> D(diag, m.d);
    const int64_t cs[3] = {1, m.d.n[0], (int64_t)m.d.n[0] * m.d.n[1]};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int u = (c + 1) % 3, v = (c + 2) % 3;
        if (nd[c] >= m.d.n[c]) continue;
        T val = zero_<T>();
        if (nd[u] >= 1 && nd[u] < m.d.n[u] && nd[v] >= 1 && nd[v] < m.d.n[v]) {
            const int64_t c00 = cs[c] * nd[c] + cs[u] * (nd[u] - 1) + cs[v] * (nd[v] - 1);
            const double z00 = ldg(m.zeta + c00), z10 = ldg(m.zeta + c00 + cs[u]),
                         z01 = ldg(m.zeta + c00 + cs[v]), z11 = ldg(m.zeta + c00 + cs[u] + cs[v]);
            const double ru0 = ldg(m.rh[u] + nd[u] - 1), ru1 = ldg(m.rh[u] + nd[u]);
            const double rv0 = ldg(m.rh[v] + nd[v] - 1), rv1 = ldg(m.rh[v] + nd[v]);
            // faces spanned by (c, u): at u-cell 0/1, zeta summed over the two v-cells
            const double acc = 0.5 * (z00 + z01) * ru0 * ru0 + 0.5 * (z10 + z11) * ru1 * ru1 +
                               0.5 * (z00 + z10) * rv0 * rv0 + 0.5 * (z01 + z11) * rv1 * rv1;
            const T st = ldg(m.eta[c] + c00) + ldg(m.eta[c] + c00 + cs[u]) +
                         ldg(m.eta[c] + c00 + cs[v]) + ldg(m.eta[c] + c00 + cs[u] + cs[v]);
            val = -0.25 * st;
            add_real(val, acc);
        }
        D.p[c][D.idx(c, nd)] = val;
    }
}

template <typename T>
void launch_edge_diag(const Model<T>& m, T* diag, cudaStream_t st) {
    dim3 b(32, 4, 2);
    dim3 g((m.d.n[0] + 1 + b.x - 1) / b.x, (m.d.n[1] + 1 + b.y - 1) / b.y,
           (m.d.n[2] + 1 + b.z - 1) / b.z);
    ++g_launch_count; edge_diag_kernel<T><<<g, b, 0, st>>>(m, diag);
}
template void launch_edge_diag<double>(const Model<double>&, double*, cudaStream_t);
template void launch_edge_diag<cplx>(const Model<cplx>&, cplx*, cudaStream_t);

// ---- schedules ---------------------------------------------------------------

// one parity class: ix = fx + 2 i, iy = fy + 2 j, iz = fz + 2 k
template <typename T>
__global__ void __launch_bounds__(128)
gs_point_color_kernel(Model<T> m, T* e, const T* s, int fx, int fy, int fz, int cx, int cy, int cz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= cx || j >= cy || k >= cz) return;
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    node_update<T>(m, E, S, fx + 2 * i, fy + 2 * j, fz + 2 * k);
}

// Tile-fused multicolour sweep for grids that do not fit the L2: the nodes are
// cut into tiles of TX x TY x TZ nodes; tiles are coloured by the parity of
// their tile index (8 tile colours, one launch each: tiles of one colour are at
// least a full tile apart and cannot interact), and one thread block relaxes
// all 8 node colours of its tile back to back, with a block barrier between
// colours.  The tile's edges, sources and coefficients are then read from HBM
// once per sweep instead of once per node colour (the re-reads hit L1/L2).
// Tile shape (nodes).  Measured at 256^3 (tools/tune_tiles128.sh, r1): 64 x 4 x 4 runs a
// tile-colour launch in 0.159 ms, 32 x 8 x 4 in 0.165, 16 x 8 x 8 in 0.181, 8 x 8 x 8 in 0.246:
// long x-rows make every warp-wide access one contiguous 1 KB run per row.
#ifndef EMG_TILE_X
#define EMG_TILE_X 64
#define EMG_TILE_Y 4
#define EMG_TILE_Z 4
#endif
#ifndef EMG_TILE_MINB
#define EMG_TILE_MINB 3
#endif
constexpr int TX = EMG_TILE_X, TY = EMG_TILE_Y, TZ = EMG_TILE_Z;
constexpr int TILE_THREADS = (TX / 2) * (TY / 2) * (TZ / 2);

template <typename T>
__global__ void __launch_bounds__(TILE_THREADS, EMG_TILE_MINB)
gs_point_tile_kernel(Model<T> m, T* e, const T* s, int tcx, int tcy, int tcz, int back) {
    // tile index of this block within its tile colour
    const int x0 = 1 + (2 * blockIdx.x + tcx) * TX;
    const int y0 = 1 + (2 * blockIdx.y + tcy) * TY;
    const int z0 = 1 + (2 * blockIdx.z + tcz) * TZ;
    const int t = threadIdx.x;
    const int i = t % (TX / 2), j = (t / (TX / 2)) % (TY / 2), k = t / ((TX / 2) * (TY / 2));
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    for (int cc = 0; cc < 8; ++cc) {
        const int c = back ? 7 - cc : cc;
        const int ix = x0 + 2 * i + (c & 1), iy = y0 + 2 * j + ((c >> 1) & 1),
                  iz = z0 + 2 * k + ((c >> 2) & 1);
        if (ix < m.d.n[0] && iy < m.d.n[1] && iz < m.d.n[2]) node_update<T>(m, E, S, ix, iy, iz);
        __syncthreads();   // block-wide visibility of the global stores of this colour
    }
}

// ---- y-marching variant of the tile-fused sweep ---------------------------------
// ncu (r1): the tile kernel above is bound by the L1/LSU pipe (60 % of peak
// wavefronts, 53 % of the warp stalls are waits on the load batch of a colour
// phase): 50 loads per node, 42 of them 16-byte loads with stride 2 between
// lanes.  Marching along y inside a tile removes 19 of them.  A tile is relaxed
// by ONE warp; lane = one column (ix, iz) of the tile, columns coloured by the
// parity of (ix, iz) (4 colours; columns of one colour are two nodes apart in x
// or z and never interact, whatever their y positions), and every lane walks its
// column in y, node after node -- a sequential Gauss-Seidel order along y.  What
// two consecutive nodes of a column share stays in registers (MarchCarry): the
// previous node's solved x- and z-edges are the outer edges of the next node's
// low faces, the outer y-edges / zeta / 1/h_y of the cell row between them are
// loaded once, and the y-edge between them is re-solved without being re-read.
// Per node: 16 E loads instead of 24, 5 + 5 source / diagonal loads instead of
// 6 + 6, 4 zeta instead of 8, 5 stores instead of 6; no block barriers (one
// __syncwarp per column colour).  Odd sweeps run the colours and the columns in
// the opposite direction, so that a pair of sweeps is a symmetric Gauss-Seidel.
template <typename T>
struct MarchCarry {
    T xs[2];           // solved ex of the previous node, by side flag (0: x+, 1: x-)
    T zs[2];           // solved ez of the previous node, by side flag (0: z+, 1: z-)
    T eyx[2];          // ey of the cell row behind at x-node ix+1 (flag 0) / ix-1 (flag 1)
    T eyz[2];          // ey of the cell row behind at z-node iz+1 (flag 0) / iz-1 (flag 1)
    T sY, dY;          // source and diagonal of the y-edge behind
    double zb[2][2];   // zeta of the cell row behind, [x-cell][z-cell]
    double rhb;        // 1/h_y of the cell row behind
};

// same arithmetic as Faces::quad, outer edges passed by value
template <typename T, int P, int Q, int SP, int SQ>
__device__ __forceinline__ void face_acc(const double (&z)[2][2][2], const double (&rh)[3][2],
                                         const T ep, const T eq, NodeSys<T>& n) {
    constexpr int W = 3 - P - Q;
    constexpr int i0x = P == 0 ? 1 - SP : Q == 0 ? 1 - SQ : 0;
    constexpr int i0y = P == 1 ? 1 - SP : Q == 1 ? 1 - SQ : 0;
    constexpr int i0z = P == 2 ? 1 - SP : Q == 2 ? 1 - SQ : 0;
    constexpr int i1x = W == 0 ? 1 : i0x, i1y = W == 1 ? 1 : i0y, i1z = W == 2 ? 1 : i0z;
    const double g = 0.5 * (z[i0x][i0y][i0z] + z[i1x][i1y][i1z]);
    const double rp = rh[P][1 - SP], rq = rh[Q][1 - SQ];
    const double al_p = SQ ? -rq : rq;
    const double al_q = SP ? rp : -rp;
    const T out = al_p * ep + al_q * eq;
    constexpr int kq = 2 * (Q - 1) + (1 - SQ);
    if (P == 0) {
        constexpr int j = 1 - SP;
        n.B[j][kq] += g * al_p * al_q;
        n.bX[j] += (g * al_p) * out;
    } else {
        constexpr int kp = 1 - SP;
        n.cyz[1 - SQ][1 - SP] += g * al_p * al_q;
        n.bT[kp] += (g * al_p) * out;
    }
    n.bT[kq] += (g * al_q) * out;
}

#define SS(r, c) s4[((r) * ((r) + 1)) / 2 + (c)]
// solve the assembled node system in place: n.bX, n.bT become the solution
// (same elimination as node_update: x-edges first, then LDL^T of the 4x4 Schur
// complement)
template <typename T>
__device__ __forceinline__ void node_eliminate(NodeSys<T>& n) {
    T rX[2], tX[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        rX[j] = rcp(n.dX[j]);
        tX[j] = rX[j] * n.bX[j];
    }
    T s4[10];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (l <= k) {
                T v = (n.B[0][k] * n.B[0][l]) * rX[0] + (n.B[1][k] * n.B[1][l]) * rX[1];
                T cc = zero_<T>();
                if (k == l) cc = n.dT[k];
                else if (k >= 2 && l < 2) add_real(cc, n.cyz[k - 2][l]);
                SS(k, l) = cc - v;
            }
        }
        n.bT[k] -= n.B[0][k] * tX[0] + n.B[1][k] * tX[1];
    }
    T dinv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        T v[4];
        T dj = SS(j, j);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < j) {
                v[k] = SS(j, k) * SS(k, k);
                dj -= SS(j, k) * v[k];
            }
        }
        SS(j, j) = dj;
        const T r = rcp(dj);
        dinv[j] = r;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i > j) {
                T t = SS(i, j);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < j) t -= SS(i, k) * v[k];
                SS(i, j) = t * r;
            }
        }
    }
#pragma unroll
    for (int j = 1; j < 4; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < j) n.bT[j] -= SS(j, k) * n.bT[k];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) n.bT[j] = n.bT[j] * dinv[j];
#pragma unroll
    for (int j = 2; j >= 0; --j) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k > j) n.bT[j] -= SS(k, j) * n.bT[k];
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        T acc = n.bX[j];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc -= n.B[j][k] * n.bT[k];
        n.bX[j] = rX[j] * acc;
    }
}
#undef SS

// ---- tile-fused sweep with the tile's E box in shared memory (r2) -----------------
// ncu on the kernel above (r1 / r2): 50 loads per node go to global memory, 42 of them 16-byte
// loads with stride 2 between lanes (L1 wavefronts 68 % of peak, half of every sector unused per
// colour phase), and every colour phase starts with a long-scoreboard wait on its load batch.
// Here a block first copies the tile's box of E values (the edges of its nodes plus one layer of
// outer edges: 6336 values = 101 KB for 64 x 4 x 4 nodes) into shared memory, the x index split by
// parity (even entries of a row first, then the odd ones) so that the stride-2 accesses of a
// colour phase become conflict-free unit-stride 16-byte accesses; the eight colour phases read
// and update E there (results also go straight to global memory, fire and forget).  What does
// not depend on E -- sources, diagonal, zeta, 1/h of the NEXT phase's node -- is loaded into
// registers while the current phase computes (NodeIn), so a phase no longer starts with a trip to
// DRAM.  Same node order and arithmetic as gs_point_tile_kernel.
template <typename T>
struct TileBox {
    static constexpr int HX = (TX + 2 + 1) / 2;          // half a row: rows hold up to TX + 2 entries
    static constexpr int ROW = 2 * HX;
    static constexpr int NY0 = TY + 2, NZ0 = TZ + 2;     // ex: y-nodes, z-nodes (x-cells along the row)
    static constexpr int NY1 = TY + 1, NZ1 = TZ + 2;     // ey: y-cells, z-nodes
    static constexpr int NY2 = TY + 2, NZ2 = TZ + 1;     // ez: y-nodes, z-cells
    static constexpr int OFF1 = ROW * NY0 * NZ0;
    static constexpr int OFF2 = OFF1 + ROW * NY1 * NZ1;
    static constexpr int SIZE = OFF2 + ROW * NY2 * NZ2;
    T* base;
    int o[3];                                            // global index of local index 0: (x0-1, y0-1, z0-1)
    __device__ __forceinline__ static int pos(int x) { return (x & 1) * HX + (x >> 1); }
    __device__ __forceinline__ static int ny(int c) { return c == 0 ? NY0 : c == 1 ? NY1 : NY2; }
    __device__ __forceinline__ static int off(int c) { return c == 0 ? 0 : c == 1 ? OFF1 : OFF2; }
    // element of component c at GLOBAL index q (cell index along c, node indices otherwise)
    __device__ __forceinline__ T& at(int c, const int* q) const {
        return base[off(c) + ROW * ((q[1] - o[1]) + ny(c) * (q[2] - o[2])) + pos(q[0] - o[0])];
    }
};

// E-independent inputs of one node
template <typename T>
struct NodeIn {
    T s[6], dg[6];            // source and diagonal of [x-, x+, y-, y+, z-, z+]
    double z[2][2][2];        // zeta of the 8 cells around the node
    double rh[3][2];
};

template <typename T>
__device__ __forceinline__ void tile_node_load(const Model<T>& m, const FieldView<const T>& S, int ix, int iy,
                                               int iz, NodeIn<T>& in) {
    const int nd[3] = {ix, iy, iz};
    const int64_t cs[3] = {1, m.d.n[0], (int64_t)m.d.n[0] * m.d.n[1]};
    const int64_t c0 = (ix - 1) + cs[1] * (iy - 1) + cs[2] * (iz - 1);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        in.rh[a][0] = ldg(m.rh[a] + nd[a] - 1);
        in.rh[a][1] = ldg(m.rh[a] + nd[a]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) in.z[i][j][k] = ldg(m.zeta + c0 + i + cs[1] * j + cs[2] * k);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            int q[3] = {ix, iy, iz};
            q[c] += sg - 1;
            const int64_t id = S.idx(c, q);
            in.s[2 * c + sg] = ldg(S.p[c] + id);
            in.dg[2 * c + sg] = ldg(m.diag + (S.p[c] - S.p[0]) + id);
        }
}

// outer edges of the face in plane (P, Q), quadrant (SP, SQ) from the box (see Faces::quad)
template <typename T, int P, int Q, int SP, int SQ>
__device__ __forceinline__ void tile_face(const TileBox<T>& box, int ix, int iy, int iz, const NodeIn<T>& in,
                                          NodeSys<T>& n) {
    int qo[3] = {ix, iy, iz};
    qo[P] -= SP;
    qo[Q] += SQ ? -1 : 1;
    int ro[3] = {ix, iy, iz};
    ro[Q] -= SQ;
    ro[P] += SP ? -1 : 1;
    face_acc<T, P, Q, SP, SQ>(in.z, in.rh, box.at(P, qo), box.at(Q, ro), n);
}

template <typename T>
__device__ __forceinline__ void tile_node_relax(const TileBox<T>& box, const FieldView<T>& E, int ix, int iy,
                                                int iz, const NodeIn<T>& in) {
    NodeSys<T> n;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) n.B[j][k] = 0.0;
    n.cyz[0][0] = n.cyz[0][1] = n.cyz[1][0] = n.cyz[1][1] = 0.0;
#pragma unroll
    for (int sg = 0; sg < 2; ++sg) {
        n.dX[sg] = in.dg[sg];
        n.bX[sg] = in.s[sg];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        n.dT[k] = in.dg[2 + k];
        n.bT[k] = in.s[2 + k];
    }
    // the three coordinate planes, four quadrants each, in the order of node_update
    tile_face<T, 0, 1, 0, 0>(box, ix, iy, iz, in, n);
    tile_face<T, 0, 1, 0, 1>(box, ix, iy, iz, in, n);
    tile_face<T, 0, 1, 1, 0>(box, ix, iy, iz, in, n);
    tile_face<T, 0, 1, 1, 1>(box, ix, iy, iz, in, n);
    tile_face<T, 0, 2, 0, 0>(box, ix, iy, iz, in, n);
    tile_face<T, 0, 2, 0, 1>(box, ix, iy, iz, in, n);
    tile_face<T, 0, 2, 1, 0>(box, ix, iy, iz, in, n);
    tile_face<T, 0, 2, 1, 1>(box, ix, iy, iz, in, n);
    tile_face<T, 1, 2, 0, 0>(box, ix, iy, iz, in, n);
    tile_face<T, 1, 2, 0, 1>(box, ix, iy, iz, in, n);
    tile_face<T, 1, 2, 1, 0>(box, ix, iy, iz, in, n);
    tile_face<T, 1, 2, 1, 1>(box, ix, iy, iz, in, n);
    node_eliminate<T>(n);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            int q[3] = {ix, iy, iz};
            q[c] += sg - 1;
            const T v = c == 0 ? n.bX[sg] : n.bT[2 * (c - 1) + sg];
            box.at(c, q) = v;
            E.p[c][E.idx(c, q)] = v;
        }
}

// Measured (r2, 256^3, B200; profiles/r2_notes.md): correct (same parity tests) but SLOWER than
// gs_point_tile_kernel -- 0.35 ms instead of 0.159 ms per tile-colour launch.  The 101 KB box
// allows two blocks = 8 warps per SM (12 before), the prefetched inputs push the kernel to 250
// registers, and the generic shared-memory addressing doubles the instruction count (112 M warp
// instructions per launch against 56 M): 6.7 cycles per issued instruction x 2 warps per
// scheduler loses against 8.6 x 3.  DRAM traffic is unchanged (0.50 GB read + 0.10 GB written
// per launch): the tile's halo re-reads were L2 hits already.  Kept as an option
// (-DEMG_PT_SMEM=1), off by default.
#ifndef EMG_PT_SMEM
#define EMG_PT_SMEM 0             // 1: tile kernel with the E box in shared memory, 0: all loads global
#endif
#ifndef EMG_SMEM_MINB
#define EMG_SMEM_MINB 2
#endif

template <typename T>
__global__ void __launch_bounds__(TILE_THREADS, EMG_SMEM_MINB)
gs_point_tile_smem_kernel(Model<T> m, T* e, const T* s, int tcx, int tcy, int tcz, int back) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    using B = TileBox<T>;
    B box;
    box.base = reinterpret_cast<T*>(tile_smem);
    const int x0 = 1 + (2 * blockIdx.x + tcx) * TX;
    const int y0 = 1 + (2 * blockIdx.y + tcy) * TY;
    const int z0 = 1 + (2 * blockIdx.z + tcz) * TZ;
    box.o[0] = x0 - 1; box.o[1] = y0 - 1; box.o[2] = z0 - 1;
    const int t = threadIdx.x;
    const int i = t % (TX / 2), j = (t / (TX / 2)) % (TY / 2), k = t / ((TX / 2) * (TY / 2));
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    const int nx = m.d.n[0], ny = m.d.n[1], nz = m.d.n[2];

    auto node_of = [&](int cc, int& ix, int& iy, int& iz) {
        const int c = back ? 7 - cc : cc;
        ix = x0 + 2 * i + (c & 1);
        iy = y0 + 2 * j + ((c >> 1) & 1);
        iz = z0 + 2 * k + ((c >> 2) & 1);
        return ix < nx && iy < ny && iz < nz;
    };
    // inputs of the first phase: in flight while the box is copied
    NodeIn<T> cur;
    int ix, iy, iz;
    bool valid = node_of(0, ix, iy, iz);
    if (valid) tile_node_load<T>(m, S, ix, iy, iz, cur);

    // ---- the tile's box of E: asynchronous element copies (LDGSTS), all in flight at once ----
    {
        constexpr int R0 = B::NY0 * B::NZ0, R1 = B::NY1 * B::NZ1, R2 = B::NY2 * B::NZ2;
        constexpr int LEN = TX + 2;                      // entries per row (ex uses TX + 1 of them)
        for (int el = t; el < (R0 + R1 + R2) * LEN; el += TILE_THREADS) {
            const int row = el / LEN, lx = el - row * LEN;
            const int c = row < R0 ? 0 : row < R0 + R1 ? 1 : 2;
            const int rr = row - (c == 0 ? 0 : c == 1 ? R0 : R0 + R1);
            const int nyc = B::ny(c);
            const int lz = rr / nyc, ly = rr - lz * nyc;
            const int gx = box.o[0] + lx, gy = box.o[1] + ly, gz = box.o[2] + lz;
            // extents of component c: cells along c, nodes otherwise
            const bool ok = gx <= (c == 0 ? nx - 1 : nx) && gy <= (c == 1 ? ny - 1 : ny) &&
                            gz <= (c == 2 ? nz - 1 : nz) && !(c == 0 && lx > TX);
            if (ok)
                cp_async_elem<(int)sizeof(T)>(box.base + B::off(c) + B::ROW * (ly + nyc * lz) + B::pos(lx),
                                              E.p[c] + E.idx(c, gx, gy, gz));
        }
        cp_async_wait_all();
    }
    __syncthreads();

    for (int cc = 0; cc < 8; ++cc) {
        NodeIn<T> nxt;
        int jx = 0, jy = 0, jz = 0;
        bool nvalid = false;
        if (cc < 7) {
            nvalid = node_of(cc + 1, jx, jy, jz);
            if (nvalid) tile_node_load<T>(m, S, jx, jy, jz, nxt);
        }
        if (valid) tile_node_relax<T>(box, E, ix, iy, iz, cur);
        __syncthreads();
        cur = nxt;
        valid = nvalid; ix = jx; iy = jy; iz = jz;
    }
}

// One column (ix, iz), nodes iy = y_first, y_first + DIR, ... (count nodes).
template <typename T, int DIR>
__device__ __forceinline__ void march_column(const Model<T>& m, const FieldView<T>& E,
                                             const FieldView<const T>& S, int ix, int iz,
                                             int y_first, int count) {
    constexpr int JA = DIR > 0 ? 1 : 0;     // index of the cell row / y-edge ahead (0: iy-1, 1: iy)
    constexpr int JB = 1 - JA;
    constexpr int SA = 1 - JA;              // side flag of the faces ahead (0: plus side)
    constexpr int SB = 1 - SA;
    const int64_t cs1 = m.d.n[0], cs2 = (int64_t)m.d.n[0] * m.d.n[1];
    const T* const dg = m.diag;             // field layout, relative to the x-component
    const int64_t dof[3] = {0, (int64_t)(S.p[1] - S.p[0]), (int64_t)(S.p[2] - S.p[0])};

    double rhx[2], rhz[2];
    rhx[0] = ldg(m.rh[0] + ix - 1); rhx[1] = ldg(m.rh[0] + ix);
    rhz[0] = ldg(m.rh[2] + iz - 1); rhz[1] = ldg(m.rh[2] + iz);

    // prologue: everything the first node needs from the row / cell row behind it
    MarchCarry<T> cr;
    {
        const int iy = y_first, yb = iy - DIR, cb = iy - 1 + JB;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            cr.xs[f] = E.p[0][E.idx(0, ix - f, yb, iz)];
            cr.zs[f] = E.p[2][E.idx(2, ix, yb, iz - f)];
            cr.eyx[f] = E.p[1][E.idx(1, ix + (f ? -1 : 1), cb, iz)];
            cr.eyz[f] = E.p[1][E.idx(1, ix, cb, iz + (f ? -1 : 1))];
        }
        const int64_t idb = S.idx(1, ix, cb, iz);
        cr.sY = ldg(S.p[1] + idb);
        cr.dY = ldg(dg + dof[1] + idb);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                cr.zb[i][k] = ldg(m.zeta + (ix - 1 + i) + cs1 * cb + cs2 * (iz - 1 + k));
        cr.rhb = ldg(m.rh[1] + cb);
    }

    for (int step = 0; step < count; ++step) {
        const int iy = y_first + DIR * step;
        const int ya = iy + DIR, ca = iy - 1 + JA, cb = iy - 1 + JB;
        const bool last = step == count - 1;

        // ---- loads -------------------------------------------------------------------
        T exA[2], eyxA[2], eyzA[2], ezA[2], exZ[2][2], ezX[2][2];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            exA[f] = E.p[0][E.idx(0, ix - f, ya, iz)];
            eyxA[f] = E.p[1][E.idx(1, ix + (f ? -1 : 1), ca, iz)];
            eyzA[f] = E.p[1][E.idx(1, ix, ca, iz + (f ? -1 : 1))];
            ezA[f] = E.p[2][E.idx(2, ix, ya, iz - f)];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                exZ[f][q] = E.p[0][E.idx(0, ix - f, iy, iz + (q ? -1 : 1))];
                ezX[f][q] = E.p[2][E.idx(2, ix + (f ? -1 : 1), iy, iz - q)];
            }
        }
        const int64_t idx0 = S.idx(0, ix - 1, iy, iz), idya = S.idx(1, ix, ca, iz),
                      idz0 = S.idx(2, ix, iy, iz - 1);
        const int64_t zst = S.s2[2];
        NodeSys<T> n;
        n.bX[0] = ldg(S.p[0] + idx0);      n.dX[0] = ldg(dg + idx0);
        n.bX[1] = ldg(S.p[0] + idx0 + 1);  n.dX[1] = ldg(dg + idx0 + 1);
        const T sYa = ldg(S.p[1] + idya), dYa = ldg(dg + dof[1] + idya);
        n.bT[JA] = sYa;                    n.dT[JA] = dYa;
        n.bT[JB] = cr.sY;                  n.dT[JB] = cr.dY;
        n.bT[2] = ldg(S.p[2] + idz0);        n.dT[2] = ldg(dg + dof[2] + idz0);
        n.bT[3] = ldg(S.p[2] + idz0 + zst);  n.dT[3] = ldg(dg + dof[2] + idz0 + zst);
        double za[2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                za[i][k] = ldg(m.zeta + (ix - 1 + i) + cs1 * ca + cs2 * (iz - 1 + k));
        const double rha = ldg(m.rh[1] + ca);

        // ---- assemble ----------------------------------------------------------------
        double z[2][2][2], rh[3][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                z[i][JA][k] = za[i][k];
                z[i][JB][k] = cr.zb[i][k];
            }
        rh[0][0] = rhx[0]; rh[0][1] = rhx[1];
        rh[2][0] = rhz[0]; rh[2][1] = rhz[1];
        rh[1][JA] = rha;   rh[1][JB] = cr.rhb;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) n.B[j][k] = 0.0;
        n.cyz[0][0] = n.cyz[0][1] = n.cyz[1][0] = n.cyz[1][1] = 0.0;

        // plane (x, y): faces ahead use loaded edges, faces behind the carried ones
        face_acc<T, 0, 1, 0, SA>(z, rh, exA[0], eyxA[0], n);
        face_acc<T, 0, 1, 1, SA>(z, rh, exA[1], eyxA[1], n);
        face_acc<T, 0, 1, 0, SB>(z, rh, cr.xs[0], cr.eyx[0], n);
        face_acc<T, 0, 1, 1, SB>(z, rh, cr.xs[1], cr.eyx[1], n);
        // plane (x, z)
        face_acc<T, 0, 2, 0, 0>(z, rh, exZ[0][0], ezX[0][0], n);
        face_acc<T, 0, 2, 0, 1>(z, rh, exZ[0][1], ezX[0][1], n);
        face_acc<T, 0, 2, 1, 0>(z, rh, exZ[1][0], ezX[1][0], n);
        face_acc<T, 0, 2, 1, 1>(z, rh, exZ[1][1], ezX[1][1], n);
        // plane (y, z)
        face_acc<T, 1, 2, SA, 0>(z, rh, eyzA[0], ezA[0], n);
        face_acc<T, 1, 2, SA, 1>(z, rh, eyzA[1], ezA[1], n);
        face_acc<T, 1, 2, SB, 0>(z, rh, cr.eyz[0], cr.zs[0], n);
        face_acc<T, 1, 2, SB, 1>(z, rh, cr.eyz[1], cr.zs[1], n);

        node_eliminate<T>(n);

        // ---- write back, carry ---------------------------------------------------------
        E.p[0][idx0] = n.bX[0];
        E.p[0][idx0 + 1] = n.bX[1];
        E.p[2][idz0] = n.bT[2];
        E.p[2][idz0 + zst] = n.bT[3];
        E.p[1][E.idx(1, ix, cb, iz)] = n.bT[JB];
        if (last) E.p[1][idya] = n.bT[JA];     // otherwise re-solved by the next node
        cr.xs[0] = n.bX[1]; cr.xs[1] = n.bX[0];
        cr.zs[0] = n.bT[3]; cr.zs[1] = n.bT[2];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            cr.eyx[f] = eyxA[f];
            cr.eyz[f] = eyzA[f];
        }
        cr.sY = sYa; cr.dY = dYa;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 2; ++k) cr.zb[i][k] = za[i][k];
        cr.rhb = rha;
    }
}

// Measured (r1, 256^3, B200; profiles/r1_notes.md): correct (parity-tested against
// the oracle in the same order) but SLOWER than the 8-colour tile kernel, 0.298 ms
// instead of 0.181 ms per tile-colour launch.  One warp per tile puts 8 tiles per
// SM in flight, 1184 x 252 KB = 300 MB of tile footprints against 126 MB of L2: the
// L2 hit rate drops from 47 % to 19 %, every step of a column waits for DRAM
// (long-scoreboard stall 8.1 of 11 cycles per issued instruction) and nothing is
// prefetched across steps.  Kept as an option (-DEMG_PT_MARCH=1), off by default.
#ifndef EMG_PT_MARCH
#define EMG_PT_MARCH 0            // 1: y-marching tile kernel, 0: 8-colour tile kernel
#endif
#ifndef EMG_MARCH_MINB
#define EMG_MARCH_MINB 8
#endif
static_assert(!EMG_PT_MARCH || (TX / 2) * (TZ / 2) == 32,
              "the marching kernel maps one tile to one warp");

template <typename T>
__global__ void __launch_bounds__(32, EMG_MARCH_MINB)
gs_point_march_kernel(Model<T> m, T* e, const T* s, int tcx, int tcy, int tcz, int back) {
    const int x0 = 1 + (2 * blockIdx.x + tcx) * TX;
    const int y0 = 1 + (2 * blockIdx.y + tcy) * TY;
    const int z0 = 1 + (2 * blockIdx.z + tcz) * TZ;
    const int i = threadIdx.x % (TX / 2), k = threadIdx.x / (TX / 2);
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    const int y1 = min(y0 + TY, m.d.n[1]) - 1;           // last node of the tile in y
    const int count = y1 - y0 + 1;
    for (int cc = 0; cc < 4; ++cc) {
        const int c = back ? 3 - cc : cc;
        const int ix = x0 + 2 * i + (c & 1), iz = z0 + 2 * k + (c >> 1);
        if (ix < m.d.n[0] && iz < m.d.n[2] && count > 0) {
            if (back) march_column<T, -1>(m, E, S, ix, iz, y1, count);
            else march_column<T, 1>(m, E, S, ix, iz, y0, count);
        }
        __syncwarp();      // the next column colour reads what this one wrote
    }
}

// one hyperplane ix + 2 iy + 3 iz = t of the lexicographic sweep
template <typename T>
__global__ void __launch_bounds__(128)
gs_point_front_kernel(Model<T> m, T* e, const T* s, int t) {
    const int iy = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int iz = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (iy >= m.d.n[1] || iz >= m.d.n[2]) return;
    const int ix = t - 2 * iy - 3 * iz;
    if (ix < 1 || ix >= m.d.n[0]) return;
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    node_update<T>(m, E, S, ix, iy, iz);
}

// whole smoothing call (all sweeps, all parity classes or all hyperplanes) in
// one block, for grids small enough that launch latency would dominate.
template <typename T>
__global__ void __launch_bounds__(256)
gs_point_small_kernel(Model<T> m, T* e, const T* s, int nu, int order) {
    FieldView<T> E(e, m.d);
    FieldView<const T> S(s, m.d);
    const int nx = m.d.n[0], ny = m.d.n[1], nz = m.d.n[2];
    const int nint = (nx - 1) * (ny - 1) * (nz - 1);
    bool back = (order >> 8) & 1;   // bits 8+ of `order`: sweeps already done (phase)
    order &= 0xff;
    for (int sw = 0; sw < nu; ++sw) {
        back = !back;                       // first sweep runs descending (core.py:301,311)
        if (order == ORDER_LEX) {
            const int tmin = 6, tmax = (nx - 1) + 2 * (ny - 1) + 3 * (nz - 1);
            const int nyz = (ny - 1) * (nz - 1);
            for (int tt = tmin; tt <= tmax; ++tt) {
                const int t = back ? tmax + tmin - tt : tt;
                for (int q = threadIdx.x; q < nyz; q += blockDim.x) {
                    const int iy = 1 + q % (ny - 1), iz = 1 + q / (ny - 1);
                    const int ix = t - 2 * iy - 3 * iz;
                    if (ix >= 1 && ix < nx) node_update<T>(m, E, S, ix, iy, iz);
                }
                __syncthreads();
            }
        } else {
            for (int cc = 0; cc < 8; ++cc) {
                const int c = back ? 7 - cc : cc;
                if (sw > 0 && cc == 0) continue;   // idempotent repeat, see launch_gs_point
                const int px = c & 1, py = (c >> 1) & 1, pz = (c >> 2) & 1;
                for (int q = threadIdx.x; q < nint; q += blockDim.x) {
                    const int ix = 1 + q % (nx - 1);
                    const int iy = 1 + (q / (nx - 1)) % (ny - 1);
                    const int iz = 1 + q / ((nx - 1) * (ny - 1));
                    if (((ix - 1) & 1) == px && ((iy - 1) & 1) == py && ((iz - 1) & 1) == pz)
                        node_update<T>(m, E, S, ix, iy, iz);
                }
                __syncthreads();
            }
        }
    }
}

template <typename T>
void launch_gs_point(const Model<T>& m, T* e, const T* s, int nu, int order, cudaStream_t st) {
    const int nx = m.d.n[0], ny = m.d.n[1], nz = m.d.n[2];
    if (nx < 2 || ny < 2 || nz < 2) return;
    const int64_t nint = (int64_t)(nx - 1) * (ny - 1) * (nz - 1);
    if (nint <= SMALL_GRID_NODES) {
        int threads = 32;
        while (threads < 256 && threads < nint) threads <<= 1;
        ++g_launch_count; gs_point_small_kernel<T><<<1, threads, 0, st>>>(m, e, s, nu, order);
        return;
    }
    bool back = (order >> 8) & 1;   // bit 8 of `order`: sweeps already done (phase)
    // bits 16-17: multicolour order, z-half of a sweep (multi-GPU z-slabs): 1 = only the colour
    // classes of even z-parity, 2 = only those of odd z-parity, 0 = all
    const int zsel = (order >> 16) & 3;
    order &= 0xff;
    for (int sw = 0; sw < nu; ++sw) {
        back = !back;
        if (order == ORDER_LEX) {
            const int tmin = 6, tmax = (nx - 1) + 2 * (ny - 1) + 3 * (nz - 1);
            dim3 b(32, 4);
            dim3 g((ny - 1 + b.x - 1) / b.x, (nz - 1 + b.y - 1) / b.y);
            for (int tt = tmin; tt <= tmax; ++tt) {
                const int t = back ? tmax + tmin - tt : tt;
                ++g_launch_count; gs_point_front_kernel<T><<<g, b, 0, st>>>(m, e, s, t);
            }
        } else if (nint > TILE_MIN_NODES) {
            const int ntx = (nx - 1 + TX - 1) / TX, nty = (ny - 1 + TY - 1) / TY,
                      ntz = (nz - 1 + TZ - 1) / TZ;
            for (int cc = 0; cc < 8; ++cc) {
                const int c = back ? 7 - cc : cc;
                const int tcx = c & 1, tcy = (c >> 1) & 1, tcz = (c >> 2) & 1;
                if (zsel && tcz != zsel - 1) continue;
                dim3 g((ntx - tcx + 1) / 2, (nty - tcy + 1) / 2, (ntz - tcz + 1) / 2);
                if (g.x == 0 || g.y == 0 || g.z == 0) continue;
#if EMG_PT_MARCH
                ++g_launch_count; gs_point_march_kernel<T><<<g, 32, 0, st>>>(m, e, s, tcx, tcy, tcz, back ? 1 : 0);
#elif EMG_PT_SMEM
                {
                    constexpr int smem = TileBox<T>::SIZE * (int)sizeof(T);
                    static bool attr_set = false;
                    if (!attr_set) {
                        cudaFuncSetAttribute(gs_point_tile_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                        cudaFuncSetAttribute(gs_point_tile_smem_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             cudaSharedmemCarveoutMaxShared);
                        attr_set = true;
                        if (getenv("EMG3D_B200_DEBUG")) {
                            int nb = 0;
                            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gs_point_tile_smem_kernel<T>,
                                                                          TILE_THREADS, smem);
                            fprintf(stderr, "gs_point_tile_smem_kernel: %d bytes of shared memory, %d blocks per SM\n",
                                    smem, nb);
                        }
                    }
                    ++g_launch_count;
                    gs_point_tile_smem_kernel<T><<<g, TILE_THREADS, smem, st>>>(m, e, s, tcx, tcy, tcz, back ? 1 : 0);
                }
#else
                ++g_launch_count; gs_point_tile_kernel<T><<<g, TILE_THREADS, 0, st>>>(m, e, s, tcx, tcy, tcz, back ? 1 : 0);
#endif
            }
        } else {
            for (int cc = 0; cc < 8; ++cc) {
                const int c = back ? 7 - cc : cc;
                // the first colour of a sweep is the last colour of the previous sweep
                // (opposite order); nodes of one colour do not interact and nothing
                // changed in between, so relaxing them again reproduces the same values
                if (sw > 0 && cc == 0) continue;
                if (zsel && ((c >> 2) & 1) != zsel - 1) continue;
                const int fx = 1 + (c & 1), fy = 1 + ((c >> 1) & 1), fz = 1 + (((c >> 2) & 1) ^ (m.d.zflip & 1));
                const int cx = (nx - fx + 1) / 2, cy = (ny - fy + 1) / 2, cz = (nz - fz + 1) / 2;
                if (cx <= 0 || cy <= 0 || cz <= 0) continue;
                dim3 b(32, 4, 1);
                dim3 g((cx + b.x - 1) / b.x, (cy + b.y - 1) / b.y, cz);
                ++g_launch_count; gs_point_color_kernel<T><<<g, b, 0, st>>>(m, e, s, fx, fy, fz, cx, cy, cz);
            }
        }
    }
}

int point_tile_schedule() { return EMG_PT_MARCH ? 1 : 0; }
void point_tile_shape(int* t) { t[0] = TX; t[1] = TY; t[2] = TZ; }

template void launch_gs_point<double>(const Model<double>&, double*, const double*, int, int, cudaStream_t);
template void launch_gs_point<cplx>(const Model<cplx>&, cplx*, const cplx*, int, int, cudaStream_t);

}  // namespace emg
