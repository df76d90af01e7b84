// Internal launch interface between the C-ABI (api.cu) and the kernel files.
#pragma once
#include "common.cuh"

namespace emg {

enum { ORDER_LEX = 0, ORDER_COLOR = 1 };

extern long long g_launch_count;   // kernel launches issued (host-side counter)

// grids with at most this many interior nodes are smoothed by one thread block in
// a single launch (8^3 and below; a colour phase of one block costs ~4 us, and
// beyond ~500 nodes one launch per colour with more blocks is faster)
constexpr int64_t SMALL_GRID_NODES = 512;
// point smoother, multicolour order: grids with more interior nodes than this
// (working set beyond the 126 MB L2) use the tile-fused schedule
constexpr int64_t TILE_MIN_NODES = 300000;

// r = s - A e (r may be null, r may alias s), or r = A e when apply_only (s unused);
// optional ||r||^2 into norm2_out[0]
template <typename T>
void launch_residual(const Model<T>& m, const T* s, const T* e, T* r, double* norm2_out,
                     double* scratch, int apply_only, cudaStream_t st);
int64_t residual_scratch_doubles(const Dims& d);

template <typename T>
void launch_gs_point(const Model<T>& m, T* e, const T* s, int nu, int order, cudaStream_t st);
// order inside a tile of the tile-fused schedule: 0 = 8 node colours, 1 = 4 column
// colours with a sequential march along y (tests build the oracle's sequence from it)
int point_tile_schedule();
void point_tile_shape(int* txyz);      // tile shape in nodes
// diagonal of A per edge (field layout), read by the point smoother through m.diag
template <typename T>
void launch_edge_diag(const Model<T>& m, T* diag, cudaStream_t st);

// line smoothers: factor once per (level, direction), then sweep
int64_t line_factor_elems(const Dims& d, int dir);     // number of T elements
template <typename T>
void launch_line_factor(const Model<T>& m, int dir, T* fac, const T* xin, T* xout, cudaStream_t st);
int64_t line_chain_elems(const Dims& d, int dir);
// `fac`: factors in the one-thread-per-line layout (lexicographic order, small grids, and the
// multicolour order wherever `fac2` is null); `fac2`: cached data of the segment-parallel
// kernels (gs_line_seg.cu; multicolour order on long lines) or null
template <typename T>
void launch_gs_line(const Model<T>& m, int dir, const T* fac, const T* fac2, T* e, const T* s, int nu,
                    int order, cudaStream_t st);
// segment-parallel line smoother (one warp per line, factors staged on chip by TMA)
int line_seg_mask(int mask);                           // set (mask >= 0) / query the direction mask
int line_seg_qp(const Dims& d, int dir);               // lanes per line; 0 = kernel not used for this shape
int64_t line_seg_elems(const Dims& d, int dir);        // number of T elements of its cached data
template <typename T>
void launch_line_seg_factor(const Model<T>& m, int dir, T* fac2, cudaStream_t st);
template <typename T>
void launch_gs_line_seg_color(const Model<T>& m, int dir, const T* fac2, T* e, const T* s, int c,
                              cudaStream_t st);

// transfer operators; cflag[a] = 1 if axis a is coarsened
template <typename T>
void launch_restrict(const Dims& fine, const int* cflag, const T* r, T* cr, const double* const* wl,
                     const double* const* w0, const double* const* wr, cudaStream_t st);
template <typename T>
void launch_prolong(const Dims& fine, const int* cflag, T* e, const T* ce, const int* const* lo,
                    const double* const* frac, cudaStream_t st);
template <typename T>
void launch_restrict_cells(const Dims& fine, const int* cflag, const T* p, T* cp, cudaStream_t st);

// banded LDL^T solve of one system (interface of core.solve)
template <typename T>
void launch_band_solve(int n, T* a, T* b, cudaStream_t st);

// eta / zeta from property arrays (models.VolumeModel)
template <typename T>
void launch_volume_model(const Dims& d, const double* hx, const double* hy, const double* hz,
                         double cr, double ci, double sr, double si, double eps0, int map,
                         const double* px, const double* py, const double* pz, const double* mu,
                         const double* eps, T* ex, T* ey, T* ez, double* zeta, cudaStream_t st);

// H on the faces from E on the edges: curl E times the two-cell average of a per-cell
// factor (Z = double: zeta of the level, or Z = T: the caller's array) times `scale`,
// divided by the dual-cell measure (fields._edge_curl_factor); hf is zeroed first
int64_t n_faces(const Dims& d);
template <typename T, typename Z>
void launch_edge_curl(const Dims& d, const T* e, T* hf, const double* hx, const double* hy,
                      const double* hz, const Z* zeta, T scale, cudaStream_t st);

// interpolation next to the solve (interp.cu)
void launch_volume_average(const double* values, int nx, int ny, double* out, int mx, int my, int mz,
                           const double* const* w, const int* const* iin, const int* const* start,
                           const double* const* hnew, int log_scale, int add, cudaStream_t st);
template <typename T>
void launch_edges_to_vol(const Dims& d, const T* e, const double* hx, const double* hy, const double* hz, T* out,
                         cudaStream_t st);
void launch_gradient_field(const Dims& d, const cplx* e, const cplx* b, cplx smu0, const double* hx,
                           const double* hy, const double* hz, double* out, cudaStream_t st);
template <typename T>
void launch_spline_filter(T* data, int n0, int n1, int n2, int reflect, cudaStream_t st);
template <typename T>
void launch_copy_box(const T* src, int n0, int n1, const int* lo, const int* m, T* dst, cudaStream_t st);
template <typename T>
void launch_pad_edge(const T* src, int n0, int n1, int n2, int npad, T* dst, cudaStream_t st);
template <typename T>
void launch_spline_eval(const T* coef, int n0, int n1, int n2, int npad, int mode, T cval, const double* cx,
                        const double* cy, const double* cz, int64_t npts, int tensor, int m0, int m1, T scale,
                        int accumulate, T* out, cudaStream_t st);
template <typename T>
void launch_linear_eval(const T* data, int n0, int n1, int n2, T fill, const double* cx, const double* cy,
                        const double* cz, int64_t npts, int tensor, int m0, int m1, T scale, int accumulate, T* out,
                        cudaStream_t st);

// vector helpers
template <typename T>
void launch_pec_zero(const Dims& d, T* e, cudaStream_t st);
template <typename T>
void launch_dot(int64_t n, const T* x, const T* y, int conj_x, double* out2, double* scratch,
                cudaStream_t st);
int64_t dot_scratch_doubles(int64_t n);
template <typename T>
void launch_axpby(int64_t n, T a, const T* x, T b, T* y, cudaStream_t st);   // y = a x + b y

}  // namespace emg
