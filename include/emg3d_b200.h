/* emg3d_b200 -- C ABI of the B200-native multigrid hot path.
 *
 * Drop-in boundary for the module-level calls that emg3d/solver.py makes into
 * emg3d/core.py (the reference has no plugin registry).
 * Plain C types only: ints, sizes, raw pointers.  Unless stated otherwise all
 * data pointers are DEVICE pointers obtained from emg3d_b200_malloc.
 *
 * Layout contract (emg3d/fields.py:116, 201-259; emg3d/models.py:662-691):
 *   field  = [fx | fy | fz], fx (nx, ny+1, nz+1), fy (nx+1, ny, nz+1),
 *            fz (nx+1, ny+1, nz), x fastest; complex128 (cplx = 1, interleaved
 *            re/im) or float64 (cplx = 0, Laplace domain);
 *   eta_x, eta_y, eta_z : (nx, ny, nz), x fastest, same dtype as the fields,
 *            may alias each other (isotropic / VTI / HTI);
 *   zeta, hx, hy, hz    : float64.
 *
 * Every function returns 0 on success and a non-zero code on failure; the
 * message is available from emg3d_b200_last_error().  One device per process
 * (one process per GPU); all work is issued on one library-owned CUDA stream.
 */
#ifndef EMG3D_B200_H
#define EMG3D_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct emg3d_b200_level emg3d_b200_level;

enum { EMG3D_B200_ORDER_LEX = 0, EMG3D_B200_ORDER_COLOR = 1 };

/* ---- runtime ------------------------------------------------------------- */
int emg3d_b200_abi_version(void);
const char* emg3d_b200_last_error(void);
int emg3d_b200_device_count(int* count);
int emg3d_b200_init(int device);
int emg3d_b200_device_name(char* buf, int buflen);
int emg3d_b200_mem_info(size_t* free_bytes, size_t* total_bytes);
int emg3d_b200_sync(void);
/* number of kernel launches issued by the library so far */
int emg3d_b200_launch_count(long long* count);

/* ---- memory -------------------------------------------------------------- */
int emg3d_b200_malloc(void** dptr, size_t nbytes);
int emg3d_b200_free(void* dptr);
/* Short-lived work arrays: stream-ordered allocations from the device's memory pool (no device
 * synchronisation; not exportable to other processes -- arrays exchanged between GPUs come from
 * emg3d_b200_malloc). */
int emg3d_b200_malloc_scratch(void** dptr, size_t nbytes);
int emg3d_b200_free_scratch(void* dptr);
int emg3d_b200_memset(void* dptr, int byte, size_t nbytes);
int emg3d_b200_h2d(void* dst_dev, const void* src_host, size_t nbytes);
int emg3d_b200_d2h(void* dst_host, const void* src_dev, size_t nbytes);
int emg3d_b200_d2d(void* dst_dev, const void* src_dev, size_t nbytes);
/* Upload of a mostly-zero host array of n elements of elsize 8 or 16 bytes (source
 * fields, emg3d/fields.py:386-519): the host array is scanned and, if fewer than
 * ~1.5 % of the elements have a non-zero bit pattern, only (index, value) pairs
 * cross PCIe and are scattered into the cleared device array; otherwise a plain
 * copy.  The device array is bit-identical to emg3d_b200_h2d either way.       */
int emg3d_b200_h2d_sparse(void* dst_dev, const void* src_host, size_t n_elems, int elsize,
                          int* used_sparse);
/* Device array = `fill_value` (one element, host) everywhere except idx[k] <- val[k], k < count
 * (host arrays; indices must be distinct): a dipole / wire source field assembled from its few
 * non-zero edges (emg3d/fields.py:386-519) without a dense host array. */
int emg3d_b200_fill_scatter(void* dst_dev, size_t n_elems, int elsize, const void* fill_value,
                            const long long* idx, const void* val, size_t count);
int emg3d_b200_host_alloc(void** hptr, size_t nbytes);   /* pinned host memory */
int emg3d_b200_host_free(void* hptr);

/* ---- timing and CUDA graphs (all on the library stream) ------------------- */
int emg3d_b200_event_create(void** ev);
int emg3d_b200_event_record(void* ev);
int emg3d_b200_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms); /* syncs on stop */
int emg3d_b200_event_destroy(void* ev);
int emg3d_b200_graph_begin(void);
int emg3d_b200_graph_end(void** graph_exec);
int emg3d_b200_graph_launch(void* graph_exec);
int emg3d_b200_graph_destroy(void* graph_exec);

/* ---- grid levels ----------------------------------------------------------
 * A level = one grid of the multigrid hierarchy: cell counts, widths (HOST
 * arrays, copied), the coefficient arrays (device, NOT owned) and cached line
 * factorisations (owned).  Replaces the (grid, VolumeModel) pair that
 * emg3d/solver.py passes around (solver.py:827-830, 1060-1062).               */
int emg3d_b200_level_create(emg3d_b200_level** out, int nx, int ny, int nz,
                            const double* hx_host, const double* hy_host,
                            const double* hz_host);
int emg3d_b200_level_destroy(emg3d_b200_level* lv);
int emg3d_b200_level_set_model(emg3d_b200_level* lv, int cplx, const void* eta_x,
                               const void* eta_y, const void* eta_z, const double* zeta);
/* z-window of a level that has its model set (multi-GPU slabs): a
 * view on cells [z0, z0 + nz) along z of the parent's arrays.  Field pointers
 * passed to kernels on the window are the PARENT's full arrays; the kernels
 * touch the window only, its first and last node plane acting as fixed boundary
 * data (halo planes).  Owns its cached factorisations; destroy before the parent. */
int emg3d_b200_level_window(emg3d_b200_level** out, const emg3d_b200_level* parent, int z0,
                            int nz);
/* Norms on this level (emg3d_b200_residual with norm2_dev) count only the edges a
 * multi-GPU rank owns: x/y-edges on node planes [plane0, plane1) of the level (or
 * window) and z-edges of the cell layers whose upper plane is in that range.
 * plane1 == 0 (default): all edges, i.e. the reference's np.linalg.norm
 * (emg3d/solver.py:1066). */
int emg3d_b200_level_set_owned(emg3d_b200_level* lv, int plane0, int plane1);
/* Multi-GPU z-slabs.  flip = 1 declares that local node plane 1 of this (window) level is an
 * even global node plane: the z-parity of the multicolour classes of the node-colour point
 * schedule and of the x- / y-line colours is then the global one (class bit cz relaxes local
 * planes of parity cz ^ flip), so that the "z-half" selection of emg3d_b200_gauss_seidel
 * (bits 16-17 of `order`) names the same global planes on every rank.                      */
int emg3d_b200_level_set_zflip(emg3d_b200_level* lv, int flip);
/* Multi-GPU z-slabs, z-line relaxation (emg3d/core.py:1071-1348 `gauss_seidel_z`): the z-lines
 * are cut by the slabs, every rank holds one piece of each line.  The block recurrence of the
 * line factorisation runs through the cuts: the first rank factorises its pieces (chain_in = NULL)
 * and hands the factors of their last blocks (chain_out, [line slot][10] elements of the level's
 * dtype, *n_elems in total) to the next rank, which passes them as chain_in, and so on.  Both
 * NULL: size query only.  Sweeps then run piecewise through emg3d_b200_gauss_seidel with the
 * phase / colour bits of `order` (bits 18-19: 1 = forward pass only, 2 = backward pass only;
 * bits 20-22: 1 + colour class; bits 23-26: batch index, bits 27-30: number of batches - 1, the
 * lines of the class cut into equal batches): forward passes rank after rank upwards, backward
 * passes downwards, the interface planes exchanged in between; with batches the ranks work as a
 * pipeline (emg3d_b200/parallel.py). */
int emg3d_b200_level_line_chain(emg3d_b200_level* lv, int ldir, const void* chain_in, void* chain_out,
                                size_t* n_elems);
/* Multicolour point schedule used for this level's shape: 0 = one block (<= 512 interior
 * nodes), 1 = one launch per node colour, 2 = tile-fused (tiles coloured by LOCAL tile index). */
int emg3d_b200_point_schedule_kind(const emg3d_b200_level* lv, int* kind);
/* bytes of the cached factorisation of line direction ldir (1, 2, 3) */
int emg3d_b200_level_factor_bytes(const emg3d_b200_level* lv, int ldir, size_t* nbytes);
int emg3d_b200_level_drop_factors(emg3d_b200_level* lv);
/* Link a coarse level to its parent.  cflag[a] = 1 if axis a is coarsened
 * (emg3d/solver.py:891-897).  weights[3*a + {0,1,2}] = wl, w0, wr of axis a as
 * returned by core.restrict_weights (core.py:2004-2076), HOST arrays of
 * length n_coarse_nodes(a), ignored (may be NULL) when axis a is not
 * coarsened.  lo[a], frac[a]: per FINE node of axis a the lower coarse node
 * and the fraction inside that coarse cell (the bilinear weights of
 * solver.py:1447-1473), HOST arrays of length n_fine_nodes(a).                */
int emg3d_b200_level_link(emg3d_b200_level* coarse, const emg3d_b200_level* fine,
                          const int* cflag, const double* const* weights,
                          const int* const* lo, const double* const* frac);

/* ---- kernels -------------------------------------------------------------- */
/* r -= A e.  Replaces core.amat_x (core.py:57-58; call sites solver.py:695-699,
 * 1060-1062).                                                                  */
int emg3d_b200_amat_x(emg3d_b200_level* lv, void* r, const void* e);
/* out = A e: the matvec of the Krylov wrapper (solver.py:686-702, where the
 * reference calls amat_x on a zero field and negates).                        */
int emg3d_b200_apply(emg3d_b200_level* lv, const void* e, void* out);
/* r = s - A e (r may be NULL) and, if norm2_dev != NULL, norm2_dev[0] =
 * ||r||_2^2 over ALL edges.  Replaces solver.residual (solver.py:1022-1070).  */
int emg3d_b200_residual(emg3d_b200_level* lv, const void* s, const void* e, void* r,
                        double* norm2_dev);
/* same, synchronous, returning ||r||_2 to the host */
int emg3d_b200_residual_norm(emg3d_b200_level* lv, const void* s, const void* e, void* r,
                             double* norm_host);
/* nu Gauss-Seidel sweeps, in place on e.  ldir 0 = point smoother
 * (core.gauss_seidel, core.py:210-212), 1/2/3 = x/y/z line relaxation
 * (core.gauss_seidel_x/_y/_z, core.py:506-508, 786-788, 1071-1073).
 * order: LEX = sequentially equivalent to the reference's lexicographic
 * sweeps; COLOR = multicolour ordering (8 colours point, 4 colours lines).
 * Higher bits of `order`: bit 8 = an odd number of sweeps precedes this call (sweeps
 * alternate direction); bits 16-17 (COLOR, point / x- / y-lines) = z-half of a sweep:
 * 1 = only the colour classes of even z-parity, 2 = only the odd ones, 0 = all.  A
 * z-slab solver runs the two halves with a halo exchange after each: planes on either
 * side of an interface are never relaxed in the same half, which makes the distributed
 * sweep a true multicolour Gauss-Seidel sweep (exact variant).      */
int emg3d_b200_gauss_seidel(emg3d_b200_level* lv, void* e, const void* s, int nu, int ldir,
                            int order);
/* Node order inside a tile of the point smoother's tile-fused multicolour schedule
 * (grids beyond the L2): 0 = the 8 parity classes one after the other, 1 = 4 column
 * colours (parity of ix, iz), every column relaxed sequentially along y
 * (csrc/gs_point.cu).  Tests use it to run the oracle in the same order. */
int emg3d_b200_point_tile_schedule(int* variant);
/* tile shape (nodes along x, y, z) of that schedule */
int emg3d_b200_point_tile_shape(int* txyz);
/* Line smoothers in multicolour order: bit a of the mask selects, for lines along axis a
 * (0 = x, 1 = y, 2 = z), the segment-parallel kernel (one warp per line, factors staged on
 * chip by TMA bulk copies, csrc/gs_line_seg.cu) on lines of 66 .. 257 cells instead of the
 * one-thread-per-line kernel (csrc/gs_line.cu).  Same colour sequence, same results to
 * rounding.  mask < 0 only queries.  Default 1 (x-lines), or the environment variable
 * EMG3D_B200_LINE_SEG.  Cached factorisations of existing levels are not converted: set the
 * mask before the first smoothing call or drop the factors.  *previous (may be NULL)
 * receives the mask in force before the call. */
int emg3d_b200_line_seg_mask(int mask, int* previous);
/* coarse_s = R r_fine.  Replaces core.restrict (core.py:1620-1621).            */
int emg3d_b200_restrict(emg3d_b200_level* coarse, const void* r_fine, void* s_coarse);
/* e_fine += P e_coarse on interior edges.  Replaces solver.prolongation
 * (solver.py:947-1019).                                                        */
int emg3d_b200_prolong(emg3d_b200_level* coarse, void* e_fine, const void* e_coarse);
/* coarse cell array = sum of the fine cells.  Replaces
 * solver._restrict_model_parameters (solver.py:1667-1718).                     */
int emg3d_b200_restrict_cells(emg3d_b200_level* coarse, int cplx, const void* p_fine,
                              void* p_coarse);
/* Volume-averaged coefficients from property arrays, on the device.  Replaces
 * models.VolumeModel.__init__ (emg3d/models.py:654-691):
 *   eta_a = c * V * (sigma_a + (s eps_0) * eps_r),  zeta = V / mu_r,
 * c = -s mu_0 (c_re, c_im), (s_re, s_im) = s * eps_0, V from the level's widths,
 * sigma_a = backward map of prop_a (map_code 0 Conductivity, 1 Resistivity,
 * 2 LgConductivity, 3 LnConductivity, 4 LgResistivity, 5 LnResistivity).
 * prop_y, prop_z, mu_r, eps_r may be NULL; eta_y / eta_z are written only when
 * the corresponding property is given.  All arrays float64 (nx, ny, nz).       */
int emg3d_b200_volume_model(emg3d_b200_level* lv, int cplx, double c_re, double c_im,
                            double s_re, double s_im, int map_code, const double* prop_x,
                            const double* prop_y, const double* prop_z, const double* mu_r,
                            const double* eps_r, void* eta_x, void* eta_y, void* eta_z,
                            double* zeta);
/* zero the tangential boundary edges (solver.py:350-355) */
int emg3d_b200_pec_zero(emg3d_b200_level* lv, void* e);

/* ---- vector helpers for the Krylov wrapper and the termination tests -------
 * n counts elements of the given dtype.  dot2_dev[0..1] = (re, im) of
 * sum conj?(x) y.  Scalars are (re, im) pairs; im ignored when cplx = 0.       */
int emg3d_b200_dot(int cplx, long long n, const void* x, const void* y, int conj_x,
                   double* dot2_dev);
int emg3d_b200_dot_host(int cplx, long long n, const void* x, const void* y, int conj_x,
                        double* dot2_host);
int emg3d_b200_axpby(int cplx, long long n, double a_re, double a_im, const void* x,
                     double b_re, double b_im, void* y);

/* ---- multi-GPU: z-slab decomposition (new functionality) ------
 * One process per GPU.  Rank 0 obtains a 128-byte NCCL unique id, distributes it
 * by any means (e.g. torch.distributed / a file), every rank calls comm_init.
 * comm_sendrecv posts n point-to-point transfers of raw bytes (device pointers)
 * in one NCCL group on the library stream: the halo exchange of the E-field
 * edges between smoothing sweeps.  comm_allreduce_sum sums doubles in place
 * (norms, dot products).  libnccl.so.2 is opened at run time.                  */
int emg3d_b200_comm_unique_id(void* out128);
int emg3d_b200_comm_init(const void* unique_id128, int nranks, int rank);
int emg3d_b200_comm_size(int* nranks, int* rank);
int emg3d_b200_comm_destroy(void);
int emg3d_b200_comm_sendrecv(int n, void* const* ptrs, const size_t* nbytes, const int* peers,
                             const int* is_send);
int emg3d_b200_comm_allreduce_sum(double* dev, int n);

/* ---- halo exchange over peer memory (NVLink, CUDA IPC) ------------------------
 * Replaces the NCCL send/recv group of a halo exchange by ONE kernel that pulls the
 * z-neighbours' boundary planes with remote loads and synchronises with them
 * through flags in peer memory (csrc/comm.cu).  p2p_init and p2p_register are
 * collective (they use the NCCL communicator to publish IPC handles); every rank
 * registers the arrays it exchanges in the same order.  *enabled == 0 / *slot < 0:
 * peer mapping is not possible here, keep using comm_sendrecv.  p2p_exchange pulls
 * n (<= 8) byte ranges from the array the neighbour registered in the same slot.
 * New functionality (the reference has no distributed solve). */
int emg3d_b200_p2p_init(int* enabled);
int emg3d_b200_p2p_register(void* dev_ptr, int* slot);
/* Forget all registered arrays and unmap the neighbours' copies.  Call on every rank before
 * the registered arrays are freed; synchronise the ranks before registering again.        */
int emg3d_b200_p2p_release(void);
/* push = 0: pull nbytes[i] from byte offset peer_off[i] of the neighbour's array (upper neighbour
 * if from_upper[i]) to my_off[i] of mine; push = 1: write nbytes[i] from my_off[i] of mine to
 * peer_off[i] of that neighbour's (posted stores over NVLink).  All ranks use the same mode. */
int emg3d_b200_p2p_exchange(int slot, int n, const size_t* my_off, const size_t* peer_off,
                            const size_t* nbytes, const int* from_upper, int push);
int emg3d_b200_p2p_status(int* status);
int emg3d_b200_p2p_shutdown(void);

/* ---- host-array convenience entry points ----------------------------------
 * Exact signatures of the reference kernels on HOST arrays (upload, run,
 * download); these are what a ctypes/cffi shim inside emg3d/core.py would
 * bind, see INTEGRATION.md.                                                    */
int emg3d_b200_host_amat_x(int cplx, int nx, int ny, int nz, void* rx, void* ry, void* rz,
                           const void* ex, const void* ey, const void* ez, const void* eta_x,
                           const void* eta_y, const void* eta_z, const double* zeta,
                           const double* hx, const double* hy, const double* hz);
int emg3d_b200_host_gauss_seidel(int cplx, int ldir, int order, int nx, int ny, int nz, void* ex,
                                 void* ey, void* ez, const void* sx, const void* sy,
                                 const void* sz, const void* eta_x, const void* eta_y,
                                 const void* eta_z, const double* zeta, const double* hx,
                                 const double* hy, const double* hz, int nu);
/* core.restrict(crx, cry, crz, rx, ry, rz, wx, wy, wz, sc_dir) (core.py:1620-1621; call site
 * solver.py:937-938) on host arrays.  (nx, ny, nz): FINE cell counts; the coarse arrays have the
 * shapes sc_dir implies (0: x, y, z halved; 1: y, z; 2: x, z; 3: x, y; 4: x; 5: y; 6: z;
 * solver.py:891-897).  weights[3 a + {0, 1, 2}] = (wl, w0, wr) of axis a as returned by
 * core.restrict_weights, length n_coarse_nodes(a); NULL for an axis that is not coarsened. */
int emg3d_b200_host_restrict(int cplx, int nx, int ny, int nz, int sc_dir, void* crx, void* cry,
                             void* crz, const void* rx, const void* ry, const void* rz,
                             const double* const* weights);

/* ---- interpolation on either side of a solve; device pointers ---------
 *
 * emg3d/maps.py:556-617 `interp_volume_average(nodes_x, nodes_y, nodes_z, values, new_nodes_x,
 * new_nodes_y, new_nodes_z, new_values, new_vol)`, the numba kernel behind
 * maps.interpolate(method='volume') / Model.interpolate_to_grid (emg3d/models.py:322-380).
 * values (nx, ny, nz), out (mx, my, mz), x fastest.  Per axis a: the merged segments of
 * maps.py:620-665 `_volume_average_weights` sorted by output cell -- w[a] (weights), iin[a]
 * (input cell), start[a] (mx + 1 offsets: the segments of output cell o are start[o] ..
 * start[o + 1]) -- and hnew[a], the widths of the new grid (new_vol = their product).
 * log_scale = 1: log10 on the way in, 10** on the way out (maps.py:306-308, 366-367).
 * add = 1: the sums are added to the content of `out` before the division, like the reference. */
int emg3d_b200_volume_average(const double* values, int nx, int ny, int nz, double* out, int mx, int my,
                              int mz, const double* const* w, const int* const* iin, const int* const* start,
                              const double* const* hnew, int log_scale, int add);
/* emg3d/maps.py:668-720 `interp_edges_to_vol_averages(ex, ey, ez, volumes, ox, oy, oz)`: field in
 * the [fx | fy | fz] layout, out = [ox | oy | oz], three arrays of nx ny nz values of the field's
 * dtype; hx, hy, hz: cell widths (device). */
int emg3d_b200_edges_to_vol_averages(int is_cplx, int nx, int ny, int nz, const void* field, const double* hx,
                                     const double* hy, const double* hz, void* out);
/* The same fused with emg3d/simulations.py:1028-1031: the edge values are Re(b * smu0 * e) of the
 * forward and the back-propagated field (complex128); out: three real arrays (the gradient on the
 * computational grid). */
int emg3d_b200_gradient_field(int nx, int ny, int nz, const void* efield, const void* bfield, double smu0_re,
                              double smu0_im, const double* hx, const double* hy, const double* hz, double* out);
/* Cubic B-spline coefficients of an (n0, n1, n2) array (x fastest), in place: what
 * scipy.ndimage.spline_filter(order=3) computes inside map_coordinates, which the reference calls
 * through emg3d/maps.py:500-553 `interp_spline_3d`.  reflect = 0: mirror boundaries (modes
 * 'constant', 'mirror'); 1: reflect (mode 'nearest', after emg3d_b200_pad_edge3 with npad = 12). */
int emg3d_b200_spline_filter3(int is_cplx, int n0, int n1, int n2, void* data, int reflect);
int emg3d_b200_pad_edge3(int is_cplx, int n0, int n1, int n2, const void* src, int npad, void* dst);
/* dst (m[0], m[1], m[2]) = the sub-box src[lo : lo + m] of an (n0, n1, n2) array (lo, m: host).  The
 * spline prefilter decays like 0.268^k: coefficients computed on a box that extends 48 samples
 * beyond the requested points equal those of the whole array to 1e-27 (emg3d_b200/maps.py). */
int emg3d_b200_copy_box3(int is_cplx, int n0, int n1, int n2, const void* src, const int* lo, const int* m,
                         void* dst);
/* Values at points: method 3 = cubic spline (data = coefficients of emg3d_b200_spline_filter3;
 * mode 0 'constant': `fill` outside the data; mode 1 'nearest': data is the padded array, npad
 * samples per side), method 1 = linear (scipy RegularGridInterpolator, maps.py:355-362; a NaN
 * coordinate marks a point that gets `fill`).  Coordinates in index units of the (unpadded) array:
 * tensor = 0: point p = (cx[p], cy[p], cz[p]); tensor = 1: p = a + m0 (b + m1 c) = (cx[a], cy[b],
 * cz[c]).  out[p] = scale * value, added to out[p] if accumulate (get_receiver sums the three
 * components weighted by the receiver's direction cosines, emg3d/fields.py:596-604). */
int emg3d_b200_interp_points(int is_cplx, int method, int n0, int n1, int n2, const void* data, int npad, int mode,
                             double fill_re, double fill_im, const double* cx, const double* cy, const double* cz,
                             long long npts, int tensor, int m0, int m1, double scale_re, double scale_im,
                             int accumulate, void* out);

/* core.solve (core.py:1481-1482): amat has 6 n entries, bvec n; both in place. */
/* Magnetic field on the faces from the electric field on the edges (Faraday's law):
 * emg3d/fields.py:617-659 `get_magnetic_field` (device pointers; the level's model must
 * be set; hfield holds (nx+1) ny nz + nx (ny+1) nz + nx ny (nz+1) values, laid out
 * [hx | hy | hz], x fastest; scale = 1 / (s mu_0)) and the kernel it calls,
 * fields.py:941-1009 `_edge_curl_factor(mx, my, mz, ex, ey, ez, hx, hy, hz, zeta)`,
 * with that argument list on host arrays (`zeta` has the dtype of the fields). */
int emg3d_b200_magnetic_field(emg3d_b200_level* lv, const void* e, void* hfield, double scale_re,
                              double scale_im);
int emg3d_b200_host_edge_curl_factor(int is_cplx, int nx, int ny, int nz, void* mx, void* my,
                                     void* mz, const void* ex, const void* ey, const void* ez,
                                     const double* hx, const double* hy, const double* hz,
                                     const void* zeta);
int emg3d_b200_host_solve(int cplx, int n, void* amat, void* bvec);

#ifdef __cplusplus
}
#endif
#endif /* EMG3D_B200_H */
