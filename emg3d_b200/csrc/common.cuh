// emg3d_b200 -- shared device helpers for the multigrid hot path (sm_100a).
//
// Data layout in HBM (identical to the reference's host layout, so transfers
// are plain copies): a field is one array [fx | fy | fz], fx (nx, ny+1, nz+1),
// fy (nx+1, ny, nz+1), fz (nx+1, ny+1, nz), all x-fastest; eta_x/y/z and zeta
// are (nx, ny, nz) x-fastest.  Fields and eta are complex128 (16-byte aligned
// re,im pairs -> one LDG.128 per value) in the frequency domain and float64 in
// the Laplace domain; zeta and the widths are always float64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace emg {

// ---- scalar types ---------------------------------------------------------
struct __align__(16) cplx {
    double re, im;
};

__host__ __device__ __forceinline__ cplx make_c(double re, double im) { cplx r; r.re = re; r.im = im; return r; }
__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return make_c(a.re + b.re, a.im + b.im); }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return make_c(a.re - b.re, a.im - b.im); }
__host__ __device__ __forceinline__ cplx operator-(cplx a) { return make_c(-a.re, -a.im); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return make_c(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
__host__ __device__ __forceinline__ cplx operator*(double a, cplx b) { return make_c(a * b.re, a * b.im); }
__host__ __device__ __forceinline__ cplx operator*(cplx b, double a) { return make_c(a * b.re, a * b.im); }
__host__ __device__ __forceinline__ cplx& operator+=(cplx& a, cplx b) { a.re += b.re; a.im += b.im; return a; }
__host__ __device__ __forceinline__ cplx& operator-=(cplx& a, cplx b) { a.re -= b.re; a.im -= b.im; return a; }
__host__ __device__ __forceinline__ cplx& operator*=(cplx& a, cplx b) { a = a * b; return a; }
__host__ __device__ __forceinline__ cplx& operator*=(cplx& a, double b) { a.re *= b; a.im *= b; return a; }

// reciprocal of a pivot.  On the device: hardware seed (MUFU.RCP64H, ~20 bits) and
// two Newton steps -- full double precision to within an ulp, a third of the
// instructions of the IEEE division sequence and no slow-path branch (pivots are
// finite, normal numbers).  Off by default (measured r1: no gain, the smoothers are bound by the memory system); -DEMG_FAST_RCP=1 enables it.
#ifndef EMG_FAST_RCP
#define EMG_FAST_RCP 0
#endif
__host__ __device__ __forceinline__ double rcp(double a) {
#if defined(__CUDA_ARCH__) && EMG_FAST_RCP
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / a;
#endif
}
__host__ __device__ __forceinline__ cplx rcp(cplx a) {
    double d = rcp(a.re * a.re + a.im * a.im);
    return make_c(a.re * d, -a.im * d);
}
__host__ __device__ __forceinline__ double abs2(double a) { return a * a; }
__host__ __device__ __forceinline__ double abs2(cplx a) { return a.re * a.re + a.im * a.im; }
__host__ __device__ __forceinline__ double conj_(double a) { return a; }
__host__ __device__ __forceinline__ cplx conj_(cplx a) { return make_c(a.re, -a.im); }

__host__ __device__ __forceinline__ void add_real(double& a, double v) { a += v; }
__host__ __device__ __forceinline__ void add_real(cplx& a, double v) { a.re += v; }

template <typename T> __host__ __device__ __forceinline__ T zero_();
template <> __host__ __device__ __forceinline__ double zero_<double>() { return 0.0; }
template <> __host__ __device__ __forceinline__ cplx zero_<cplx>() { return make_c(0.0, 0.0); }

// read-only (non-coherent) loads
__device__ __forceinline__ double ldg(const double* p) { return __ldg(p); }
__device__ __forceinline__ int ldg(const int* p) { return __ldg(p); }
__device__ __forceinline__ cplx ldg(const cplx* p) {
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return make_c(v.x, v.y);
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier ---------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// nbytes (multiple of 16) from global to this CTA's shared memory; 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned nbytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(dst)), "l"(src), "r"(nbytes), "r"(smem_u32(bar)) : "memory");
}
// order generic-proxy accesses to shared memory before later async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
// asynchronous copy of one 8- or 16-byte element from global to shared memory (SASS LDGSTS)
template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem_u32(dst_smem)), "l"(src), "n"(BYTES)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
// nbytes (multiple of 16) from global memory into the L2, no destination
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned nbytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src), "r"(nbytes) : "memory");
}

// ---- grid / model description passed by value to kernels -------------------
struct Dims {
    int n[3];          // cells per axis
    // z-window (multi-GPU slabs): the arrays hold `nzf` cells along z (0 = no
    // window, i.e. n[2]); this view starts at cell / node plane `zoff` of them and
    // spans n[2] cells.  z is the slowest axis, so a window of a component or of a
    // cell array is a contiguous range: only base pointers move.
    int zoff, nzf;
    // planes of this view whose edges count in norms (multi-GPU: the planes a rank
    // owns; own1 == 0: all).  x/y-edges on node plane k count iff own0 <= k < own1,
    // z-edges of cell layer k iff own0 <= k + 1 < own1 (a layer belongs to the
    // owner of its upper plane).
    int own0, own1;
    // multi-GPU z-slabs: 1 if local node plane 1 of this view is an EVEN global node plane, so
    // that the z-parity of the multicolour classes (node colours, x-/y-line colours) is the
    // global one on every rank: class bit cz relaxes the local planes of parity cz ^ zflip.
    // (The tile-fused point schedule colours by local tile index and ignores it.)
    int zflip;
};
__host__ __device__ __forceinline__ bool owns_plane(const Dims& d, int k) {
    return d.own1 == 0 || (k >= d.own0 && k < d.own1);
}
__host__ __device__ __forceinline__ bool owns_layer(const Dims& d, int k) {
    return d.own1 == 0 || (k + 1 >= d.own0 && k + 1 < d.own1);
}
__host__ __device__ __forceinline__ int full_nz(const Dims& d) { return d.nzf > 0 ? d.nzf : d.n[2]; }

template <typename T>
struct Model {
    Dims d;
    const T* eta[3];       // eta_x, eta_y, eta_z (may alias)
    const double* zeta;
    const double* h[3];    // widths
    const double* rh[3];   // 1 / widths
    const T* diag;         // diagonal of A per edge, field layout (point smoother; may be null)
};

// F-order extents of the three field components
__host__ __device__ __forceinline__ int64_t comp_d0(const Dims& d, int c) { return d.n[0] + (c != 0); }
__host__ __device__ __forceinline__ int64_t comp_d1(const Dims& d, int c) { return d.n[1] + (c != 1); }
__host__ __device__ __forceinline__ int64_t comp_d2(const Dims& d, int c) { return full_nz(d) + (c != 2); }
__host__ __device__ __forceinline__ int64_t comp_size(const Dims& d, int c) {
    return comp_d0(d, c) * comp_d1(d, c) * comp_d2(d, c);
}
__host__ __device__ __forceinline__ int64_t comp_offset(const Dims& d, int c) {
    int64_t o = 0;
    for (int k = 0; k < c; ++k) o += comp_size(d, k);
    return o;
}
__host__ __device__ __forceinline__ int64_t n_edges(const Dims& d) { return comp_offset(d, 3); }
__host__ __device__ __forceinline__ int64_t n_cells(const Dims& d) {
    return (int64_t)d.n[0] * d.n[1] * d.n[2];
}

// Strides of component c and of the cell arrays, and pointers to components.
template <typename T>
struct FieldView {
    T* p[3];
    int64_t s1[3], s2[3];   // stride of index 1 and 2 (stride of index 0 is 1)
    __host__ __device__ FieldView() {}
    __host__ __device__ FieldView(T* base, const Dims& d) {
        int64_t o = 0;
        for (int c = 0; c < 3; ++c) {
            s1[c] = comp_d0(d, c);
            s2[c] = comp_d0(d, c) * comp_d1(d, c);
            p[c] = base ? base + o + s2[c] * d.zoff : nullptr;
            o += comp_size(d, c);
        }
    }
    __host__ __device__ __forceinline__ int64_t idx(int c, int i, int j, int k) const {
        return i + s1[c] * j + s2[c] * k;
    }
    __host__ __device__ __forceinline__ int64_t idx(int c, const int* q) const {
        return q[0] + s1[c] * q[1] + s2[c] * q[2];
    }
};

}  // namespace emg
