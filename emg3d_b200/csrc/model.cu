// Volume-averaged coefficients on the device (what emg3d/models.py:654-691
// `VolumeModel.__init__` computes with NumPy on the host):
//     eta_a = -s mu_0 V (sigma_a + s eps_0 eps_r),   zeta = V / mu_r,
// V = hx hy hz per cell, sigma_a = map.backward(property_a).  Moving this to the
// GPU removes the O(cells) host work from every solve and halves the H2D
// volume (real property arrays instead of complex eta arrays).
// The arithmetic order follows NumPy's evaluation ((-s mu0) * V) * (...) without
// fused multiply-adds, so the result is bitwise the one the reference feeds its
// kernels.
#include "common.cuh"
#include "kernels.h"

namespace emg {

// property -> conductivity (emg3d/maps.py:52-227)
__device__ __forceinline__ double to_cond(double p, int map) {
    switch (map) {
        case 0: return p;                 // Conductivity
        case 1: return 1.0 / p;           // Resistivity
        case 2: return pow(10.0, p);      // LgConductivity
        case 3: return exp(p);            // LnConductivity
        case 4: return pow(10.0, -p);     // LgResistivity
        default: return exp(-p);          // LnResistivity
    }
}

__device__ __forceinline__ void store_eta(cplx* out, int64_t i, double cr, double ci, double v,
                                          double cond, bool has_eps, double er, double ei) {
    // t = c * v (complex * real), then t * (cond [+ (er + i ei)])
    const double tr = __dmul_rn(cr, v), ti = __dmul_rn(ci, v);
    if (!has_eps) {
        out[i] = make_c(__dmul_rn(tr, cond), __dmul_rn(ti, cond));
    } else {
        const double wr = __dadd_rn(cond, er), wi = ei;
        out[i] = make_c(__dsub_rn(__dmul_rn(tr, wr), __dmul_rn(ti, wi)),
                        __dadd_rn(__dmul_rn(tr, wi), __dmul_rn(ti, wr)));
    }
}
__device__ __forceinline__ void store_eta(double* out, int64_t i, double cr, double, double v,
                                          double cond, bool has_eps, double er, double) {
    const double t = __dmul_rn(cr, v);
    out[i] = __dmul_rn(t, has_eps ? __dadd_rn(cond, er) : cond);
}

template <typename T>
__global__ void __launch_bounds__(256)
volume_model_kernel(Dims d, const double* __restrict__ hx, const double* __restrict__ hy,
                    const double* __restrict__ hz, double cr, double ci, double sr, double si,
                    double eps0, int map, const double* __restrict__ px, const double* __restrict__ py,
                    const double* __restrict__ pz, const double* __restrict__ mu,
                    const double* __restrict__ eps, T* ex, T* ey, T* ez, double* zeta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= d.n[0] || j >= d.n[1] || k >= d.n[2]) return;
    const int64_t id = i + (int64_t)d.n[0] * (j + (int64_t)d.n[1] * k);
    const double v = __dmul_rn(__dmul_rn(hx[i], hy[j]), hz[k]);
    double er = 0.0, ei = 0.0;
    if (eps) {                            // s * eps_0 * eps_r, evaluated left to right
        const double e = eps[id];
        er = __dmul_rn(__dmul_rn(sr, eps0), e);
        ei = __dmul_rn(__dmul_rn(si, eps0), e);
    }
    store_eta(ex, id, cr, ci, v, to_cond(px[id], map), eps != nullptr, er, ei);
    if (py) store_eta(ey, id, cr, ci, v, to_cond(py[id], map), eps != nullptr, er, ei);
    if (pz) store_eta(ez, id, cr, ci, v, to_cond(pz[id], map), eps != nullptr, er, ei);
    zeta[id] = mu ? v / mu[id] : v;
}

template <typename T>
void launch_volume_model(const Dims& d, const double* hx, const double* hy, const double* hz,
                         double cr, double ci, double sr, double si, double eps0, int map,
                         const double* px, const double* py, const double* pz, const double* mu,
                         const double* eps, T* ex, T* ey, T* ez, double* zeta, cudaStream_t st) {
    dim3 b(32, 4, 2);
    dim3 g((d.n[0] + b.x - 1) / b.x, (d.n[1] + b.y - 1) / b.y, (d.n[2] + b.z - 1) / b.z);
    ++g_launch_count; volume_model_kernel<T><<<g, b, 0, st>>>(d, hx, hy, hz, cr, ci, sr, si, eps0, map, px, py, pz, mu, eps, ex, ey, ez, zeta);
}

template void launch_volume_model<double>(const Dims&, const double*, const double*, const double*,
                                          double, double, double, double, double, int, const double*,
                                          const double*, const double*, const double*, const double*,
                                          double*, double*, double*, double*, cudaStream_t);
template void launch_volume_model<cplx>(const Dims&, const double*, const double*, const double*,
                                        double, double, double, double, double, int, const double*,
                                        const double*, const double*, const double*, const double*,
                                        cplx*, cplx*, cplx*, double*, cudaStream_t);

}  // namespace emg
