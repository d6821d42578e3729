// gauss.cuh -- itk::GaussianOperator coefficients (host) and separable clamp-boundary convolutions (device).
//   N1  DiscreteGaussianImageFilter        reference utils.py:226, fusion.py:168,279
//   N6  PDE field smoothing (update/displacement fields)  reference deformable.py:249-257
#pragma once
#include "common.cuh"

namespace b200 {

// ---- GaussianOperator::GenerateCoefficients ---------------------------------------------------------
// Discrete Gaussian e^{-t} I_n(t); the modified Bessel functions are the polynomial approximations ITK
// uses (Abramowitz & Stegun 9.8), so coefficients agree with ITK's to the last bit, not just to 1e-7.
inline double bessel_i0(double y)
{
    const double d = std::fabs(y);
    if (d < 3.75) {
        double m = y / 3.75;
        m *= m;
        return 1.0 + m * (3.5156229 + m * (3.0899424 + m * (1.2067492 + m * (0.2659732 + m * (0.360768e-1 + m * 0.45813e-2)))));
    }
    const double m = 3.75 / d;
    return (std::exp(d) / std::sqrt(d)) *
           (0.39894228 + m * (0.1328592e-1 + m * (0.225319e-2 + m * (-0.157565e-2 + m * (0.916281e-2 + m * (-0.2057706e-1 + m * (0.2635537e-1 + m * (-0.1647633e-1 + m * 0.392377e-2))))))));
}
inline double bessel_i1(double y)
{
    const double d = std::fabs(y);
    double acc;
    if (d < 3.75) {
        double m = y / 3.75;
        m *= m;
        acc = d * (0.5 + m * (0.87890594 + m * (0.51498869 + m * (0.15084934 + m * (0.2658733e-1 + m * (0.301532e-2 + m * 0.32411e-3))))));
    } else {
        const double m = 3.75 / d;
        acc = 0.2282967e-1 + m * (-0.2895312e-1 + m * (0.1787654e-1 - m * 0.420059e-2));
        acc = 0.39894228 + m * (-0.3988024e-1 + m * (-0.362018e-2 + m * (0.163801e-2 + m * (-0.1031555e-1 + m * acc))));
        acc *= (std::exp(d) / std::sqrt(d));
    }
    return y < 0.0 ? -acc : acc;
}
inline double bessel_in(int n, double y)
{
    if (y == 0.0) return 0.0;
    const double toy = 2.0 / std::fabs(y);
    double qip = 0.0, acc = 0.0, qi = 1.0;
    for (int j = 2 * (n + (int)std::sqrt(40.0 * n)); j > 0; j--) {
        const double qim = qip + j * toy * qi;
        qip = qi;
        qi = qim;
        if (std::fabs(qi) > 1.0e10) {
            acc *= 1.0e-10;
            qi *= 1.0e-10;
            qip *= 1.0e-10;
        }
        if (j == n) acc = qip;
    }
    acc *= bessel_i0(y) / qi;
    return (y < 0.0 && (n & 1)) ? -acc : acc;
}

// symmetric, normalised kernel of 2r+1 taps; terms are added until the running sum reaches 1 - max_error,
// a term drops below sum * DBL_EPSILON, or the one-sided length exceeds max_width.
inline std::vector<double> gaussian_operator(double variance, double max_error, int max_width)
{
    std::vector<double> c;
    const double et = std::exp(-variance), cap = 1.0 - max_error;
    double sum = 0.0;
    c.push_back(et * bessel_i0(variance));
    sum += c[0];
    c.push_back(et * bessel_i1(variance));
    sum += c[1] * 2.0;
    for (int i = 2; sum < cap; ++i) {
        c.push_back(et * bessel_in(i, variance));
        sum += c[i] * 2.0;
        if (c[i] < sum * DBL_EPSILON) break;
        if ((int)c.size() > max_width) break;
    }
    for (double& v : c) v /= sum;
    const int r = (int)c.size() - 1;
    std::vector<double> k(2 * r + 1);
    for (int i = 0; i <= r; ++i) {
        k[r + i] = c[i];
        k[r - i] = c[i];
    }
    return k;
}

// Coefficients travel to kernels by value (constant bank), up to radius KMAX_R.
constexpr int KMAX_R = 48;
struct KernelCoeffs {
    int r;
    double k[2 * KMAX_R + 1];
};
inline int make_coeffs(const std::vector<double>& k, KernelCoeffs* out)
{
    const int r = ((int)k.size() - 1) / 2;
    if (r > KMAX_R) return set_error(B200REG_ERR_UNSUPPORTED, "Gaussian kernel radius %d exceeds the supported %d", r, KMAX_R);
    out->r = r;
    for (size_t i = 0; i < k.size(); ++i) out->k[i] = k[i];
    return B200REG_OK;
}

// Demons loop control block (device resident; see demons.cuh).  Kernels launched for iteration `it` do
// nothing once it >= halt_iter, so a whole level is enqueued without any host synchronisation.
struct DemonsCtrl {
    int halt_iter;
    int elapsed;
    double metric;
    double rms;
};

// ---- 1-D convolution along AXIS, ZeroFluxNeumann (index clamp) boundary --------------------------------
// Inner product accumulated in double from offset -r to +r (itk::NeighborhoodInnerProduct order).
// ADD: the input is a + b evaluated on the fly (D + U of FastSymmetricForcesDemons::ApplyUpdate).
// gridDim.z = nz * nplanes (SoA planes are contiguous volumes).
template <typename T, int AXIS, bool ADD>
__global__ void __launch_bounds__(BX* BY) conv_axis_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                              int nx, int ny, int nz, const __grid_constant__ KernelCoeffs kc,
                                                              const DemonsCtrl* __restrict__ ctrl, int it)
{
    if (ctrl && it >= ctrl->halt_iter) return;
    const int i = blockIdx.x * BX + threadIdx.x;
    const int j = blockIdx.y * BY + threadIdx.y;
    const int kz = blockIdx.z;  // plane * nz + k
    if (i >= nx || j >= ny) return;
    const int k = kz % nz;
    const size_t plane_off = (size_t)(kz / nz) * nx * ny * nz;
    const int n = AXIS == 0 ? nx : (AXIS == 1 ? ny : nz);
    const int pos = AXIS == 0 ? i : (AXIS == 1 ? j : k);
    const size_t sa = AXIS == 0 ? 1 : (AXIS == 1 ? (size_t)nx : (size_t)nx * ny);
    const size_t o = plane_off + ((size_t)k * ny + j) * nx + i;
    const size_t base = o - (size_t)pos * sa;
    const int r = kc.r;
    double sum = 0.0;
    for (int t = -r; t <= r; ++t) {
        int q = pos + t;
        q = q < 0 ? 0 : (q > n - 1 ? n - 1 : q);
        const size_t idx = base + (size_t)q * sa;
        double v;
        if (ADD) v = (double)a[idx] + (double)b[idx];
        else v = (double)a[idx];
        sum += kc.k[t + r] * v;
    }
    out[o] = (T)sum;
}

template <typename T, bool ADD>
inline int launch_conv_axis(b200reg_ctx* ctx, int axis, const T* a, const T* b, T* out, int nx, int ny, int nz, int nplanes,
                            const KernelCoeffs& kc, const DemonsCtrl* ctrl, int it)
{
    dim3 g((nx + BX - 1) / BX, (ny + BY - 1) / BY, nz * nplanes), blk(BX, BY, 1);
    if (axis == 0) conv_axis_kernel<T, 0, ADD><<<g, blk, 0, ctx->stream>>>(a, b, out, nx, ny, nz, kc, ctrl, it);
    else if (axis == 1) conv_axis_kernel<T, 1, ADD><<<g, blk, 0, ctx->stream>>>(a, b, out, nx, ny, nz, kc, ctrl, it);
    else conv_axis_kernel<T, 2, ADD><<<g, blk, 0, ctx->stream>>>(a, b, out, nx, ny, nz, kc, ctrl, it);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// DiscreteGaussianImageFilter on a Float32 image: variance (mm^2) -> voxel^2 per axis when
// use_image_spacing, passes z -> y -> x, float32 intermediates.
inline int discrete_gaussian_f32(b200reg_ctx* ctx, const float* d_in, float* d_out, const b200reg_geom& g, const double* variance,
                                 int max_width, double max_error, int use_spacing)
{
    const int nx = g.size[0], ny = g.size[1], nz = g.size[2];
    const size_t n = nvox(g);
    TempBuf t1, t2;
    B200_TRY(t1.alloc(ctx, n * sizeof(float)));
    B200_TRY(t2.alloc(ctx, n * sizeof(float)));
    const float* src = d_in;
    float* dsts[3] = { t1.as<float>(), t2.as<float>(), d_out };
    int pass = 0;
    for (int axis = 2; axis >= 0; --axis, ++pass) {
        double t = variance[axis];
        if (use_spacing) t = t / (g.spacing[axis] * g.spacing[axis]);
        KernelCoeffs kc;
        B200_TRY(make_coeffs(gaussian_operator(t, max_error, max_width), &kc));
        B200_TRY((launch_conv_axis<float, false>(ctx, axis, src, nullptr, dsts[pass], nx, ny, nz, 1, kc, nullptr, 0)));
        src = dsts[pass];
    }
    return B200REG_OK;
}

}  // namespace b200
