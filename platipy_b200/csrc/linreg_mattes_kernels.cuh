// linreg_mattes_kernels.cuh -- Mattes mutual information for linear_registration (reference linear.py:145-146,
// SetMetricAsMattesMutualInformation() -> itk::MattesMutualInformationImageToImageMetricv4, 50 bins): a joint histogram with a
// zero-order Parzen window on the fixed intensity (one bin) and a cubic B-spline window on the moving intensity (four bins).
//
// Two passes over the same samples as the other metrics (linreg_sample_point):
//   1. histogram   hist[f][m] += B3(m - u),  u = M / binsize_M - normalised_min_M, bins clamped to [2, bins - 3] like ITK.
//                  Weights are accumulated as 2^-32 fixed-point integers with integer atomics: exactly associative, so the
//                  histogram is deterministic whatever order the atomics land in.
//   2. derivative  with the host's table L[f][m] = log(p(f, m) / p_M(m)) (Mattes et al., eq. 27; the terms in the marginals
//                  cancel), every sample has ONE scalar weight  w = sum_m L[f][m] * B3'(m - u) * (-1 / binsize_M)  and the
//                  derivative sums are  s = sum w h,  S = sum w h (x - c)^T  -- the 12 numbers every transform model needs.
// Per-sample code is plain C++ (also run by tests/emu); the block reduction of pass 2 is CUDA-specific.
#pragma once
#include "linreg_corr_kernels.cuh"

namespace b200 {

struct MattesBins {
    int n;             // bins per axis (50)
    double fbin, fmin; // fixed: bin size, normalised minimum (min / binsize - padding)
    double mbin, mmin; // moving
};
constexpr int MATTES_NV = 12;
constexpr double MATTES_FIXED_POINT = 4294967296.0;  // 2^32

__device__ __forceinline__ double bspline3(double u)
{
    const double a = u < 0.0 ? -u : u;
    if (a < 1.0) return (4.0 - 6.0 * a * a + 3.0 * a * a * a) / 6.0;
    if (a < 2.0) return (2.0 - a) * (2.0 - a) * (2.0 - a) / 6.0;
    return 0.0;
}
__device__ __forceinline__ double bspline3_derivative(double u)
{
    const double a = u < 0.0 ? -u : u;
    double d;
    if (a < 1.0) d = -2.0 * a + 1.5 * a * a;
    else if (a < 2.0) d = -0.5 * (2.0 - a) * (2.0 - a);
    else d = 0.0;
    return u < 0.0 ? -d : d;
}
// fixed bin, first moving bin and the Parzen-window term of one sample
__device__ __forceinline__ void mattes_bins(const MattesBins& mb, double fval, double mval, int& fi, int& mi, double& term)
{
    const int lo = 2, hi = mb.n - 3;
    fi = (int)floor(fval / mb.fbin - mb.fmin);
    fi = fi < lo ? lo : (fi > hi ? hi : fi);
    term = mval / mb.mbin - mb.mmin;
    mi = (int)floor(term);
    mi = mi < lo ? lo : (mi > hi ? hi : mi);
}

// hist: [replicas][n][n] fixed-point weights (zeroed by the caller); count: number of valid samples.  Neighbouring blocks add into
// different replicas (block index modulo `replicas`): a CT-like image puts most samples into a handful of bins, and the replicas
// divide the contention on those addresses; the caller sums them (integers: still exactly associative).
__global__ void __launch_bounds__(256) linreg_mattes_hist_kernel(const float* __restrict__ F, const float* __restrict__ M, const uint8_t* __restrict__ fmask,
                                                                 const uint8_t* __restrict__ mmask, const __grid_constant__ CorrGeom gf,
                                                                 const __grid_constant__ CorrGeom gm, const __grid_constant__ CorrPose ps,
                                                                 const __grid_constant__ MattesBins mb, int stride, size_t nsamples, int replicas,
                                                                 unsigned long long* __restrict__ hist, unsigned long long* __restrict__ count)
{
    unsigned long long local = 0;
    hist += (size_t)(blockIdx.x % (unsigned)replicas) * mb.n * mb.n;
    for (size_t sidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < nsamples; sidx += (size_t)gridDim.x * blockDim.x) {
        LinregPoint pt;
        if (!linreg_sample_point(F, M, fmask, mmask, gf, gm, ps, sidx * (size_t)stride, pt)) continue;
        int fi, mi;
        double term;
        mattes_bins(mb, pt.fval, pt.mval, fi, mi, term);
        for (int b = mi - 1; b <= mi + 2; ++b) {
            const double w = bspline3((double)b - term);
            if (w > 0.0) atomicAdd(&hist[(size_t)fi * mb.n + b], (unsigned long long)(w * MATTES_FIXED_POINT + 0.5));
        }
        ++local;
    }
    if (local) atomicAdd(count, local);
}

// table: [n][n] doubles, L[f][m] = log(p(f, m) / p_M(m)) (0 where the ratio is undefined); partials as in linreg_corr_kernel
__global__ void __launch_bounds__(128) linreg_mattes_deriv_kernel(const float* __restrict__ F, const float* __restrict__ M, const uint8_t* __restrict__ fmask,
                                                                  const uint8_t* __restrict__ mmask, const __grid_constant__ CorrGeom gf,
                                                                  const __grid_constant__ CorrGeom gm, const __grid_constant__ CorrPose ps,
                                                                  const __grid_constant__ MattesBins mb, int stride, size_t nsamples,
                                                                  const double* __restrict__ table, double* __restrict__ partials)
{
    double acc[MATTES_NV];
#pragma unroll
    for (int v = 0; v < MATTES_NV; ++v) acc[v] = 0.0;
    for (size_t sidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < nsamples; sidx += (size_t)gridDim.x * blockDim.x) {
        LinregPoint pt;
        if (!linreg_sample_point(F, M, fmask, mmask, gf, gm, ps, sidx * (size_t)stride, pt)) continue;
        int fi, mi;
        double term;
        mattes_bins(mb, pt.fval, pt.mval, fi, mi, term);
        double w = 0.0;
        for (int b = mi - 1; b <= mi + 2; ++b) w += table[(size_t)fi * mb.n + b] * bspline3_derivative((double)b - term);
        w = -w / mb.mbin;  // d(b - u)/dM = -1 / binsize
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double wr = w * pt.h[r];
            acc[r] += wr;
            acc[3 + r * 3 + 0] += wr * pt.xc[0];
            acc[3 + r * 3 + 1] += wr * pt.xc[1];
            acc[3 + r * 3 + 2] += wr * pt.xc[2];
        }
    }
#ifdef B200_HOST_EMU
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int v = 0; v < MATTES_NV; ++v) partials[tid * MATTES_NV + v] = acc[v];
#else
    __shared__ double sh[MATTES_NV][4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < MATTES_NV; ++v) {
        const double t = warp_sum(acc[v]);
        if (lane == 0) sh[v][wid] = t;
    }
    __syncthreads();
    if (threadIdx.x < MATTES_NV) {
        double t = 0.0;
        for (int w4 = 0; w4 < 4; ++w4) t += sh[threadIdx.x][w4];
        partials[(size_t)blockIdx.x * MATTES_NV + threadIdx.x] = t;
    }
#endif
}

__global__ void linreg_mattes_final_kernel(const double* __restrict__ partials, int nb, double* __restrict__ out)
{
    const int v = threadIdx.x;
    if (v >= MATTES_NV) return;
    double t = 0.0;
    for (int q = 0; q < nb; ++q) t += partials[(size_t)q * MATTES_NV + v];
    out[v] = t;
}

}  // namespace b200
