// linreg_corr_kernels.cuh -- sums behind the correlation metric of linear_registration (reference linear.py:141-146,
// SetMetricAsCorrelation -> itk::CorrelationImageToImageMetricv4: value = -(sum (F - mF)(M - mM))^2 / (sum (F - mF)^2 sum (M - mM)^2)).
//
// Same sampling as the mean-squares kernel (linreg.cuh): every `stride`-th fixed voxel in raster order, x -> y = A x + b,
// trilinear value and analytic gradient of the moving image, optional masks.  The value and its derivative are rational
// functions of sums that can all be taken in ONE pass:
//     [0..5]   N, sum F, sum M, sum F^2, sum M^2, sum F M
//     for each weight w in (1, F, M):  s_w = sum w h  (3),  S_w = sum w h (x - c)^T  (9),   h = A_i^T grad_y M
// (d value / d p = sum_i [alpha (F_i - mF) + beta (M_i - mM)] dM_i/dp with alpha, beta made of the six scalar sums; the host
// forms alpha s_F + beta s_M - (alpha mF + beta mM) s_1 and hands the result to the same optimiser as mean squares).
//
// The per-sample function uses nothing but plain C++ so that tests/emu can run it on the host; only the block reduction
// at the end of the kernel is CUDA-specific (B200_HOST_EMU: per-thread partials are written out instead).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace b200 {

struct CorrGeom {
    int nx, ny, nz;
    double origin[3];
    double i2p[9];  // Direction * diag(Spacing)
    double p2i[9];  // inverse
};
struct CorrPose {
    double A[9], b[3];  // fixed physical point -> moving physical point
    double Bt[9];       // transpose of the moving-initial matrix
    double c[3];        // centre of the optimised transform
};
constexpr int LINREG_CORR_NV = 42;

// One sampled fixed voxel q: maps it into the moving image and, when it is a valid sample (inside the masks and the moving buffer),
// returns the fixed value, the trilinear moving value, h = A_i^T grad_y M and x - c.  Shared by the correlation and Mattes kernels.
struct LinregPoint {
    double fval, mval, h[3], xc[3];
};
__device__ __forceinline__ bool linreg_sample_point(const float* __restrict__ F, const float* __restrict__ M, const uint8_t* __restrict__ fmask,
                                                    const uint8_t* __restrict__ mmask, const CorrGeom& gf, const CorrGeom& gm, const CorrPose& ps, size_t q,
                                                    LinregPoint& pt)
{
    if (fmask && fmask[q] == 0) return false;
    const size_t plane = (size_t)gf.nx * gf.ny;
    const int k = (int)(q / plane), j = (int)((q % plane) / gf.nx), i = (int)(q % gf.nx);
    double x[3], y[3], c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {  // ImageBase::TransformIndexToPhysicalPoint
        double sum = 0.0;
        sum += gf.i2p[r * 3 + 0] * (double)i;
        sum += gf.i2p[r * 3 + 1] * (double)j;
        sum += gf.i2p[r * 3 + 2] * (double)k;
        x[r] = sum + gf.origin[r];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) y[r] = ps.A[r * 3 + 0] * x[0] + ps.A[r * 3 + 1] * x[1] + ps.A[r * 3 + 2] * x[2] + ps.b[r];
    const double v0 = y[0] - gm.origin[0], v1 = y[1] - gm.origin[1], v2 = y[2] - gm.origin[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {  // TransformPhysicalPointToContinuousIndex
        double sum = 0.0;
        sum += gm.p2i[r * 3 + 0] * v0;
        sum += gm.p2i[r * 3 + 1] * v1;
        sum += gm.p2i[r * 3 + 2] * v2;
        c[r] = sum;
    }
    // ImageFunction::IsInsideBuffer: [-0.5, size - 0.5); NaN -> outside
    if (!(c[0] >= -0.5 && c[0] < (double)gm.nx - 0.5 && c[1] >= -0.5 && c[1] < (double)gm.ny - 0.5 && c[2] >= -0.5 && c[2] < (double)gm.nz - 0.5)) return false;
    if (mmask) {
        const int i0 = (int)floor(c[0] + 0.5), i1 = (int)floor(c[1] + 0.5), i2 = (int)floor(c[2] + 0.5);
        if (mmask[((size_t)i2 * gm.ny + i1) * gm.nx + i0] == 0) return false;
    }
    // LinearInterpolateImageFunction: base clamped up to 0 (distance 0 there), upper neighbour clamped at the far edge
    int b[3], u[3];
    double d[3];
    const int nmax[3] = { gm.nx - 1, gm.ny - 1, gm.nz - 1 };
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double cc = c[r] < 0.0 ? 0.0 : c[r];
        const double fl = floor(cc);
        b[r] = (int)fl;
        d[r] = cc - fl;
        u[r] = b[r] + 1 < nmax[r] ? b[r] + 1 : nmax[r];
    }
    const size_t sy = (size_t)gm.nx, sz = (size_t)gm.nx * gm.ny;
    const size_t r00 = (size_t)b[2] * sz + (size_t)b[1] * sy, r10 = (size_t)b[2] * sz + (size_t)u[1] * sy;
    const size_t r01 = (size_t)u[2] * sz + (size_t)b[1] * sy, r11 = (size_t)u[2] * sz + (size_t)u[1] * sy;
    const double v000 = (double)M[r00 + b[0]], v100 = (double)M[r00 + u[0]], v010 = (double)M[r10 + b[0]], v110 = (double)M[r10 + u[0]];
    const double v001 = (double)M[r01 + b[0]], v101 = (double)M[r01 + u[0]], v011 = (double)M[r11 + b[0]], v111 = (double)M[r11 + u[0]];
    const double a00 = v100 - v000, a10 = v110 - v010, a01 = v101 - v001, a11 = v111 - v011;
    const double vx00 = v000 + a00 * d[0], vx10 = v010 + a10 * d[0], vx01 = v001 + a01 * d[0], vx11 = v011 + a11 * d[0];
    const double vxx0 = vx00 + (vx10 - vx00) * d[1], vxx1 = vx01 + (vx11 - vx01) * d[1];
    pt.mval = vxx0 + (vxx1 - vxx0) * d[2];
    const double gx0 = a00 + (a10 - a00) * d[1], gx1 = a01 + (a11 - a01) * d[1];
    const double gi[3] = { gx0 + (gx1 - gx0) * d[2], (vx10 - vx00) + ((vx11 - vx01) - (vx10 - vx00)) * d[2], vxx1 - vxx0 };
    double gy[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) gy[r] = gi[0] * gm.p2i[0 * 3 + r] + gi[1] * gm.p2i[1 * 3 + r] + gi[2] * gm.p2i[2 * 3 + r];
#pragma unroll
    for (int r = 0; r < 3; ++r) pt.h[r] = ps.Bt[r * 3 + 0] * gy[0] + ps.Bt[r * 3 + 1] * gy[1] + ps.Bt[r * 3 + 2] * gy[2];
    pt.fval = (double)F[q];
#pragma unroll
    for (int r = 0; r < 3; ++r) pt.xc[r] = x[r] - ps.c[r];
    return true;
}

__device__ __forceinline__ void corr_sample(const float* __restrict__ F, const float* __restrict__ M, const uint8_t* __restrict__ fmask,
                                            const uint8_t* __restrict__ mmask, const CorrGeom& gf, const CorrGeom& gm, const CorrPose& ps, size_t q,
                                            double* acc)
{
    LinregPoint pt;
    if (!linreg_sample_point(F, M, fmask, mmask, gf, gm, ps, q, pt)) return;
    const double fval = pt.fval, mval = pt.mval;
    acc[0] += 1.0;
    acc[1] += fval;
    acc[2] += mval;
    acc[3] += fval * fval;
    acc[4] += mval * mval;
    acc[5] += fval * mval;
    const double wt[3] = { 1.0, fval, mval };
#pragma unroll
    for (int w = 0; w < 3; ++w) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double wr = wt[w] * pt.h[r];
            acc[6 + 12 * w + r] += wr;
            acc[6 + 12 * w + 3 + r * 3 + 0] += wr * pt.xc[0];
            acc[6 + 12 * w + 3 + r * 3 + 1] += wr * pt.xc[1];
            acc[6 + 12 * w + 3 + r * 3 + 2] += wr * pt.xc[2];
        }
    }
}

// partials: [gridDim.x][LINREG_CORR_NV] block sums (under the host emulation: one row per thread)
__global__ void __launch_bounds__(128) linreg_corr_kernel(const float* __restrict__ F, const float* __restrict__ M, const uint8_t* __restrict__ fmask,
                                                          const uint8_t* __restrict__ mmask, const __grid_constant__ CorrGeom gf,
                                                          const __grid_constant__ CorrGeom gm, const __grid_constant__ CorrPose ps, int stride,
                                                          size_t nsamples, double* __restrict__ partials)
{
    double acc[LINREG_CORR_NV];
#pragma unroll
    for (int v = 0; v < LINREG_CORR_NV; ++v) acc[v] = 0.0;
    for (size_t sidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < nsamples; sidx += (size_t)gridDim.x * blockDim.x)
        corr_sample(F, M, fmask, mmask, gf, gm, ps, sidx * (size_t)stride, acc);
#ifdef B200_HOST_EMU
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int v = 0; v < LINREG_CORR_NV; ++v) partials[tid * LINREG_CORR_NV + v] = acc[v];
#else
    __shared__ double sh[LINREG_CORR_NV][4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < LINREG_CORR_NV; ++v) {
        const double t = warp_sum(acc[v]);
        if (lane == 0) sh[v][wid] = t;
    }
    __syncthreads();
    if (threadIdx.x < LINREG_CORR_NV) {
        double t = 0.0;
        for (int w4 = 0; w4 < 4; ++w4) t += sh[threadIdx.x][w4];
        partials[(size_t)blockIdx.x * LINREG_CORR_NV + threadIdx.x] = t;
    }
#endif
}

// fixed-order sum of the block partials: one thread per accumulator
__global__ void linreg_corr_final_kernel(const double* __restrict__ partials, int nb, double* __restrict__ out)
{
    const int v = threadIdx.x;
    if (v >= LINREG_CORR_NV) return;
    double t = 0.0;
    for (int q = 0; q < nb; ++q) t += partials[(size_t)q * LINREG_CORR_NV + v];
    out[v] = t;
}

}  // namespace b200
