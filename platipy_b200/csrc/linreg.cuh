// linreg.cuh -- mean-squares similarity of linear_registration (reference linear.py:50-260:
// sitk.ImageRegistrationMethod with SetMetricAsMeanSquares, linear interpolation, REGULAR sampling; the step before
// Demons in every atlas pipeline, multiatlas/run.py:229,271).
//
// For any transform of the family the reference offers (translation, versor-rigid, similarity, affine, scale...)
// the optimised map is u = R(p) (x - c) + c + t, followed by the fixed moving-initial transform y = A_i u + b_i.
// One pass over the sampled fixed voxels therefore accumulates everything every parameterisation needs:
//     value   sum (M(y) - F(x))^2, count
//     s       sum w            with w = 2 (M - F) A_i^T grad_y M        (derivative w.r.t. the translation)
//     S       sum w (x - c)^T                                           (dV/dp_k = <dR/dp_k, S>_F for matrix parameters)
// and the (tiny) parameter algebra stays on the host.  Sums are fixed-order block partials -> deterministic.
#pragma once
#include "common.cuh"
#include "resample.cuh"

namespace b200 {

struct LinRegPose {
    double A[9], b[3];  // fixed physical point -> moving physical point (optimised transform, then moving-initial transform)
    double Bt[9];       // transpose of the moving-initial matrix
    double c[3];        // centre of the optimised transform
};
constexpr int LINREG_NV = 14;

__global__ void __launch_bounds__(256) linreg_meansq_kernel(const float* __restrict__ F, const float* __restrict__ M, const uint8_t* __restrict__ fmask,
                                                            const uint8_t* __restrict__ mmask, const __grid_constant__ GeomD gf,
                                                            const __grid_constant__ GeomD gm, const __grid_constant__ LinRegPose ps, int stride,
                                                            size_t nsamples, double* __restrict__ partials)
{
    double acc[LINREG_NV];
#pragma unroll
    for (int v = 0; v < LINREG_NV; ++v) acc[v] = 0.0;
    const size_t plane = (size_t)gf.nx * gf.ny;
    for (size_t sidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < nsamples; sidx += (size_t)gridDim.x * blockDim.x) {
        const size_t q = sidx * (size_t)stride;
        if (fmask && fmask[q] == 0) continue;
        const int k = (int)(q / plane), j = (int)((q % plane) / gf.nx), i = (int)(q % gf.nx);
        double x[3], y[3], c[3];
        idx2pt(gf, (double)i, (double)j, (double)k, x);
#pragma unroll
        for (int r = 0; r < 3; ++r) y[r] = ps.A[r * 3 + 0] * x[0] + ps.A[r * 3 + 1] * x[1] + ps.A[r * 3 + 2] * x[2] + ps.b[r];
        pt2cidx(gm, y, c);
        if (!inside_buffer(gm, c)) continue;
        if (mmask) {
            const int i0 = (int)floor(c[0] + 0.5), i1 = (int)floor(c[1] + 0.5), i2 = (int)floor(c[2] + 0.5);
            if (mmask[((size_t)i2 * gm.ny + i1) * gm.nx + i0] == 0) continue;
        }
        const LinW w = lin_setup(gm, c);
        const size_t sy = (size_t)gm.nx, sz = (size_t)gm.nx * gm.ny;
        const size_t r00 = (size_t)w.b2 * sz + (size_t)w.b1 * sy, r10 = (size_t)w.b2 * sz + (size_t)w.u1 * sy;
        const size_t r01 = (size_t)w.u2 * sz + (size_t)w.b1 * sy, r11 = (size_t)w.u2 * sz + (size_t)w.u1 * sy;
        const double v000 = (double)M[r00 + w.b0], v100 = (double)M[r00 + w.u0], v010 = (double)M[r10 + w.b0], v110 = (double)M[r10 + w.u0];
        const double v001 = (double)M[r01 + w.b0], v101 = (double)M[r01 + w.u0], v011 = (double)M[r11 + w.b0], v111 = (double)M[r11 + w.u0];
        const double a00 = v100 - v000, a10 = v110 - v010, a01 = v101 - v001, a11 = v111 - v011;
        const double vx00 = v000 + a00 * w.d0, vx10 = v010 + a10 * w.d0, vx01 = v001 + a01 * w.d0, vx11 = v011 + a11 * w.d0;
        const double vxx0 = vx00 + (vx10 - vx00) * w.d1, vxx1 = vx01 + (vx11 - vx01) * w.d1;
        const double mval = vxx0 + (vxx1 - vxx0) * w.d2;
        // gradient of the trilinear interpolant w.r.t. the continuous index
        const double gx0 = a00 + (a10 - a00) * w.d1, gx1 = a01 + (a11 - a01) * w.d1;
        const double gi[3] = { gx0 + (gx1 - gx0) * w.d2, (vx10 - vx00) + ((vx11 - vx01) - (vx10 - vx00)) * w.d2, vxx1 - vxx0 };
        // -> moving physical space (c = P2I (y - o))  -> space of the optimised transform's output (h = A_i^T grad_y)
        double gy[3], h[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) gy[r] = gi[0] * gm.p2i[0 * 3 + r] + gi[1] * gm.p2i[1 * 3 + r] + gi[2] * gm.p2i[2 * 3 + r];
#pragma unroll
        for (int r = 0; r < 3; ++r) h[r] = ps.Bt[r * 3 + 0] * gy[0] + ps.Bt[r * 3 + 1] * gy[1] + ps.Bt[r * 3 + 2] * gy[2];
        const double d = mval - (double)F[q];
        acc[0] += d * d;
        acc[1] += 1.0;
        const double xc[3] = { x[0] - ps.c[0], x[1] - ps.c[1], x[2] - ps.c[2] };
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double wr = 2.0 * d * h[r];
            acc[2 + r] += wr;
            acc[5 + r * 3 + 0] += wr * xc[0];
            acc[5 + r * 3 + 1] += wr * xc[1];
            acc[5 + r * 3 + 2] += wr * xc[2];
        }
    }
    __shared__ double sh[LINREG_NV][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < LINREG_NV; ++v) {
        const double t = warp_sum(acc[v]);
        if (lane == 0) sh[v][wid] = t;
    }
    __syncthreads();
    if (threadIdx.x < LINREG_NV) {
        double t = 0.0;
        for (int w8 = 0; w8 < 8; ++w8) t += sh[threadIdx.x][w8];
        partials[(size_t)blockIdx.x * LINREG_NV + threadIdx.x] = t;
    }
}
__global__ void linreg_final_kernel(const double* __restrict__ partials, int nb, double* __restrict__ out)
{
    const int v = threadIdx.x;
    if (v >= LINREG_NV) return;
    double t = 0.0;
    for (int q = 0; q < nb; ++q) t += partials[(size_t)q * LINREG_NV + v];
    out[v] = t;
}

inline int linreg_meansq(b200reg_ctx* ctx, const float* F, const b200reg_geom& gF, const float* M, const b200reg_geom& gM, const LinRegPose& ps,
                         const uint8_t* fmask, const uint8_t* mmask, int stride, double* h_out)
{
    const size_t n = nvox(gF);
    const size_t nsamples = (n + (size_t)stride - 1) / (size_t)stride;
    int nb = (int)((nsamples + 255) / 256);
    if (nb > ctx->sm_count * 8) nb = ctx->sm_count * 8;
    if (nb < 1) nb = 1;
    TempBuf part, out;
    B200_TRY(part.alloc(ctx, sizeof(double) * LINREG_NV * (size_t)nb));
    B200_TRY(out.alloc(ctx, sizeof(double) * LINREG_NV));
    linreg_meansq_kernel<<<nb, 256, 0, ctx->stream>>>(F, M, fmask, mmask, make_geomd(gF), make_geomd(gM), ps, stride, nsamples, part.as<double>());
    linreg_final_kernel<<<1, 32, 0, ctx->stream>>>(part.as<double>(), nb, out.as<double>());
    ctx->launches += 2;
    B200_CHECK_LAUNCH();
    B200_CUDA(small_d2h(ctx, ctx->h_scratch, out.p, sizeof(double) * LINREG_NV));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int v = 0; v < LINREG_NV; ++v) h_out[v] = ctx->h_scratch[v];
    return B200REG_OK;
}

}  // namespace b200
