// deriche.cuh -- itk::SmoothingRecursiveGaussianImageFilter (Deriche 4th-order IIR, zero order) on a
// 3-component f64 field, reference deformable.py:157-159.
#pragma once
#include "common.cuh"

namespace b200 {

struct DericheC {
    double N0, N1, N2, N3, D1, D2, D3, D4, M1, M2, M3, M4, BN1, BN2, BN3, BN4, BM1, BM2, BM3, BM4;
};

// RecursiveGaussianImageFilter::SetUp (ZeroOrder, NormalizeAcrossScale off); sigma physical.
inline DericheC deriche_setup(double sigma, double spacing)
{
    const double A1 = 1.3530, B1 = 1.8151, W1 = 0.6681, L1 = -1.3932;
    const double A2 = -0.3531, B2 = 0.0902, W2 = 2.0787, L2 = -1.3732;
    if (spacing < 0.0) spacing = -spacing;
    const double sigmad = sigma / spacing;
    const double Sin1 = std::sin(W1 / sigmad), Sin2 = std::sin(W2 / sigmad), Cos1 = std::cos(W1 / sigmad), Cos2 = std::cos(W2 / sigmad);
    const double Exp1 = std::exp(L1 / sigmad), Exp2 = std::exp(L2 / sigmad);
    DericheC c;
    c.D4 = Exp1 * Exp1 * Exp2 * Exp2;
    c.D3 = -2 * Cos1 * Exp1 * Exp2 * Exp2;
    c.D3 += -2 * Cos2 * Exp2 * Exp1 * Exp1;
    c.D2 = 4 * Cos2 * Cos1 * Exp1 * Exp2;
    c.D2 += Exp1 * Exp1 + Exp2 * Exp2;
    c.D1 = -2 * (Exp2 * Cos2 + Exp1 * Cos1);
    const double SD = 1.0 + c.D1 + c.D2 + c.D3 + c.D4;
    c.N0 = A1 + A2;
    c.N1 = Exp2 * (B2 * Sin2 - (A2 + 2 * A1) * Cos2);
    c.N1 += Exp1 * (B1 * Sin1 - (A1 + 2 * A2) * Cos1);
    c.N2 = (A1 + A2) * Cos2 * Cos1;
    c.N2 -= B1 * Cos2 * Sin1 + B2 * Cos1 * Sin2;
    c.N2 *= 2 * Exp1 * Exp2;
    c.N2 += A2 * Exp1 * Exp1 + A1 * Exp2 * Exp2;
    c.N3 = Exp2 * Exp1 * Exp1 * (B2 * Sin2 - A2 * Cos2);
    c.N3 += Exp1 * Exp2 * Exp2 * (B1 * Sin1 - A1 * Cos1);
    const double SN = c.N0 + c.N1 + c.N2 + c.N3;
    const double alpha0 = 2 * SN / SD - c.N0;
    c.N0 *= 1.0 / alpha0;
    c.N1 *= 1.0 / alpha0;
    c.N2 *= 1.0 / alpha0;
    c.N3 *= 1.0 / alpha0;
    c.M1 = c.N1 - c.D1 * c.N0;
    c.M2 = c.N2 - c.D2 * c.N0;
    c.M3 = c.N3 - c.D3 * c.N0;
    c.M4 = -c.D4 * c.N0;
    const double SN2 = c.N0 + c.N1 + c.N2 + c.N3;
    const double SM = c.M1 + c.M2 + c.M3 + c.M4;
    const double SD2 = 1.0 + c.D1 + c.D2 + c.D3 + c.D4;
    c.BN1 = c.D1 * SN2 / SD2;
    c.BN2 = c.D2 * SN2 / SD2;
    c.BN3 = c.D3 * SN2 / SD2;
    c.BN4 = c.D4 * SN2 / SD2;
    c.BM1 = c.D1 * SM / SD2;
    c.BM2 = c.D2 * SM / SD2;
    c.BM3 = c.D3 * SM / SD2;
    c.BM4 = c.D4 * SM / SD2;
    return c;
}

// RecursiveSeparableImageFilter::FilterDataArray, one thread per (line, component).  The line runs along
// AXIS with stride `sa`; threads are laid out over the two other axes (a: fastest).
// Causal pass writes out, anticausal pass re-reads the input and adds.
__global__ void __launch_bounds__(128) deriche_line_kernel(const double* __restrict__ in, double* __restrict__ out, int ln, size_t sa, int na,
                                                            size_t stride_a, int nb, size_t stride_b, size_t plane,
                                                            const __grid_constant__ DericheC c)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= na || b >= nb) return;
    const size_t base = (size_t)blockIdx.z * plane + (size_t)a * stride_a + (size_t)b * stride_b;
    const double* x = in + base;
    double* y = out + base;
    // causal
    {
        const double v1 = x[0];
        double xm1 = v1, xm2 = v1, xm3 = v1;
        double sm1 = 0, sm2 = 0, sm3 = 0, sm4 = 0;
        for (int i = 0; i < ln; ++i) {
            const double xi = x[(size_t)i * sa];
            double acc = xi * c.N0 + xm1 * c.N1 + xm2 * c.N2 + xm3 * c.N3;
            const double t1 = i >= 1 ? sm1 * c.D1 : v1 * c.BN1;
            const double t2 = i >= 2 ? sm2 * c.D2 : v1 * c.BN2;
            const double t3 = i >= 3 ? sm3 * c.D3 : v1 * c.BN3;
            const double t4 = i >= 4 ? sm4 * c.D4 : v1 * c.BN4;
            acc -= t1 + t2 + t3 + t4;
            y[(size_t)i * sa] = acc;
            xm3 = xm2; xm2 = xm1; xm1 = xi;
            sm4 = sm3; sm3 = sm2; sm2 = sm1; sm1 = acc;
        }
    }
    // anti-causal
    {
        const double v2 = x[(size_t)(ln - 1) * sa];
        double xp1 = v2, xp2 = v2, xp3 = v2, xp4 = v2;
        double sp1 = 0, sp2 = 0, sp3 = 0, sp4 = 0;
        for (int i = ln - 1; i >= 0; --i) {
            const int m = ln - 1 - i;  // number of valid outputs after i
            double acc = xp1 * c.M1 + xp2 * c.M2 + xp3 * c.M3 + xp4 * c.M4;
            const double t1 = m >= 1 ? sp1 * c.D1 : v2 * c.BM1;
            const double t2 = m >= 2 ? sp2 * c.D2 : v2 * c.BM2;
            const double t3 = m >= 3 ? sp3 * c.D3 : v2 * c.BM3;
            const double t4 = m >= 4 ? sp4 * c.D4 : v2 * c.BM4;
            acc -= t1 + t2 + t3 + t4;
            const double xi = x[(size_t)i * sa];
            y[(size_t)i * sa] += acc;
            xp4 = xp3; xp3 = xp2; xp2 = xp1; xp1 = xi;
            sp4 = sp3; sp3 = sp2; sp2 = sp1; sp1 = acc;
        }
    }
}

// Lines along x (the contiguous axis): one WARP owns 32 lines (32 consecutive y of one z / component) and walks
// them in chunks of 32 samples staged through a padded shared-memory tile, so that global loads and stores are
// coalesced along x while each lane still runs the sequential recurrence of its own line.  Same arithmetic as
// deriche_line_kernel.
constexpr int DX_WARPS = 4;
__global__ void __launch_bounds__(32 * DX_WARPS, 6) deriche_x_kernel(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, size_t nlines,
                                                                    const __grid_constant__ DericheC c)
{
    __shared__ double tile[DX_WARPS][32][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t line0 = ((size_t)blockIdx.x * DX_WARPS + wid) * 32;  // first line of this warp (lines = rows of nx samples)
    if (line0 >= nlines) return;
    double(*T)[33] = tile[wid];
    const size_t myline = line0 + lane;
    const bool mine = myline < nlines;
    const int nchunk = (nx + 31) / 32;
    // ---- causal
    {
        double v1 = 0.0, xm1 = 0.0, xm2 = 0.0, xm3 = 0.0, sm1 = 0, sm2 = 0, sm3 = 0, sm4 = 0;
        for (int ch = 0; ch < nchunk; ++ch) {
            const int x0 = ch * 32, x = x0 + lane;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const size_t l = line0 + i;
                T[i][lane] = (l < nlines && x < nx) ? in[l * nx + x] : 0.0;
            }
            __syncwarp();
            if (mine) {
                if (ch == 0) {
                    v1 = T[lane][0];
                    xm1 = xm2 = xm3 = v1;
                }
                const int lim = min(32, nx - x0);
#pragma unroll 4
                for (int j = 0; j < lim; ++j) {
                    const int i = x0 + j;
                    const double xi = T[lane][j];
                    double acc = xi * c.N0 + xm1 * c.N1 + xm2 * c.N2 + xm3 * c.N3;
                    const double t1 = i >= 1 ? sm1 * c.D1 : v1 * c.BN1;
                    const double t2 = i >= 2 ? sm2 * c.D2 : v1 * c.BN2;
                    const double t3 = i >= 3 ? sm3 * c.D3 : v1 * c.BN3;
                    const double t4 = i >= 4 ? sm4 * c.D4 : v1 * c.BN4;
                    acc -= t1 + t2 + t3 + t4;
                    T[lane][j] = acc;
                    xm3 = xm2; xm2 = xm1; xm1 = xi;
                    sm4 = sm3; sm3 = sm2; sm2 = sm1; sm1 = acc;
                }
            }
            __syncwarp();
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const size_t l = line0 + i;
                if (l < nlines && x < nx) out[l * nx + x] = T[i][lane];
            }
            __syncwarp();
        }
    }
    // ---- anti-causal (right to left); out += anticausal
    {
        double v2 = 0.0, xp1 = 0.0, xp2 = 0.0, xp3 = 0.0, xp4 = 0.0, sp1 = 0, sp2 = 0, sp3 = 0, sp4 = 0;
        for (int ch = nchunk - 1; ch >= 0; --ch) {
            const int x0 = ch * 32, x = x0 + lane;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const size_t l = line0 + i;
                T[i][lane] = (l < nlines && x < nx) ? in[l * nx + x] : 0.0;
            }
            __syncwarp();
            if (mine) {
                const int lim = min(32, nx - x0);
                if (ch == nchunk - 1) {
                    v2 = T[lane][lim - 1];
                    xp1 = xp2 = xp3 = xp4 = v2;
                }
#pragma unroll 4
                for (int j = lim - 1; j >= 0; --j) {
                    const int i = x0 + j;
                    const int m = nx - 1 - i;
                    double acc = xp1 * c.M1 + xp2 * c.M2 + xp3 * c.M3 + xp4 * c.M4;
                    const double t1 = m >= 1 ? sp1 * c.D1 : v2 * c.BM1;
                    const double t2 = m >= 2 ? sp2 * c.D2 : v2 * c.BM2;
                    const double t3 = m >= 3 ? sp3 * c.D3 : v2 * c.BM3;
                    const double t4 = m >= 4 ? sp4 * c.D4 : v2 * c.BM4;
                    acc -= t1 + t2 + t3 + t4;
                    const double xi = T[lane][j];
                    T[lane][j] = acc;
                    xp4 = xp3; xp3 = xp2; xp2 = xp1; xp1 = xi;
                    sp4 = sp3; sp3 = sp2; sp2 = sp1; sp1 = acc;
                }
            }
            __syncwarp();
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const size_t l = line0 + i;
                if (l < nlines && x < nx) out[l * nx + x] += T[i][lane];
            }
            __syncwarp();
        }
    }
}

// axis order z, x, y (SmoothingRecursiveGaussianImageFilter: first filter on the last dimension, then
// dimensions 0 .. N-2); sigma is physical, ITK divides by the spacing of each axis.
inline int recursive_gaussian_vec3(b200reg_ctx* ctx, double* field, const b200reg_geom& g, const double* sigma)
{
    const int nx = g.size[0], ny = g.size[1], nz = g.size[2];
    if (nx < 4 || ny < 4 || nz < 4)
        return set_error(B200REG_ERR_RUNTIME, "RecursiveGaussianImageFilter: the number of pixels along a direction is less than 4");
    const size_t n = nvox(g);
    // field -> t1 -> t2 -> field: the third pass writes the result in place, no copy back
    TempBuf t1, t2;
    B200_TRY(t1.alloc(ctx, 3 * n * sizeof(double)));
    B200_TRY(t2.alloc(ctx, 3 * n * sizeof(double)));
    double* bufs[4] = { field, t1.as<double>(), t2.as<double>(), field };
    const int order_zxy[3] = { 2, 0, 1 }, order_xyz[3] = { 0, 1, 2 };  // semantic switch recursive_gaussian_axis_order
    const int* order = semantics().recursive_gaussian_axis_order ? order_xyz : order_zxy;
    const size_t strides[3] = { 1, (size_t)nx, (size_t)nx * ny };
    const int dims[3] = { nx, ny, nz };
    for (int pass = 0; pass < 3; ++pass) {
        const double* src = bufs[pass];
        double* dst = bufs[pass + 1];
        const int axis = order[pass];
        const DericheC c = deriche_setup(sigma[axis], g.spacing[axis]);
        // thread axes: a = the lowest remaining axis (x unless axis == 0), b = the other
        const int aa = axis == 0 ? 1 : 0;
        const int ab = axis == 2 ? 1 : 2;
        if (axis == 0) {
            const size_t nlines = (size_t)ny * nz * 3;  // the three component volumes are contiguous: lines are rows of nx samples
            const unsigned nblk = (unsigned)((nlines + 32 * DX_WARPS - 1) / (32 * DX_WARPS));
            deriche_x_kernel<<<nblk, 32 * DX_WARPS, 0, ctx->stream>>>(src, dst, nx, ny, nlines, c);
        } else {
            dim3 blk(128, 1, 1), grd((dims[aa] + 127) / 128, dims[ab], 3);
            deriche_line_kernel<<<grd, blk, 0, ctx->stream>>>(src, dst, dims[axis], strides[axis], dims[aa], strides[aa], dims[ab],
                                                              strides[ab], n, c);
        }
        ctx->launches++;
        B200_CHECK_LAUNCH();
    }
    return B200REG_OK;
}

}  // namespace b200
