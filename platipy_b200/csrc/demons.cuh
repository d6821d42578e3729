// demons.cuh -- the Fast-Symmetric-Forces Demons inner loop (itk::FastSymmetricForcesDemonsRegistrationFilter
// with itk::ESMDemonsRegistrationFunction), reference deformable.py:244-257,143-149.
//
// Per iteration (all fields f64 SoA, images f32):
//   warp     W = float32(trilinear(M, x + D(x)))  or FLT_MAX outside the moving buffer      (WarpImageFilter)
//   force    U = ESM symmetric update from F, W and their finite differences; per-block partial sums
//            of SSD, voxel count, |U|^2                                                       (CalculateChange)
//   finish   fixed-order sum of the partials -> metric, RMS change, halt decision (device resident)
//   smooth   U <- G_u * U (x,y,z) ; D <- G_d * (D + U) (x,y,z)                                (ApplyUpdate)
// The whole level is enqueued without host synchronisation: kernels of iteration `it` return
// immediately when it >= ctrl->halt_iter (DenseFiniteDifferenceImageFilter::Halt evaluated on device).
#pragma once
#include "common.cuh"
#include "gauss.cuh"
#include "gauss_fast.cuh"
#include "resample.cuh"
#include "demons_split.cuh"

namespace b200 {

// WarpImageFilter with the field on the output (fixed) grid: point = index->physical + D; linear
// interpolation of the moving image; edge padding NumericTraits<float>::max().
template <bool SMALL>
__global__ void __launch_bounds__(BX* BY) demons_warp_kernel(const float* __restrict__ M, const double* __restrict__ D, float* __restrict__ W,
                                                              const __grid_constant__ GeomD gf, const __grid_constant__ GeomD gm,
                                                              const DemonsCtrl* __restrict__ ctrl, int it)
{
    if (it >= ctrl->halt_iter) return;
    const int i = blockIdx.x * BX + threadIdx.x;
    const int j = blockIdx.y * BY + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= gf.nx || j >= gf.ny) return;
    const size_t n = (size_t)gf.nx * gf.ny * gf.nz;
    const size_t o = ((size_t)k * gf.ny + j) * gf.nx + i;
    double p[3], c[3];
    idx2pt(gf, (double)i, (double)j, (double)k, p);
    p[0] += D[o];
    p[1] += D[o + n];
    p[2] += D[o + 2 * n];
    pt2cidx(gm, p, c);
    float w = FLT_MAX;
    if (inside_buffer(gm, c)) {
        const LinW lw = lin_setup(gm, c);
        w = (float)(SMALL ? lin_eval_i32<float>(M, gm.nx, gm.nx * gm.ny, lw) : lin_eval<float>(M, gm, lw));
    }
    W[o] = w;
}

// ESMDemonsRegistrationFunction::ComputeUpdate, UseGradientType = Symmetric.
__global__ void __launch_bounds__(BX* BY) demons_force_kernel(const float* __restrict__ F, const float* __restrict__ W, double* __restrict__ U,
                                                               double* __restrict__ partials, const __grid_constant__ GeomD gf,
                                                               const __grid_constant__ ForceParams fp, const DemonsCtrl* __restrict__ ctrl, int it)
{
    if (it >= ctrl->halt_iter) return;
    const int i = blockIdx.x * BX + threadIdx.x;
    const int j = blockIdx.y * BY + threadIdx.y;
    const int k = blockIdx.z;
    double ssd = 0.0, cnt = 0.0, ssc = 0.0;
    if (i < gf.nx && j < gf.ny) {
        const size_t n = (size_t)gf.nx * gf.ny * gf.nz;
        const size_t o = ((size_t)k * gf.ny + j) * gf.nx + i;
        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
        const float mv = W[o];
        if (mv != FLT_MAX) {
            const double fixedValue = (double)F[o];
            const double movingValue = (double)mv;
            const int idx[3] = { i, j, k };
            const int dims[3] = { gf.nx, gf.ny, gf.nz };
            const size_t strides[3] = { 1, (size_t)gf.nx, (size_t)gf.nx * gf.ny };
            double g2[3];
#pragma unroll
            for (int dim = 0; dim < 3; ++dim) {
                const int nd = dims[dim];
                const size_t s = strides[dim];
                double wg;
                if (idx[dim] == 0) {
                    if (nd < 2) wg = 0.0;
                    else {
                        const float nb = W[o + s];
                        if (nb == FLT_MAX) wg = 0.0;
                        else {
                            wg = (double)nb - movingValue;
                            wg /= gf.spacing[dim];
                        }
                    }
                } else if (idx[dim] == nd - 1) {
                    const float nb = W[o - s];
                    if (nb == FLT_MAX) wg = 0.0;
                    else {
                        wg = movingValue - (double)nb;
                        wg /= gf.spacing[dim];
                    }
                } else {
                    const float nb = W[o + s];
                    const float pb = W[o - s];
                    if (nb == FLT_MAX) {
                        if (pb == FLT_MAX) wg = 0.0;
                        else {
                            wg = movingValue - (double)pb;  // backward difference
                            wg /= gf.spacing[dim];
                        }
                    } else if (pb == FLT_MAX) {
                        wg = (double)nb - movingValue;  // forward difference
                        wg /= gf.spacing[dim];
                    } else {
                        wg = (double)nb - (double)pb;  // central difference
                        wg *= 0.5 / gf.spacing[dim];
                    }
                }
                // CentralDifferenceImageFunction::EvaluateAtIndex (UseImageDirection off)
                double fg;
                if (idx[dim] < 1 || idx[dim] > nd - 2) fg = 0.0;
                else {
                    fg = (double)F[o + s] - (double)F[o - s];
                    fg *= 0.5 / gf.spacing[dim];
                }
                g2[dim] = fg + wg;
            }
            // TransformLocalVectorToPhysicalVector
            double J[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double sum = 0.0;
                sum += gf.direction[r * 3 + 0] * g2[0];
                sum += gf.direction[r * 3 + 1] * g2[1];
                sum += gf.direction[r * 3 + 2] * g2[2];
                J[r] = sum;
            }
            const double gm2 = J[0] * J[0] + J[1] * J[1] + J[2] * J[2];
            const double speed = fixedValue - movingValue;
            if (!(fabs(speed) < fp.intensity_thresh)) {
                const double denom = (fp.normalizer > 0.0) ? gm2 + (speed * speed) / fp.normalizer : gm2;
                if (!(denom < fp.denom_thresh)) {
                    const double factor = 2.0 * speed / denom;
                    u0 = factor * J[0];
                    u1 = factor * J[1];
                    u2 = factor * J[2];
                }
            }
            ssd = speed * speed;
            cnt = 1.0;
            ssc = u0 * u0 + u1 * u1 + u2 * u2;
        }
        U[o] = u0;
        U[o + n] = u1;
        U[o + 2 * n] = u2;
    }
    // block reduction: warp shuffles, then warp 0 over the per-warp partials (fixed order)
    __shared__ double sh[3][BX * BY / 32];
    const int tid = threadIdx.y * BX + threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    ssd = warp_sum(ssd);
    cnt = warp_sum(cnt);
    ssc = warp_sum(ssc);
    if (lane == 0) {
        sh[0][wid] = ssd;
        sh[1][wid] = cnt;
        sh[2][wid] = ssc;
    }
    __syncthreads();
    if (wid == 0) {
        constexpr int NW = BX * BY / 32;
        double a = lane < NW ? sh[0][lane] : 0.0, b = lane < NW ? sh[1][lane] : 0.0, c = lane < NW ? sh[2][lane] : 0.0;
        a = warp_sum(a);
        b = warp_sum(b);
        c = warp_sum(c);
        if (lane == 0) {
            const size_t bid = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            partials[bid * 3 + 0] = a;
            partials[bid * 3 + 1] = b;
            partials[bid * 3 + 2] = c;
        }
    }
}

// Single block: sums the per-block partials in a fixed order, publishes metric / RMS change, advances the
// elapsed count and evaluates DenseFiniteDifferenceImageFilter::Halt for the NEXT iteration.
__global__ void __launch_bounds__(1024) demons_finish_kernel(const double* __restrict__ partials, size_t nblocks, DemonsCtrl* ctrl,
                                                              double max_rms_error, int it, int n_iters, double* __restrict__ trace)
{
    pdl_launch_dependents();
    pdl_wait();
    if (it >= ctrl->halt_iter) return;
    double a = 0.0, b = 0.0, c = 0.0;
    for (size_t q = threadIdx.x; q < nblocks; q += 1024) {
        a += partials[q * 3 + 0];
        b += partials[q * 3 + 1];
        c += partials[q * 3 + 2];
    }
    __shared__ double sh[3][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    if (lane == 0) {
        sh[0][wid] = a;
        sh[1][wid] = b;
        sh[2][wid] = c;
    }
    __syncthreads();
    if (wid == 0) {
        a = warp_sum(sh[0][lane]);
        b = warp_sum(sh[1][lane]);
        c = warp_sum(sh[2][lane]);
        if (lane == 0) {
            double rms = ctrl->rms;
            if (b > 0.0) {
                ctrl->metric = a / b;
                rms = sqrt(c / b);
                ctrl->rms = rms;
            }
            ctrl->elapsed = it + 1;
            if (trace) {
                trace[2 * it] = ctrl->metric;
                trace[2 * it + 1] = rms;
            }
            // Halt() before iteration it+1: elapsed >= NumberOfIterations, or MaximumRMSError > RMSChange
            if (it + 1 >= n_iters || max_rms_error > rms) ctrl->halt_iter = it + 1;
        }
    }
}

// (The fused z-marching warp + force kernel and its warp-specialised variant lived here until round 2; both measured slower than the
// two high-occupancy kernels of demons_split.cuh -- 1.86-2.1 ms against 0.56 + 0.76 ms per full-resolution iteration, DVF identical,
// profiles/r01_summary.md and profiles/r01_s2_ab_split_vs_fused.log -- and were removed from the product headers; git history has
// them.)

__global__ void demons_ctrl_init_kernel(DemonsCtrl* ctrl, int n_iters)
{
    ctrl->halt_iter = n_iters <= 0 ? 0 : 0x7fffffff;
    ctrl->elapsed = 0;
    ctrl->metric = DBL_MAX;
    ctrl->rms = 0.0;
}

struct DemonsWorkspace {
    TempBuf U, T1, T2, P1, W, partials, ctrl, trace;
    TempBuf fP0, fP1, fU, fT1;  // fast mode: float32 fields
    size_t nblocks = 0;
};

inline int demons_prepare(b200reg_ctx* ctx, const b200reg_geom& gF, int n_iters, DemonsWorkspace* ws, bool want_trace)
{
    const size_t n = nvox(gF);
    B200_TRY(ws->U.alloc(ctx, 3 * n * sizeof(double)));
    B200_TRY(ws->T1.alloc(ctx, 3 * n * sizeof(double)));
    B200_TRY(ws->T2.alloc(ctx, 3 * n * sizeof(double)));
    B200_TRY(ws->P1.alloc(ctx, 3 * n * sizeof(double)));
    B200_TRY(ws->W.alloc(ctx, n * sizeof(float)));
    const dim3 g = grid3(gF.size[0], gF.size[1], gF.size[2]);
    // either update path; the split path appends whole z slices of border blocks (demons_split.cuh)
    ws->nblocks = (size_t)g.x * g.y * g.z + border_blocks(gF.size[0], gF.size[1], gF.size[2]) + (size_t)g.x * g.y + 1;
    B200_TRY(ws->partials.alloc(ctx, ws->nblocks * 3 * sizeof(double)));
    B200_TRY(ws->ctrl.alloc(ctx, sizeof(DemonsCtrl)));
    if (want_trace) B200_TRY(ws->trace.alloc(ctx, sizeof(double) * 2 * (size_t)(n_iters > 0 ? n_iters : 1)));
    return B200REG_OK;
}

inline ForceParams make_force_params(const b200reg_geom& gF, const b200reg_demons_params& p)
{
    ForceParams fp;
    if (p.max_update_step_length > 0.0) {
        double nrm = 0.0;
        for (int k = 0; k < 3; ++k) nrm += gF.spacing[k] * gF.spacing[k];
        nrm *= p.max_update_step_length * p.max_update_step_length / 3.0;
        fp.normalizer = nrm;
    } else fp.normalizer = -1.0;
    for (int k = 0; k < 3; ++k) fp.half_inv_sp[k] = 0.5 / gF.spacing[k];
    {
        // x / 2^k == x * 2^-k exactly (barring overflow/underflow), so a power-of-two normalizer lets the kernel
        // replace the division by a multiplication without changing a single bit
        int e = 0;
        const double m = std::frexp(fp.normalizer, &e);
        fp.inv_normalizer = (fp.normalizer > 0.0 && m == 0.5 && e > -500 && e < 500) ? 1.0 / fp.normalizer : 0.0;
    }
    fp.intensity_thresh = p.intensity_difference_threshold;
    fp.denom_thresh = p.denominator_threshold;
    fp.max_rms_error = p.max_rms_error;
    return fp;
}

inline int demons_calculate_change(b200reg_ctx* ctx, const float* F, const GeomD& gf, const float* M, const GeomD& gm, const double* D,
                                   const ForceParams& fp, DemonsWorkspace* ws, int it, int n_iters, bool want_w = false)
{
    DemonsCtrl* ctrl = ws->ctrl.as<DemonsCtrl>();
    size_t nblocks;
    (void)want_w;  // both paths leave W in the workspace
    const bool fits32 = (size_t)gf.nx * gf.ny * gf.nz * 3 < (1ull << 31) && (size_t)gm.nx * gm.ny * gm.nz < (1ull << 31);
    if (fits32 && !ctx->unfused_force) {
        // two high-occupancy kernels with 32-bit offsets, W through HBM (demons_split.cuh)
        B200_TRY(launch_update_split(ctx, F, gf, M, gm, D, ws->W.as<float>(), ws->U.as<double>(), ws->partials.as<double>(), fp,
                                     geom_is_diag(gf) && geom_is_diag(gm), ctrl, it, &nblocks));
    } else {
        // volumes of 2^31 / 3 voxels and more (64-bit offsets), and B200REG_UNFUSED_FORCE=1: one thread per voxel
        const dim3 g = grid3(gf.nx, gf.ny, gf.nz), b = block3();
        nblocks = (size_t)g.x * g.y * g.z;
        if (gm.small) demons_warp_kernel<true><<<g, b, 0, ctx->stream>>>(M, D, ws->W.as<float>(), gf, gm, ctrl, it);
        else demons_warp_kernel<false><<<g, b, 0, ctx->stream>>>(M, D, ws->W.as<float>(), gf, gm, ctrl, it);
        demons_force_kernel<<<g, b, 0, ctx->stream>>>(F, ws->W.as<float>(), ws->U.as<double>(), ws->partials.as<double>(), gf, fp, ctrl, it);
        ctx->launches += 2;
    }
    B200_CUDA(launch_pdl(ctx, demons_finish_kernel, dim3(1), dim3(1024), 0, ws->partials.as<double>(), nblocks, ctrl, fp.max_rms_error, it, n_iters,
                         ws->trace.as<double>()));
    ctx->launches += 1;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// PDEDeformableRegistrationFilter::Smooth{Update,Displacement}Field: x -> y -> z, variance = sd^2 (voxel
// units), clamp boundary, f64.  `add` != nullptr: the first pass reads src + add (AddImageFilter fused).
// Out of place: src/add -> dst, using s1/s2 as scratch for the separable fallback (dst, s1, s2, src, add
// all distinct).  Small radii take the fused z-marching kernel (one read + one write per voxel).
inline int pde_smooth(b200reg_ctx* ctx, const double* src, const double* add, double* dst, double* s1, double* s2, int nx, int ny, int nz,
                      const KernelCoeffs kc[3], const DemonsCtrl* ctrl, int it)
{
    if (zmarch_supported(kc) && !ctx->force_separable) return launch_conv3d_zmarch(ctx, src, add, dst, nx, ny, nz, 3, kc, ctrl, it);
    if (add) B200_TRY((launch_conv_axis<double, true>(ctx, 0, src, add, s1, nx, ny, nz, 3, kc[0], ctrl, it)));
    else B200_TRY((launch_conv_axis<double, false>(ctx, 0, src, nullptr, s1, nx, ny, nz, 3, kc[0], ctrl, it)));
    B200_TRY((launch_conv_axis<double, false>(ctx, 1, s1, nullptr, s2, nx, ny, nz, 3, kc[1], ctrl, it)));
    B200_TRY((launch_conv_axis<double, false>(ctx, 2, s2, nullptr, dst, nx, ny, nz, 3, kc[2], ctrl, it)));
    return B200REG_OK;
}

__global__ void add_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, size_t n,
                           const DemonsCtrl* __restrict__ ctrl, int it)
{
    if (ctrl && it >= ctrl->halt_iter) return;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = a[q] + b[q];
}
// The field ping-pongs between two buffers; starting in buffer `start` it lives in buffer (start + elapsed) % 2 after `elapsed`
// iterations.  `start` is chosen so that a level that runs all its iterations ends in the caller's buffer P0; only an early halt
// after an odd number of remaining iterations leaves it in P1, and then -- decided on the device -- it is copied.
__global__ void select_copy_kernel(const double* __restrict__ p1, double* __restrict__ p0, size_t n, const DemonsCtrl* __restrict__ ctrl, int start)
{
    if (((start + ctrl->elapsed) & 1) == 0) return;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) p0[q] = p1[q];
}

inline int make_pde_coeffs(const double sd[3], double max_error, int max_width, KernelCoeffs kc[3])
{
    for (int a = 0; a < 3; ++a) B200_TRY(make_coeffs(gaussian_operator(sd[a] * sd[a], max_error, max_width), &kc[a]));
    return B200REG_OK;
}

// ---- fast mode (float32 fields, 92 algorithmic B/voxel/iteration) ------------------------------------------------------
// Same loop, same kernels for warp / force / finish (double arithmetic per voxel, float32 storage of D and U), float32 FMA smoothing
// (gauss_fast.cuh).  NOT a parity path -- selected explicitly (b200reg_demons_params::field_precision = 1) and reported with its
// error against the parity path; whatever a level cannot run this way (radii, row length, size) runs in parity mode.
__global__ void select_cast_kernel(const float* __restrict__ p0, const float* __restrict__ p1, double* __restrict__ out, size_t n,
                                   const DemonsCtrl* __restrict__ ctrl)
{
    const float* __restrict__ src = (ctrl->elapsed & 1) ? p1 : p0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = (double)src[q];
}
inline bool demons_fast_possible(b200reg_ctx* ctx, const b200reg_geom& gF, const b200reg_geom& gM, const b200reg_demons_params& p, KernelCoeffs kd[3],
                                 KernelCoeffs ku[3])
{
    if (!p.smooth_displacement_field || !p.smooth_update_field) return false;
    if (make_pde_coeffs(p.std_dev, p.max_error, p.max_kernel_width, kd) != B200REG_OK) return false;
    if (make_pde_coeffs(p.update_std_dev, p.max_error, p.max_kernel_width, ku) != B200REG_OK) return false;
    const size_t nf = nvox(gF), nm = nvox(gM);
    (void)ctx;
    return zmarch_fast_supported(kd, gF.size[0]) && zmarch_fast_supported(ku, gF.size[0]) && nf * 3 < (1ull << 31) && nm < (1ull << 31);
}
inline int demons_enqueue_fast(b200reg_ctx* ctx, const float* F, const b200reg_geom& gF, const float* M, const b200reg_geom& gM,
                               const b200reg_demons_params& p, const KernelCoeffs kd[3], const KernelCoeffs ku[3], double* D, DemonsWorkspace* ws)
{
    const int nx = gF.size[0], ny = gF.size[1], nz = gF.size[2];
    const size_t n = nvox(gF);
    const GeomD gf = make_geomd(gF), gm = make_geomd(gM);
    const ForceParams fp = make_force_params(gF, p);
    const int n_iters = p.number_of_iterations;
    B200_TRY(ws->fP0.alloc(ctx, 3 * n * sizeof(float)));
    B200_TRY(ws->fP1.alloc(ctx, 3 * n * sizeof(float)));
    B200_TRY(ws->fU.alloc(ctx, 3 * n * sizeof(float)));
    B200_TRY(ws->fT1.alloc(ctx, 3 * n * sizeof(float)));
    float* P[2] = { ws->fP0.as<float>(), ws->fP1.as<float>() };
    float* U = ws->fU.as<float>();
    float* T1 = ws->fT1.as<float>();
    B200_CUDA(cudaMemsetAsync(P[0], 0, 3 * n * sizeof(float), ctx->stream));
    DemonsCtrl* ctrl = ws->ctrl.as<DemonsCtrl>();
    demons_ctrl_init_kernel<<<1, 1, 0, ctx->stream>>>(ctrl, n_iters);
    ctx->launches++;
    const bool diag = geom_is_diag(gf) && geom_is_diag(gm);
    for (int it = 0; it < n_iters; ++it) {
        float* cur = P[it & 1];
        float* nxt = P[(it + 1) & 1];
        size_t nblocks;
        B200_TRY(launch_update_split<float>(ctx, F, gf, M, gm, cur, ws->W.as<float>(), U, ws->partials.as<double>(), fp, diag, ctrl, it, &nblocks));
        B200_CUDA(launch_pdl(ctx, demons_finish_kernel, dim3(1), dim3(1024), 0, ws->partials.as<double>(), nblocks, ctrl, fp.max_rms_error, it, n_iters,
                             ws->trace.as<double>()));
        ctx->launches += 1;
        B200_TRY(launch_conv3d_zmarch_fast(ctx, U, cur, T1, nx, ny, nz, 3, ku, ctrl, it));      // T1 = D + G_u * U
        B200_TRY(launch_conv3d_zmarch_fast(ctx, T1, nullptr, nxt, nx, ny, nz, 3, kd, ctrl, it));  // D' = G_d * T1
    }
    select_cast_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(P[0], P[1], D, 3 * n, ctrl);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// registration_algorithm.Execute(f_image, m_image): zero initial field, FiniteDifferenceImageFilter loop.
// Everything is enqueued on the stream; stats are read back by the caller after synchronising.
// Per iteration: CalculateChange (P[it%2] -> U), SmoothUpdateField (U -> T1), Add + SmoothDisplacementField
// ((P[it%2] + T1) -> P[(it+1)%2]).
inline int demons_enqueue(b200reg_ctx* ctx, const float* F, const b200reg_geom& gF, const float* M, const b200reg_geom& gM,
                          const b200reg_demons_params& p, double* D, DemonsWorkspace* ws)
{
    const int nx = gF.size[0], ny = gF.size[1], nz = gF.size[2];
    const size_t n = nvox(gF);
    const GeomD gf = make_geomd(gF), gm = make_geomd(gM);
    const ForceParams fp = make_force_params(gF, p);
    const int n_iters = p.number_of_iterations;
    KernelCoeffs kd[3], ku[3];
    if (p.field_precision == 1 && demons_fast_possible(ctx, gF, gM, p, kd, ku)) return demons_enqueue_fast(ctx, F, gF, M, gM, p, kd, ku, D, ws);
    if (p.smooth_displacement_field) B200_TRY(make_pde_coeffs(p.std_dev, p.max_error, p.max_kernel_width, kd));
    if (p.smooth_update_field) B200_TRY(make_pde_coeffs(p.update_std_dev, p.max_error, p.max_kernel_width, ku));
    double* P[2] = { D, ws->P1.as<double>() };
    const int start = n_iters & 1;  // an odd number of iterations starts in P1 and ends in the caller's buffer
    B200_CUDA(cudaMemsetAsync(P[start], 0, 3 * n * sizeof(double), ctx->stream));
    DemonsCtrl* ctrl = ws->ctrl.as<DemonsCtrl>();
    demons_ctrl_init_kernel<<<1, 1, 0, ctx->stream>>>(ctrl, n_iters);
    ctx->launches++;
    double* U = ws->U.as<double>();
    double* T1 = ws->T1.as<double>();
    double* T2 = ws->T2.as<double>();
    for (int it = 0; it < n_iters; ++it) {
        double* cur = P[(start + it) & 1];
        double* nxt = P[(start + it + 1) & 1];
        B200_TRY(demons_calculate_change(ctx, F, gf, M, gm, cur, fp, ws, it, n_iters));
        const double* upd = U;
        if (p.smooth_update_field && p.smooth_displacement_field && !ctx->force_separable && ctx->zm_addout && zmarch2_supported(ctx, ku) &&
            zmarch2_supported(ctx, kd)) {
            // T1 = D + G_u * U (the sum is formed when the smoothed update is stored), then D' = G_d * T1: the same IEEE
            // additions as AddImageFilter, and the displacement smoothing reads one operand instead of two
            B200_TRY(launch_conv3d_zmarch(ctx, U, cur, T1, nx, ny, nz, 3, ku, ctrl, it, true));
            B200_TRY(launch_conv3d_zmarch(ctx, T1, nullptr, nxt, nx, ny, nz, 3, kd, ctrl, it));
            continue;
        }
        if (p.smooth_update_field) {
            B200_TRY(pde_smooth(ctx, U, nullptr, T1, T2, nxt, nx, ny, nz, ku, ctrl, it));  // nxt is free scratch here
            upd = T1;
        }
        if (p.smooth_displacement_field) {
            // scratch: T2 and whichever of U / T1 does not hold the update
            B200_TRY(pde_smooth(ctx, cur, upd, nxt, T2, upd == U ? T1 : U, nx, ny, nz, kd, ctrl, it));
        } else {
            add_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(cur, upd, nxt, 3 * n, ctrl, it);
            ctx->launches++;
            B200_CHECK_LAUNCH();
        }
    }
    select_copy_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(P[1], P[0], 3 * n, ctrl, start);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

}  // namespace b200
