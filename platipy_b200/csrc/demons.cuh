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

// ---- fused warp + force, z-marching -------------------------------------------------------------------------------
// One CTA owns a 64 x 16 column of the fixed grid and marches along z.  Per step every thread produces the
// warped-moving value W (rounded through float32 exactly as WarpImageFilter stores it) and the fixed value F of
// its voxels on plane z+1 -- plus one voxel of the one-voxel x/y halo -- into 4-deep shared-memory rings held as
// doubles (FLT_MAX sentinel kept), then computes the ESM update of plane z from the rings.  W never travels to
// HBM, every F / D value is read once, index -> physical arithmetic that does not depend on z is hoisted out of
// the loop, and the SSD / count / |U|^2 partial sums stay in registers until one block reduction at the end.
// Arithmetic per voxel is the same sequence of IEEE operations as demons_warp_kernel + demons_force_kernel.
#ifndef UP_MINB
#define UP_MINB 2
#endif
#ifndef UP_RING_DEPTH
#define UP_RING_DEPTH 4
#endif
constexpr int UP_TX = 64, UP_TY = 16, UP_NT = 256, UP_HW = UP_TX + 2, UP_HH = UP_TY + 2, UP_NP = UP_HW * UP_HH, UP_RING = UP_RING_DEPTH;
constexpr int UP_NHALO = UP_NP - UP_TX * UP_TY;  // 164
#ifndef UP_GROUP
#define UP_GROUP 2
#endif
constexpr int UP_G = UP_GROUP;  // voxels of a thread whose force-phase dependency chains are interleaved (1, 2 or 4)
constexpr size_t UP_SMEM = (size_t)2 * UP_RING * UP_NP * sizeof(double);

// CONSUME 0: per-voxel branches in the force phase (first version).  CONSUME 1..3: branch-free force phase, the four
// voxels of a thread interleaved; the intensity normalisation is a compile-time case (1: none, 2: multiplication by
// the exact reciprocal of a power-of-two normalizer, 3: division).
template <bool DIAG, int CONSUME>
__global__ void __launch_bounds__(UP_NT, UP_MINB) demons_update_kernel(const float* __restrict__ F, const float* __restrict__ M, const double* __restrict__ D,
                                                                  double* __restrict__ U, double* __restrict__ partials,
                                                                  const __grid_constant__ GeomD gf, const __grid_constant__ GeomD gm,
                                                                  const __grid_constant__ ForceParams fp, int zchunk, int nchunks,
                                                                  const DemonsCtrl* __restrict__ ctrl, int it)
{
    if (it >= ctrl->halt_iter) return;
    extern __shared__ __align__(16) double up_smem[];
    double* Wr = up_smem;                     // [UP_RING][UP_NP]
    double* Fr = up_smem + UP_RING * UP_NP;   // [UP_RING][UP_NP]
    // sent[slot] == p  <=>  plane p (held in that slot) contains at least one FLT_MAX sentinel inside this CTA's
    // tile + halo.  Stale entries name older planes and never match, so no reset is needed.
    __shared__ int sent[UP_RING];
    const int tid = threadIdx.x;
    if (tid < UP_RING) sent[tid] = -0x7fffffff;
    __syncthreads();
    const int nx = gf.nx, ny = gf.ny, nz = gf.nz;
    const int x0 = blockIdx.x * UP_TX, y0 = blockIdx.y * UP_TY;
    const int z0 = blockIdx.z * zchunk, z1 = min(nz, z0 + zchunk);
    const int plane = nx * ny;
    const size_t n = (size_t)plane * nz;
    const double WMAX = (double)FLT_MAX;
    const double* __restrict__ D0 = D;
    const double* __restrict__ D1 = D + n;
    const double* __restrict__ D2 = D + 2 * n;
    double* __restrict__ U0 = U;
    double* __restrict__ U1 = U + n;
    double* __restrict__ U2 = U + 2 * n;

    // positions this thread produces: 4 own voxels (column ox, rows 4*yb..4*yb+3) and at most one halo voxel
    const int ox = tid & (UP_TX - 1), yb = tid >> 6;
    int hx[5], hy[5];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hx[j] = ox + 1;
        hy[j] = 4 * yb + j + 1;
    }
    {
        const int e = tid;
        if (e < UP_HW) { hx[4] = e; hy[4] = 0; }
        else if (e < 2 * UP_HW) { hx[4] = e - UP_HW; hy[4] = UP_HH - 1; }
        else if (e < 2 * UP_HW + UP_TY) { hx[4] = 0; hy[4] = e - 2 * UP_HW + 1; }
        else { hx[4] = UP_HW - 1; hy[4] = e - 2 * UP_HW - UP_TY + 1; }
    }
    const int npos = tid < UP_NHALO ? 5 : 4;
    int poff[5];      // offset within a plane, or -1 when the position lies outside the image
    double px[5], py[5];  // DIAG: z-independent part of the physical point
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int gx = x0 - 1 + hx[q], gy = y0 - 1 + hy[q];
        const bool ok = q < npos && gx >= 0 && gx < nx && gy >= 0 && gy < ny;
        poff[q] = ok ? gy * nx + gx : -1;
        if (DIAG) {
            // idx2pt with a diagonal index-to-physical matrix: the zero terms add exact zeros
            px[q] = gf.i2p[0] * (double)gx + gf.origin[0];
            py[q] = gf.i2p[4] * (double)gy + gf.origin[1];
        }
    }

    int soff[5];  // always-in-bounds offset for unconditional loads
#pragma unroll
    for (int q = 0; q < 5; ++q) soff[q] = poff[q] < 0 ? 0 : poff[q];

    // W / F of plane z for this thread's positions.  Branch-free up to the shared-memory stores so that the
    // loads of all positions are in flight together (field loads first, then the 8-point gathers).
    auto produce = [&](int z) {
        if (z < 0 || z >= nz) return;
        const int slot = (z + UP_RING) % UP_RING;
        const int zo = z * plane;  // volumes handled by this kernel have fewer than 2^31 voxels (checked on the host)
        double pz = 0.0;
        if (DIAG) pz = gf.i2p[8] * (double)z + gf.origin[2];
        double dd[5][3];
        float fv[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int o = zo + soff[q];
            dd[q][0] = D0[o];
            dd[q][1] = D1[o];
            dd[q][2] = D2[o];
            fv[q] = F[o];
        }
        LinW lw[5];
        bool ins[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            double p[3], c[3];
            if (DIAG) {
                p[0] = px[q];
                p[1] = py[q];
                p[2] = pz;
            } else {
                const int gy = soff[q] / nx, gx = soff[q] - gy * nx;
                idx2pt(gf, (double)gx, (double)gy, (double)z, p);
            }
            p[0] += dd[q][0];
            p[1] += dd[q][1];
            p[2] += dd[q][2];
            if (DIAG) {
                c[0] = gm.p2i[0] * (p[0] - gm.origin[0]);
                c[1] = gm.p2i[4] * (p[1] - gm.origin[1]);
                c[2] = gm.p2i[8] * (p[2] - gm.origin[2]);
            } else {
                pt2cidx(gm, p, c);
            }
            ins[q] = inside_buffer(gm, c);
            lw[q] = lin_setup(gm, c);
            // points outside the moving buffer are never interpolated; keep their (unused) gather in bounds
            lw[q].b0 = (int)min((unsigned)lw[q].b0, (unsigned)(gm.nx - 1));
            lw[q].b1 = (int)min((unsigned)lw[q].b1, (unsigned)(gm.ny - 1));
            lw[q].b2 = (int)min((unsigned)lw[q].b2, (unsigned)(gm.nz - 1));
            lw[q].u0 = min(lw[q].b0 + 1, gm.nx - 1);
            lw[q].u1 = min(lw[q].b1 + 1, gm.ny - 1);
            lw[q].u2 = min(lw[q].b2 + 1, gm.nz - 1);
        }
        double wv[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) wv[q] = lin_eval_i32<float>(M, gm.nx, gm.nx * gm.ny, lw[q]);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            if (poff[q] >= 0) {
                const int si = slot * UP_NP + hy[q] * UP_HW + hx[q];
                Wr[si] = ins[q] ? (double)(float)wv[q] : WMAX;
                Fr[si] = (double)fv[q];
                if (CONSUME == 0 && !ins[q]) sent[slot] = z;
            }
        }
    };

    double ssd = 0.0, cnt = 0.0, ssc = 0.0;
    const int gxo = x0 + ox;
    // voxels whose x/y neighbours all exist (z is checked per step)
    bool inner_xy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int gy = y0 + 4 * yb + j;
        inner_xy[j] = poff[j] >= 0 && gxo >= 1 && gxo <= nx - 2 && gy >= 1 && gy <= ny - 2;
    }

    // generic ESM update of one voxel (all border / sentinel cases), identical to demons_force_kernel
    auto slow_voxel = [&](int j, int z, int sc, int sm1, int sp1, double& u0, double& u1, double& u2, double& ds, double& dc, double& du) {
        const int gy = y0 + 4 * yb + j;
        const int ci = hy[j] * UP_HW + hx[j];
        const double movingValue = Wr[sc + ci];
        if (movingValue == WMAX) return;
        const double fixedValue = Fr[sc + ci];
        const int idx[3] = { gxo, gy, z };
        const int dims[3] = { nx, ny, nz };
        const int nb_p[3] = { sc + ci + 1, sc + ci + UP_HW, sp1 + ci };
        const int nb_m[3] = { sc + ci - 1, sc + ci - UP_HW, sm1 + ci };
        double g2[3];
#pragma unroll
        for (int dim = 0; dim < 3; ++dim) {
            const int nd = dims[dim];
            double wg;
            if (idx[dim] == 0) {
                if (nd < 2) wg = 0.0;
                else {
                    const double nb = Wr[nb_p[dim]];
                    if (nb == WMAX) wg = 0.0;
                    else {
                        wg = nb - movingValue;
                        wg /= gf.spacing[dim];
                    }
                }
            } else if (idx[dim] == nd - 1) {
                const double nb = Wr[nb_m[dim]];
                if (nb == WMAX) wg = 0.0;
                else {
                    wg = movingValue - nb;
                    wg /= gf.spacing[dim];
                }
            } else {
                const double nb = Wr[nb_p[dim]];
                const double pb = Wr[nb_m[dim]];
                if (nb == WMAX) {
                    if (pb == WMAX) wg = 0.0;
                    else {
                        wg = movingValue - pb;
                        wg /= gf.spacing[dim];
                    }
                } else if (pb == WMAX) {
                    wg = nb - movingValue;
                    wg /= gf.spacing[dim];
                } else {
                    wg = nb - pb;
                    wg *= fp.half_inv_sp[dim];
                }
            }
            double fg;
            if (idx[dim] < 1 || idx[dim] > nd - 2) fg = 0.0;
            else {
                fg = Fr[nb_p[dim]] - Fr[nb_m[dim]];
                fg *= fp.half_inv_sp[dim];
            }
            g2[dim] = fg + wg;
        }
        double J[3];
        if (DIAG) {
            J[0] = g2[0];
            J[1] = g2[1];
            J[2] = g2[2];
        } else {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double sum = 0.0;
                sum += gf.direction[r * 3 + 0] * g2[0];
                sum += gf.direction[r * 3 + 1] * g2[1];
                sum += gf.direction[r * 3 + 2] * g2[2];
                J[r] = sum;
            }
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
        ds = speed * speed;
        dc = 1.0;
        du = u0 * u0 + u1 * u1 + u2 * u2;
    };

    produce(z0 - 1);
    produce(z0);
    for (int z = z0; z < z1; ++z) {
        produce(z + 1);
        __syncthreads();
        const int sc = ((z + UP_RING) % UP_RING) * UP_NP, sm1 = ((z - 1 + UP_RING) % UP_RING) * UP_NP, sp1 = ((z + 1 + UP_RING) % UP_RING) * UP_NP;
        const int zo = z * plane;
        const bool inner_z = z >= 1 && z <= nz - 2;
        // interior planes without any sentinel in the three ring planes take the branch-free path
        const bool clean = inner_z && sent[sc / UP_NP] != z && sent[sm1 / UP_NP] != z - 1 && sent[sp1 / UP_NP] != z + 1;
        const bool all_fast = clean && inner_xy[0] && inner_xy[1] && inner_xy[2] && inner_xy[3];
        if (CONSUME > 0) {
            // Straight-line force phase.  Every voxel is first computed with central differences (the same IEEE
            // operations as the generic path); a voxel whose stencil touches the image border or a FLT_MAX sentinel
            // is flagged and redone by the generic path afterwards -- per voxel, not per CTA.  The sentinel test
            // compares high words only: (double)FLT_MAX is 0x47EFFFFF'E0000000, so a hit may also be one of the
            // seven next-largest floats, which merely sends that voxel through the (always correct) generic path.
#pragma unroll
            for (int h = 0; h < 4; h += UP_G) {
                double u0[UP_G], u1[UP_G], u2[UP_G], ds[UP_G], dc[UP_G], du[UP_G];
                bool bad[UP_G], anybad = false;
                {
                    double g0[UP_G], g1[UP_G], g2[UP_G], sp[UP_G], den[UP_G], num[UP_G], fac[UP_G];
                    bool live[UP_G], okd[UP_G];
#pragma unroll
                    for (int q = 0; q < UP_G; ++q) {
                        const int j = h + q;
                        const int ci = hy[j] * UP_HW + hx[j];
                        const double fc = Fr[sc + ci], wc = Wr[sc + ci];
                        const double wxp = Wr[sc + ci + 1], wxm = Wr[sc + ci - 1], wyp = Wr[sc + ci + UP_HW], wym = Wr[sc + ci - UP_HW];
                        const double wzp = Wr[sp1 + ci], wzm = Wr[sm1 + ci];
                        constexpr int SH = 0x47EFFFFF;
                        const bool snt = __double2hiint(wc) == SH || __double2hiint(wxp) == SH || __double2hiint(wxm) == SH ||
                                         __double2hiint(wyp) == SH || __double2hiint(wym) == SH || __double2hiint(wzp) == SH || __double2hiint(wzm) == SH;
                        bad[q] = snt || !inner_z || !inner_xy[j];
                        anybad = anybad || bad[q];
                        g0[q] = (Fr[sc + ci + 1] - Fr[sc + ci - 1]) * fp.half_inv_sp[0] + (wxp - wxm) * fp.half_inv_sp[0];
                        g1[q] = (Fr[sc + ci + UP_HW] - Fr[sc + ci - UP_HW]) * fp.half_inv_sp[1] + (wyp - wym) * fp.half_inv_sp[1];
                        g2[q] = (Fr[sp1 + ci] - Fr[sm1 + ci]) * fp.half_inv_sp[2] + (wzp - wzm) * fp.half_inv_sp[2];
                        if (!DIAG) {
                            const double a0 = g0[q], a1 = g1[q], a2 = g2[q];
                            g0[q] = ((0.0 + gf.direction[0] * a0) + gf.direction[1] * a1) + gf.direction[2] * a2;
                            g1[q] = ((0.0 + gf.direction[3] * a0) + gf.direction[4] * a1) + gf.direction[5] * a2;
                            g2[q] = ((0.0 + gf.direction[6] * a0) + gf.direction[7] * a1) + gf.direction[8] * a2;
                        }
                        sp[q] = fc - wc;
                        ds[q] = sp[q] * sp[q];
                        dc[q] = 1.0;
                        den[q] = g0[q] * g0[q] + g1[q] * g1[q] + g2[q] * g2[q];
                    }
                    if (CONSUME == 2) {
#pragma unroll
                        for (int q = 0; q < UP_G; ++q) den[q] = den[q] + ds[q] * fp.inv_normalizer;
                    } else if (CONSUME == 3) {
                        double qn[UP_G];
                        bool okn[UP_G], redo = false;
#pragma unroll
                        for (int q = 0; q < UP_G; ++q) {
                            qn[q] = div_fast_path(ds[q], fp.normalizer, okn[q]);
                            redo = redo || (!okn[q] && !bad[q]);
                        }
                        if (redo) {
#pragma unroll
                            for (int q = 0; q < UP_G; ++q)
                                if (!okn[q] && !bad[q]) qn[q] = ds[q] / fp.normalizer;
                        }
#pragma unroll
                        for (int q = 0; q < UP_G; ++q) den[q] = den[q] + qn[q];
                    }
                    bool redo = false;
#pragma unroll
                    for (int q = 0; q < UP_G; ++q) {
                        live[q] = !(fabs(sp[q]) < fp.intensity_thresh) && !(den[q] < fp.denom_thresh);
                        num[q] = 2.0 * sp[q];
                        fac[q] = div_fast_path(num[q], den[q], okd[q]);
                        redo = redo || (live[q] && !okd[q] && !bad[q]);
                    }
                    if (redo) {
#pragma unroll
                        for (int q = 0; q < UP_G; ++q)
                            if (live[q] && !okd[q] && !bad[q]) fac[q] = num[q] / den[q];
                    }
#pragma unroll
                    for (int q = 0; q < UP_G; ++q) {
                        u0[q] = live[q] ? fac[q] * g0[q] : 0.0;
                        u1[q] = live[q] ? fac[q] * g1[q] : 0.0;
                        u2[q] = live[q] ? fac[q] * g2[q] : 0.0;
                        du[q] = u0[q] * u0[q] + u1[q] * u1[q] + u2[q] * u2[q];
                    }
                }
                if (anybad) {
#pragma unroll
                    for (int q = 0; q < UP_G; ++q) {
                        if (bad[q]) {
                            u0[q] = u1[q] = u2[q] = 0.0;
                            ds[q] = dc[q] = du[q] = 0.0;
                            if (poff[h + q] >= 0) slow_voxel(h + q, z, sc, sm1, sp1, u0[q], u1[q], u2[q], ds[q], dc[q], du[q]);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < UP_G; ++q) {
                    ssd += ds[q];
                    cnt += dc[q];
                    ssc += du[q];
                    if (poff[h + q] >= 0) {
                        const int o = zo + poff[h + q];
                        U0[o] = u0[q];
                        U1[o] = u1[q];
                        U2[o] = u2[q];
                    }
                }
            }
        } else if (all_fast) {
            // branch-free interior path: central differences everywhere (same operations as the generic path)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = hy[j] * UP_HW + hx[j];
                const double fc = Fr[sc + ci], wcj = Wr[sc + ci];
                double g0 = (Fr[sc + ci + 1] - Fr[sc + ci - 1]) * fp.half_inv_sp[0] + (Wr[sc + ci + 1] - Wr[sc + ci - 1]) * fp.half_inv_sp[0];
                double g1 = (Fr[sc + ci + UP_HW] - Fr[sc + ci - UP_HW]) * fp.half_inv_sp[1] + (Wr[sc + ci + UP_HW] - Wr[sc + ci - UP_HW]) * fp.half_inv_sp[1];
                double g2 = (Fr[sp1 + ci] - Fr[sm1 + ci]) * fp.half_inv_sp[2] + (Wr[sp1 + ci] - Wr[sm1 + ci]) * fp.half_inv_sp[2];
                if (!DIAG) {
                    const double a0 = g0, a1 = g1, a2 = g2;
                    g0 = ((0.0 + gf.direction[0] * a0) + gf.direction[1] * a1) + gf.direction[2] * a2;
                    g1 = ((0.0 + gf.direction[3] * a0) + gf.direction[4] * a1) + gf.direction[5] * a2;
                    g2 = ((0.0 + gf.direction[6] * a0) + gf.direction[7] * a1) + gf.direction[8] * a2;
                }
                const double gm2 = g0 * g0 + g1 * g1 + g2 * g2;
                const double speed = fc - wcj;
                const double s2 = speed * speed;
                double denom = gm2;
                if (fp.normalizer > 0.0) denom = gm2 + (fp.inv_normalizer != 0.0 ? s2 * fp.inv_normalizer : s2 / fp.normalizer);
                const bool live = !(fabs(speed) < fp.intensity_thresh) && !(denom < fp.denom_thresh);
                const double factor = live ? 2.0 * speed / denom : 0.0;
                const double u0 = live ? factor * g0 : 0.0, u1 = live ? factor * g1 : 0.0, u2 = live ? factor * g2 : 0.0;
                ssd += s2;
                cnt += 1.0;
                ssc += u0 * u0 + u1 * u1 + u2 * u2;
                const int o = zo + poff[j];
                U0[o] = u0;
                U1[o] = u1;
                U2[o] = u2;
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                if (poff[j] < 0) continue;
                double u0 = 0.0, u1 = 0.0, u2 = 0.0, ds = 0.0, dc = 0.0, du = 0.0;
                slow_voxel(j, z, sc, sm1, sp1, u0, u1, u2, ds, dc, du);
                ssd += ds;
                cnt += dc;
                ssc += du;
                const int o = zo + poff[j];
                U0[o] = u0;
                U1[o] = u1;
                U2[o] = u2;
            }
        }
        if (UP_RING == 3) __syncthreads();  // 3-deep ring: plane z-1's slot is refilled by the next step
    }
    // one block reduction per CTA
    __shared__ double sh[3][UP_NT / 32];
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
        constexpr int NW = UP_NT / 32;
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

#ifdef B200REG_AB_VARIANTS
// ---- fused warp + force, warp-specialised (producer / consumer) ---------------------------------------------------
// Same tile, rings and arithmetic as demons_update_kernel, but the two phases run on different warps of one 512-thread
// CTA: 10 producer warps keep filling the W / F rings (field loads, 8-point gathers, XU conversions) up to three planes
// ahead while 6 consumer warps compute the ESM update (FP64 chains, two IEEE divisions) of the planes already complete.
// Hand-off through named barriers (bar.arrive / bar.sync), one "full" and one "empty" barrier per ring slot, so the
// memory-latency phase and the FP64-latency phase overlap inside every SM instead of alternating.
constexpr int WS_NT = 512, WS_PW = 10, WS_PT = WS_PW * 32, WS_CT = WS_NT - WS_PT;   // 320 producer / 192 consumer threads
constexpr int WS_PPOS = (UP_NP + WS_PT - 1) / WS_PT;                                 // 4 positions per producer thread
constexpr int WS_CVOX = (UP_TX * UP_TY + WS_CT - 1) / WS_CT;                         // 6 voxel slots per consumer thread
constexpr int WS_RING = 4;
constexpr size_t WS_SMEM = (size_t)2 * WS_RING * UP_NP * sizeof(double);

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <bool DIAG>
__global__ void __launch_bounds__(WS_NT, 1) demons_update_ws_kernel(const float* __restrict__ F, const float* __restrict__ M, const double* __restrict__ D,
                                                                     double* __restrict__ U, double* __restrict__ partials,
                                                                     const __grid_constant__ GeomD gf, const __grid_constant__ GeomD gm,
                                                                     const __grid_constant__ ForceParams fp, int zchunk, int nchunks,
                                                                     const DemonsCtrl* __restrict__ ctrl, int it)
{
    if (it >= ctrl->halt_iter) return;
    extern __shared__ __align__(16) double up_smem[];
    double* Wr = up_smem;
    double* Fr = up_smem + WS_RING * UP_NP;
    __shared__ int sent[WS_RING];
    __shared__ double sh[3][WS_NT / 32];
    const int tid = threadIdx.x;
    if (tid < WS_RING) sent[tid] = -0x7fffffff;
    __syncthreads();
    const int nx = gf.nx, ny = gf.ny, nz = gf.nz;
    const int x0 = blockIdx.x * UP_TX, y0 = blockIdx.y * UP_TY;
    const int z0 = blockIdx.z * zchunk, z1 = min(nz, z0 + zchunk);
    const int plane = nx * ny;
    const size_t n = (size_t)plane * nz;
    const double WMAX = (double)FLT_MAX;
    // barrier ids: 1..4 full[slot], 5..8 empty[slot]   (0 is __syncthreads)
    double ssd = 0.0, cnt = 0.0, ssc = 0.0;

    if (tid < WS_PT) {
        // ================================ producer warps ================================
        int hxy[WS_PPOS], poff[WS_PPOS], soff[WS_PPOS];
        double px[WS_PPOS], py[WS_PPOS];
#pragma unroll
        for (int q = 0; q < WS_PPOS; ++q) {
            const int e = q * WS_PT + tid;
            const int hy = e / UP_HW, hx = e - hy * UP_HW;
            const int gx = x0 - 1 + hx, gy = y0 - 1 + hy;
            const bool ok = e < UP_NP && gx >= 0 && gx < nx && gy >= 0 && gy < ny;
            hxy[q] = e < UP_NP ? e : 0;
            poff[q] = ok ? gy * nx + gx : -1;
            soff[q] = ok ? gy * nx + gx : 0;
            if (DIAG) {
                px[q] = gf.i2p[0] * (double)gx + gf.origin[0];
                py[q] = gf.i2p[4] * (double)gy + gf.origin[1];
            }
        }
        for (int pz = z0 - 1; pz <= z1; ++pz) {
            const int slot = (pz + WS_RING) & (WS_RING - 1);
            if (pz - (z0 - 1) >= WS_RING) named_bar_sync(5 + slot, WS_NT);  // consumers are done with plane pz - 4
            if (pz >= 0 && pz < nz) {
                const size_t zo = (size_t)pz * plane;
                double pzc = 0.0;
                if (DIAG) pzc = gf.i2p[8] * (double)pz + gf.origin[2];
                double dd[WS_PPOS][3];
                float fv[WS_PPOS];
#pragma unroll
                for (int q = 0; q < WS_PPOS; ++q) {
                    const size_t o = zo + soff[q];
                    dd[q][0] = D[o];
                    dd[q][1] = D[o + n];
                    dd[q][2] = D[o + 2 * n];
                    fv[q] = F[o];
                }
                LinW lw[WS_PPOS];
                bool ins[WS_PPOS];
#pragma unroll
                for (int q = 0; q < WS_PPOS; ++q) {
                    double p[3], c[3];
                    if (DIAG) {
                        p[0] = px[q];
                        p[1] = py[q];
                        p[2] = pzc;
                    } else {
                        const int gy = soff[q] / nx, gx = soff[q] - gy * nx;
                        idx2pt(gf, (double)gx, (double)gy, (double)pz, p);
                    }
                    p[0] += dd[q][0];
                    p[1] += dd[q][1];
                    p[2] += dd[q][2];
                    if (DIAG) {
                        c[0] = gm.p2i[0] * (p[0] - gm.origin[0]);
                        c[1] = gm.p2i[4] * (p[1] - gm.origin[1]);
                        c[2] = gm.p2i[8] * (p[2] - gm.origin[2]);
                    } else {
                        pt2cidx(gm, p, c);
                    }
                    ins[q] = inside_buffer(gm, c);
                    lw[q] = lin_setup(gm, c);
                    lw[q].b0 = (int)min((unsigned)lw[q].b0, (unsigned)(gm.nx - 1));
                    lw[q].b1 = (int)min((unsigned)lw[q].b1, (unsigned)(gm.ny - 1));
                    lw[q].b2 = (int)min((unsigned)lw[q].b2, (unsigned)(gm.nz - 1));
                    lw[q].u0 = min(lw[q].b0 + 1, gm.nx - 1);
                    lw[q].u1 = min(lw[q].b1 + 1, gm.ny - 1);
                    lw[q].u2 = min(lw[q].b2 + 1, gm.nz - 1);
                }
                double wv[WS_PPOS];
#pragma unroll
                for (int q = 0; q < WS_PPOS; ++q) wv[q] = lin_eval<float>(M, gm, lw[q]);
#pragma unroll
                for (int q = 0; q < WS_PPOS; ++q) {
                    if (poff[q] >= 0) {
                        const int si = slot * UP_NP + hxy[q];
                        Wr[si] = ins[q] ? (double)(float)wv[q] : WMAX;
                        Fr[si] = (double)fv[q];
                        if (!ins[q]) sent[slot] = pz;
                    }
                }
            }
            __threadfence_block();              // ring stores visible before the hand-off
            named_bar_arrive(1 + slot, WS_NT);  // plane pz is complete
        }
    } else {
        // ================================ consumer warps ================================
        const int ct = tid - WS_PT;
        const int ox = ct & (UP_TX - 1), yb = ct >> 6;  // column ox, rows yb, yb + 3, yb + 6, ...
        const int gxo = x0 + ox;
        // wait for the first three planes
        named_bar_sync(1 + ((z0 - 1 + WS_RING) & (WS_RING - 1)), WS_NT);
        named_bar_sync(1 + ((z0 + WS_RING) & (WS_RING - 1)), WS_NT);
        for (int z = z0; z < z1; ++z) {
            named_bar_sync(1 + ((z + 1 + WS_RING) & (WS_RING - 1)), WS_NT);
            const int sc = ((z + WS_RING) & (WS_RING - 1)) * UP_NP, sm1 = ((z - 1 + WS_RING) & (WS_RING - 1)) * UP_NP,
                      sp1 = ((z + 1 + WS_RING) & (WS_RING - 1)) * UP_NP;
            const size_t zo = (size_t)z * plane;
            const bool clean = z >= 1 && z <= nz - 2 && sent[sc / UP_NP] != z && sent[sm1 / UP_NP] != z - 1 && sent[sp1 / UP_NP] != z + 1;
#pragma unroll 2
            for (int r = 0; r < WS_CVOX; ++r) {
                const int ly = yb + 3 * r;
                if (ly >= UP_TY) break;
                const int gy = y0 + ly;
                if (gxo >= nx || gy >= ny) continue;
                const int ci = (ly + 1) * UP_HW + ox + 1;
                const size_t o = zo + (size_t)gy * nx + gxo;
                double u0 = 0.0, u1 = 0.0, u2 = 0.0;
                const bool inner = clean && gxo >= 1 && gxo <= nx - 2 && gy >= 1 && gy <= ny - 2;
                if (inner) {
                    const double fc = Fr[sc + ci], wcj = Wr[sc + ci];
                    double g0 = (Fr[sc + ci + 1] - Fr[sc + ci - 1]) * fp.half_inv_sp[0] + (Wr[sc + ci + 1] - Wr[sc + ci - 1]) * fp.half_inv_sp[0];
                    double g1 = (Fr[sc + ci + UP_HW] - Fr[sc + ci - UP_HW]) * fp.half_inv_sp[1] + (Wr[sc + ci + UP_HW] - Wr[sc + ci - UP_HW]) * fp.half_inv_sp[1];
                    double g2 = (Fr[sp1 + ci] - Fr[sm1 + ci]) * fp.half_inv_sp[2] + (Wr[sp1 + ci] - Wr[sm1 + ci]) * fp.half_inv_sp[2];
                    if (!DIAG) {
                        const double a0 = g0, a1 = g1, a2 = g2;
                        g0 = ((0.0 + gf.direction[0] * a0) + gf.direction[1] * a1) + gf.direction[2] * a2;
                        g1 = ((0.0 + gf.direction[3] * a0) + gf.direction[4] * a1) + gf.direction[5] * a2;
                        g2 = ((0.0 + gf.direction[6] * a0) + gf.direction[7] * a1) + gf.direction[8] * a2;
                    }
                    const double gm2 = g0 * g0 + g1 * g1 + g2 * g2;
                    const double speed = fc - wcj;
                    const double s2 = speed * speed;
                    double denom = gm2;
                    if (fp.normalizer > 0.0) denom = gm2 + (fp.inv_normalizer != 0.0 ? s2 * fp.inv_normalizer : s2 / fp.normalizer);
                    const bool live = !(fabs(speed) < fp.intensity_thresh) && !(denom < fp.denom_thresh);
                    const double factor = live ? 2.0 * speed / denom : 0.0;
                    u0 = live ? factor * g0 : 0.0;
                    u1 = live ? factor * g1 : 0.0;
                    u2 = live ? factor * g2 : 0.0;
                    ssd += s2;
                    cnt += 1.0;
                    ssc += u0 * u0 + u1 * u1 + u2 * u2;
                } else {
                    // generic path (borders, FLT_MAX sentinels): identical to demons_force_kernel
                    const double movingValue = Wr[sc + ci];
                    if (movingValue != WMAX) {
                        const double fixedValue = Fr[sc + ci];
                        const int idx[3] = { gxo, gy, z };
                        const int dims[3] = { nx, ny, nz };
                        const int nb_p[3] = { sc + ci + 1, sc + ci + UP_HW, sp1 + ci };
                        const int nb_m[3] = { sc + ci - 1, sc + ci - UP_HW, sm1 + ci };
                        double g2v[3];
#pragma unroll
                        for (int dim = 0; dim < 3; ++dim) {
                            const int nd = dims[dim];
                            double wg;
                            if (idx[dim] == 0) {
                                if (nd < 2) wg = 0.0;
                                else {
                                    const double nb = Wr[nb_p[dim]];
                                    if (nb == WMAX) wg = 0.0;
                                    else {
                                        wg = nb - movingValue;
                                        wg /= gf.spacing[dim];
                                    }
                                }
                            } else if (idx[dim] == nd - 1) {
                                const double nb = Wr[nb_m[dim]];
                                if (nb == WMAX) wg = 0.0;
                                else {
                                    wg = movingValue - nb;
                                    wg /= gf.spacing[dim];
                                }
                            } else {
                                const double nb = Wr[nb_p[dim]];
                                const double pb = Wr[nb_m[dim]];
                                if (nb == WMAX) {
                                    if (pb == WMAX) wg = 0.0;
                                    else {
                                        wg = movingValue - pb;
                                        wg /= gf.spacing[dim];
                                    }
                                } else if (pb == WMAX) {
                                    wg = nb - movingValue;
                                    wg /= gf.spacing[dim];
                                } else {
                                    wg = nb - pb;
                                    wg *= fp.half_inv_sp[dim];
                                }
                            }
                            double fg;
                            if (idx[dim] < 1 || idx[dim] > nd - 2) fg = 0.0;
                            else {
                                fg = Fr[nb_p[dim]] - Fr[nb_m[dim]];
                                fg *= fp.half_inv_sp[dim];
                            }
                            g2v[dim] = fg + wg;
                        }
                        double J[3];
                        if (DIAG) {
                            J[0] = g2v[0];
                            J[1] = g2v[1];
                            J[2] = g2v[2];
                        } else {
#pragma unroll
                            for (int rr = 0; rr < 3; ++rr) {
                                double sum = 0.0;
                                sum += gf.direction[rr * 3 + 0] * g2v[0];
                                sum += gf.direction[rr * 3 + 1] * g2v[1];
                                sum += gf.direction[rr * 3 + 2] * g2v[2];
                                J[rr] = sum;
                            }
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
                        ssd += speed * speed;
                        cnt += 1.0;
                        ssc += u0 * u0 + u1 * u1 + u2 * u2;
                    }
                }
                U[o] = u0;
                U[o + n] = u1;
                U[o + 2 * n] = u2;
            }
            named_bar_arrive(5 + ((z - 1 + WS_RING) & (WS_RING - 1)), WS_NT);  // plane z-1 may be overwritten
        }
    }
    // one block reduction per CTA (producer threads contribute zeros)
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
        constexpr int NW = WS_NT / 32;
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

#endif  // B200REG_AB_VARIANTS

inline bool geom_is_diag(const GeomD& g)
{
    const double* d = g.direction;
    return d[0] == 1.0 && d[4] == 1.0 && d[8] == 1.0 && d[1] == 0.0 && d[2] == 0.0 && d[3] == 0.0 && d[5] == 0.0 && d[6] == 0.0 && d[7] == 0.0;
}

__global__ void demons_ctrl_init_kernel(DemonsCtrl* ctrl, int n_iters)
{
    ctrl->halt_iter = n_iters <= 0 ? 0 : 0x7fffffff;
    ctrl->elapsed = 0;
    ctrl->metric = DBL_MAX;
    ctrl->rms = 0.0;
}

struct DemonsWorkspace {
    TempBuf U, T1, T2, P1, W, partials, ctrl, trace;
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

inline dim3 update_grid(b200reg_ctx* ctx, const GeomD& gf, int* zchunk)
{
    const int tx = (gf.nx + UP_TX - 1) / UP_TX, ty = (gf.ny + UP_TY - 1) / UP_TY;
    int nchunks = (ctx->sm_count * 8 + tx * ty - 1) / (tx * ty);
    const int max_chunks = (gf.nz + 15) / 16;
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks < 1) nchunks = 1;
    *zchunk = (gf.nz + nchunks - 1) / nchunks;
    nchunks = (gf.nz + *zchunk - 1) / *zchunk;
    return dim3(tx, ty, nchunks);
}

inline int demons_calculate_change(b200reg_ctx* ctx, const float* F, const GeomD& gf, const float* M, const GeomD& gm, const double* D,
                                   const ForceParams& fp, DemonsWorkspace* ws, int it, int n_iters, bool want_w = false)
{
    DemonsCtrl* ctrl = ws->ctrl.as<DemonsCtrl>();
    size_t nblocks;
    // the z-marching kernel needs a few CTAs per SM to hide its per-plane latency: small (coarse-level) grids
    // run the one-thread-per-voxel kernels instead
    int zchunk_probe;
    const dim3 gprobe = update_grid(ctx, gf, &zchunk_probe);
    const bool small_grid = (size_t)gprobe.x * gprobe.y * gprobe.z < (size_t)ctx->sm_count * 4;
    const bool huge = (size_t)gf.nx * gf.ny * gf.nz >= (1ull << 31) || (size_t)gm.nx * gm.ny * gm.nz >= (1ull << 31);  // 32-bit offsets inside
    const bool fits32 = (size_t)gf.nx * gf.ny * gf.nz * 3 < (1ull << 31) && (size_t)gm.nx * gm.ny * gm.nz < (1ull << 31);
    if (ctx->update_split && fits32) {
        // two high-occupancy kernels, W through HBM (demons_split.cuh)
        B200_TRY(launch_update_split(ctx, F, gf, M, gm, D, ws->W.as<float>(), ws->U.as<double>(), ws->partials.as<double>(), fp,
                                     geom_is_diag(gf) && geom_is_diag(gm), ctrl, it, &nblocks));
    } else if (want_w || ctx->unfused_force || small_grid || huge) {
        const dim3 g = grid3(gf.nx, gf.ny, gf.nz), b = block3();
        nblocks = (size_t)g.x * g.y * g.z;
        if (gm.small) demons_warp_kernel<true><<<g, b, 0, ctx->stream>>>(M, D, ws->W.as<float>(), gf, gm, ctrl, it);
        else demons_warp_kernel<false><<<g, b, 0, ctx->stream>>>(M, D, ws->W.as<float>(), gf, gm, ctrl, it);
        demons_force_kernel<<<g, b, 0, ctx->stream>>>(F, ws->W.as<float>(), ws->U.as<double>(), ws->partials.as<double>(), gf, fp, ctrl, it);
        ctx->launches += 2;
    } else {
        int zchunk;
        const dim3 g = update_grid(ctx, gf, &zchunk);
        nblocks = (size_t)g.x * g.y * g.z;
        const bool diag = geom_is_diag(gf) && geom_is_diag(gm);
#ifdef B200REG_AB_VARIANTS
        if (ctx->update_ws) {
            if (diag) B200_TRY(ensure_dynamic_smem(ctx, demons_update_ws_kernel<true>, WS_SMEM));
            else B200_TRY(ensure_dynamic_smem(ctx, demons_update_ws_kernel<false>, WS_SMEM));
            if (diag)
                demons_update_ws_kernel<true><<<g, WS_NT, WS_SMEM, ctx->stream>>>(F, M, D, ws->U.as<double>(), ws->partials.as<double>(), gf, gm, fp, zchunk,
                                                                                  (int)g.z, ctrl, it);
            else
                demons_update_ws_kernel<false><<<g, WS_NT, WS_SMEM, ctx->stream>>>(F, M, D, ws->U.as<double>(), ws->partials.as<double>(), gf, gm, fp, zchunk,
                                                                                   (int)g.z, ctrl, it);
        } else {
#else
        {
#endif
        // force-phase variant: 0 = per-voxel branches, else the straight-line form specialised on the normalisation
        int consume = fp.normalizer > 0.0 ? (fp.inv_normalizer != 0.0 ? 2 : 3) : 1;
#ifdef B200REG_AB_VARIANTS
        if (ctx->update_branchy) consume = 0;
#endif
#define UP_LAUNCH(DG, CS)                                                                                                                     \
    do {                                                                                                                                      \
        B200_TRY(ensure_dynamic_smem(ctx, demons_update_kernel<DG, CS>, UP_SMEM));                                                            \
        demons_update_kernel<DG, CS><<<g, UP_NT, UP_SMEM, ctx->stream>>>(F, M, D, ws->U.as<double>(), ws->partials.as<double>(), gf, gm, fp, \
                                                                         zchunk, (int)g.z, ctrl, it);                                         \
    } while (0)
        if (diag) {
            switch (consume) {
#ifdef B200REG_AB_VARIANTS
            case 0: UP_LAUNCH(true, 0); break;
#endif
            case 1: UP_LAUNCH(true, 1); break;
            case 2: UP_LAUNCH(true, 2); break;
            default: UP_LAUNCH(true, 3); break;
            }
        } else {
            switch (consume) {
#ifdef B200REG_AB_VARIANTS
            case 0: UP_LAUNCH(false, 0); break;
#endif
            case 1: UP_LAUNCH(false, 1); break;
            case 2: UP_LAUNCH(false, 2); break;
            default: UP_LAUNCH(false, 3); break;
            }
        }
#undef UP_LAUNCH
        }
        ctx->launches += 1;
    }
    demons_finish_kernel<<<1, 1024, 0, ctx->stream>>>(ws->partials.as<double>(), nblocks, ctrl, fp.max_rms_error, it, n_iters, ws->trace.as<double>());
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
// The field ping-pongs between two buffers; after `elapsed` iterations it lives in buffer elapsed % 2.
// Decided on the device: copy P1 -> P0 when the elapsed count is odd.
__global__ void select_copy_kernel(const double* __restrict__ p1, double* __restrict__ p0, size_t n, const DemonsCtrl* __restrict__ ctrl)
{
    if ((ctrl->elapsed & 1) == 0) return;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) p0[q] = p1[q];
}

inline int make_pde_coeffs(const double sd[3], double max_error, int max_width, KernelCoeffs kc[3])
{
    for (int a = 0; a < 3; ++a) B200_TRY(make_coeffs(gaussian_operator(sd[a] * sd[a], max_error, max_width), &kc[a]));
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
    if (p.smooth_displacement_field) B200_TRY(make_pde_coeffs(p.std_dev, p.max_error, p.max_kernel_width, kd));
    if (p.smooth_update_field) B200_TRY(make_pde_coeffs(p.update_std_dev, p.max_error, p.max_kernel_width, ku));
    B200_CUDA(cudaMemsetAsync(D, 0, 3 * n * sizeof(double), ctx->stream));
    DemonsCtrl* ctrl = ws->ctrl.as<DemonsCtrl>();
    demons_ctrl_init_kernel<<<1, 1, 0, ctx->stream>>>(ctrl, n_iters);
    ctx->launches++;
    double* U = ws->U.as<double>();
    double* T1 = ws->T1.as<double>();
    double* T2 = ws->T2.as<double>();
    double* P[2] = { D, ws->P1.as<double>() };
    for (int it = 0; it < n_iters; ++it) {
        double* cur = P[it & 1];
        double* nxt = P[(it + 1) & 1];
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
    select_copy_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(P[1], P[0], 3 * n, ctrl);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

}  // namespace b200
