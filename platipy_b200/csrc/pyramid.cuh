// pyramid.cuh -- smooth_and_resample (utils.py:195-267) for a SHRINKING pyramid level, computed only where the level reads it.
//
// The reference blurs the full-resolution image (sitk.DiscreteGaussian, utils.py:216-226) and then resamples it onto the level's grid
// with a linear interpolator (utils.py:257-267).  A level that shrinks by s reads, along each axis, the two input indices around every
// output index: 2 / s of the planes, rows and columns.  Because the Gaussian is separable and ITK runs it one axis at a time with
// float32 intermediates, each pass needs to produce only the positions the NEXT stage reads:
//
//   pass over axis p1 (z by default): all x, all y, the needed z planes          -> work  f1       (f = needed fraction of an axis)
//   pass over axis p2 (y):            all x, needed y rows of those planes       ->       f1 f2
//   pass over axis p3 (x):            needed columns of those rows               ->       f1 f2 f3
//   gather: the trilinear interpolation of the level's voxels from the compact result.
//
// Shrink 4: 0.5 + 0.25 + 0.125 of a pass instead of 3 passes; shrink 8: 0.33 instead of 3.  Every value that is computed is computed
// by the same operations in the same order as in the full passes (double accumulation over ascending taps of clamped float32 inputs,
// one rounding to float32 per pass) and the interpolation is the resampler's own (lin_eval's nested lerps on the same base index and
// distance), so the level is bit-identical to blur-everything-then-resample.
//
// Which positions are needed is decided ON THE DEVICE by the resampler's own index arithmetic (out_to_in_cidx + base_and_distance),
// evaluated per axis by a one-block kernel that also compacts the lists -- no host replay of floating-point expressions, no copies.
// The host only proves what makes the axes separable (identity direction cosines on both grids, identity transform) and that the level
// lies inside the input buffer with a margin; anything else takes the generic path.
#pragma once
#include "gauss.cuh"
#include "resample.cuh"

namespace b200 {

struct ShrinkAxis {
    int* need;   // [cap]   input positions the level reads along this axis, ascending
    int* lo;     // [n_out] compact position of the base index of output index i
    int* hi;     // [n_out] compact position of the upper neighbour
    double* d;   // [n_out] interpolation distance
    int* count;  // number of needed positions
    int n_in, n_out, cap;
};
struct ShrinkTabs {
    ShrinkAxis ax[3];
};

// one block per axis; the flags of the input positions live in shared memory (n_in <= SHRINK_MAX_AXIS)
constexpr int SHRINK_MAX_AXIS = 12288;
__global__ void __launch_bounds__(256) shrink_tables_kernel(const __grid_constant__ GeomD go, const __grid_constant__ GeomD gi, const __grid_constant__ ChainD ch,
                                                             const __grid_constant__ ShrinkTabs tabs)
{
    extern __shared__ int sflag[];
    const int a = blockIdx.x;
    const ShrinkAxis& t = tabs.ax[a];
    for (int p = threadIdx.x; p < t.n_in; p += blockDim.x) sflag[p] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < t.n_out; i += blockDim.x) {
        double c[3];
        // with identity direction cosines and an identity transform component a of the continuous index depends on index a only
        out_to_in_cidx(go, gi, ch, a == 0 ? i : 0, a == 1 ? i : 0, a == 2 ? i : 0, c);
        int b;
        double d;
        base_and_distance(c[a], b, d);
        b = min(b, t.n_in - 1);  // never taken for a point inside the buffer (the host checked that)
        const int u = min(b + 1, t.n_in - 1);
        t.lo[i] = b;
        t.hi[i] = u;
        t.d[i] = d;
        sflag[b] = 1;
        sflag[u] = 1;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // ordered compaction by one warp: 32 positions per step
        const int lane = threadIdx.x;
        int m = 0;
        for (int base = 0; base < t.n_in; base += 32) {
            const int p = base + lane;
            const bool f = p < t.n_in && sflag[p] != 0;
            const unsigned mask = __ballot_sync(0xffffffffu, f);
            const int pos = m + __popc(mask & ((1u << lane) - 1u));
            if (f && pos < t.cap) {
                t.need[pos] = p;
                sflag[p] = pos;
            }
            m += __popc(mask);
        }
        if (lane == 0) *t.count = min(m, t.cap);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < t.n_out; i += blockDim.x) {
        t.lo[i] = sflag[t.lo[i]];
        t.hi[i] = sflag[t.hi[i]];
    }
}

// One separable pass that produces only the needed positions along its axis.  in: [iz][iy][ix] (full along AXIS), out: the same with
// the AXIS dimension replaced by `cap` (compact; the first *count entries are valid).  One thread per pair of consecutive needed
// positions (the level reads base index and base + 1: usually adjacent, so the pair shares all but one of its taps).
template <int AXIS>
__global__ void __launch_bounds__(128) conv_sel_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int ix, int iy, int iz, int cap,
                                                            const int* __restrict__ need, const int* __restrict__ count, const int* __restrict__ lim_y,
                                                            const int* __restrict__ lim_z, const __grid_constant__ KernelCoeffs kc)
{
    const int m = *count;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    int x, y, z, q0;
    if (AXIS == 0) {
        q0 = 2 * t0; x = 0; y = blockIdx.y; z = blockIdx.z;
    } else if (AXIS == 1) {
        x = t0; q0 = 2 * blockIdx.y; y = 0; z = blockIdx.z;
    } else {
        x = t0; y = blockIdx.y; q0 = 2 * blockIdx.z; z = 0;
    }
    if (q0 >= m || x >= ix) return;
    // rows / planes beyond the valid extent of an already compacted axis hold nothing
    if (AXIS != 1 && lim_y && y >= *lim_y) return;
    if (AXIS != 2 && lim_z && z >= *lim_z) return;
    const int n = AXIS == 0 ? ix : (AXIS == 1 ? iy : iz);
    const size_t sa = AXIS == 0 ? 1 : (AXIS == 1 ? (size_t)ix : (size_t)ix * iy);
    const float* line = in + ((size_t)z * iy + y) * ix + x;
    const int ox = AXIS == 0 ? cap : ix, oy = AXIS == 1 ? cap : iy;
    const size_t so = AXIS == 0 ? 1 : (AXIS == 1 ? (size_t)ox : (size_t)ox * oy);
    float* oline = out + ((size_t)z * oy + y) * ox + x;
    const int r = kc.r;
    const bool two = q0 + 1 < m;
    const int p0 = need[q0], p1 = two ? need[q0 + 1] : p0;
    double a0 = 0.0, a1 = 0.0;
    if (two && p1 == p0 + 1 && p0 - r >= 0 && p1 + r <= n - 1) {
        // the usual case -- base index and base + 1, window inside the image: no clamping, and the second output sees every value one tap
        // later than the first, so one coefficient fetch per value serves both (ascending taps for each, as in the tiled kernels)
        const float* ptr = line + (size_t)(p0 - r) * sa;
        const int ntap = 2 * r + 2;
        double kprev = 0.0;
        for (int eb = 0; eb < ntap; eb += 8) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = ptr[(size_t)min(eb + e, ntap - 1) * sa];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int t = eb + e;
                const double vd = (double)v[e];
                const double kcur = kc.k[min(t, 2 * r)];
                if (t <= 2 * r) a0 += kcur * vd;
                if (t >= 1 && t < ntap) a1 += kprev * vd;
                kprev = kcur;
            }
        }
        oline[(size_t)q0 * so] = (float)a0;
        oline[(size_t)(q0 + 1) * so] = (float)a1;
        return;
    }
    // taps in ascending order, eight loads in flight at a time (the radius is a run-time value: without the explicit batches every load
    // would wait for the one before it)
    const int ulo = p0 - r, uhi = p1 + r;
    for (int ub = ulo; ub <= uhi; ub += 8) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int u = ub + e;
            const int uc = u < 0 ? 0 : (u > n - 1 ? n - 1 : u);
            v[e] = line[(size_t)uc * sa];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int u = ub + e;
            const int k0 = u - ulo, k1 = u - (p1 - r);
            const double vd = (double)v[e];
            if (k0 <= 2 * r) a0 += kc.k[k0] * vd;
            if (two && k1 >= 0 && u <= uhi) a1 += kc.k[k1] * vd;
        }
    }
    oline[(size_t)q0 * so] = (float)a0;
    if (two) oline[(size_t)(q0 + 1) * so] = (float)a1;
}

// the level's voxels: LinearInterpolateImageFunction on the compact result (lin_eval's expression, same operand order)
__global__ void __launch_bounds__(256) shrink_gather_kernel(const float* __restrict__ src, float* __restrict__ out, int cx, int cy, int nxo, int nyo, int nzo,
                                                             const __grid_constant__ ShrinkTabs tabs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= nxo) return;
    const int b0 = tabs.ax[0].lo[i], u0 = tabs.ax[0].hi[i];
    const int b1 = tabs.ax[1].lo[j], u1 = tabs.ax[1].hi[j];
    const int b2 = tabs.ax[2].lo[k], u2 = tabs.ax[2].hi[k];
    const double d0 = tabs.ax[0].d[i], d1 = tabs.ax[1].d[j], d2 = tabs.ax[2].d[k];
    const size_t sy = (size_t)cx, sz = (size_t)cx * cy;
    const size_t r00 = (size_t)b2 * sz + (size_t)b1 * sy, r10 = (size_t)b2 * sz + (size_t)u1 * sy;
    const size_t r01 = (size_t)u2 * sz + (size_t)b1 * sy, r11 = (size_t)u2 * sz + (size_t)u1 * sy;
    const double v000 = (double)src[r00 + b0], v100 = (double)src[r00 + u0];
    const double v010 = (double)src[r10 + b0], v110 = (double)src[r10 + u0];
    const double v001 = (double)src[r01 + b0], v101 = (double)src[r01 + u0];
    const double v011 = (double)src[r11 + b0], v111 = (double)src[r11 + u0];
    const double vx00 = v000 + (v100 - v000) * d0;
    const double vx10 = v010 + (v110 - v010) * d0;
    const double vxx0 = vx00 + (vx10 - vx00) * d1;
    const double vx01 = v001 + (v101 - v001) * d0;
    const double vx11 = v011 + (v111 - v011) * d0;
    const double vxx1 = vx01 + (vx11 - vx01) * d1;
    out[((size_t)k * nyo + j) * nxo + i] = (float)(vxx0 + (vxx1 - vxx0) * d2);
}

// *used = false: the conditions do not hold (or the level does not shrink enough to pay) and nothing was launched.
inline int smooth_and_shrink_f32(b200reg_ctx* ctx, const float* d_in, const b200reg_geom& gin, const double variance[3], int max_width, double max_error,
                                 const b200reg_geom& gout, float* d_out, bool* used, bool whenever_possible = false)
{
    *used = false;
    if (!ctx->pyramid_restrict && !whenever_possible) return B200REG_OK;
    const GeomD gi = make_geomd(gin), go = make_geomd(gout);
    for (int q = 0; q < 9; ++q) {
        const double want = (q % 4 == 0) ? 1.0 : 0.0;
        if (gi.direction[q] != want || go.direction[q] != want) return B200REG_OK;
    }
    int order[3];
    for (int p = 0; p < 3; ++p) order[p] = semantics().discrete_gaussian_axis_order ? p : 2 - p;
    int cap[3];
    double f[3];
    for (int a = 0; a < 3; ++a) {
        // the level must lie inside the input buffer with a margin (ITK's test is [-0.5, n - 0.5)): approximate continuous indices of the
        // first and last output index (the exact ones are evaluated on the device)
        const double c_first = (gout.origin[a] - gin.origin[a]) / gin.spacing[a];
        const double c_last = (gout.origin[a] + (gout.size[a] - 1) * gout.spacing[a] - gin.origin[a]) / gin.spacing[a];
        if (!(c_first > -0.25 && c_last < gin.size[a] - 0.75 && c_last >= c_first)) return B200REG_OK;
        cap[a] = std::min(gin.size[a], 2 * gout.size[a]);
        f[a] = (double)cap[a] / gin.size[a];
    }
    const double cost = f[order[0]] + f[order[0]] * f[order[1]] + f[order[0]] * f[order[1]] * f[order[2]];
    // three full passes cost 3, but the tiled kernels spend a third of the instructions per output and tap (four adjacent outputs share
    // every staged value): measured at 512 x 512 x 256, shrink 4 (cost 0.875) breaks even, shrink 8 (0.33) wins
    if (cost > ctx->pyramid_restrict_cost && !whenever_possible) return B200REG_OK;
    if (gin.size[1] > 65535 || gin.size[2] > 65535) return B200REG_OK;
    int max_axis = 0;
    for (int a = 0; a < 3; ++a) max_axis = std::max(max_axis, gin.size[a]);
    if (max_axis > SHRINK_MAX_AXIS) return B200REG_OK;

    ChainD ch;
    B200_TRY(make_chain(nullptr, 0, &ch));
    // tables: per axis need[cap] lo[n_out] hi[n_out] count[1] (ints), d[n_out] (doubles, first in the buffer)
    size_t n_dbl = 0, n_int = 0;
    for (int a = 0; a < 3; ++a) {
        n_dbl += (size_t)gout.size[a];
        n_int += (size_t)cap[a] + 2 * (size_t)gout.size[a] + 1;
    }
    TempBuf tb;
    B200_TRY(tb.alloc(ctx, n_dbl * sizeof(double) + n_int * sizeof(int)));
    ShrinkTabs tabs;
    {
        double* pd = tb.as<double>();
        int* pi = reinterpret_cast<int*>(pd + n_dbl);
        for (int a = 0; a < 3; ++a) {
            ShrinkAxis& t = tabs.ax[a];
            t.n_in = gin.size[a];
            t.n_out = gout.size[a];
            t.cap = cap[a];
            t.d = pd; pd += t.n_out;
            t.need = pi; pi += t.cap;
            t.lo = pi; pi += t.n_out;
            t.hi = pi; pi += t.n_out;
            t.count = pi; pi += 1;
        }
    }
    shrink_tables_kernel<<<3, 256, (size_t)max_axis * sizeof(int), ctx->stream>>>(go, gi, ch, tabs);
    ctx->launches++;
    B200_CHECK_LAUNCH();

    int cur[3] = { gin.size[0], gin.size[1], gin.size[2] };
    bool compacted[3] = { false, false, false };
    const float* src = d_in;
    TempBuf bufs[3];
    for (int p = 0; p < 3; ++p) {
        const int a = order[p];
        double t = variance[a] / (gin.spacing[a] * gin.spacing[a]);  // use_image_spacing, as utils.py:216-226 calls the filter
        KernelCoeffs kc;
        B200_TRY(make_coeffs(gaussian_operator(t, max_error, max_width), &kc));
        int od[3] = { cur[0], cur[1], cur[2] };
        od[a] = cap[a];
        B200_TRY(bufs[p].alloc(ctx, (size_t)od[0] * od[1] * od[2] * sizeof(float)));
        const int* lim_y = compacted[1] ? tabs.ax[1].count : nullptr;
        const int* lim_z = compacted[2] ? tabs.ax[2].count : nullptr;
        const int pairs = (cap[a] + 1) / 2;
        if (a == 0) {
            conv_sel_f32_kernel<0><<<dim3((pairs + 127) / 128, cur[1], cur[2]), 128, 0, ctx->stream>>>(src, bufs[p].as<float>(), cur[0], cur[1], cur[2], cap[a],
                                                                                                       tabs.ax[a].need, tabs.ax[a].count, lim_y, lim_z, kc);
        } else if (a == 1) {
            conv_sel_f32_kernel<1><<<dim3((cur[0] + 127) / 128, pairs, cur[2]), 128, 0, ctx->stream>>>(src, bufs[p].as<float>(), cur[0], cur[1], cur[2], cap[a],
                                                                                                        tabs.ax[a].need, tabs.ax[a].count, lim_y, lim_z, kc);
        } else {
            conv_sel_f32_kernel<2><<<dim3((cur[0] + 127) / 128, cur[1], pairs), 128, 0, ctx->stream>>>(src, bufs[p].as<float>(), cur[0], cur[1], cur[2], cap[a],
                                                                                                        tabs.ax[a].need, tabs.ax[a].count, lim_y, lim_z, kc);
        }
        ctx->launches++;
        B200_CHECK_LAUNCH();
        cur[a] = cap[a];
        compacted[a] = true;
        src = bufs[p].as<float>();
    }
    shrink_gather_kernel<<<dim3((gout.size[0] + 255) / 256, gout.size[1], gout.size[2]), 256, 0, ctx->stream>>>(src, d_out, cur[0], cur[1], gout.size[0],
                                                                                                                 gout.size[1], gout.size[2], tabs);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    *used = true;
    return B200REG_OK;
}

}  // namespace b200
