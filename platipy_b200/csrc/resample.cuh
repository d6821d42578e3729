// resample.cuh -- itk::ResampleImageFilter for scalar volumes and 3-component f64 fields, with
// identity / affine / displacement-field transform chains, nearest-neighbour and linear interpolation.
//   N2 utils.py:257-267   N3 deformable.py:130,137,185   N5 deformable.py:140   N7 deformable.py:154
//   N9 utils.py:176-190, deformable.py:281-301
#pragma once
#include "common.cuh"

namespace b200 {

#ifndef RS_MINB
#define RS_MINB 4  // resident 256-thread blocks per SM the resampling kernels are compiled for (64 registers)
#endif

struct TfmD {
    int kind;
    double matrix[9];
    double offset[3];
    const double* dvf;  // SoA planes
    GeomD g;
    int on_grid;        // first element only: the field lives on the OUTPUT grid and every index maps to itself exactly -> read per index
};
struct ChainD {
    int n;
    int linear;     // every element affine -> ITK's scan-line path (semantic switch resample_linear_scanline)
    int dvf_lerp;   // semantic switch dvf_transform_interpolation: nested lerps instead of the weighted sum
    int vec_wsum;   // semantic switch vector_resample_interpolation: weighted sum instead of nested lerps
    TfmD t[B200REG_MAX_TRANSFORMS];
};

inline int make_chain(const b200reg_transform* chain, int n_chain, ChainD* out)
{
    if (n_chain < 0 || n_chain > B200REG_MAX_TRANSFORMS)
        return set_error(B200REG_ERR_ARG, "transform chain length %d not in [0, %d]", n_chain, B200REG_MAX_TRANSFORMS);
    if (n_chain > 0 && !chain) return set_error(B200REG_ERR_ARG, "null transform chain");
    out->n = n_chain;
    out->linear = semantics().resample_linear_scanline ? 1 : 0;
    out->dvf_lerp = semantics().dvf_transform_interpolation;
    out->vec_wsum = semantics().vector_resample_interpolation;
    for (int i = 0; i < n_chain; ++i) {
        TfmD& t = out->t[i];
        t.kind = chain[i].kind;
        for (int k = 0; k < 9; ++k) t.matrix[k] = chain[i].matrix[k];
        for (int k = 0; k < 3; ++k) t.offset[k] = chain[i].offset[k];
        t.dvf = chain[i].d_dvf;
        t.on_grid = 0;
        if (t.kind == B200REG_TFM_DVF) {
            if (!t.dvf || !valid_geom(&chain[i].dvf_geom)) return set_error(B200REG_ERR_ARG, "invalid displacement-field transform");
            t.g = make_geomd(chain[i].dvf_geom);
            out->linear = 0;
        } else if (t.kind != B200REG_TFM_AFFINE) {
            return set_error(B200REG_ERR_ARG, "unknown transform kind %d", t.kind);
        }
    }
    return B200REG_OK;
}

// ---- identity resamples that are exact copies ---------------------------------------------------------------------------------------
// platipy re-grids onto an identical grid in several places (the full-resolution pyramid level of smooth_and_resample, utils.py:257-267
// with shrink factor 1; the final sitk.Resample(dvf_total, fixed_image), deformable.py:185, when the last level is the full grid).
// ITK still runs the resampler there.  Its result equals the input exactly when every continuous index it computes is exactly the
// integer index -- then nearest neighbour and linear interpolation (distance 0: v + (w - v) * 0) return the voxel itself.  That is
// decided here by evaluating, on the host, the very arithmetic the kernel would evaluate (index -> point of the output grid,
// point -> continuous index of the input grid, the scan-line interpolation of resample_linear_scanline) for every index along every
// axis; only grids with identity direction cosines qualify (the axes then separate).  Finite voxel values assumed (an infinite
// neighbour would turn v + (inf - v) * 0 into NaN in the generic path).
inline bool identity_resample_is_exact(const b200reg_geom& gin, const b200reg_geom& gout, bool allow_scanline = true)
{
    for (int a = 0; a < 3; ++a)
        if (gin.size[a] != gout.size[a]) return false;
    const GeomD gi = make_geomd(gin), go = make_geomd(gout);
    for (int q = 0; q < 9; ++q) {
        const double want = (q % 4 == 0) ? 1.0 : 0.0;
        if (gi.direction[q] != want || go.direction[q] != want) return false;
    }
    const int n[3] = { go.nx, go.ny, go.nz };
    const bool scanline = allow_scanline && semantics().resample_linear_scanline != 0;
    for (int a = 0; a < 3; ++a) {
        // one axis of idx2pt / pt2cidx with the zero off-diagonal terms dropped (they add exact zeros)
        auto f = [&](double t) {
            const double p = go.i2p[a * 4] * t + go.origin[a];
            return gi.p2i[a * 4] * (p - gi.origin[a]);
        };
        if (a == 0 && scanline) {
            const double cs = f(0.0), ce = f((double)n[0]);
            for (int i = 0; i < n[0]; ++i) {
                const double alpha = (double)i / (double)n[0];
                if (cs + alpha * (ce - cs) != (double)i) return false;
            }
        } else {
            for (int i = 0; i < n[a]; ++i)
                if (f((double)i) != (double)i) return false;
        }
    }
    return true;
}

// A displacement field that lives on the OUTPUT grid of a resample (the usual case: the transform of a registration applied on the
// fixed grid, deformable.py:281-301, multiatlas/run.py:331-345) is read by ITK through the generic path: output index -> point ->
// continuous index of the field -> interpolation.  When the host proves that every such continuous index is exactly the integer
// index (identity_resample_is_exact, per-voxel form), the interpolation returns the node value itself (weight 1 on one neighbour)
// and the kernel reads the field per index instead.
inline void chain_mark_on_grid(b200reg_ctx* ctx, ChainD* ch, const b200reg_transform* chain, int n_chain, const b200reg_geom& gout)
{
    if (ctx->identity_copy && n_chain >= 1 && chain[0].kind == B200REG_TFM_DVF && identity_resample_is_exact(chain[0].dvf_geom, gout, false)) ch->t[0].on_grid = 1;
}

// VectorLinearInterpolateImageFunction (inside DisplacementFieldTransform): weighted sum over the 8
// neighbours, bit k of the counter selects the upper neighbour in dim k, indices clamped into the
// buffer, zero-overlap neighbours skipped, early exit when the accumulated overlap is exactly 1.
__device__ __forceinline__ void interp_wsum_vec3(const double* __restrict__ f, const GeomD& g, const double* c, double* out)
{
    // floor without XU conversions: c + 1.5 * 2^52 rounded toward -inf holds floor(c) in its low mantissa word
    // (two's complement, so negative indices down to -2^31 work too); only called for points inside the buffer.
    const double magic = 6755399441055744.0;
    const double t0 = __dadd_rd(c[0], magic), t1 = __dadd_rd(c[1], magic), t2 = __dadd_rd(c[2], magic);
    const int b[3] = { __double2loint(t0), __double2loint(t1), __double2loint(t2) };
    const double d[3] = { c[0] - (t0 - magic), c[1] - (t1 - magic), c[2] - (t2 - magic) };
    const int n[3] = { g.nx, g.ny, g.nz };
    const size_t plane = (size_t)g.nx * g.ny * g.nz;
    out[0] = out[1] = out[2] = 0.0;
    double total = 0.0;
#pragma unroll
    for (unsigned counter = 0; counter < 8; ++counter) {
        double overlap = 1.0;
        int ni[3];
#pragma unroll
        for (int dim = 0; dim < 3; ++dim) {
            if ((counter >> dim) & 1u) {
                ni[dim] = b[dim] + 1;
                if (ni[dim] > n[dim] - 1) ni[dim] = n[dim] - 1;
                overlap *= d[dim];
            } else {
                ni[dim] = b[dim];
                if (ni[dim] < 0) ni[dim] = 0;
                overlap *= 1.0 - d[dim];
            }
        }
        if (overlap != 0.0) {
            const size_t o = ((size_t)ni[2] * n[1] + ni[1]) * n[0] + ni[0];
            out[0] += overlap * __ldg(f + o);
            out[1] += overlap * __ldg(f + plane + o);
            out[2] += overlap * __ldg(f + 2 * plane + o);
            total += overlap;
        }
        if (total == 1.0) break;
    }
}

// the alternative settings of the interpolation switches: out of line, so that the default paths keep their register budgets
__device__ __noinline__ void interp_lerp_vec3_alt(const double* __restrict__ f, const GeomD& g, const double* c, double* out);
__device__ __noinline__ void interp_wsum_vec3_alt(const double* __restrict__ f, const GeomD& g, const double* c, double* out);

__device__ __forceinline__ void apply_chain(const ChainD& ch, double* p, int first = 0)
{
    for (int i = first; i < ch.n; ++i) {
        const TfmD& t = ch.t[i];
        if (t.kind == B200REG_TFM_AFFINE) {
            double q[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double sum = 0.0;
                sum += t.matrix[r * 3 + 0] * p[0];
                sum += t.matrix[r * 3 + 1] * p[1];
                sum += t.matrix[r * 3 + 2] * p[2];
                q[r] = sum + t.offset[r];
            }
            p[0] = q[0];
            p[1] = q[1];
            p[2] = q[2];
        } else {
            double c[3], dd[3];
            pt2cidx(t.g, p, c);
            if (inside_buffer(t.g, c)) {
                if (ch.dvf_lerp) interp_lerp_vec3_alt(t.dvf, t.g, c, dd);
                else interp_wsum_vec3(t.dvf, t.g, c, dd);
                p[0] += dd[0];
                p[1] += dd[1];
                p[2] += dd[2];
            }
        }
    }
}

// continuous input index of output voxel (i, j, k).  Linear chains follow
// ResampleImageFilter::LinearThreadedGenerateData: continuous index at the first index of the row and at
// one-past-the-last, interpolated with alpha = i / size_x.
__device__ __forceinline__ void out_to_in_cidx(const GeomD& go, const GeomD& gi, const ChainD& ch, int i, int j, int k, double* c)
{
    double p[3];
    if (!ch.linear) {
        idx2pt(go, (double)i, (double)j, (double)k, p);
        int first = 0;
        if (ch.n > 0 && ch.t[0].on_grid) {
            // the field on the output grid, read per index (see chain_mark_on_grid)
            const size_t n = (size_t)go.nx * go.ny * go.nz, o = ((size_t)k * go.ny + j) * go.nx + i;
            const double* __restrict__ f = ch.t[0].dvf;
            p[0] += __ldg(f + o);
            p[1] += __ldg(f + n + o);
            p[2] += __ldg(f + 2 * n + o);
            first = 1;
        }
        apply_chain(ch, p, first);
        pt2cidx(gi, p, c);
    } else {
        double cs[3], ce[3];
        idx2pt(go, 0.0, (double)j, (double)k, p);
        apply_chain(ch, p);
        pt2cidx(gi, p, cs);
        idx2pt(go, (double)go.nx, (double)j, (double)k, p);
        apply_chain(ch, p);
        pt2cidx(gi, p, ce);
        const double alpha = (double)i / (double)go.nx;
#pragma unroll
        for (int r = 0; r < 3; ++r) c[r] = cs[r] + alpha * (ce[r] - cs[r]);
    }
}

// The same for a whole (BX, BY) thread block.  In the scan-line form everything but alpha is a property of the ROW (j, k) and alpha = i / nx
// a property of the COLUMN, yet out_to_in_cidx evaluates both row ends (2 x (index -> point -> chain -> continuous index), ~90 FP64
// instructions) and the division for every voxel -- ncu of the field re-gridding kernel: FP64 pipe 56 %, issue slots 69 %, the kernel is
// bound by this arithmetic, not by its 1.6 GB of stores.  Here two threads of each row evaluate the row ends, the first row of threads
// the BX divisions, and the block shares them through shared memory: the same operations on the same operands, once.
// Every thread of the block must call this (it synchronises), also those outside the image.
__device__ __forceinline__ void out_to_in_cidx_block(const GeomD& go, const GeomD& gi, const ChainD& ch, int i, int j, int k, double* c)
{
#ifdef RS_NO_BLOCK_SCANLINE  // A/B build: every voxel evaluates its own row ends
    out_to_in_cidx(go, gi, ch, i, j, k, c);
    return;
#endif
    if (!ch.linear) {  // uniform over the grid
        out_to_in_cidx(go, gi, ch, i, j, k, c);
        return;
    }
    __shared__ double s_end[BY][2][3];
    __shared__ double s_alpha[BX];
    if (threadIdx.x < 2) {
        double p[3], e[3];
        idx2pt(go, threadIdx.x == 0 ? 0.0 : (double)go.nx, (double)j, (double)k, p);
        apply_chain(ch, p);
        pt2cidx(gi, p, e);
        s_end[threadIdx.y][threadIdx.x][0] = e[0];
        s_end[threadIdx.y][threadIdx.x][1] = e[1];
        s_end[threadIdx.y][threadIdx.x][2] = e[2];
    }
    if (threadIdx.y == 0) s_alpha[threadIdx.x] = (double)i / (double)go.nx;
    __syncthreads();
    const double alpha = s_alpha[threadIdx.x];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double cs = s_end[threadIdx.y][0][r], ce = s_end[threadIdx.y][1][r];
        c[r] = cs + alpha * (ce - cs);
    }
}

// LinearInterpolateImageFunction::EvaluateOptimized(Dispatch<3>): nested lerps x, y, z as
// a + (b - a) * d in double.  Base index clamped up to 0 with non-positive distances treated as 0,
// upper neighbours clamped to the last index: bit-identical to ITK's branchy form.
struct LinW {
    int b0, b1, b2, u0, u1, u2;
    double d0, d1, d2;
};
// Base index and distance of LinearInterpolateImageFunction::EvaluateOptimized without XU-pipe conversions
// (F2I / I2F run at 16 lanes/clk/SM on sm_100) and with three instructions per axis:
//   * a continuous index in [-0.5, 0) gives base 0 / distance 0 in ITK (base clamped up to the start index, the
//     negative distance treated as 0) -- exactly what clamping the index itself to 0 gives;
//   * floor(c) for 0 <= c < 2^31 is the low mantissa word of c + 1.5 * 2^52 added with round-toward-minus-infinity
//     (one DADD.RM), and the same sum minus the constant is floor(c) as a double (exact).
// Indices beyond 2^31 or NaN never belong to a point inside a buffer; callers that gather unconditionally clamp the
// resulting integers.
__device__ __forceinline__ void base_and_distance(double c, int& b, double& d)
{
    const double magic = 6755399441055744.0;
    const double cc = c < 0.0 ? 0.0 : c;
    const double t = __dadd_rd(cc, magic);
    b = __double2loint(t);
    d = cc - (t - magic);
}
__device__ __forceinline__ LinW lin_setup(const GeomD& g, const double* c)
{
    LinW w;
    base_and_distance(c[0], w.b0, w.d0);
    base_and_distance(c[1], w.b1, w.d1);
    base_and_distance(c[2], w.b2, w.d2);
    w.u0 = min(w.b0 + 1, g.nx - 1);
    w.u1 = min(w.b1 + 1, g.ny - 1);
    w.u2 = min(w.b2 + 1, g.nz - 1);
    return w;
}
template <typename T>
__device__ __forceinline__ double lin_eval(const T* __restrict__ img, const GeomD& g, const LinW& w)
{
    const size_t sy = (size_t)g.nx, sz = (size_t)g.nx * g.ny;
    const size_t r00 = (size_t)w.b2 * sz + (size_t)w.b1 * sy, r10 = (size_t)w.b2 * sz + (size_t)w.u1 * sy;
    const size_t r01 = (size_t)w.u2 * sz + (size_t)w.b1 * sy, r11 = (size_t)w.u2 * sz + (size_t)w.u1 * sy;
    const double v000 = (double)__ldg(img + r00 + w.b0), v100 = (double)__ldg(img + r00 + w.u0);
    const double v010 = (double)__ldg(img + r10 + w.b0), v110 = (double)__ldg(img + r10 + w.u0);
    const double v001 = (double)__ldg(img + r01 + w.b0), v101 = (double)__ldg(img + r01 + w.u0);
    const double v011 = (double)__ldg(img + r11 + w.b0), v111 = (double)__ldg(img + r11 + w.u0);
    const double vx00 = v000 + (v100 - v000) * w.d0;
    const double vx10 = v010 + (v110 - v010) * w.d0;
    const double vxx0 = vx00 + (vx10 - vx00) * w.d1;
    const double vx01 = v001 + (v101 - v001) * w.d0;
    const double vx11 = v011 + (v111 - v011) * w.d0;
    const double vxx1 = vx01 + (vx11 - vx01) * w.d1;
    return vxx0 + (vxx1 - vxx0) * w.d2;
}

// Same interpolation with 32-bit element offsets (volumes below 2^31 voxels): one base offset, three neighbour
// deltas, eight loads addressed as base pointer + 32-bit offset.  Identical values, far less integer arithmetic than
// forming eight 64-bit indices.
template <typename T>
__device__ __forceinline__ double lin_eval_i32(const T* __restrict__ img, int nx, int nxy, const LinW& w)
{
    const int o000 = (w.b2 * nxy) + (w.b1 * nx) + w.b0;
    const int dx = w.u0 - w.b0, dy = (w.u1 - w.b1) * nx, dz = (w.u2 - w.b2) * nxy;
    const int o010 = o000 + dy, o001 = o000 + dz, o011 = o010 + dz;
    // offsets are summed as 32-bit integers before they meet the pointer: one IMAD.WIDE per load instead of a
    // 64-bit add chain
    const int o100 = o000 + dx, o110 = o010 + dx, o101 = o001 + dx, o111 = o011 + dx;
    const double v000 = (double)__ldg(img + o000), v100 = (double)__ldg(img + o100);
    const double v010 = (double)__ldg(img + o010), v110 = (double)__ldg(img + o110);
    const double v001 = (double)__ldg(img + o001), v101 = (double)__ldg(img + o101);
    const double v011 = (double)__ldg(img + o011), v111 = (double)__ldg(img + o111);
    const double vx00 = v000 + (v100 - v000) * w.d0;
    const double vx10 = v010 + (v110 - v010) * w.d0;
    const double vxx0 = vx00 + (vx10 - vx00) * w.d1;
    const double vx01 = v001 + (v101 - v001) * w.d0;
    const double vx11 = v011 + (v111 - v011) * w.d0;
    const double vxx1 = vx01 + (vx11 - vx01) * w.d1;
    return vxx0 + (vxx1 - vxx0) * w.d2;
}

__device__ __noinline__ void interp_lerp_vec3_alt(const double* __restrict__ f, const GeomD& g, const double* c, double* out)
{
    const LinW w = lin_setup(g, c);
    const size_t plane = (size_t)g.nx * g.ny * g.nz;
    out[0] = lin_eval<double>(f, g, w);
    out[1] = lin_eval<double>(f + plane, g, w);
    out[2] = lin_eval<double>(f + 2 * plane, g, w);
}
__device__ __noinline__ void interp_wsum_vec3_alt(const double* __restrict__ f, const GeomD& g, const double* c, double* out)
{
    interp_wsum_vec3(f, g, c, out);
}

// One image of a batch
struct BatchItem {
    const void* in;
    void* out;
    int dtype;
    int interp;
    double default_value;
    const double* coeff;  // B-spline coefficients of `in` (interp == B200REG_INTERP_BSPLINE)
};

// ---- itk::BSplineInterpolateImageFunction, spline order 3 (sitk.sitkBSpline) ------------------------------------------
// Coefficients: itk::BSplineDecompositionImageFilter::DataToCoefficients1D along x, y, z in turn, one thread per line,
// in place on a float64 copy of the image.  The constants that involve libm (pole, gain, horizon, z^(n-1)) come from
// the host so that the recursion is the same sequence of IEEE operations as the CPU oracle's.
struct BsplinePole {
    double z, gain, z2n0, anti;  // pole, (1 - z)(1 - 1/z), z^(n-1), z / (z^2 - 1)
    int horizon;
};
template <int AXIS>
__global__ void __launch_bounds__(128) bspline3_prefilter_kernel(double* __restrict__ c, int nx, int ny, int nz, const __grid_constant__ BsplinePole pp)
{
    const int n = AXIS == 0 ? nx : (AXIS == 1 ? ny : nz);
    const size_t nlines = AXIS == 0 ? (size_t)ny * nz : (AXIS == 1 ? (size_t)nx * nz : (size_t)nx * ny);
    const size_t s = AXIS == 0 ? 1 : (AXIS == 1 ? (size_t)nx : (size_t)nx * ny);
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nlines || n == 1) return;
    double* l = AXIS == 0 ? c + r * (size_t)nx : (AXIS == 1 ? c + (r / nx) * (size_t)nx * ny + (r % nx) : c + r);
    const double z = pp.z;
    for (int k = 0; k < n; ++k) l[k * s] *= pp.gain;
    {
        double zn = z, sum;
        if (pp.horizon < n) {
            sum = l[0];
            for (int k = 1; k < pp.horizon; ++k) {
                sum += zn * l[k * s];
                zn *= z;
            }
            l[0] = sum;
        } else {
            const double iz = 1.0 / z;
            double z2n = pp.z2n0;
            sum = l[0] + z2n * l[(size_t)(n - 1) * s];
            z2n *= z2n * iz;
            for (int k = 1; k <= n - 2; ++k) {
                sum += (zn + z2n) * l[k * s];
                zn *= z;
                z2n *= iz;
            }
            l[0] = sum / (1.0 - zn * zn);
        }
    }
    for (int k = 1; k < n; ++k) l[k * s] += z * l[(k - 1) * s];
    l[(size_t)(n - 1) * s] = pp.anti * (z * l[(size_t)(n - 2) * s] + l[(size_t)(n - 1) * s]);
    for (int k = n - 2; k >= 0; --k) l[k * s] = z * (l[(k + 1) * s] - l[k * s]);
}
inline BsplinePole bspline3_pole(int n)
{
    BsplinePole p;
    p.z = std::sqrt(3.0) - 2.0;
    double c0 = 1.0;
    c0 = c0 * (1.0 - p.z) * (1.0 - 1.0 / p.z);
    p.gain = c0;
    p.horizon = (int)std::ceil(std::log(1e-10) / std::log(std::fabs(p.z)));
    p.z2n0 = std::pow(p.z, (double)(n - 1));
    p.anti = p.z / (p.z * p.z - 1.0);
    return p;
}
template <typename T>
__global__ void __launch_bounds__(256) to_f64_kernel(const T* __restrict__ in, double* __restrict__ out, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = (double)in[q];
}
// EvaluateAtContinuousIndexInternal: support floor((float)x) - 1 .. + 2, cubic weights, mirror boundary, 64-point sum with
// x fastest and weight ((1 * wx) * wy) * wz
__device__ __forceinline__ double bspline3_eval(const double* __restrict__ coef, const GeomD& g, const double* x)
{
    const int n[3] = { g.nx, g.ny, g.nz };
    int idx[3][4];
    double w[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int i0 = (int)floorf((float)x[d]) - 1;
        const double t = x[d] - (double)(i0 + 1);
        w[d][3] = (1.0 / 6.0) * t * t * t;
        w[d][0] = (1.0 / 6.0) + 0.5 * t * (t - 1.0) - w[d][3];
        w[d][2] = t + w[d][0] - 2.0 * w[d][3];
        w[d][1] = 1.0 - w[d][0] - w[d][2] - w[d][3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int q = i0 + k;
            if (n[d] == 1) q = 0;
            else {
                if (q < 0) q = -q;
                if (q > n[d] - 1) q = (n[d] - 1) - (q - (n[d] - 1));
                q = q < 0 ? 0 : (q > n[d] - 1 ? n[d] - 1 : q);
            }
            idx[d][k] = q;
        }
    }
    double v = 0.0;
#pragma unroll
    for (int kz = 0; kz < 4; ++kz)
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const size_t row = ((size_t)idx[2][kz] * n[1] + (size_t)idx[1][ky]) * n[0];
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                double ww = 1.0;
                ww *= w[0][kx];
                ww *= w[1][ky];
                ww *= w[2][kz];
                v += ww * __ldg(coef + row + idx[0][kx]);
            }
        }
    return v;
}
constexpr int RESAMPLE_BATCH = 8;
struct BatchD {
    int n;
    BatchItem item[RESAMPLE_BATCH];
};

template <typename T, bool SMALL, bool BSP>
__device__ __forceinline__ void resample_one(const BatchItem& it, const GeomD& gi, const double* c, bool inside, size_t o)
{
    const T* in = reinterpret_cast<const T*>(it.in);
    T* out = reinterpret_cast<T*>(it.out);
    if (inside) {
        double v;
        if (BSP && it.interp == B200REG_INTERP_BSPLINE) {
            v = bspline3_eval(it.coeff, gi, c);
        } else if (it.interp == B200REG_INTERP_NN) {
            // NearestNeighborInterpolateImageFunction: RoundHalfIntegerUp = floor(x + 0.5)
            const int i0 = (int)floor(c[0] + 0.5), i1 = (int)floor(c[1] + 0.5), i2 = (int)floor(c[2] + 0.5);
            v = Px<T>::ld(in, ((size_t)i2 * gi.ny + i1) * gi.nx + i0);
        } else {
            const LinW w = lin_setup(gi, c);
            v = SMALL ? lin_eval_i32<T>(in, gi.nx, gi.nx * gi.ny, w) : lin_eval<T>(in, gi, w);
        }
        out[o] = Px<T>::cast(v);
    } else {
        out[o] = Px<T>::cast(it.default_value);
    }
}

// ---- bit-packed label propagation ---------------------------------------------------------------------------------------------
// The reference propagates S binary structures through one transform with S nearest-neighbour calls (multiatlas/run.py:338-345):
// S gathers of one byte each per output voxel, i.e. S 32-byte sectors fetched for S useful bytes.  Here the UInt8 nearest-neighbour
// items of a batch are first packed into ONE 32-bit word per input voxel (bit b = item b is non-zero there; a streaming pass), and
// the resampling kernel gathers that one word per output voxel and expands it into the S outputs.  Exact for any label whose
// non-zero voxels all carry the same value (the value is found by the packing pass: min and max over the non-zero voxels); an
// item with several non-zero values is recognised on the device and gathered from its own image as before -- no host round trip.
constexpr int PACK_MAX = 32;
struct PackedD {
    const uint32_t* words;   // [input voxels]
    const unsigned* meta;    // [2 * n]: min over the non-zero values (0xFFFFFFFF if none), max value
    int n;
    const uint8_t* in[PACK_MAX];
    uint8_t* out[PACK_MAX];
    uint8_t dflt[PACK_MAX];  // DefaultPixelValue cast to UInt8
};
struct PackSrc {
    int n;
    const uint8_t* in[PACK_MAX];
};
__global__ void pack_meta_init_kernel(unsigned* meta, int n)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) {
        meta[2 * b] = 0xFFFFFFFFu;
        meta[2 * b + 1] = 0u;
    }
}
// one thread per 4 consecutive voxels (VEC) or per voxel; the per-item min / max go through a warp reduction and an atomic that
// is skipped once the published value already covers the warp's (after the first few warps nothing is left to publish)
template <bool VEC>
__global__ void __launch_bounds__(256) pack_u8_bits_kernel(const __grid_constant__ PackSrc src, uint32_t* __restrict__ words, unsigned* __restrict__ meta,
                                                            size_t nvox)
{
    constexpr int V = VEC ? 4 : 1;
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    const bool live = q < nvox;
    uint32_t w[V];
#pragma unroll
    for (int v = 0; v < V; ++v) w[v] = 0;
    for (int b = 0; b < src.n; ++b) {
        unsigned vals[V];
        if (VEC) {
            const uchar4 u = live ? __ldg(reinterpret_cast<const uchar4*>(src.in[b] + q)) : make_uchar4(0, 0, 0, 0);
            vals[0] = u.x;
            if (V > 1) {
                vals[V > 1 ? 1 : 0] = u.y;
                vals[V > 2 ? 2 : 0] = u.z;
                vals[V > 3 ? 3 : 0] = u.w;
            }
        } else {
            vals[0] = live ? __ldg(src.in[b] + q) : 0u;
        }
        unsigned mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (vals[v]) {
                w[v] |= 1u << b;
                mn = min(mn, vals[v]);
                mx = max(mx, vals[v]);
            }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        if ((threadIdx.x & 31) == 0 && mx != 0u) {
            if (mn < *((volatile unsigned*)(meta + 2 * b))) atomicMin(meta + 2 * b, mn);
            if (mx > *((volatile unsigned*)(meta + 2 * b + 1))) atomicMax(meta + 2 * b + 1, mx);
        }
    }
    if (live) {
        if (VEC) *reinterpret_cast<uint4*>(words + q) = make_uint4(w[0], w[V > 1 ? 1 : 0], w[V > 2 ? 2 : 0], w[V > 3 ? 3 : 0]);
        else words[q] = w[0];
    }
}

// BSP: the batch contains a B-spline item (the 64-point evaluation needs far more registers than the other two
// interpolators, so it lives in its own instantiation and the common one keeps 4 blocks per SM)
template <bool SMALL, bool BSP, bool PACKED = false>
__global__ void __launch_bounds__(BX* BY, BSP ? 2 : RS_MINB) resample_batch_kernel(const __grid_constant__ BatchD batch, const __grid_constant__ GeomD gi,
                                                                 const __grid_constant__ GeomD go, const __grid_constant__ ChainD ch,
                                                                 const __grid_constant__ PackedD pk)
{
    const int i = blockIdx.x * BX + threadIdx.x;
    const int j = blockIdx.y * BY + threadIdx.y;
    const int k = blockIdx.z;
    double c[3];
    out_to_in_cidx_block(go, gi, ch, i, j, k, c);
    if (i >= go.nx || j >= go.ny) return;
    const bool inside = inside_buffer(gi, c);
    const size_t o = ((size_t)k * go.ny + j) * go.nx + i;
    for (int b = 0; b < batch.n; ++b) {
        const BatchItem& it = batch.item[b];
        switch (it.dtype) {
        case B200REG_I8: resample_one<int8_t, SMALL, BSP>(it, gi, c, inside, o); break;
        case B200REG_U8: resample_one<uint8_t, SMALL, BSP>(it, gi, c, inside, o); break;
        case B200REG_I16: resample_one<int16_t, SMALL, BSP>(it, gi, c, inside, o); break;
        case B200REG_U16: resample_one<uint16_t, SMALL, BSP>(it, gi, c, inside, o); break;
        case B200REG_I32: resample_one<int32_t, SMALL, BSP>(it, gi, c, inside, o); break;
        case B200REG_U32: resample_one<uint32_t, SMALL, BSP>(it, gi, c, inside, o); break;
        case B200REG_I64: resample_one<int64_t, SMALL, BSP>(it, gi, c, inside, o); break;
        case B200REG_U64: resample_one<uint64_t, SMALL, BSP>(it, gi, c, inside, o); break;
        case B200REG_F32: resample_one<float, SMALL, BSP>(it, gi, c, inside, o); break;
        default: resample_one<double, SMALL, BSP>(it, gi, c, inside, o); break;
        }
    }
    if (PACKED) {
        // the packed UInt8 nearest-neighbour items: one gathered word instead of pk.n gathered bytes
        uint32_t w = 0;
        size_t src = 0;
        if (inside) {
            // NearestNeighborInterpolateImageFunction: RoundHalfIntegerUp = floor(x + 0.5), as in resample_one
            const int i0 = (int)floor(c[0] + 0.5), i1 = (int)floor(c[1] + 0.5), i2 = (int)floor(c[2] + 0.5);
            src = ((size_t)i2 * gi.ny + i1) * gi.nx + i0;
            w = __ldg(pk.words + src);
        }
        for (int b = 0; b < pk.n; ++b) {
            uint8_t v;
            if (!inside) {
                v = pk.dflt[b];
            } else {
                const unsigned mn = __ldg(pk.meta + 2 * b), mx = __ldg(pk.meta + 2 * b + 1);
                if (mn >= mx) v = ((w >> b) & 1u) ? (uint8_t)mx : (uint8_t)0;  // one non-zero value (or none): the bit says it all
                else v = __ldg(pk.in[b] + src);                                 // several non-zero values: the item's own voxel
            }
            pk.out[b][o] = v;
        }
    }
}

inline int resample_batch(b200reg_ctx* ctx, int n, const void* const* d_in, const int* dtypes, const b200reg_geom& gin,
                          void* const* d_out, const b200reg_geom& gout, const b200reg_transform* chain, int n_chain,
                          const int* interps, const double* defaults)
{
    ChainD ch;
    B200_TRY(make_chain(chain, n_chain, &ch));
    chain_mark_on_grid(ctx, &ch, chain, n_chain, gout);
    const GeomD gi = make_geomd(gin), go = make_geomd(gout);
    for (int i = 0; i < n; ++i) {
        if (!d_in[i] || !d_out[i]) return set_error(B200REG_ERR_ARG, "null image pointer in resample batch");
        if (dtype_size(dtypes[i]) == 0) return set_error(B200REG_ERR_ARG, "unsupported pixel type %d", dtypes[i]);
        if (interps[i] != B200REG_INTERP_NN && interps[i] != B200REG_INTERP_LINEAR && interps[i] != B200REG_INTERP_BSPLINE)
            return set_error(B200REG_ERR_UNSUPPORTED, "interpolator %d is not supported (nearest neighbour = 1, linear = 2, B-spline = 3)", interps[i]);
    }
    const size_t n_in = nvox(gin);
    if (n_chain == 0 && ctx->identity_copy) {
        bool plain = true;
        for (int i = 0; i < n; ++i) plain = plain && (interps[i] == B200REG_INTERP_NN || interps[i] == B200REG_INTERP_LINEAR);
        if (plain && identity_resample_is_exact(gin, gout)) {
            for (int i = 0; i < n; ++i)
                if (d_in[i] != d_out[i])
                    B200_CUDA(cudaMemcpyAsync(d_out[i], d_in[i], n_in * dtype_size(dtypes[i]), cudaMemcpyDeviceToDevice, ctx->stream));
            return B200REG_OK;
        }
    }
    // UInt8 nearest-neighbour items (propagated structures) travel bit-packed when there are enough of them to pay for the packing pass
    std::vector<int> plain, packed;
    for (int i = 0; i < n; ++i) {
        if (dtypes[i] == B200REG_U8 && interps[i] == B200REG_INTERP_NN && d_in[i] != d_out[i] && ctx->pack_labels) packed.push_back(i);
        else plain.push_back(i);
    }
    if ((int)packed.size() < 4) {
        plain.clear();
        packed.clear();
        for (int i = 0; i < n; ++i) plain.push_back(i);
    }
    TempBuf words, meta;
    std::vector<PackedD> groups;
    const bool vec_ok = (n_in % 4) == 0;
    for (size_t g0 = 0; g0 < packed.size(); g0 += PACK_MAX) {
        PackedD pk;
        memset(&pk, 0, sizeof(pk));
        PackSrc src;
        memset(&src, 0, sizeof(src));
        pk.n = src.n = (int)std::min<size_t>(PACK_MAX, packed.size() - g0);
        bool vec = vec_ok;
        for (int b = 0; b < pk.n; ++b) {
            const int i = packed[g0 + b];
            pk.in[b] = src.in[b] = (const uint8_t*)d_in[i];
            pk.out[b] = (uint8_t*)d_out[i];
            pk.dflt[b] = (uint8_t)Px<uint8_t>::cast_host(defaults[i]);
            vec = vec && (reinterpret_cast<uintptr_t>(d_in[i]) % 4) == 0;
        }
        if (g0 == 0) {
            const size_t ngroups = (packed.size() + PACK_MAX - 1) / PACK_MAX;
            B200_TRY(words.alloc(ctx, ngroups * n_in * sizeof(uint32_t)));
            B200_TRY(meta.alloc(ctx, ngroups * 2 * PACK_MAX * sizeof(unsigned)));
        }
        uint32_t* wptr = words.as<uint32_t>() + (g0 / PACK_MAX) * n_in;
        unsigned* mptr = meta.as<unsigned>() + (g0 / PACK_MAX) * 2 * PACK_MAX;
        pack_meta_init_kernel<<<1, 64, 0, ctx->stream>>>(mptr, pk.n);
        if (vec) pack_u8_bits_kernel<true><<<(unsigned)((n_in / 4 + 255) / 256), 256, 0, ctx->stream>>>(src, wptr, mptr, n_in);
        else pack_u8_bits_kernel<false><<<(unsigned)((n_in + 255) / 256), 256, 0, ctx->stream>>>(src, wptr, mptr, n_in);
        ctx->launches += 2;
        B200_CHECK_LAUNCH();
        pk.words = wptr;
        pk.meta = mptr;
        groups.push_back(pk);
    }
    const int n_plain = (int)plain.size();
    PackedD no_pack;
    memset(&no_pack, 0, sizeof(no_pack));
    size_t next_group = 0;
    // every launch carries up to RESAMPLE_BATCH plain items and one packed group: CT + structures share one evaluation of the transform
    for (int start = 0; start < n_plain || next_group < groups.size(); start += RESAMPLE_BATCH) {
        BatchD b;
        b.n = start < n_plain ? ((n_plain - start) < RESAMPLE_BATCH ? (n_plain - start) : RESAMPLE_BATCH) : 0;
        TempBuf coef[RESAMPLE_BATCH];
        for (int q = 0; q < b.n; ++q) {
            const int src_i = plain[start + q];
            b.item[q] = BatchItem{ d_in[src_i], d_out[src_i], dtypes[src_i], interps[src_i], defaults[src_i], nullptr };
            if (interps[src_i] == B200REG_INTERP_BSPLINE) {
                B200_TRY(coef[q].alloc(ctx, n_in * sizeof(double)));
                double* c = coef[q].as<double>();
                const int nb = ctx->sm_count * 8;
                B200_DISPATCH_DTYPE(dtypes[src_i], T, { to_f64_kernel<T><<<nb, 256, 0, ctx->stream>>>((const T*)d_in[src_i], c, n_in); });
                const size_t l0 = (size_t)gi.ny * gi.nz, l1 = (size_t)gi.nx * gi.nz, l2 = (size_t)gi.nx * gi.ny;
                bspline3_prefilter_kernel<0><<<(unsigned)((l0 + 127) / 128), 128, 0, ctx->stream>>>(c, gi.nx, gi.ny, gi.nz, bspline3_pole(gi.nx));
                bspline3_prefilter_kernel<1><<<(unsigned)((l1 + 127) / 128), 128, 0, ctx->stream>>>(c, gi.nx, gi.ny, gi.nz, bspline3_pole(gi.ny));
                bspline3_prefilter_kernel<2><<<(unsigned)((l2 + 127) / 128), 128, 0, ctx->stream>>>(c, gi.nx, gi.ny, gi.nz, bspline3_pole(gi.nz));
                ctx->launches += 4;
                B200_CHECK_LAUNCH();
                b.item[q].coeff = c;
            }
        }
        bool bsp = false;
        for (int q = 0; q < b.n; ++q) bsp = bsp || b.item[q].interp == B200REG_INTERP_BSPLINE;
        const dim3 g3 = grid3(go.nx, go.ny, go.nz);
        const bool with_pack = next_group < groups.size();
        const PackedD& pk = with_pack ? groups[next_group] : no_pack;
        if (with_pack) ++next_group;
        if (bsp) {
            if (gi.small) {
                if (with_pack) resample_batch_kernel<true, true, true><<<g3, block3(), 0, ctx->stream>>>(b, gi, go, ch, pk);
                else resample_batch_kernel<true, true, false><<<g3, block3(), 0, ctx->stream>>>(b, gi, go, ch, pk);
            } else {
                if (with_pack) resample_batch_kernel<false, true, true><<<g3, block3(), 0, ctx->stream>>>(b, gi, go, ch, pk);
                else resample_batch_kernel<false, true, false><<<g3, block3(), 0, ctx->stream>>>(b, gi, go, ch, pk);
            }
        } else {
            if (gi.small) {
                if (with_pack) resample_batch_kernel<true, false, true><<<g3, block3(), 0, ctx->stream>>>(b, gi, go, ch, pk);
                else resample_batch_kernel<true, false, false><<<g3, block3(), 0, ctx->stream>>>(b, gi, go, ch, pk);
            } else {
                if (with_pack) resample_batch_kernel<false, false, true><<<g3, block3(), 0, ctx->stream>>>(b, gi, go, ch, pk);
                else resample_batch_kernel<false, false, false><<<g3, block3(), 0, ctx->stream>>>(b, gi, go, ch, pk);
            }
        }
        ctx->launches++;
        B200_CHECK_LAUNCH();
    }
    return B200REG_OK;
}

// ---- vector (f64 x 3, SoA) --------------------------------------------------------------------------
// LinearInterpolateImageFunction on a VectorImage: the same nested-lerp form, per component.
// ACCUM: out = acc + value (dvf_total + Resample(dvf_iter, tfm_total), deformable.py:154).
template <bool ACCUM, bool SMALL>
__global__ void __launch_bounds__(BX* BY, RS_MINB) resample_vec3_kernel(const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ acc,
                                                                const __grid_constant__ GeomD gi, const __grid_constant__ GeomD go,
                                                                const __grid_constant__ ChainD ch, double default_value)
{
    const int i = blockIdx.x * BX + threadIdx.x;
    const int j = blockIdx.y * BY + threadIdx.y;
    const int k = blockIdx.z;
    double c[3];
    out_to_in_cidx_block(go, gi, ch, i, j, k, c);
    if (i >= go.nx || j >= go.ny) return;
    const size_t o = ((size_t)k * go.ny + j) * go.nx + i;
    const size_t po = (size_t)go.nx * go.ny * go.nz, pi = (size_t)gi.nx * gi.ny * gi.nz;
    double v[3];
    if (inside_buffer(gi, c) && ch.vec_wsum) {
        interp_wsum_vec3_alt(in, gi, c, v);
    } else if (inside_buffer(gi, c)) {
        const LinW w = lin_setup(gi, c);
        if (SMALL) {
            v[0] = lin_eval_i32<double>(in, gi.nx, gi.nx * gi.ny, w);
            v[1] = lin_eval_i32<double>(in + pi, gi.nx, gi.nx * gi.ny, w);
            v[2] = lin_eval_i32<double>(in + 2 * pi, gi.nx, gi.nx * gi.ny, w);
        } else {
            v[0] = lin_eval<double>(in, gi, w);
            v[1] = lin_eval<double>(in + pi, gi, w);
            v[2] = lin_eval<double>(in + 2 * pi, gi, w);
        }
    } else {
        v[0] = v[1] = v[2] = default_value;
    }
    if (ACCUM) {
        out[o] = acc[o] + v[0];
        out[o + po] = acc[o + po] + v[1];
        out[o + 2 * po] = acc[o + 2 * po] + v[2];
    } else {
        out[o] = v[0];
        out[o + po] = v[1];
        out[o + 2 * po] = v[2];
    }
}

inline int resample_vec3(b200reg_ctx* ctx, const double* d_in, const b200reg_geom& gin, double* d_out, const b200reg_geom& gout,
                         const b200reg_transform* chain, int n_chain, double default_value, const double* d_acc = nullptr)
{
    if (n_chain == 0 && !d_acc && ctx->identity_copy && identity_resample_is_exact(gin, gout)) {
        if (d_in != d_out) B200_CUDA(cudaMemcpyAsync(d_out, d_in, 3 * nvox(gin) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        return B200REG_OK;
    }
    ChainD ch;
    B200_TRY(make_chain(chain, n_chain, &ch));
    chain_mark_on_grid(ctx, &ch, chain, n_chain, gout);
    const GeomD gi = make_geomd(gin), go = make_geomd(gout);
    const dim3 g = grid3(go.nx, go.ny, go.nz), b = block3();
    if (d_acc) {
        if (gi.small) resample_vec3_kernel<true, true><<<g, b, 0, ctx->stream>>>(d_in, d_out, d_acc, gi, go, ch, default_value);
        else resample_vec3_kernel<true, false><<<g, b, 0, ctx->stream>>>(d_in, d_out, d_acc, gi, go, ch, default_value);
    } else {
        if (gi.small) resample_vec3_kernel<false, true><<<g, b, 0, ctx->stream>>>(d_in, d_out, nullptr, gi, go, ch, default_value);
        else resample_vec3_kernel<false, false><<<g, b, 0, ctx->stream>>>(d_in, d_out, nullptr, gi, go, ch, default_value);
    }
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// TransformToDisplacementFieldFilter: D(x) = T(x) - x (per-voxel TransformPoint), SoA f64 output.
__global__ void __launch_bounds__(BX* BY) transform_to_dvf_kernel(double* __restrict__ out, const __grid_constant__ GeomD go, const __grid_constant__ ChainD ch)
{
    const int i = blockIdx.x * BX + threadIdx.x;
    const int j = blockIdx.y * BY + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= go.nx || j >= go.ny) return;
    double p[3], q[3];
    idx2pt(go, (double)i, (double)j, (double)k, p);
    q[0] = p[0];
    q[1] = p[1];
    q[2] = p[2];
    apply_chain(ch, q);
    const size_t o = ((size_t)k * go.ny + j) * go.nx + i, n = (size_t)go.nx * go.ny * go.nz;
    out[o] = q[0] - p[0];
    out[o + n] = q[1] - p[1];
    out[o + 2 * n] = q[2] - p[2];
}

}  // namespace b200
