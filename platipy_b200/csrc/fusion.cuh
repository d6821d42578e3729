// fusion.cuh -- label fusion (reference platipy/imaging/label/fusion.py:56-292) and small utility kernels.
#pragma once
#include "common.cuh"
#include "gauss.cuh"

namespace b200 {

// ---- layout / cast / reductions ---------------------------------------------------------------------
__global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        soa[q] = aos[3 * q];
        soa[q + n] = aos[3 * q + 1];
        soa[q + 2 * n] = aos[3 * q + 2];
    }
}
__global__ void soa_to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        aos[3 * q] = soa[q];
        aos[3 * q + 1] = soa[q + n];
        aos[3 * q + 2] = soa[q + 2 * n];
    }
}

// sitk.Cast: static_cast (no clamping; float -> int truncates toward zero)
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = (TO)in[q];
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_partial_kernel(const T* __restrict__ in, size_t n, double* __restrict__ partial)
{
    double mn = DBL_MAX, mx = -DBL_MAX;
    // four independent loads per step (one load per step left the kernel waiting on memory: 1.3 TB/s at 268 MB); min / max are
    // order-independent, so the result is the same
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; q + 3 * stride < n; q += 4 * stride) {
        const double v0 = (double)in[q], v1 = (double)in[q + stride], v2 = (double)in[q + 2 * stride], v3 = (double)in[q + 3 * stride];
        mn = fmin(fmin(mn, v0), fmin(v1, fmin(v2, v3)));
        mx = fmax(fmax(mx, v0), fmax(v1, fmax(v2, v3)));
    }
    for (; q < n; q += stride) {
        const double v = (double)in[q];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    __shared__ double sh[2][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) {
        sh[0][wid] = mn;
        sh[1][wid] = mx;
    }
    __syncthreads();
    if (wid == 0) {
        mn = lane < 8 ? sh[0][lane] : DBL_MAX;
        mx = lane < 8 ? sh[1][lane] : -DBL_MAX;
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) {
            partial[2 * blockIdx.x] = mn;
            partial[2 * blockIdx.x + 1] = mx;
        }
    }
}
__global__ void minmax_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out)
{
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (int q = threadIdx.x; q < nb; q += 32) {
        mn = fmin(mn, partial[2 * q]);
        mx = fmax(mx, partial[2 * q + 1]);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (threadIdx.x == 0) {
        out[0] = mn;
        out[1] = mx;
    }
}

// device-resident min/max into d_out[0..1]
template <typename T>
inline int minmax_device(b200reg_ctx* ctx, const T* d_in, size_t n, double* d_out, TempBuf* partial)
{
    const int nb = ctx->sm_count * 8;
    B200_TRY(partial->alloc(ctx, sizeof(double) * 2 * (size_t)nb));
    minmax_partial_kernel<T><<<nb, 256, 0, ctx->stream>>>(d_in, n, partial->as<double>());
    minmax_final_kernel<<<1, 32, 0, ctx->stream>>>(partial->as<double>(), nb, d_out);
    ctx->launches += 2;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// fixed-order sum of doubles: per-block partials then one block
template <typename T, typename F>
__global__ void __launch_bounds__(256) sum_partial_kernel(const T* __restrict__ in, size_t n, double* __restrict__ partial, F f)
{
    double s = 0.0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) s += f(in[q]);
    __shared__ double sh[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s = warp_sum(s);
    if (lane == 0) sh[wid] = s;
    __syncthreads();
    if (wid == 0) {
        s = lane < 8 ? sh[lane] : 0.0;
        s = warp_sum(s);
        if (lane == 0) partial[blockIdx.x] = s;
    }
}

// ---- compute_weight_map (fusion.py:148-177) -----------------------------------------------------------
// SquaredDifferenceImageFilter: (double(a) - double(b))^2 cast to float
__global__ void sqdiff_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const double d = (double)a[q] - (double)b[q];
        out[q] = (float)(d * d);
    }
}
// target * 0.0 + w  (keeps NaN/Inf propagation of the reference expression)
__global__ void const_weight_kernel(const float* __restrict__ target, float* __restrict__ out, size_t n, const double* __restrict__ d_sum,
                                    double factor, int use_sum)
{
    const float w = use_sum ? (float)(factor / d_sum[0]) : 1.0f;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = target[q] * 0.0f + w;
}
// sitk.Pow(raw + eps, -1.0): std::pow in double, stored as float
__global__ void local_weight_kernel(float* __restrict__ io, size_t n, double eps)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const float r = (float)((double)io[q] + eps);
        io[q] = (float)pow((double)r, -1.0);
    }
}
__global__ void sum_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out)
{
    double s = 0.0;
    for (int q = threadIdx.x; q < nb; q += 32) s += partial[q];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s;
}

struct IdentF {
    __device__ double operator()(float v) const { return (double)v; }
};

inline int weight_map(b200reg_ctx* ctx, const float* d_target, const float* d_moving, const b200reg_geom& g, int vote_type, double factor,
                      double sigma, double epsilon, float* d_weight)
{
    const size_t n = nvox(g);
    const int nb = ctx->sm_count * 8;
    if (vote_type == 0) {
        const_weight_kernel<<<nb, 256, 0, ctx->stream>>>(d_target, d_weight, n, nullptr, 0.0, 0);
        ctx->launches++;
    } else if (vote_type == 1) {
        TempBuf sq, part, sum;
        B200_TRY(sq.alloc(ctx, n * sizeof(float)));
        B200_TRY(part.alloc(ctx, sizeof(double) * (size_t)nb));
        B200_TRY(sum.alloc(ctx, sizeof(double)));
        sqdiff_kernel<<<nb, 256, 0, ctx->stream>>>(d_target, d_moving, sq.as<float>(), n);
        sum_partial_kernel<float, IdentF><<<nb, 256, 0, ctx->stream>>>(sq.as<float>(), n, part.as<double>(), IdentF());
        sum_final_kernel<<<1, 32, 0, ctx->stream>>>(part.as<double>(), nb, sum.as<double>());
        const_weight_kernel<<<nb, 256, 0, ctx->stream>>>(d_target, d_weight, n, sum.as<double>(), factor, 1);
        ctx->launches += 4;
    } else if (vote_type == 2) {
        TempBuf sq;
        B200_TRY(sq.alloc(ctx, n * sizeof(float)));
        sqdiff_kernel<<<nb, 256, 0, ctx->stream>>>(d_target, d_moving, sq.as<float>(), n);
        ctx->launches++;
        const double var[3] = { sigma * sigma, sigma * sigma, sigma * sigma };
        B200_TRY(discrete_gaussian_f32(ctx, sq.as<float>(), d_weight, g, var, 32, 0.01, 1));
        local_weight_kernel<<<nb, 256, 0, ctx->stream>>>(d_weight, n, epsilon);
        ctx->launches++;
    } else {
        return set_error(B200REG_ERR_UNSUPPORTED, "vote type %d is not supported (0 unweighted, 1 global, 2 local)", vote_type);
    }
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// ---- compute_weight_map, vote_type "block" (fusion.py:179-200) ------------------------------------------
// sitk.BoxMean(square_difference, radius): mean over the window [i - r, i + r] cropped to the image (itk::BoxMeanImageFilter
// divides by the number of pixels inside), accumulated in double (NumericTraits<float>::RealType).  ITK sums through an
// integral image; here the window is summed directly, separably, which differs from it only by double rounding.
template <typename TIN>
__global__ void __launch_bounds__(256) box_sum_axis_kernel(const TIN* __restrict__ in, double* __restrict__ out, int nx, int ny, int nz, int axis, int r)
{
    const size_t n = (size_t)nx * ny * nz;
    const int dims[3] = { nx, ny, nz };
    const size_t strides[3] = { 1, (size_t)nx, (size_t)nx * ny };
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int idx[3] = { (int)(q % nx), (int)((q / nx) % ny), (int)(q / ((size_t)nx * ny)) };
        const int c = idx[axis], lo = max(c - r, 0), hi = min(c + r, dims[axis] - 1);
        const size_t base = q - (size_t)c * strides[axis];
        double s = 0.0;
        for (int t = lo; t <= hi; ++t) s += (double)in[base + (size_t)t * strides[axis]];
        out[q] = s;
    }
}
// raw = float(sum / count); weight = factor * Pow(raw, -1.0) ** |gain / 2|, every stage a Float32 image (fusion.py:191-192)
__global__ void __launch_bounds__(256) block_weight_kernel(const double* __restrict__ sum, float* __restrict__ out, int nx, int ny, int nz, int rx, int ry,
                                                           int rz, double factor, double half_gain)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(q % nx), y = (int)((q / nx) % ny), z = (int)(q / ((size_t)nx * ny));
        const double cnt = (double)(min(x + rx, nx - 1) - max(x - rx, 0) + 1) * (double)(min(y + ry, ny - 1) - max(y - ry, 0) + 1) *
                           (double)(min(z + rz, nz - 1) - max(z - rz, 0) + 1);
        const float raw = (float)(sum[q] / cnt);
        const float inv = (float)pow((double)raw, -1.0);
        const float pw = (float)pow((double)inv, half_gain);
        out[q] = (float)((double)pw * factor);
    }
}
inline int weight_map_block(b200reg_ctx* ctx, const float* d_target, const float* d_moving, const b200reg_geom& g, const int32_t radius[3],
                            double factor, double gain, float* d_weight)
{
    const size_t n = nvox(g);
    const int nb = ctx->sm_count * 8;
    const int nx = g.size[0], ny = g.size[1], nz = g.size[2];
    TempBuf sq, s1, s2;
    B200_TRY(sq.alloc(ctx, n * sizeof(float)));
    B200_TRY(s1.alloc(ctx, n * sizeof(double)));
    B200_TRY(s2.alloc(ctx, n * sizeof(double)));
    sqdiff_kernel<<<nb, 256, 0, ctx->stream>>>(d_target, d_moving, sq.as<float>(), n);
    box_sum_axis_kernel<float><<<nb, 256, 0, ctx->stream>>>(sq.as<float>(), s1.as<double>(), nx, ny, nz, 0, radius[0]);
    box_sum_axis_kernel<double><<<nb, 256, 0, ctx->stream>>>(s1.as<double>(), s2.as<double>(), nx, ny, nz, 1, radius[1]);
    box_sum_axis_kernel<double><<<nb, 256, 0, ctx->stream>>>(s2.as<double>(), s1.as<double>(), nx, ny, nz, 2, radius[2]);
    block_weight_kernel<<<nb, 256, 0, ctx->stream>>>(s1.as<double>(), d_weight, nx, ny, nz, radius[0], radius[1], radius[2], factor, fabs(gain / 2.0));
    ctx->launches += 5;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// weight_map / max(weight_map) or / max(sitk.Mask(weight_map, mask)) (fusion.py:171-177,196-200): Float32 / double constant
__global__ void __launch_bounds__(256) mask_f32_kernel(const float* __restrict__ in, const uint8_t* __restrict__ mask, float* __restrict__ out, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = mask[q] ? in[q] : 0.0f;
}
__global__ void __launch_bounds__(256) divide_by_max_kernel(float* __restrict__ io, const double* __restrict__ d_minmax, size_t n)
{
    const double mx = d_minmax[1];
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x)
        io[q] = mx != 0.0 ? (float)((double)io[q] / mx) : FLT_MAX;  // itk::Functor::Div: B == 0 -> NumericTraits<float>::max()
}
inline int normalise_by_max(b200reg_ctx* ctx, float* d_weight, const uint8_t* d_mask, size_t n)
{
    const int nb = ctx->sm_count * 8;
    TempBuf part, mm, masked;
    B200_TRY(mm.alloc(ctx, 2 * sizeof(double)));
    const float* src = d_weight;
    if (d_mask) {
        B200_TRY(masked.alloc(ctx, n * sizeof(float)));
        mask_f32_kernel<<<nb, 256, 0, ctx->stream>>>(d_weight, d_mask, masked.as<float>(), n);
        ctx->launches++;
        src = masked.as<float>();
    }
    B200_TRY(minmax_device<float>(ctx, src, n, mm.as<double>(), &part));
    divide_by_max_kernel<<<nb, 256, 0, ctx->stream>>>(d_weight, mm.as<double>(), n);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// ---- combine_labels (fusion.py:253-288) ---------------------------------------------------------------
// float32 arithmetic in atlas order: num = sum_a w_a * float(L_a), den = sum_a w_a
__global__ void vote_accumulate_kernel(const uint8_t* __restrict__ label, const float* __restrict__ w, float* __restrict__ num,
                                       float* __restrict__ den, size_t n, int first)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const float wv = w[q];
        const float t = wv * (float)label[q];
        if (first) {
            num[q] = t;
            if (den) den[q] = wv;
        } else {
            num[q] = num[q] + t;
            if (den) den[q] = den[q] + wv;
        }
    }
}
// sitk.Mask(den, den == 0, maskingValue=1, outsideValue=1) then num / den
__global__ void vote_divide_kernel(float* __restrict__ num, const float* __restrict__ den, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        float d = den[q];
        if (d == 0.0f) d = 1.0f;
        num[q] = num[q] / d;
    }
}
// RescaleIntensityImageFilter(0, 1) + ThresholdImageFilter(lower, 1, outside 0) with min/max on device
template <typename T>
__global__ void rescale_threshold_kernel(const T* in, T* out, size_t n, const double* __restrict__ mm, double threshold,
                                         double type_eps, int rescale)
{
    const double mn = mm[0], mx = mm[1];
    double scale, shift;
    if (fabs((double)((T)mx - (T)mn)) > type_eps) scale = 1.0 / (mx - mn);
    else if (mx != 0.0) scale = 1.0 / mx;
    else scale = 0.0;
    shift = 0.0 - mn * scale;
    const T lower = (T)threshold;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        T r = in[q];
        if (rescale) {
            const double v = (double)r * scale + shift;
            r = (T)v;
            r = (r > (T)1) ? (T)1 : r;
            r = (r < (T)0) ? (T)0 : r;
        }
        if (threshold != 0.0) {
            if (!(r >= lower && r <= (T)1)) r = (T)0;
        }
        out[q] = r;
    }
}

inline int vote_finalize(b200reg_ctx* ctx, float* d_num, const float* d_den, const b200reg_geom& g, double smooth_variance, double threshold,
                         float* d_out)
{
    const size_t n = nvox(g);
    const int nb = ctx->sm_count * 8;
    if (d_den) {
        vote_divide_kernel<<<nb, 256, 0, ctx->stream>>>(d_num, d_den, n);
        ctx->launches++;
    }
    TempBuf sm, part, mm;
    B200_TRY(sm.alloc(ctx, n * sizeof(float)));
    const double var[3] = { smooth_variance, smooth_variance, smooth_variance };
    B200_TRY(discrete_gaussian_f32(ctx, d_num, sm.as<float>(), g, var, 32, 0.01, 1));
    B200_TRY(mm.alloc(ctx, 2 * sizeof(double)));
    B200_TRY(minmax_device<float>(ctx, sm.as<float>(), n, mm.as<double>(), &part));
    rescale_threshold_kernel<float><<<nb, 256, 0, ctx->stream>>>(sm.as<float>(), d_out, n, mm.as<double>(), threshold, (double)FLT_EPSILON, 1);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// BinaryThresholdImageFilter: (lower <= x <= upper) ? 1 : 0, compared as real numbers.
// (SimpleITK may cast the double thresholds to the pixel type first, which would turn 0.5 into 0 for integer
// label images; the real-valued comparison is what fusion.py:217-220 intends and is identical for float images.
// See DESIGN.md "uncertain ITK semantics".)
template <typename T>
__global__ void binary_threshold_kernel(const T* __restrict__ in, uint8_t* __restrict__ out, size_t n, double lower, double upper)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const double v = (double)in[q];
        out[q] = (lower <= v && v <= upper) ? 1 : 0;
    }
}

__global__ void pack_decision_kernel(const uint8_t* __restrict__ label, int bit, int32_t* __restrict__ packed, size_t n, int first)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int32_t v = (label[q] != 0 ? 1 : 0) << bit;
        packed[q] = first ? v : (packed[q] | v);
    }
}
__global__ void unpack_decision_kernel(const int32_t* __restrict__ packed, int bit, uint8_t* __restrict__ out, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = (uint8_t)((packed[q] >> bit) & 1);
}

// ---- sitk.STAPLE (itk::STAPLEImageFilter), binary EM ---------------------------------------------------
constexpr int STAPLE_MAX_RATERS = 32;
struct StaplePtrs {
    const uint8_t* d[STAPLE_MAX_RATERS];
    int n;
};
struct StapleState {
    double p[STAPLE_MAX_RATERS], q[STAPLE_MAX_RATERS], last_p[STAPLE_MAX_RATERS], last_q[STAPLE_MAX_RATERS];
    double g;
    int converged;
    int elapsed;
};

// W = mean_j D_j ; partial sums of W for g
__global__ void __launch_bounds__(256) staple_init_kernel(const __grid_constant__ StaplePtrs ptrs, double* __restrict__ W, size_t n,
                                                           double* __restrict__ partial)
{
    double s = 0.0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
        double w = 0.0;
        for (int j = 0; j < ptrs.n; ++j)
            if (ptrs.d[j][v] == 1) w = w + 1.0;
        w = w / (double)ptrs.n;
        W[v] = w;
        s += w;
    }
    __shared__ double sh[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s = warp_sum(s);
    if (lane == 0) sh[wid] = s;
    __syncthreads();
    if (wid == 0) {
        s = lane < 8 ? sh[lane] : 0.0;
        s = warp_sum(s);
        if (lane == 0) partial[blockIdx.x] = s;
    }
}
__global__ void staple_g_kernel(const double* __restrict__ partial, int nb, size_t n, double confidence, StapleState* st)
{
    double s = 0.0;
    for (int q = threadIdx.x; q < nb; q += 32) s += partial[q];
    s = warp_sum(s);
    if (threadIdx.x == 0) {
        st->g = (s / (double)n) * confidence;
        st->converged = 0;
        st->elapsed = 0;
        for (int j = 0; j < STAPLE_MAX_RATERS; ++j) {
            st->last_p[j] = -10.0;
            st->last_q[j] = -10.0;
        }
    }
}
// M-step partial sums: for every rater sum_{D=1} W and sum_{D=0} (1 - W); plus sum W and sum (1 - W)
__global__ void __launch_bounds__(256) staple_mstep_kernel(const __grid_constant__ StaplePtrs ptrs, const double* __restrict__ W, size_t n,
                                                            double* __restrict__ partial, const StapleState* st)
{
    if (st->converged) return;
    double pn[STAPLE_MAX_RATERS], qn[STAPLE_MAX_RATERS];
    for (int j = 0; j < ptrs.n; ++j) pn[j] = qn[j] = 0.0;
    double sw = 0.0, s1w = 0.0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
        const double w = W[v], w1 = 1.0 - w;
        sw += w;
        s1w += w1;
        for (int j = 0; j < ptrs.n; ++j) {
            if (ptrs.d[j][v] == 1) pn[j] += w;
            else qn[j] += w1;
        }
    }
    __shared__ double sh[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int stride = 2 * ptrs.n + 2;
    for (int t = 0; t < stride; ++t) {
        double val = t < ptrs.n ? pn[t] : (t < 2 * ptrs.n ? qn[t - ptrs.n] : (t == 2 * ptrs.n ? sw : s1w));
        val = warp_sum(val);
        if (lane == 0) sh[wid] = val;
        __syncthreads();
        if (wid == 0) {
            double x = lane < 8 ? sh[lane] : 0.0;
            x = warp_sum(x);
            if (lane == 0) partial[(size_t)blockIdx.x * stride + t] = x;
        }
        __syncthreads();
    }
}
__global__ void staple_pq_kernel(const double* __restrict__ partial, int nb, int n_raters, StapleState* st)
{
    if (st->converged) return;
    const int stride = 2 * n_raters + 2;
    __shared__ double tot[2 * STAPLE_MAX_RATERS + 2];
    for (int t = threadIdx.x >> 5; t < stride; t += blockDim.x >> 5) {
        double s = 0.0;
        for (int q = threadIdx.x & 31; q < nb; q += 32) s += partial[(size_t)q * stride + t];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) tot[t] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 0; j < n_raters; ++j) {
            st->p[j] = tot[j] / tot[2 * n_raters];
            st->q[j] = tot[n_raters + j] / tot[2 * n_raters + 1];
        }
    }
}
// E-step
__global__ void __launch_bounds__(256) staple_estep_kernel(const __grid_constant__ StaplePtrs ptrs, double* __restrict__ W, size_t n,
                                                            const StapleState* st)
{
    if (st->converged) return;
    const double g = st->g;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
        double alpha1 = 1.0, beta1 = 1.0;
        for (int j = 0; j < ptrs.n; ++j) {
            if (ptrs.d[j][v] == 1) {
                alpha1 = alpha1 * st->p[j];
                beta1 = beta1 * (1.0 - st->q[j]);
            } else {
                alpha1 = alpha1 * (1.0 - st->p[j]);
                beta1 = beta1 * st->q[j];
            }
        }
        W[v] = g * alpha1 / (g * alpha1 + (1.0 - g) * beta1);
    }
}
// convergence test after the E-step of iteration `iter`
__global__ void staple_converge_kernel(StapleState* st, int n_raters, unsigned iter)
{
    if (st->converged) return;
    bool flag = false;
    if (iter != 0) {
        flag = true;
        for (int j = 0; j < n_raters; ++j) {
            if (((st->p[j] - st->last_p[j]) * (st->p[j] - st->last_p[j])) > 1.0e-14) { flag = false; break; }
            if (((st->q[j] - st->last_q[j]) * (st->q[j] - st->last_q[j])) > 1.0e-14) { flag = false; break; }
        }
    }
    for (int j = 0; j < n_raters; ++j) {
        st->last_p[j] = st->p[j];
        st->last_q[j] = st->q[j];
    }
    st->elapsed = (int)iter;
    if (flag) st->converged = 1;
}


// ---- STAPLE by decision pattern ------------------------------------------------------------------------------
// In binary STAPLE the posterior W_i depends on voxel i only through its decision pattern (D_i1 .. D_iN): after the
// initialisation W = mean_j D_j (also a function of the pattern) every E-step gives W_i = f(pattern_i).  The EM over
// 67 M voxels is therefore an EM over the 2^N-bin histogram of patterns: one pass over the volume to build the
// histogram, the whole iteration loop on the table inside ONE thread block (no host round trips), one pass to write
// W = table[pattern].  Same formulas and rater order as itk::STAPLEImageFilter; only the association of the sums
// differs (histogram-weighted instead of voxel-serial), which the reference itself does not fix across thread
// counts.  Used for N <= STAPLE_PATTERN_MAX raters.
constexpr int STAPLE_PATTERN_MAX = 16;

__global__ void __launch_bounds__(256) staple_pattern_kernel(const __grid_constant__ StaplePtrs ptrs, uint32_t* __restrict__ pattern, size_t n,
                                                              unsigned long long* __restrict__ hist)
{
    // block-private histogram for up to 2^12 bins, direct (warp-aggregated) atomics above
    extern __shared__ unsigned int shist[];
    const int nbins = 1 << ptrs.n;
    const bool priv = ptrs.n <= 12;
    if (priv) {
        for (int b = threadIdx.x; b < nbins; b += blockDim.x) shist[b] = 0;
        __syncthreads();
    }
    unsigned long long zero_count = 0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
        uint32_t pat = 0;
        for (int j = 0; j < ptrs.n; ++j) pat |= (ptrs.d[j][v] == 1 ? 1u : 0u) << j;
        pattern[v] = pat;
        if (priv) atomicAdd(&shist[pat], 1u);
        else if (pat == 0) ++zero_count;
        else atomicAdd(&hist[pat], 1ull);
    }
    if (priv) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += blockDim.x)
            if (shist[b]) atomicAdd(&hist[b], (unsigned long long)shist[b]);
    } else {
        // the all-background pattern dominates: one atomic per warp
        for (int o = 16; o > 0; o >>= 1) zero_count += __shfl_down_sync(0xffffffffu, zero_count, o);
        if ((threadIdx.x & 31) == 0 && zero_count) atomicAdd(&hist[0], zero_count);
    }
}
// same, from an already packed decision mask (the multi-GPU exchange format)
__global__ void __launch_bounds__(256) staple_hist_packed_kernel(const int32_t* __restrict__ packed, uint32_t mask, size_t n,
                                                                  unsigned long long* __restrict__ hist, uint32_t* __restrict__ pattern)
{
    unsigned long long zero_count = 0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
        const uint32_t pat = (uint32_t)packed[v] & mask;
        pattern[v] = pat;
        if (pat == 0) ++zero_count;
        else atomicAdd(&hist[pat], 1ull);
    }
    for (int o = 16; o > 0; o >>= 1) zero_count += __shfl_down_sync(0xffffffffu, zero_count, o);
    if ((threadIdx.x & 31) == 0 && zero_count) atomicAdd(&hist[0], zero_count);
}

__device__ __forceinline__ double block_sum_1024(double v, double* sh)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = warp_sum(sh[lane]);
        if (lane == 0) sh[0] = r;
    }
    __syncthreads();
    r = sh[0];
    return r;
}

// One block runs the whole EM on the table.  table_w[b]: posterior of pattern b; out_tab[b]: value written to voxels
// (optionally rescaled to [0,1] over the patterns that occur, and thresholded).
__global__ void __launch_bounds__(1024) staple_em_table_kernel(const unsigned long long* __restrict__ hist, int n_raters, double confidence,
                                                                unsigned max_iter, double threshold, int rescale, double* __restrict__ table_w,
                                                                double* __restrict__ out_tab, StapleState* st)
{
    __shared__ double sh[32];
    __shared__ double p[STAPLE_MAX_RATERS], q[STAPLE_MAX_RATERS], lp[STAPLE_MAX_RATERS], lq[STAPLE_MAX_RATERS];
    __shared__ int flag;
    const int nbins = 1 << n_raters;
    const int tid = threadIdx.x;
    // W = mean_j D_j ; g = mean(W) * confidence
    double sw = 0.0, tot = 0.0;
    for (int b = tid; b < nbins; b += 1024) {
        const double h = (double)hist[b];
        const double w = (double)__popc(b) / (double)n_raters;
        table_w[b] = w;
        sw += h * w;
        tot += h;
    }
    sw = block_sum_1024(sw, sh);
    tot = block_sum_1024(tot, sh);
    const double g = (sw / tot) * confidence;
    if (tid < n_raters) {
        lp[tid] = -10.0;
        lq[tid] = -10.0;
    }
    unsigned iter = 0;
    for (; iter < max_iter; ++iter) {
        // M step
        double s_w = 0.0, s_1w = 0.0;
        for (int b = tid; b < nbins; b += 1024) {
            const double h = (double)hist[b], w = table_w[b];
            s_w += h * w;
            s_1w += h * (1.0 - w);
        }
        s_w = block_sum_1024(s_w, sh);
        s_1w = block_sum_1024(s_1w, sh);
        for (int j = 0; j < n_raters; ++j) {
            double pn = 0.0, qn = 0.0;
            for (int b = tid; b < nbins; b += 1024) {
                const double h = (double)hist[b], w = table_w[b];
                if ((b >> j) & 1) pn += h * w;
                else qn += h * (1.0 - w);
            }
            pn = block_sum_1024(pn, sh);
            qn = block_sum_1024(qn, sh);
            if (tid == 0) {
                p[j] = pn / s_w;
                q[j] = qn / s_1w;
            }
        }
        __syncthreads();
        // E step
        for (int b = tid; b < nbins; b += 1024) {
            double alpha1 = 1.0, beta1 = 1.0;
            for (int j = 0; j < n_raters; ++j) {
                if ((b >> j) & 1) {
                    alpha1 = alpha1 * p[j];
                    beta1 = beta1 * (1.0 - q[j]);
                } else {
                    alpha1 = alpha1 * (1.0 - p[j]);
                    beta1 = beta1 * q[j];
                }
            }
            table_w[b] = g * alpha1 / (g * alpha1 + (1.0 - g) * beta1);
        }
        __syncthreads();
        if (tid == 0) {
            int f = 0;
            if (iter != 0) {
                f = 1;
                for (int j = 0; j < n_raters; ++j) {
                    if (((p[j] - lp[j]) * (p[j] - lp[j])) > 1.0e-14) { f = 0; break; }
                    if (((q[j] - lq[j]) * (q[j] - lq[j])) > 1.0e-14) { f = 0; break; }
                }
            }
            for (int j = 0; j < n_raters; ++j) {
                lp[j] = p[j];
                lq[j] = q[j];
            }
            flag = f;
        }
        __syncthreads();
        if (flag) break;
    }
    // RescaleIntensity(0,1) + Threshold over the values that actually occur in the volume
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (int b = tid; b < nbins; b += 1024)
        if (hist[b]) {
            mn = fmin(mn, table_w[b]);
            mx = fmax(mx, table_w[b]);
        }
    mn = warp_min(mn);
    mx = warp_max(mx);
    __syncthreads();
    if ((tid & 31) == 0) {
        sh[tid >> 5] = mn;
    }
    __syncthreads();
    if (tid < 32) {
        double v = warp_min(sh[tid]);
        if (tid == 0) sh[0] = v;
    }
    __syncthreads();
    mn = sh[0];
    __syncthreads();
    if ((tid & 31) == 0) sh[tid >> 5] = mx;
    __syncthreads();
    if (tid < 32) {
        double v = warp_max(sh[tid]);
        if (tid == 0) sh[0] = v;
    }
    __syncthreads();
    mx = sh[0];
    double scale, shift;
    if (fabs(mx - mn) > DBL_EPSILON) scale = 1.0 / (mx - mn);
    else if (mx != 0.0) scale = 1.0 / mx;
    else scale = 0.0;
    shift = 0.0 - mn * scale;
    for (int b = tid; b < nbins; b += 1024) {
        double r = table_w[b];
        if (rescale) {
            r = r * scale + shift;
            r = (r > 1.0) ? 1.0 : r;
            r = (r < 0.0) ? 0.0 : r;
        }
        if (threshold != 0.0) {
            if (!(r >= threshold && r <= 1.0)) r = 0.0;
        }
        out_tab[b] = r;
    }
    if (tid == 0) {
        st->elapsed = (int)iter;
        st->converged = 1;
        st->g = g;
        for (int j = 0; j < n_raters; ++j) {
            st->p[j] = p[j];
            st->q[j] = q[j];
        }
    }
}
__global__ void staple_write_kernel(const uint32_t* __restrict__ pattern, const double* __restrict__ out_tab, double* __restrict__ out, size_t n)
{
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) out[v] = __ldg(out_tab + pattern[v]);
}

// ---- compact exchange formats of the sharded fusion ---------------------------------------------------
// STAPLE: one decision mask per structure, bit a = BinaryThreshold(label of atlas a, 0.5, 255) -- u8 for up to 8 atlases,
// u16 up to 16, u32 beyond.  Ranks own disjoint bits, so a SUM reduce-scatter of the masks is their bitwise OR.  The mask IS the
// decision pattern the pattern-histogram EM works on: the owner of a structure goes from the reduced mask to the
// fused probability in two passes (histogram; table look-up) without unpacking the decisions.
template <typename TM>
__global__ void pack_label_kernel(const uint8_t* __restrict__ label, int bit, TM* __restrict__ packed, size_t n, int first)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const TM v = (TM)((TM)(label[q] != 0 ? 1 : 0) << bit);
        packed[q] = first ? v : (TM)(packed[q] | v);
    }
}
// software pext: the bits of `v` selected by `mask`, packed towards bit 0 in ascending order (rater j = j-th holder atlas)
__device__ __forceinline__ uint32_t gather_bits(uint32_t v, uint32_t mask)
{
    uint32_t out = 0, k = 0;
    while (mask) {
        const uint32_t low = mask & (0u - mask);
        if (v & low) out |= 1u << k;
        ++k;
        mask ^= low;
    }
    return out;
}
template <typename TM, bool DENSE>
__global__ void __launch_bounds__(256) staple_hist_mask_kernel(const TM* __restrict__ packed, uint32_t holder_mask, int n_raters, size_t n,
                                                                unsigned long long* __restrict__ hist)
{
    extern __shared__ unsigned int shist[];
    const int nbins = 1 << n_raters;
    const bool priv = n_raters <= 12;
    if (priv) {
        for (int b = threadIdx.x; b < nbins; b += blockDim.x) shist[b] = 0;
        __syncthreads();
    }
    unsigned long long zero_count = 0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
        const uint32_t raw = (uint32_t)packed[v];
        const uint32_t pat = DENSE ? (raw & holder_mask) : gather_bits(raw, holder_mask);
        if (pat == 0) ++zero_count;  // the all-background pattern dominates: counted in a register
        else if (priv) atomicAdd(&shist[pat], 1u);
        else atomicAdd(&hist[pat], 1ull);
    }
    for (int o = 16; o > 0; o >>= 1) zero_count += __shfl_down_sync(0xffffffffu, zero_count, o);
    if ((threadIdx.x & 31) == 0 && zero_count) atomicAdd(&hist[0], zero_count);
    if (priv) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += blockDim.x)
            if (shist[b]) atomicAdd(&hist[b], (unsigned long long)shist[b]);
    }
}
template <typename TM, bool DENSE>
__global__ void staple_write_mask_kernel(const TM* __restrict__ packed, uint32_t holder_mask, const double* __restrict__ out_tab,
                                         double* __restrict__ out, size_t n)
{
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
        const uint32_t raw = (uint32_t)packed[v];
        out[v] = __ldg(out_tab + (DENSE ? (raw & holder_mask) : gather_bits(raw, holder_mask)));
    }
}

// Unweighted vote (fusion.py:253-276 with every weight map == 1): num = sum_a float(L_a) is a small integer, so the float32
// accumulation of the reference is exact and, for binary labels and up to 255 atlases, a u8 count carries the same information
// in a quarter of the bytes (NCCL reduces 8-bit integers natively; it has no 16-bit integer type); den = number of atlases
// holding the structure, known on the host.  A label value above 1 raises `*flag`: the caller then takes the float32 path.
__global__ void count_accumulate_kernel(const uint8_t* __restrict__ label, uint8_t* __restrict__ counts, size_t n, int first, int* __restrict__ flag)
{
    bool big = false;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const uint8_t v = label[q];
        big |= v > 1;
        counts[q] = first ? v : (uint8_t)(counts[q] + v);
    }
    if (__any_sync(0xffffffffu, big) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
__global__ void counts_to_prob_kernel(const uint8_t* __restrict__ counts, float den, float* __restrict__ out, size_t n)
{
    // the guarded division of vote_divide_kernel, with num = float(count) and a constant den
    const float d = den == 0.0f ? 1.0f : den;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = (float)counts[q] / d;
}

}  // namespace b200
