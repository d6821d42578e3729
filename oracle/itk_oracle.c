/*
 * itk_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * PARITY UNPINNED: the arithmetic of the reference hot path lives in SimpleITK 2.3.1 / ITK 5.3
 * (reference pyproject.toml:24, poetry.lock:4523-4524), which is neither vendored in /root/reference
 * nor installable here (no network).  The reference's own tests hold no golden vectors for this path
 * (only Dice thresholds, platipy/imaging/tests/test_cardiac.py:142,231,237).  This file is therefore a
 * restatement of the *published* ITK 5.3 filter algorithms, written from the filter semantics, anchored
 * on the reference's call sites.  Nothing here was checked against a running SimpleITK.
 *
 * EXCEPTION (pinned): orc_signed_maurer and orc_label_contour reproduce, through oracle/comparison_ref.py, the eleven golden numbers
 * of the reference's own known-answer tests for the surface metrics (platipy/imaging/tests/test_metrics.py:6-67), which were
 * produced by the real SimpleITK -- see tests/test_reference_golden_metrics.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (platipy_b200) never imports it.
 *
 * Conventions: arrays are C-order [z][y][x] (x fastest) = numpy view of a SimpleITK image; vector images
 * are AoS [z][y][x][3] of double (sitkVectorFloat64), components (dx,dy,dz) in physical mm.
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC  (no FMA contraction: the CUDA
 * parity kernels are compiled with -fmad=false and are expected to agree bit-for-bit with this file
 * wherever the operation order is the same).
 *
 * Each function cites the reference call site it serves (paths relative to /root/reference) and the
 * ITK class whose algorithm it restates.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* Named semantic switches.  Every ITK behaviour this restatement could only RECALL (confidence M)   */
/* and that is cheap to state both ways is a switch with a name: the default is the recalled behaviour, the other      */
/* value is the plausible alternative.  libb200reg.so has the same switches under the same names                       */
/* (b200reg_set_semantic), so a correction found against a real SimpleITK is a flag flip on both sides, not an edit    */
/* of CUDA code.  tools/validate_against_sitk.py reports which setting of each switch matches SimpleITK.               */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { const char* name; int value; const char* meaning; } orc_semantic;
static orc_semantic g_semantics[] = {
    { "discrete_gaussian_axis_order", 0, "DiscreteGaussianImageFilter pass order: 0 = z, y, x (recalled); 1 = x, y, z" },
    { "recursive_gaussian_axis_order", 0, "SmoothingRecursiveGaussianImageFilter pass order: 0 = z, x, y (recalled); 1 = x, y, z" },
    { "resample_linear_scanline", 1, "ResampleImageFilter with a linear transform: 1 = scan-line continuous index (recalled); 0 = per-voxel" },
    { "dvf_transform_interpolation", 0, "interpolator inside DisplacementFieldTransform: 0 = weighted sum of 8 neighbours (recalled); 1 = nested lerps" },
    { "vector_resample_interpolation", 0, "linear interpolation of a vector image in ResampleImageFilter: 0 = nested lerps (recalled); 1 = weighted sum" },
    { "binary_threshold_in_pixel_type", 0, "BinaryThreshold bounds: 0 = compared as real numbers (recalled); 1 = cast to the pixel type first" },
};
#define ORC_N_SEMANTICS ((int)(sizeof(g_semantics) / sizeof(g_semantics[0])))
enum { SEM_DG_ORDER = 0, SEM_RG_ORDER, SEM_SCANLINE, SEM_DVF_INTERP, SEM_VEC_INTERP, SEM_BT_PIXEL };
#define SEM(i) (g_semantics[i].value)
ORC_API int orc_set_semantic(const char* name, int value)
{
    for (int i = 0; i < ORC_N_SEMANTICS; ++i)
        if (strcmp(g_semantics[i].name, name) == 0) { g_semantics[i].value = value; return 0; }
    return -1;
}
ORC_API int orc_get_semantic(const char* name)
{
    for (int i = 0; i < ORC_N_SEMANTICS; ++i)
        if (strcmp(g_semantics[i].name, name) == 0) return g_semantics[i].value;
    return -1;
}
ORC_API int orc_num_semantics(void) { return ORC_N_SEMANTICS; }
ORC_API const char* orc_semantic_name(int i) { return (i >= 0 && i < ORC_N_SEMANTICS) ? g_semantics[i].name : NULL; }
ORC_API const char* orc_semantic_meaning(int i) { return (i >= 0 && i < ORC_N_SEMANTICS) ? g_semantics[i].meaning : NULL; }

/* ------------------------------------------------------------------------------------------------ */
/* Geometry (itk::ImageBase): index<->physical point transforms                                      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t size[3];      /* x, y, z */
    double spacing[3];
    double origin[3];
    double direction[9];  /* row-major 3x3 */
} orc_geom;

typedef struct {
    int nx, ny, nz;
    double origin[3];
    double i2p[9]; /* Direction * diag(Spacing) */
    double p2i[9]; /* inverse */
} geomx;

static void inv3(const double* m, double* o)
{
    /* diagonal matrices (identity direction): exact reciprocal, as an SVD pseudo-inverse gives */
    if (m[1] == 0 && m[2] == 0 && m[3] == 0 && m[5] == 0 && m[6] == 0 && m[7] == 0) {
        memset(o, 0, 9 * sizeof(double));
        o[0] = 1.0 / m[0]; o[4] = 1.0 / m[4]; o[8] = 1.0 / m[8];
        return;
    }
    double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    o[0] = c00 / det; o[1] = (m[2] * m[7] - m[1] * m[8]) / det; o[2] = (m[1] * m[5] - m[2] * m[4]) / det;
    o[3] = c01 / det; o[4] = (m[0] * m[8] - m[2] * m[6]) / det; o[5] = (m[2] * m[3] - m[0] * m[5]) / det;
    o[6] = c02 / det; o[7] = (m[1] * m[6] - m[0] * m[7]) / det; o[8] = (m[0] * m[4] - m[1] * m[3]) / det;
}

static void geomx_init(geomx* g, const orc_geom* s)
{
    g->nx = s->size[0]; g->ny = s->size[1]; g->nz = s->size[2];
    for (int r = 0; r < 3; ++r) {
        g->origin[r] = s->origin[r];
        for (int c = 0; c < 3; ++c) g->i2p[r * 3 + c] = s->direction[r * 3 + c] * s->spacing[c];
    }
    inv3(g->i2p, g->p2i);
}

/* ImageBase::TransformIndexToPhysicalPoint: sum over columns, then + origin */
static inline void idx2pt(const geomx* g, double i0, double i1, double i2, double* p)
{
    for (int r = 0; r < 3; ++r) {
        double sum = 0.0;
        sum += g->i2p[r * 3 + 0] * i0;
        sum += g->i2p[r * 3 + 1] * i1;
        sum += g->i2p[r * 3 + 2] * i2;
        p[r] = sum + g->origin[r];
    }
}
/* ImageBase::TransformPhysicalPointToContinuousIndex */
static inline void pt2cidx(const geomx* g, const double* p, double* c)
{
    double v0 = p[0] - g->origin[0], v1 = p[1] - g->origin[1], v2 = p[2] - g->origin[2];
    for (int r = 0; r < 3; ++r) {
        double sum = 0.0;
        sum += g->p2i[r * 3 + 0] * v0;
        sum += g->p2i[r * 3 + 1] * v1;
        sum += g->p2i[r * 3 + 2] * v2;
        c[r] = sum;
    }
}
/* ImageFunction::IsInsideBuffer(ContinuousIndex): [start-0.5, start+size-0.5) ; NaN -> outside */
static inline int inside_buffer(const geomx* g, const double* c)
{
    return (c[0] >= -0.5 && c[0] < g->nx - 0.5 && c[1] >= -0.5 && c[1] < g->ny - 0.5 && c[2] >= -0.5 && c[2] < g->nz - 0.5);
}

/* ------------------------------------------------------------------------------------------------ */
/* Pixel access in double for the supported scalar types (SimpleITK pixel IDs)                       */
/* ------------------------------------------------------------------------------------------------ */
enum { ORC_I8 = 0, ORC_U8 = 1, ORC_I16 = 2, ORC_U16 = 3, ORC_I32 = 4, ORC_U32 = 5, ORC_I64 = 6, ORC_U64 = 7, ORC_F32 = 8, ORC_F64 = 9 };

static inline double ld(const void* p, int dt, size_t i)
{
    switch (dt) {
    case ORC_I8: return (double)((const int8_t*)p)[i];
    case ORC_U8: return (double)((const uint8_t*)p)[i];
    case ORC_I16: return (double)((const int16_t*)p)[i];
    case ORC_U16: return (double)((const uint16_t*)p)[i];
    case ORC_I32: return (double)((const int32_t*)p)[i];
    case ORC_U32: return (double)((const uint32_t*)p)[i];
    case ORC_I64: return (double)((const int64_t*)p)[i];
    case ORC_U64: return (double)((const uint64_t*)p)[i];
    case ORC_F32: return (double)((const float*)p)[i];
    default: return ((const double*)p)[i];
    }
}
/* ResampleImageFilter::CastPixelWithBoundsChecking: clamp to [lowest,max] of the output type, then C cast */
static inline void st(void* p, int dt, size_t i, double v)
{
#define CLAMPST(T, LO, HI) { double w = v; if (w < (double)(LO)) w = (double)(LO); if (w > (double)(HI)) w = (double)(HI); ((T*)p)[i] = (T)w; } break
    switch (dt) {
    case ORC_I8: CLAMPST(int8_t, INT8_MIN, INT8_MAX);
    case ORC_U8: CLAMPST(uint8_t, 0, UINT8_MAX);
    case ORC_I16: CLAMPST(int16_t, INT16_MIN, INT16_MAX);
    case ORC_U16: CLAMPST(uint16_t, 0, UINT16_MAX);
    case ORC_I32: CLAMPST(int32_t, INT32_MIN, INT32_MAX);
    case ORC_U32: CLAMPST(uint32_t, 0, UINT32_MAX);
    case ORC_I64: {
        if (v <= -9223372036854775808.0) ((int64_t*)p)[i] = INT64_MIN;
        else if (v >= 9223372036854775808.0) ((int64_t*)p)[i] = INT64_MAX;
        else ((int64_t*)p)[i] = (int64_t)v;
    } break;
    case ORC_U64: {
        if (v <= 0.0) ((uint64_t*)p)[i] = 0;
        else if (v >= 18446744073709551616.0) ((uint64_t*)p)[i] = UINT64_MAX;
        else ((uint64_t*)p)[i] = (uint64_t)v;
    } break;
    case ORC_F32: {
        double w = v; if (w < -(double)FLT_MAX) w = -(double)FLT_MAX; if (w > (double)FLT_MAX) w = (double)FLT_MAX;
        ((float*)p)[i] = (float)w;
    } break;
    default: ((double*)p)[i] = v; break;
    }
#undef CLAMPST
}
/* default pixel value: static_cast<PixelType>(double) without the interpolator clamp path */
static inline void st_default(void* p, int dt, size_t i, double v) { st(p, dt, i, v); }

/* ------------------------------------------------------------------------------------------------ */
/* Interpolators                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
/* LinearInterpolateImageFunction::EvaluateOptimized(Dispatch<3>): nested lerps x, y, z written
 * a + (b - a) * d in double; base index clamped up to start (distance <= 0 -> axis skipped), upper
 * neighbour beyond the last index -> lower value alone.  Skipping an axis equals lerping with d = 0 and
 * clamping equals lerping two equal values, so the branch-free form below is bit-identical. */
static inline double interp_linear_scalar(const void* img, int dt, const geomx* g, const double* c)
{
    int b0 = (int)floor(c[0]), b1 = (int)floor(c[1]), b2 = (int)floor(c[2]);
    if (b0 < 0) b0 = 0;
    if (b1 < 0) b1 = 0;
    if (b2 < 0) b2 = 0;
    double d0 = c[0] - (double)b0, d1 = c[1] - (double)b1, d2 = c[2] - (double)b2;
    if (d0 <= 0.) d0 = 0.;
    if (d1 <= 0.) d1 = 0.;
    if (d2 <= 0.) d2 = 0.;
    int u0 = b0 + 1 > g->nx - 1 ? g->nx - 1 : b0 + 1;
    int u1 = b1 + 1 > g->ny - 1 ? g->ny - 1 : b1 + 1;
    int u2 = b2 + 1 > g->nz - 1 ? g->nz - 1 : b2 + 1;
    size_t sx = 1, sy = (size_t)g->nx, sz = (size_t)g->nx * g->ny;
#define PX(i, j, k) ld(img, dt, (size_t)(k) * sz + (size_t)(j) * sy + (size_t)(i) * sx)
    double v000 = PX(b0, b1, b2), v100 = PX(u0, b1, b2), v010 = PX(b0, u1, b2), v110 = PX(u0, u1, b2);
    double v001 = PX(b0, b1, u2), v101 = PX(u0, b1, u2), v011 = PX(b0, u1, u2), v111 = PX(u0, u1, u2);
#undef PX
    double vx00 = v000 + (v100 - v000) * d0;
    double vx10 = v010 + (v110 - v010) * d0;
    double vxx0 = vx00 + (vx10 - vx00) * d1;
    double vx01 = v001 + (v101 - v001) * d0;
    double vx11 = v011 + (v111 - v011) * d0;
    double vxx1 = vx01 + (vx11 - vx01) * d1;
    return vxx0 + (vxx1 - vxx0) * d2;
}
/* NearestNeighborInterpolateImageFunction: Math::RoundHalfIntegerUp = floor(x + 0.5) */
static inline double interp_nn_scalar(const void* img, int dt, const geomx* g, const double* c)
{
    int i0 = (int)floor(c[0] + 0.5), i1 = (int)floor(c[1] + 0.5), i2 = (int)floor(c[2] + 0.5);
    return ld(img, dt, ((size_t)i2 * g->ny + i1) * g->nx + i0);
}
/* Same nested-lerp form applied component-wise: LinearInterpolateImageFunction on a VectorImage
 * (sitk.Resample of a VectorFloat64 image; reference deformable.py:130,137,154,185) */
static inline void interp_linear_vec3(const double* img, const geomx* g, const double* c, double* out)
{
    int b0 = (int)floor(c[0]), b1 = (int)floor(c[1]), b2 = (int)floor(c[2]);
    if (b0 < 0) b0 = 0;
    if (b1 < 0) b1 = 0;
    if (b2 < 0) b2 = 0;
    double d0 = c[0] - (double)b0, d1 = c[1] - (double)b1, d2 = c[2] - (double)b2;
    if (d0 <= 0.) d0 = 0.;
    if (d1 <= 0.) d1 = 0.;
    if (d2 <= 0.) d2 = 0.;
    int u0 = b0 + 1 > g->nx - 1 ? g->nx - 1 : b0 + 1;
    int u1 = b1 + 1 > g->ny - 1 ? g->ny - 1 : b1 + 1;
    int u2 = b2 + 1 > g->nz - 1 ? g->nz - 1 : b2 + 1;
    size_t sy = (size_t)g->nx, sz = (size_t)g->nx * g->ny;
    for (int k = 0; k < 3; ++k) {
#define PX(i, j, l) img[((size_t)(l) * sz + (size_t)(j) * sy + (size_t)(i)) * 3 + k]
        double v000 = PX(b0, b1, b2), v100 = PX(u0, b1, b2), v010 = PX(b0, u1, b2), v110 = PX(u0, u1, b2);
        double v001 = PX(b0, b1, u2), v101 = PX(u0, b1, u2), v011 = PX(b0, u1, u2), v111 = PX(u0, u1, u2);
#undef PX
        double vx00 = v000 + (v100 - v000) * d0;
        double vx10 = v010 + (v110 - v010) * d0;
        double vxx0 = vx00 + (vx10 - vx00) * d1;
        double vx01 = v001 + (v101 - v001) * d0;
        double vx11 = v011 + (v111 - v011) * d0;
        double vxx1 = vx01 + (vx11 - vx01) * d1;
        out[k] = vxx0 + (vxx1 - vxx0) * d2;
    }
}
/* VectorLinearInterpolateImageFunction (the interpolator inside DisplacementFieldTransform):
 * weighted sum over the 8 neighbours, bit k of the counter = upper neighbour in dim k, indices clamped
 * into the buffer, neighbours with zero overlap skipped, early exit when the total overlap is exactly 1 */
static inline void interp_wsum_vec3(const double* img, const geomx* g, const double* c, double* out)
{
    int b[3] = { (int)floor(c[0]), (int)floor(c[1]), (int)floor(c[2]) };
    double d[3] = { c[0] - (double)b[0], c[1] - (double)b[1], c[2] - (double)b[2] };
    int n[3] = { g->nx, g->ny, g->nz };
    out[0] = out[1] = out[2] = 0.0;
    double total = 0.0;
    for (unsigned counter = 0; counter < 8; ++counter) {
        double overlap = 1.0;
        unsigned upper = counter;
        int ni[3];
        for (int dim = 0; dim < 3; ++dim) {
            if (upper & 1) {
                ni[dim] = b[dim] + 1;
                if (ni[dim] > n[dim] - 1) ni[dim] = n[dim] - 1;
                overlap *= d[dim];
            } else {
                ni[dim] = b[dim];
                if (ni[dim] < 0) ni[dim] = 0;
                overlap *= 1.0 - d[dim];
            }
            upper >>= 1;
        }
        if (overlap) {
            const double* px = img + (((size_t)ni[2] * n[1] + ni[1]) * n[0] + ni[0]) * 3;
            out[0] += overlap * px[0];
            out[1] += overlap * px[1];
            out[2] += overlap * px[2];
            total += overlap;
        }
        if (total == 1.0) break;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Transform chains                                                                                  */
/* ------------------------------------------------------------------------------------------------ */
enum { ORC_TFM_AFFINE = 0, ORC_TFM_DVF = 1 };
typedef struct {
    int32_t kind;
    int32_t pad;
    double matrix[9];   /* affine: p' = M p + offset (MatrixOffsetTransformBase::TransformPoint) */
    double offset[3];
    const double* dvf;  /* DVF: AoS f64 [z][y][x][3] */
    orc_geom dvf_geom;
} orc_transform;

typedef struct { int kind; double matrix[9], offset[3]; const double* dvf; geomx g; } tfmx;

/* DisplacementFieldTransform::TransformPoint (reference deformable.py:139,296): p + D(p) if p is inside
 * the field buffer, p unchanged otherwise.  Chain is given in application order (first applied first),
 * i.e. the reverse of sitk.CompositeTransform's add order. */
static inline void apply_chain(const tfmx* t, int nt, double* p)
{
    for (int i = 0; i < nt; ++i) {
        if (t[i].kind == ORC_TFM_AFFINE) {
            double q[3];
            for (int r = 0; r < 3; ++r) {
                double sum = 0.0;
                sum += t[i].matrix[r * 3 + 0] * p[0];
                sum += t[i].matrix[r * 3 + 1] * p[1];
                sum += t[i].matrix[r * 3 + 2] * p[2];
                q[r] = sum + t[i].offset[r];
            }
            p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
        } else {
            double c[3], dd[3];
            pt2cidx(&t[i].g, p, c);
            if (inside_buffer(&t[i].g, c)) {
                if (SEM(SEM_DVF_INTERP)) interp_linear_vec3(t[i].dvf, &t[i].g, c, dd);
                else interp_wsum_vec3(t[i].dvf, &t[i].g, c, dd);
                p[0] += dd[0]; p[1] += dd[1]; p[2] += dd[2];
            }
        }
    }
}
static int chain_is_linear(const tfmx* t, int nt)
{
    if (!SEM(SEM_SCANLINE)) return 0;
    for (int i = 0; i < nt; ++i) if (t[i].kind != ORC_TFM_AFFINE) return 0;
    return 1;
}
static tfmx* chain_prepare(const orc_transform* t, int nt)
{
    tfmx* x = (tfmx*)calloc(nt > 0 ? nt : 1, sizeof(tfmx));
    for (int i = 0; i < nt; ++i) {
        x[i].kind = t[i].kind;
        memcpy(x[i].matrix, t[i].matrix, sizeof(x[i].matrix));
        memcpy(x[i].offset, t[i].offset, sizeof(x[i].offset));
        x[i].dvf = t[i].dvf;
        if (t[i].kind == ORC_TFM_DVF) geomx_init(&x[i].g, &t[i].dvf_geom);
    }
    return x;
}

/* ------------------------------------------------------------------------------------------------ */
/* ResampleImageFilter (reference utils.py:176-190, 257-267; deformable.py:130,137,140,154,185,281-301) */
/* ------------------------------------------------------------------------------------------------ */
enum { ORC_INTERP_NN = 1, ORC_INTERP_LINEAR = 2 };

/* continuous input index of output voxel (i,j,k).  Linear transforms take ITK's scan-line path
 * (ResampleImageFilter::LinearThreadedGenerateData): the continuous index is evaluated at the first
 * index of the row and one-past-the-last, and interpolated with alpha = i / size_x. */
static inline void out_to_in_cidx(const geomx* go, const geomx* gi, const tfmx* t, int nt, int linear, int i, int j, int k, double* c)
{
    double p[3];
    if (!linear) {
        idx2pt(go, (double)i, (double)j, (double)k, p);
        apply_chain(t, nt, p);
        pt2cidx(gi, p, c);
    } else {
        double cs[3], ce[3];
        idx2pt(go, 0.0, (double)j, (double)k, p);
        apply_chain(t, nt, p);
        pt2cidx(gi, p, cs);
        idx2pt(go, (double)go->nx, (double)j, (double)k, p);
        apply_chain(t, nt, p);
        pt2cidx(gi, p, ce);
        double alpha = (double)i / (double)go->nx;
        for (int r = 0; r < 3; ++r) c[r] = cs[r] + alpha * (ce[r] - cs[r]);
    }
}

/* ---- itk::BSplineInterpolateImageFunction, spline order 3 (sitk.sitkBSpline; reference deformable.py:221-224
 * interp_order, utils.py:148-192 interpolator) ------------------------------------------------------------------ */
/* itk::BSplineDecompositionImageFilter::DataToCoefficients1D on one line (stride s, length n), in place:
 * gain, causal initialisation (truncated at the horizon ceil(log(1e-10) / log|z|), exact mirror sum for short
 * lines), causal and anti-causal recursions with the cubic pole z = sqrt(3) - 2. */
static void bspline3_line(double* c, size_t s, int n)
{
    if (n == 1) return;
    const double z = sqrt(3.0) - 2.0, tol = 1e-10;
    double c0 = 1.0;
    c0 = c0 * (1.0 - z) * (1.0 - 1.0 / z);
    for (int k = 0; k < n; ++k) c[k * s] *= c0;
    {
        double zn = z, sum;
        long horizon = (long)ceil(log(tol) / log(fabs(z)));
        if (horizon < n) {
            sum = c[0];
            for (long k = 1; k < horizon; ++k) { sum += zn * c[k * s]; zn *= z; }
            c[0] = sum;
        } else {
            const double iz = 1.0 / z;
            double z2n = pow(z, (double)(n - 1));
            sum = c[0] + z2n * c[(size_t)(n - 1) * s];
            z2n *= z2n * iz;
            for (int k = 1; k <= n - 2; ++k) { sum += (zn + z2n) * c[k * s]; zn *= z; z2n *= iz; }
            c[0] = sum / (1.0 - zn * zn);
        }
    }
    for (int k = 1; k < n; ++k) c[k * s] += z * c[(k - 1) * s];
    c[(size_t)(n - 1) * s] = (z / (z * z - 1.0)) * (z * c[(size_t)(n - 2) * s] + c[(size_t)(n - 1) * s]);
    for (int k = n - 2; k >= 0; --k) c[k * s] = z * (c[(k + 1) * s] - c[k * s]);
}
/* coefficients of the whole volume: dimensions x, y, z in turn (DataToCoefficientsND) */
static double* bspline3_coefficients(const void* in, int dtype, const geomx* g)
{
    const size_t n = (size_t)g->nx * g->ny * g->nz;
    double* c = (double*)malloc(n * sizeof(double));
    if (!c) return NULL;
    for (size_t i = 0; i < n; ++i) c[i] = ld(in, dtype, i);
    const size_t sy = (size_t)g->nx, sz = (size_t)g->nx * g->ny;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)g->ny * g->nz; ++r) bspline3_line(c + (size_t)r * sy, 1, g->nx);
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)g->nx * g->nz; ++r) bspline3_line(c + (size_t)(r / g->nx) * sz + (size_t)(r % g->nx), sy, g->ny);
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)g->nx * g->ny; ++r) bspline3_line(c + (size_t)r, sz, g->nz);
    return c;
}
/* EvaluateAtContinuousIndexInternal: support floor((float)x) - 1 .. + 2, cubic weights, mirror boundary, sum over the
 * 64 points with x fastest, weight = ((1 * wx) * wy) * wz */
static inline double interp_bspline3(const double* coef, const geomx* g, const double* x)
{
    const int n[3] = { g->nx, g->ny, g->nz };
    long idx[3][4];
    double w[3][4];
    for (int d = 0; d < 3; ++d) {
        long indx = (long)floor((float)x[d]) - 1;
        for (int k = 0; k < 4; ++k) idx[d][k] = indx++;
        const double t = x[d] - (double)idx[d][1];
        w[d][3] = (1.0 / 6.0) * t * t * t;
        w[d][0] = (1.0 / 6.0) + 0.5 * t * (t - 1.0) - w[d][3];
        w[d][2] = t + w[d][0] - 2.0 * w[d][3];
        w[d][1] = 1.0 - w[d][0] - w[d][2] - w[d][3];
        for (int k = 0; k < 4; ++k) {
            if (n[d] == 1) idx[d][k] = 0;
            else {
                if (idx[d][k] < 0) idx[d][k] = -idx[d][k];
                if (idx[d][k] > n[d] - 1) idx[d][k] = (n[d] - 1) - (idx[d][k] - (n[d] - 1));
                /* lines shorter than the support: keep the single reflection inside the buffer */
                if (idx[d][k] < 0) idx[d][k] = 0;
                if (idx[d][k] > n[d] - 1) idx[d][k] = n[d] - 1;
            }
        }
    }
    double v = 0.0;
    for (int kz = 0; kz < 4; ++kz)
        for (int ky = 0; ky < 4; ++ky)
            for (int kx = 0; kx < 4; ++kx) {
                double ww = 1.0;
                ww *= w[0][kx];
                ww *= w[1][ky];
                ww *= w[2][kz];
                v += ww * coef[((size_t)idx[2][kz] * n[1] + (size_t)idx[1][ky]) * n[0] + (size_t)idx[0][kx]];
            }
    return v;
}

ORC_API int orc_bspline3_coefficients(const void* in, int dtype, const orc_geom* gin, double* out)
{
    geomx gi;
    geomx_init(&gi, gin);
    double* c = bspline3_coefficients(in, dtype, &gi);
    if (!c) return -1;
    memcpy(out, c, (size_t)gi.nx * gi.ny * gi.nz * sizeof(double));
    free(c);
    return 0;
}

ORC_API int orc_resample_scalar(const void* in, int dtype, const orc_geom* gin, void* out, const orc_geom* gout,
                                const orc_transform* tf, int ntf, int interp, double default_value)
{
    geomx gi, go;
    geomx_init(&gi, gin);
    geomx_init(&go, gout);
    tfmx* t = chain_prepare(tf, ntf);
    int linear = chain_is_linear(t, ntf);
    if (interp == 3) { /* sitk.sitkBSpline */
        double* coef = bspline3_coefficients(in, dtype, &gi);
        if (!coef) { free(t); return -1; }
#pragma omp parallel for schedule(static)
        for (int k = 0; k < go.nz; ++k)
            for (int j = 0; j < go.ny; ++j)
                for (int i = 0; i < go.nx; ++i) {
                    double c[3];
                    size_t o = ((size_t)k * go.ny + j) * go.nx + i;
                    out_to_in_cidx(&go, &gi, t, ntf, linear, i, j, k, c);
                    if (inside_buffer(&gi, c)) st(out, dtype, o, interp_bspline3(coef, &gi, c));
                    else st_default(out, dtype, o, default_value);
                }
        free(coef);
        free(t);
        return 0;
    }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < go.nz; ++k)
        for (int j = 0; j < go.ny; ++j)
            for (int i = 0; i < go.nx; ++i) {
                double c[3];
                size_t o = ((size_t)k * go.ny + j) * go.nx + i;
                out_to_in_cidx(&go, &gi, t, ntf, linear, i, j, k, c);
                if (inside_buffer(&gi, c)) {
                    double v = (interp == ORC_INTERP_NN) ? interp_nn_scalar(in, dtype, &gi, c) : interp_linear_scalar(in, dtype, &gi, c);
                    st(out, dtype, o, v);
                } else {
                    st_default(out, dtype, o, default_value);
                }
            }
    free(t);
    return 0;
}

ORC_API int orc_resample_vec3(const double* in, const orc_geom* gin, double* out, const orc_geom* gout,
                              const orc_transform* tf, int ntf, double default_value)
{
    geomx gi, go;
    geomx_init(&gi, gin);
    geomx_init(&go, gout);
    tfmx* t = chain_prepare(tf, ntf);
    int linear = chain_is_linear(t, ntf);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < go.nz; ++k)
        for (int j = 0; j < go.ny; ++j)
            for (int i = 0; i < go.nx; ++i) {
                double c[3];
                size_t o = (((size_t)k * go.ny + j) * go.nx + i) * 3;
                out_to_in_cidx(&go, &gi, t, ntf, linear, i, j, k, c);
                if (inside_buffer(&gi, c)) {
                    if (SEM(SEM_VEC_INTERP)) interp_wsum_vec3(in, &gi, c, out + o);
                    else interp_linear_vec3(in, &gi, c, out + o);
                } else {
                    out[o] = out[o + 1] = out[o + 2] = default_value;
                }
            }
    free(t);
    return 0;
}

/* TransformToDisplacementFieldFilter (reference deformable.py:101-108): D(x) = T(x) - x on the reference grid,
 * non-linear path (per-voxel TransformPoint); SimpleITK instantiates it for VectorFloat64. */
ORC_API int orc_transform_to_dvf(const orc_geom* gout, const orc_transform* tf, int ntf, double* out)
{
    geomx go;
    geomx_init(&go, gout);
    tfmx* t = chain_prepare(tf, ntf);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < go.nz; ++k)
        for (int j = 0; j < go.ny; ++j)
            for (int i = 0; i < go.nx; ++i) {
                double p[3], q[3];
                idx2pt(&go, (double)i, (double)j, (double)k, p);
                q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
                apply_chain(t, ntf, q);
                size_t o = (((size_t)k * go.ny + j) * go.nx + i) * 3;
                out[o] = q[0] - p[0]; out[o + 1] = q[1] - p[1]; out[o + 2] = q[2] - p[2];
            }
    free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* GaussianOperator (itkGaussianOperator.hxx): discrete Gaussian e^{-t} I_n(t), Numerical-Recipes     */
/* polynomial Bessel functions                                                                       */
/* ------------------------------------------------------------------------------------------------ */
static double bessel_i0(double y)
{
    double d = fabs(y), accumulator, m;
    if (d < 3.75) {
        m = y / 3.75; m *= m;
        accumulator = 1.0 + m * (3.5156229 + m * (3.0899424 + m * (1.2067492 + m * (0.2659732 + m * (0.360768e-1 + m * 0.45813e-2)))));
    } else {
        m = 3.75 / d;
        accumulator = (exp(d) / sqrt(d)) * (0.39894228 + m * (0.1328592e-1 + m * (0.225319e-2 + m * (-0.157565e-2 + m * (0.916281e-2 + m * (-0.2057706e-1 + m * (0.2635537e-1 + m * (-0.1647633e-1 + m * 0.392377e-2))))))));
    }
    return accumulator;
}
static double bessel_i1(double y)
{
    double d = fabs(y), accumulator, m;
    if (d < 3.75) {
        m = y / 3.75; m *= m;
        accumulator = d * (0.5 + m * (0.87890594 + m * (0.51498869 + m * (0.15084934 + m * (0.2658733e-1 + m * (0.301532e-2 + m * 0.32411e-3))))));
    } else {
        m = 3.75 / d;
        accumulator = 0.2282967e-1 + m * (-0.2895312e-1 + m * (0.1787654e-1 - m * 0.420059e-2));
        accumulator = 0.39894228 + m * (-0.3988024e-1 + m * (-0.362018e-2 + m * (0.163801e-2 + m * (-0.1031555e-1 + m * accumulator))));
        accumulator *= (exp(d) / sqrt(d));
    }
    return y < 0.0 ? -accumulator : accumulator;
}
static double bessel_in(int n, double y)
{
    const double ACCURACY = 40.0;
    if (y == 0.0) return 0.0;
    double toy = 2.0 / fabs(y), qip = 0.0, accumulator = 0.0, qi = 1.0, qim;
    for (int j = 2 * (n + (int)sqrt(ACCURACY * n)); j > 0; j--) {
        qim = qip + j * toy * qi;
        qip = qi;
        qi = qim;
        if (fabs(qi) > 1.0e10) { accumulator *= 1.0e-10; qi *= 1.0e-10; qip *= 1.0e-10; }
        if (j == n) accumulator = qip;
    }
    accumulator *= bessel_i0(y) / qi;
    return (y < 0.0 && (n & 1)) ? -accumulator : accumulator;
}

/* GaussianOperator::GenerateCoefficients.  Returns the radius r; kernel[0..2r] symmetric, normalised.
 * Terms are added until the running sum reaches 1 - maximumError, a term drops below sum*DBL_EPSILON,
 * or the one-sided length exceeds maximumKernelWidth. */
ORC_API int orc_gaussian_operator(double variance, double max_error, int max_width, double* kernel, int cap)
{
    double* c = (double*)malloc(sizeof(double) * (size_t)(max_width + 8));
    int n = 0;
    const double et = exp(-variance), lim = 1.0 - max_error;
    double sum = 0.0;
    c[n++] = et * bessel_i0(variance); sum += c[0];
    c[n++] = et * bessel_i1(variance); sum += c[1] * 2.0;
    for (int i = 2; sum < lim; ++i) {
        c[n++] = et * bessel_in(i, variance);
        sum += c[i] * 2.0;
        if (c[i] < sum * DBL_EPSILON) break;
        if (n > max_width) break;
    }
    for (int i = 0; i < n; ++i) c[i] /= sum;
    int r = n - 1;
    if (2 * r + 1 > cap) { free(c); return -1; }
    for (int i = 0; i <= r; ++i) { kernel[r + i] = c[i]; kernel[r - i] = c[i]; }
    free(c);
    return r;
}

/* 1-D convolution along `axis` with ZeroFluxNeumann (index clamp) boundary; inner product accumulated
 * in double from offset -r to +r (NeighborhoodInnerProduct order). */
static void conv_axis_f32(const float* in, float* out, int nx, int ny, int nz, int axis, const double* kern, int r)
{
    const size_t strides[3] = { 1, (size_t)nx, (size_t)nx * ny };
    const int dims[3] = { nx, ny, nz };
    const int n = dims[axis];
    const size_t sa = strides[axis];
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                int pos = axis == 0 ? i : (axis == 1 ? j : k);
                size_t base = ((size_t)k * ny + j) * nx + i - (size_t)pos * sa;
                double sum = 0.0;
                for (int t = -r; t <= r; ++t) {
                    int q = pos + t;
                    if (q < 0) q = 0;
                    if (q > n - 1) q = n - 1;
                    sum += kern[t + r] * (double)in[base + (size_t)q * sa];
                }
                out[((size_t)k * ny + j) * nx + i] = (float)sum;
            }
}
static void conv_axis_f64c(const double* in, double* out, int nx, int ny, int nz, int ncomp, int axis, const double* kern, int r)
{
    const size_t strides[3] = { 1, (size_t)nx, (size_t)nx * ny };
    const int dims[3] = { nx, ny, nz };
    const int n = dims[axis];
    const size_t sa = strides[axis];
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                int pos = axis == 0 ? i : (axis == 1 ? j : k);
                size_t base = ((size_t)k * ny + j) * nx + i - (size_t)pos * sa;
                for (int cc = 0; cc < ncomp; ++cc) {
                    double sum = 0.0;
                    for (int t = -r; t <= r; ++t) {
                        int q = pos + t;
                        if (q < 0) q = 0;
                        if (q > n - 1) q = n - 1;
                        sum += kern[t + r] * in[(base + (size_t)q * sa) * ncomp + cc];
                    }
                    out[(((size_t)k * ny + j) * nx + i) * ncomp + cc] = sum;
                }
            }
}

/* DiscreteGaussianImageFilter (reference utils.py:226; fusion.py:168,279): variance in mm^2 converted to
 * voxel^2 per axis (useImageSpacing), passes applied z -> y -> x, intermediates stored as float32. */
ORC_API int orc_discrete_gaussian_f32(const float* in, float* out, const orc_geom* g, const double* variance,
                                      int max_width, double max_error, int use_spacing)
{
    int nx = g->size[0], ny = g->size[1], nz = g->size[2];
    size_t nvox = (size_t)nx * ny * nz;
    float* a = (float*)malloc(nvox * sizeof(float));
    float* b = (float*)malloc(nvox * sizeof(float));
    int cap = 2 * (max_width + 8) + 1;
    double* kern = (double*)malloc(sizeof(double) * (size_t)cap);
    const float* src = in;
    float* dsts[3] = { a, b, out };
    int pass = 0;
    for (pass = 0; pass < 3; ++pass) {
        const int axis = SEM(SEM_DG_ORDER) ? pass : 2 - pass;
        double t = variance[axis];
        if (use_spacing) t = t / (g->spacing[axis] * g->spacing[axis]);
        int r = orc_gaussian_operator(t, max_error, max_width, kern, cap);
        if (r < 0) { free(a); free(b); free(kern); return -1; }
        conv_axis_f32(src, dsts[pass], nx, ny, nz, axis, kern, r);
        src = dsts[pass];
    }
    free(a); free(b); free(kern);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* FastSymmetricForcesDemonsRegistrationFilter (reference deformable.py:244-257,143-149)             */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    double std_dev[3];          /* SetStandardDeviations, voxel units (deformable.py:253-257) */
    double update_std_dev[3];   /* UpdateFieldStandardDeviations (1.0) */
    int32_t smooth_displacement_field;
    int32_t smooth_update_field;
    double max_error;           /* 0.1 */
    int32_t max_kernel_width;   /* 30 */
    int32_t number_of_iterations;
    double max_rms_error;       /* 0.02 */
    double max_update_step_length; /* 0.5 */
    double intensity_difference_threshold; /* 0.001 */
    double denominator_threshold;          /* 1e-9 */
} orc_demons_params;

typedef struct { int32_t elapsed_iterations; int32_t pad; double metric; double rms_change; } orc_demons_stats;

/* PDEDeformableRegistrationFilter::SmoothDisplacementField / SmoothUpdateField: passes x -> y -> z,
 * variance = sd^2 in voxel units, clamp boundary, all double, ping-pong with a temp field. */
static void pde_smooth_field(double* field, double* tmp, int nx, int ny, int nz, const double* sd, double max_error, int max_width)
{
    int cap = 2 * (max_width + 8) + 1;
    double* kern = (double*)malloc(sizeof(double) * (size_t)cap);
    double* src = field; double* dst = tmp;
    for (int axis = 0; axis < 3; ++axis) {
        int r = orc_gaussian_operator(sd[axis] * sd[axis], max_error, max_width, kern, cap);
        conv_axis_f64c(src, dst, nx, ny, nz, 3, axis, kern, r);
        double* sw = src; src = dst; dst = sw;
    }
    /* after 3 passes the result is in `src` == tmp; copy back */
    if (src != field) memcpy(field, src, sizeof(double) * 3 * (size_t)nx * ny * nz);
    free(kern);
}

ORC_API int orc_pde_smooth_field(double* field, const orc_geom* g, const double* sd, double max_error, int max_width)
{
    size_t n = (size_t)g->size[0] * g->size[1] * g->size[2];
    double* tmp = (double*)malloc(sizeof(double) * 3 * n);
    pde_smooth_field(field, tmp, g->size[0], g->size[1], g->size[2], sd, max_error, max_width);
    free(tmp);
    return 0;
}

/* One InitializeIteration + CalculateChange of ESMDemonsRegistrationFunction.
 * W: warped moving image (float32, FLT_MAX where x + D(x) leaves the moving buffer).
 * U: raw update (AoS f64).  Partials are accumulated per z-slice in scan order and merged in z order. */
static void esm_iteration(const float* F, const geomx* gf, const double* dirF, const double* spF,
                          const float* M, const geomx* gm, const double* D, float* W, double* U,
                          const orc_demons_params* p, double* metric, double* rms)
{
    const int nx = gf->nx, ny = gf->ny, nz = gf->nz;
    /* normalizer = mean(spacing^2) * MaximumUpdateStepLength^2, or -1 (unrestricted) */
    double normalizer;
    if (p->max_update_step_length > 0.0) {
        normalizer = 0.0;
        for (int k = 0; k < 3; ++k) normalizer += spF[k] * spF[k];
        normalizer *= p->max_update_step_length * p->max_update_step_length / 3.0;
    } else normalizer = -1.0;

    /* WarpImageFilter, linear interpolation, edge padding = NumericTraits<float>::max() */
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                size_t o = ((size_t)k * ny + j) * nx + i;
                double pt[3], c[3];
                idx2pt(gf, (double)i, (double)j, (double)k, pt);
                pt[0] += D[o * 3 + 0]; pt[1] += D[o * 3 + 1]; pt[2] += D[o * 3 + 2];
                pt2cidx(gm, pt, c);
                W[o] = inside_buffer(gm, c) ? (float)interp_linear_scalar(M, ORC_F32, gm, c) : FLT_MAX;
            }

    double* part = (double*)calloc((size_t)nz * 3, sizeof(double));
    const int dims[3] = { nx, ny, nz };
    const size_t strides[3] = { 1, (size_t)nx, (size_t)nx * ny };
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k) {
        double ssd = 0.0, cnt = 0.0, ssc = 0.0;
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                size_t o = ((size_t)k * ny + j) * nx + i;
                const int idx[3] = { i, j, k };
                double* u = U + o * 3;
                float mv = W[o];
                if (mv == FLT_MAX) { u[0] = u[1] = u[2] = 0.0; continue; }
                const double fixedValue = (double)F[o];
                const double movingValue = (double)mv;
                double g2[3]; /* fixedGradient + warpedMovingGradient, orientation-free */
                for (int dim = 0; dim < 3; ++dim) {
                    double wg;
                    const int n = dims[dim];
                    const size_t s = strides[dim];
                    if (n == 0) { wg = 0.0; }
                    else if (idx[dim] == 0) {
                        if (n < 2) wg = 0.0; /* single-slice dimension: neighbour does not exist */
                        else {
                            float nb = W[o + s];
                            if (nb == FLT_MAX) wg = 0.0;
                            else { wg = (double)nb - movingValue; wg /= spF[dim]; }
                        }
                    } else if (idx[dim] == n - 1) {
                        float nb = W[o - s];
                        if (nb == FLT_MAX) wg = 0.0;
                        else { wg = movingValue - (double)nb; wg /= spF[dim]; }
                    } else {
                        float nb = W[o + s];
                        if (nb == FLT_MAX) {
                            wg = movingValue;
                            float pb = W[o - s];
                            if (pb == FLT_MAX) wg = 0.0;
                            else { wg -= (double)pb; wg /= spF[dim]; }
                        } else {
                            wg = (double)nb;
                            float pb = W[o - s];
                            if (pb == FLT_MAX) { wg -= movingValue; wg /= spF[dim]; }
                            else { wg -= (double)pb; wg *= 0.5 / spF[dim]; }
                        }
                    }
                    /* CentralDifferenceImageFunction::EvaluateAtIndex, UseImageDirection off */
                    double fg;
                    if (idx[dim] < 1 || idx[dim] > n - 2) fg = 0.0;
                    else { fg = (double)F[o + s]; fg -= (double)F[o - s]; fg *= 0.5 / spF[dim]; }
                    g2[dim] = fg + wg;
                }
                /* TransformLocalVectorToPhysicalVector: direction * g */
                double J[3];
                for (int r = 0; r < 3; ++r) {
                    double sum = 0.0;
                    sum += dirF[r * 3 + 0] * g2[0];
                    sum += dirF[r * 3 + 1] * g2[1];
                    sum += dirF[r * 3 + 2] * g2[2];
                    J[r] = sum;
                }
                const double gm2 = J[0] * J[0] + J[1] * J[1] + J[2] * J[2];
                const double speed = fixedValue - movingValue;
                if (fabs(speed) < p->intensity_difference_threshold) { u[0] = u[1] = u[2] = 0.0; }
                else {
                    double denom = (normalizer > 0.0) ? gm2 + (speed * speed) / normalizer : gm2;
                    if (denom < p->denominator_threshold) { u[0] = u[1] = u[2] = 0.0; }
                    else {
                        const double factor = 2.0 * speed / denom;
                        u[0] = factor * J[0]; u[1] = factor * J[1]; u[2] = factor * J[2];
                    }
                }
                ssd += speed * speed;
                cnt += 1.0;
                ssc += u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
            }
        part[k * 3 + 0] = ssd; part[k * 3 + 1] = cnt; part[k * 3 + 2] = ssc;
    }
    double ssd = 0.0, cnt = 0.0, ssc = 0.0;
    for (int k = 0; k < nz; ++k) { ssd += part[k * 3]; cnt += part[k * 3 + 1]; ssc += part[k * 3 + 2]; }
    free(part);
    if (cnt > 0.0) { *metric = ssd / cnt; *rms = sqrt(ssc / cnt); }
}

/* registration_algorithm.Execute(f_image, m_image) (reference deformable.py:149): starts from a zero
 * field on the fixed grid; loop of FiniteDifferenceImageFilter::GenerateData with the
 * DenseFiniteDifferenceImageFilter halt rule.  D_out: AoS f64 on the fixed grid. */
ORC_API int orc_demons_execute(const float* F, const orc_geom* gF, const float* M, const orc_geom* gM,
                               const orc_demons_params* p, double* D_out, orc_demons_stats* stats,
                               double* metric_trace /* optional, number_of_iterations entries */)
{
    geomx gf, gm;
    geomx_init(&gf, gF);
    geomx_init(&gm, gM);
    const int nx = gf.nx, ny = gf.ny, nz = gf.nz;
    const size_t n = (size_t)nx * ny * nz;
    double* D = D_out;
    memset(D, 0, sizeof(double) * 3 * n);
    double* U = (double*)malloc(sizeof(double) * 3 * n);
    double* T = (double*)malloc(sizeof(double) * 3 * n);
    float* W = (float*)malloc(sizeof(float) * n);
    int elapsed = 0;
    double metric = DBL_MAX, rms = 0.0; /* m_Metric init = max, m_RMSChange init = 0 in the function; filter RMSChange = 0 */
    for (;;) {
        /* Halt() */
        if (elapsed >= p->number_of_iterations) break;
        if (elapsed != 0 && p->max_rms_error > rms) break;
        esm_iteration(F, &gf, gF->direction, gF->spacing, M, &gm, D, W, U, p, &metric, &rms);
        /* ApplyUpdate(dt = 1) */
        if (p->smooth_update_field) pde_smooth_field(U, T, nx, ny, nz, p->update_std_dev, p->max_error, p->max_kernel_width);
#pragma omp parallel for schedule(static)
        for (size_t q = 0; q < 3 * n; ++q) D[q] = D[q] + U[q];
        if (p->smooth_displacement_field) pde_smooth_field(D, T, nx, ny, nz, p->std_dev, p->max_error, p->max_kernel_width);
        if (metric_trace) metric_trace[elapsed] = metric;
        ++elapsed;
    }
    stats->elapsed_iterations = elapsed;
    stats->metric = metric;
    stats->rms_change = rms;
    free(U); free(T); free(W);
    return 0;
}

/* A single CalculateChange, exposed so tests can compare the raw update field and the warped image */
ORC_API int orc_demons_force(const float* F, const orc_geom* gF, const float* M, const orc_geom* gM, const double* D,
                             const orc_demons_params* p, float* W, double* U, double* metric, double* rms)
{
    geomx gf, gm;
    geomx_init(&gf, gF);
    geomx_init(&gm, gM);
    *metric = DBL_MAX; *rms = 0.0;
    esm_iteration(F, &gf, gF->direction, gF->spacing, M, &gm, D, W, U, p, metric, rms);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* SmoothingRecursiveGaussianImageFilter (reference deformable.py:158): Deriche zero-order IIR        */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { double N0, N1, N2, N3, D1, D2, D3, D4, M1, M2, M3, M4, BN1, BN2, BN3, BN4, BM1, BM2, BM3, BM4; } deriche_t;

ORC_API void orc_deriche_setup(double sigma, double spacing, double* coeffs20)
{
    const double A1 = 1.3530, B1 = 1.8151, W1 = 0.6681, L1 = -1.3932;
    const double A2 = -0.3531, B2 = 0.0902, W2 = 2.0787, L2 = -1.3732;
    if (spacing < 0.0) spacing = -spacing;
    const double sigmad = sigma / spacing;
    const double Sin1 = sin(W1 / sigmad), Sin2 = sin(W2 / sigmad), Cos1 = cos(W1 / sigmad), Cos2 = cos(W2 / sigmad);
    const double Exp1 = exp(L1 / sigmad), Exp2 = exp(L2 / sigmad);
    deriche_t c;
    /* ComputeDCoefficients */
    c.D4 = Exp1 * Exp1 * Exp2 * Exp2;
    c.D3 = -2 * Cos1 * Exp1 * Exp2 * Exp2;
    c.D3 += -2 * Cos2 * Exp2 * Exp1 * Exp1;
    c.D2 = 4 * Cos2 * Cos1 * Exp1 * Exp2;
    c.D2 += Exp1 * Exp1 + Exp2 * Exp2;
    c.D1 = -2 * (Exp2 * Cos2 + Exp1 * Cos1);
    const double SD = 1.0 + c.D1 + c.D2 + c.D3 + c.D4;
    /* ComputeNCoefficients (zero order) */
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
    const double across_scale_normalization = 1.0; /* NormalizeAcrossScale off */
    c.N0 *= across_scale_normalization / alpha0;
    c.N1 *= across_scale_normalization / alpha0;
    c.N2 *= across_scale_normalization / alpha0;
    c.N3 *= across_scale_normalization / alpha0;
    /* ComputeRemainingCoefficients(symmetric = true) */
    c.M1 = c.N1 - c.D1 * c.N0;
    c.M2 = c.N2 - c.D2 * c.N0;
    c.M3 = c.N3 - c.D3 * c.N0;
    c.M4 = -c.D4 * c.N0;
    const double SN2 = c.N0 + c.N1 + c.N2 + c.N3;
    const double SM = c.M1 + c.M2 + c.M3 + c.M4;
    const double SD2 = 1.0 + c.D1 + c.D2 + c.D3 + c.D4;
    c.BN1 = c.D1 * SN2 / SD2; c.BN2 = c.D2 * SN2 / SD2; c.BN3 = c.D3 * SN2 / SD2; c.BN4 = c.D4 * SN2 / SD2;
    c.BM1 = c.D1 * SM / SD2; c.BM2 = c.D2 * SM / SD2; c.BM3 = c.D3 * SM / SD2; c.BM4 = c.D4 * SM / SD2;
    memcpy(coeffs20, &c, sizeof(c));
}

/* RecursiveSeparableImageFilter::FilterDataArray */
static void deriche_line(const deriche_t* c, const double* data, double* outs, double* scratch, int ln)
{
#define EM(o, a1, b1, a2, b2, a3, b3, a4, b4) o = a1 * b1 + a2 * b2 + a3 * b3 + a4 * b4
#define SM_(o, a1, b1, a2, b2, a3, b3, a4, b4) o -= a1 * b1 + a2 * b2 + a3 * b3 + a4 * b4
    const double outV1 = data[0];
    EM(scratch[0], outV1, c->N0, outV1, c->N1, outV1, c->N2, outV1, c->N3);
    EM(scratch[1], data[1], c->N0, outV1, c->N1, outV1, c->N2, outV1, c->N3);
    EM(scratch[2], data[2], c->N0, data[1], c->N1, outV1, c->N2, outV1, c->N3);
    EM(scratch[3], data[3], c->N0, data[2], c->N1, data[1], c->N2, outV1, c->N3);
    SM_(scratch[0], outV1, c->BN1, outV1, c->BN2, outV1, c->BN3, outV1, c->BN4);
    SM_(scratch[1], scratch[0], c->D1, outV1, c->BN2, outV1, c->BN3, outV1, c->BN4);
    SM_(scratch[2], scratch[1], c->D1, scratch[0], c->D2, outV1, c->BN3, outV1, c->BN4);
    SM_(scratch[3], scratch[2], c->D1, scratch[1], c->D2, scratch[0], c->D3, outV1, c->BN4);
    for (int i = 4; i < ln; ++i) {
        EM(scratch[i], data[i], c->N0, data[i - 1], c->N1, data[i - 2], c->N2, data[i - 3], c->N3);
        SM_(scratch[i], scratch[i - 1], c->D1, scratch[i - 2], c->D2, scratch[i - 3], c->D3, scratch[i - 4], c->D4);
    }
    for (int i = 0; i < ln; ++i) outs[i] = scratch[i];
    const double outV2 = data[ln - 1];
    EM(scratch[ln - 1], outV2, c->M1, outV2, c->M2, outV2, c->M3, outV2, c->M4);
    EM(scratch[ln - 2], data[ln - 1], c->M1, outV2, c->M2, outV2, c->M3, outV2, c->M4);
    EM(scratch[ln - 3], data[ln - 2], c->M1, data[ln - 1], c->M2, outV2, c->M3, outV2, c->M4);
    EM(scratch[ln - 4], data[ln - 3], c->M1, data[ln - 2], c->M2, data[ln - 1], c->M3, outV2, c->M4);
    SM_(scratch[ln - 1], outV2, c->BM1, outV2, c->BM2, outV2, c->BM3, outV2, c->BM4);
    SM_(scratch[ln - 2], scratch[ln - 1], c->D1, outV2, c->BM2, outV2, c->BM3, outV2, c->BM4);
    SM_(scratch[ln - 3], scratch[ln - 2], c->D1, scratch[ln - 1], c->D2, outV2, c->BM3, outV2, c->BM4);
    SM_(scratch[ln - 4], scratch[ln - 3], c->D1, scratch[ln - 2], c->D2, scratch[ln - 1], c->D3, outV2, c->BM4);
    for (int i = ln - 4; i > 0; i--) {
        EM(scratch[i - 1], data[i], c->M1, data[i + 1], c->M2, data[i + 2], c->M3, data[i + 3], c->M4);
        SM_(scratch[i - 1], scratch[i], c->D1, scratch[i + 1], c->D2, scratch[i + 2], c->D3, scratch[i + 3], c->D4);
    }
    for (int i = 0; i < ln; ++i) outs[i] += scratch[i];
#undef EM
#undef SM_
}

/* SmoothingRecursiveGaussianImageFilter on a VectorFloat64 image: axis order z, x, y; sigma is physical
 * (sigma / spacing per axis); every line needs >= 4 samples (ITK throws otherwise -> return -2). */
ORC_API int orc_recursive_gaussian_vec3(double* field, const orc_geom* g, const double* sigma)
{
    const int nx = g->size[0], ny = g->size[1], nz = g->size[2];
    const int dims[3] = { nx, ny, nz };
    const size_t strides[3] = { 1, (size_t)nx, (size_t)nx * ny };
    if (nx < 4 || ny < 4 || nz < 4) return -2;
    const int order_zxy[3] = { 2, 0, 1 }, order_xyz[3] = { 0, 1, 2 };
    const int* order = SEM(SEM_RG_ORDER) ? order_xyz : order_zxy;
    for (int pass = 0; pass < 3; ++pass) {
        const int axis = order[pass];
        deriche_t c;
        orc_deriche_setup(sigma[axis], g->spacing[axis], (double*)&c);
        const int ln = dims[axis];
        const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
        const int n1 = dims[a1], n2 = dims[a2];
#pragma omp parallel
        {
            double* data = (double*)malloc(sizeof(double) * (size_t)ln * 3);
            double* outs = data + ln;
            double* scratch = data + 2 * ln;
#pragma omp for schedule(static) collapse(2)
            for (int q2 = 0; q2 < n2; ++q2)
                for (int q1 = 0; q1 < n1; ++q1) {
                    size_t base = (size_t)q1 * strides[a1] + (size_t)q2 * strides[a2];
                    for (int cc = 0; cc < 3; ++cc) {
                        for (int i = 0; i < ln; ++i) data[i] = field[(base + (size_t)i * strides[axis]) * 3 + cc];
                        deriche_line(&c, data, outs, scratch, ln);
                        for (int i = 0; i < ln; ++i) field[(base + (size_t)i * strides[axis]) * 3 + cc] = outs[i];
                    }
                }
            free(data);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Fusion tail (reference fusion.py:205-292)                                                         */
/* ------------------------------------------------------------------------------------------------ */
/* combine_labels (fusion.py:253-288) for one structure: labels[a] (u8), weights[a] (f32), N atlases.
 * All image arithmetic is per-pixel float32 as SimpleITK does it (a+b, a*b, a/b on Float32 images). */
ORC_API int orc_combine_labels_f32(const uint8_t* const* labels, const float* const* weights, int n_atlas,
                                   const orc_geom* g, double smooth_variance, double threshold, float* out)
{
    const size_t n = (size_t)g->size[0] * g->size[1] * g->size[2];
    float* comb = (float*)malloc(n * sizeof(float));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float wsum = weights[0][i];
        for (int a = 1; a < n_atlas; ++a) wsum = wsum + weights[a][i];
        /* sitk.Mask(wsum, wsum == 0, maskingValue=1, outsideValue=1): where (wsum==0) == 1 -> 1 */
        if (wsum == 0.0f) wsum = 1.0f;
        float acc = weights[0][i] * (float)labels[0][i];
        for (int a = 1; a < n_atlas; ++a) acc = acc + weights[a][i] * (float)labels[a][i];
        comb[i] = acc / wsum;
    }
    double var[3] = { smooth_variance, smooth_variance, smooth_variance };
    float* sm = (float*)malloc(n * sizeof(float));
    orc_discrete_gaussian_f32(comb, sm, g, var, 32, 0.01, 1);
    /* RescaleIntensity(0,1) then Threshold(lower=threshold, upper=1, outside=0) */
    float mn = sm[0], mx = sm[0];
    for (size_t i = 1; i < n; ++i) { if (sm[i] < mn) mn = sm[i]; if (sm[i] > mx) mx = sm[i]; }
    double scale, shift;
    if (fabsf(mx - mn) > FLT_EPSILON) scale = (1.0 - 0.0) / ((double)mx - (double)mn);
    else if ((double)mx != 0.0) scale = (1.0 - 0.0) / (double)mx;
    else scale = 0.0;
    shift = 0.0 - (double)mn * scale;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        double v = (double)sm[i] * scale + shift;
        float r = (float)v;
        r = (r > 1.0f) ? 1.0f : r;
        r = (r < 0.0f) ? 0.0f : r;
        /* ThresholdImageFilter compares in the pixel type: lower = static_cast<float>(threshold) */
        if (threshold != 0.0) { if (!(r >= (float)threshold && r <= 1.0f)) r = 0.0f; }
        out[i] = r;
    }
    free(comb); free(sm);
    return 0;
}

/* sitk.STAPLE(binary_labels) (itkSTAPLEImageFilter.hxx; reference fusion.py:217-223): binary EM,
 * foreground 1, confidence weight 1, unlimited iterations, serial double accumulation in scan order.
 * Output W (f64).  Returns elapsed iterations. */
ORC_API int orc_staple(const uint8_t* const* D, int n_raters, size_t n, double confidence_weight, unsigned max_iter,
                       double* W, double* p_out, double* q_out)
{
    double* p = (double*)malloc(sizeof(double) * n_raters * 4);
    double* q = p + n_raters; double* lp = q + n_raters; double* lq = lp + n_raters;
    for (int i = 0; i < n_raters; ++i) { lp[i] = -10.0; lq[i] = -10.0; }
    for (size_t v = 0; v < n; ++v) W[v] = 0.0;
    for (int i = 0; i < n_raters; ++i)
        for (size_t v = 0; v < n; ++v) if (D[i][v] == 1) W[v] = W[v] + 1.0;
    double g = 0.0, N = 0.0;
    for (size_t v = 0; v < n; ++v) { W[v] = W[v] / (double)n_raters; g += W[v]; N = N + 1.0; }
    g = (g / N) * confidence_weight;
    unsigned iter;
    for (iter = 0; iter < max_iter; ++iter) {
        for (int i = 0; i < n_raters; ++i) {
            double p_num = 0.0, p_denom = 0.0, q_num = 0.0, q_denom = 0.0;
            const uint8_t* d = D[i];
            for (size_t v = 0; v < n; ++v) {
                if (d[v] == 1) p_num += W[v]; else q_num += (1.0 - W[v]);
                p_denom += W[v];
                q_denom += (1.0 - W[v]);
            }
            p[i] = p_num / p_denom;
            q[i] = q_num / q_denom;
        }
#pragma omp parallel for schedule(static)
        for (size_t v = 0; v < n; ++v) {
            double alpha1 = 1.0, beta1 = 1.0;
            for (int i = 0; i < n_raters; ++i) {
                if (D[i][v] == 1) { alpha1 = alpha1 * p[i]; beta1 = beta1 * (1.0 - q[i]); }
                else { alpha1 = alpha1 * (1.0 - p[i]); beta1 = beta1 * q[i]; }
            }
            W[v] = g * alpha1 / (g * alpha1 + (1.0 - g) * beta1);
        }
        int flag = 0;
        if (iter != 0) {
            flag = 1;
            for (int i = 0; i < n_raters; ++i) {
                if (((p[i] - lp[i]) * (p[i] - lp[i])) > 1.0e-14) { flag = 0; break; }
                if (((q[i] - lq[i]) * (q[i] - lq[i])) > 1.0e-14) { flag = 0; break; }
            }
        }
        for (int i = 0; i < n_raters; ++i) { lp[i] = p[i]; lq[i] = q[i]; }
        if (flag) break;
    }
    if (p_out) memcpy(p_out, p, sizeof(double) * n_raters);
    if (q_out) memcpy(q_out, q, sizeof(double) * n_raters);
    free(p);
    return (int)iter;
}

/* RescaleIntensity(img, 0, 1) + Threshold(lower, upper=1, outside=0) on a Float64 image
 * (reference fusion.py:226-232) */
ORC_API int orc_rescale_threshold_f64(double* img, size_t n, double threshold)
{
    double mn = img[0], mx = img[0];
    for (size_t i = 1; i < n; ++i) { if (img[i] < mn) mn = img[i]; if (img[i] > mx) mx = img[i]; }
    double scale, shift;
    if (fabs(mx - mn) > DBL_EPSILON) scale = 1.0 / (mx - mn);
    else if (mx != 0.0) scale = 1.0 / mx;
    else scale = 0.0;
    shift = 0.0 - mn * scale;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        double r = img[i] * scale + shift;
        r = (r > 1.0) ? 1.0 : r;
        r = (r < 0.0) ? 0.0 : r;
        if (threshold != 0.0) { if (!(r >= threshold && r <= 1.0)) r = 0.0; }
        img[i] = r;
    }
    return 0;
}

/* ---- binary post-processing (reference fusion.py:295-328 process_probability_image) ------------------------- */
/* Face-connected flood fill with an explicit stack: every voxel of `cls` reachable from `seed` gets `label`. */
static void flood6(const uint8_t* in, int cls, int32_t* lab, int32_t label, int nx, int ny, int nz, size_t seed, size_t* stack, size_t* count)
{
    size_t top = 0, c = 0;
    const size_t pz = (size_t)nx * ny;
    stack[top++] = seed;
    lab[seed] = label;
    while (top) {
        const size_t i = stack[--top];
        ++c;
        const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / pz);
#define ORC_TRY(cond, j) if (cond) { const size_t q = (j); if (lab[q] == 0 && ((in[q] != 0) == cls)) { lab[q] = label; stack[top++] = q; } }
        ORC_TRY(x > 0, i - 1)
        ORC_TRY(x < nx - 1, i + 1)
        ORC_TRY(y > 0, i - nx)
        ORC_TRY(y < ny - 1, i + nx)
        ORC_TRY(z > 0, i - pz)
        ORC_TRY(z < nz - 1, i + pz)
#undef ORC_TRY
    }
    if (count) *count = c;
}

/* itk::BinaryFillholeImageFilter, FullyConnected = false, ForegroundValue = 1 (sitk.BinaryFillhole, fusion.py:311):
 * the image is padded with background, the background is labelled (face connectivity) and only the object on
 * the border is kept as background -- i.e. background voxels not connected to the image border become 1. */
ORC_API int orc_binary_fillhole(const uint8_t* in, int nx, int ny, int nz, uint8_t* out)
{
    const size_t n = (size_t)nx * ny * nz, pz = (size_t)nx * ny;
    int32_t* lab = (int32_t*)calloc(n, sizeof(int32_t));
    size_t* stack = (size_t*)malloc(n * sizeof(size_t));
    if (!lab || !stack) { free(lab); free(stack); return -1; }
    for (size_t i = 0; i < n; ++i) {
        const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / pz);
        const int border = x == 0 || x == nx - 1 || y == 0 || y == ny - 1 || z == 0 || z == nz - 1;
        if (border && in[i] == 0 && lab[i] == 0) flood6(in, 0, lab, 1, nx, ny, nz, i, stack, NULL);
    }
    for (size_t i = 0; i < n; ++i) out[i] = (in[i] != 0 || lab[i] == 0) ? 1 : 0;
    free(lab);
    free(stack);
    return 0;
}

/* sitk.ConnectedComponent (FullyConnected = false; objects numbered 1.. in raster order of their first voxel)
 * -> LabelShapeStatistics.GetNumberOfPixels -> np.argmax (first maximum) -> labels == k -> UInt8
 * (fusion.py:314-328).  labels_out (optional, int32) receives the ConnectedComponent image.
 * Returns the number of objects; with none, out is all zeros. */
ORC_API int orc_largest_component(const uint8_t* in, int nx, int ny, int nz, uint8_t* out, int32_t* labels_out, int64_t* largest_voxels)
{
    const size_t n = (size_t)nx * ny * nz;
    int32_t* lab = (int32_t*)calloc(n, sizeof(int32_t));
    size_t* stack = (size_t*)malloc(n * sizeof(size_t));
    if (!lab || !stack) { free(lab); free(stack); return -1; }
    int32_t nlab = 0, best = 0;
    size_t best_count = 0;
    for (size_t i = 0; i < n; ++i) {
        if (in[i] != 0 && lab[i] == 0) {
            size_t c = 0;
            flood6(in, 1, lab, ++nlab, nx, ny, nz, i, stack, &c);
            if (c > best_count) { best_count = c; best = nlab; }
        }
    }
    for (size_t i = 0; i < n; ++i) out[i] = (best != 0 && lab[i] == best) ? 1 : 0;
    if (labels_out) memcpy(labels_out, lab, n * sizeof(int32_t));
    if (largest_voxels) *largest_voxels = (int64_t)best_count;
    free(lab);
    free(stack);
    return (int)nlab;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------ */
/* Distance maps, contours, binary morphology (rows after the hot path: SURVEY 8f-3 / 8f-4)             */
/* ------------------------------------------------------------------------------------------------ */
/* itk::SignedMaurerDistanceMapImageFilter as SimpleITK instantiates it (output Float32, background value 0):
 * sitk.SignedMaurerDistanceMap(mask, insideIsPositive, squaredDistance, useImageSpacing)
 * (registration/utils.py:289-294; label/projection.py:22-31,80-82).  [ITK-recall]
 *  GenerateData: BinaryThreshold (background -> max, object -> 0), BinaryContour with FullyConnected = true (object voxels
 *  with a background voxel among their 26 neighbours stay 0, every other voxel becomes max), then one Voronoi pass per
 *  dimension (x, y, z); all of Voronoi()'s arithmetic -- line coordinate i * spacing, squared distances, Remove() -- is in the
 *  output pixel type, i.e. single precision.  Each pass stores +-d with the sign of the final result and reads |d| back.
 *  The last pass is followed by sqrt(|d|) unless squared distances were asked for. */
static int maurer_remove_f(float d1, float d2, float df, float x1, float x2, float xf)
{
    const float a = x2 - x1, b = xf - x2, c = xf - x1;
    const float value = c * fabsf(d2) - b * fabsf(d1) - a * fabsf(df) - a * b * c;
    return value > 0.0f;
}
static void maurer_line(float* out, const uint8_t* mask, size_t base, size_t stride, int nd, float sp, int inside_pos, float* g, float* h)
{
    int l = -1;
    for (int i = 0; i < nd; ++i) {
        const float di = out[base + (size_t)i * stride];
        const float iw = (float)i * sp;
        if (di != FLT_MAX) {
            if (l < 1) {
                ++l; g[l] = di; h[l] = iw;
            } else {
                while (l >= 1 && maurer_remove_f(g[l - 1], g[l], di, h[l - 1], h[l], iw)) --l;
                ++l; g[l] = di; h[l] = iw;
            }
        }
    }
    if (l == -1) return;
    const int ns = l;
    l = 0;
    for (int i = 0; i < nd; ++i) {
        const float iw = (float)i * sp;
        float d1 = fabsf(g[l]) + (h[l] - iw) * (h[l] - iw);
        while (l < ns) {
            const float d2 = fabsf(g[l + 1]) + (h[l + 1] - iw) * (h[l + 1] - iw);
            if (d1 <= d2) break;
            ++l;
            d1 = d2;
        }
        const size_t q = base + (size_t)i * stride;
        if (mask[q] != 0) out[q] = inside_pos ? d1 : -d1;
        else out[q] = inside_pos ? -d1 : d1;
    }
}
ORC_API int orc_signed_maurer(const uint8_t* mask, int nx, int ny, int nz, const double* spacing, int inside_pos, int squared, int use_spacing,
                              float* out)
{
    const size_t n = (size_t)nx * ny * nz, pz = (size_t)nx * ny;
    int nmax = nx > ny ? nx : ny;
    if (nz > nmax) nmax = nz;
    /* threshold + contour */
    for (size_t q = 0; q < n; ++q) {
        float v = FLT_MAX;
        if (mask[q]) {
            const int x = (int)(q % nx), y = (int)((q / nx) % ny), z = (int)(q / pz);
            int border = 0;
            for (int dz = -1; dz <= 1 && !border; ++dz)
                for (int dy = -1; dy <= 1 && !border; ++dy)
                    for (int dx = -1; dx <= 1 && !border; ++dx) {
                        const int xx = x + dx, yy = y + dy, zz = z + dz;
                        if (xx < 0 || yy < 0 || zz < 0 || xx >= nx || yy >= ny || zz >= nz) continue;
                        if (!mask[(size_t)zz * pz + (size_t)yy * nx + xx]) border = 1;
                    }
            if (border) v = 0.0f;
        }
        out[q] = v;
    }
    float* g = (float*)malloc(sizeof(float) * (size_t)nmax);
    float* h = (float*)malloc(sizeof(float) * (size_t)nmax);
    if (!g || !h) { free(g); free(h); return -1; }
    const float sx = use_spacing ? (float)spacing[0] : 1.0f, sy = use_spacing ? (float)spacing[1] : 1.0f, sz = use_spacing ? (float)spacing[2] : 1.0f;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y) maurer_line(out, mask, (size_t)z * pz + (size_t)y * nx, 1, nx, sx, inside_pos, g, h);
    for (int z = 0; z < nz; ++z)
        for (int x = 0; x < nx; ++x) maurer_line(out, mask, (size_t)z * pz + x, (size_t)nx, ny, sy, inside_pos, g, h);
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) maurer_line(out, mask, (size_t)y * nx + x, pz, nz, sz, inside_pos, g, h);
    if (!squared) {
        for (size_t q = 0; q < n; ++q) {
            const float r = sqrtf(fabsf(out[q]));
            if (mask[q] != 0) out[q] = inside_pos ? r : -r;
            else out[q] = inside_pos ? -r : r;
        }
    }
    free(g);
    free(h);
    return 0;
}

/* itk::LabelContourImageFilter (sitk.LabelContour, label/projection.py:33,85): background 0; a labelled voxel is kept when a
 * neighbour inside the image -- face neighbours, or all 26 when fully connected -- has another value.  [ITK-recall] */
ORC_API void orc_label_contour(const uint8_t* in, int nx, int ny, int nz, int fully, uint8_t* out)
{
    const size_t pz = (size_t)nx * ny;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t q = (size_t)z * pz + (size_t)y * nx + x;
                const uint8_t v = in[q];
                int edge = 0;
                if (v) {
                    for (int dz = -1; dz <= 1; ++dz)
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int manhattan = abs(dx) + abs(dy) + abs(dz);
                                if (manhattan == 0 || (!fully && manhattan != 1)) continue;
                                const int xx = x + dx, yy = y + dy, zz = z + dz;
                                if (xx < 0 || yy < 0 || zz < 0 || xx >= nx || yy >= ny || zz >= nz) continue;
                                if (in[(size_t)zz * pz + (size_t)yy * nx + xx] != v) edge = 1;
                            }
                }
                out[q] = edge ? v : 0;
            }
}

/* itk::BinaryDilateImageFilter / itk::BinaryErodeImageFilter (sitk.BinaryDilate / BinaryErode: registration/utils.py:331,
 * generation/dvf.py:269-287), foreground 1, background 0, flat structuring element given as offsets.  Dilation: the element is
 * painted around every foreground voxel; erosion: a foreground voxel survives when the element around it holds only
 * foreground (outside the image counts as foreground with boundaryToForeground, SimpleITK's default for the erosion). */
ORC_API void orc_binary_morph(const uint8_t* in, int nx, int ny, int nz, const int32_t* offs, int noffs, int dilate, int boundary_fg, uint8_t* out)
{
    const size_t pz = (size_t)nx * ny, n = pz * nz;
    memcpy(out, in, n);
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t q = (size_t)z * pz + (size_t)y * nx + x;
                if (dilate) {
                    if (in[q] != 1) continue;
                    for (int t = 0; t < noffs; ++t) {
                        const int xx = x + offs[3 * t], yy = y + offs[3 * t + 1], zz = z + offs[3 * t + 2];
                        if (xx < 0 || yy < 0 || zz < 0 || xx >= nx || yy >= ny || zz >= nz) continue;
                        out[(size_t)zz * pz + (size_t)yy * nx + xx] = 1;
                    }
                } else if (in[q] == 1) {
                    int keep = 1;
                    for (int t = 0; t < noffs && keep; ++t) {
                        const int xx = x + offs[3 * t], yy = y + offs[3 * t + 1], zz = z + offs[3 * t + 2];
                        if (xx < 0 || yy < 0 || zz < 0 || xx >= nx || yy >= ny || zz >= nz) keep = boundary_fg != 0;
                        else keep = in[(size_t)zz * pz + (size_t)yy * nx + xx] == 1;
                    }
                    if (!keep) out[q] = 0;
                }
            }
    if (dilate && boundary_fg) {
        /* foreground outside the image: a voxel within the element's reach of the outside is painted too */
        for (int z = 0; z < nz; ++z)
            for (int y = 0; y < ny; ++y)
                for (int x = 0; x < nx; ++x) {
                    const size_t q = (size_t)z * pz + (size_t)y * nx + x;
                    if (out[q] == 1) continue;
                    for (int t = 0; t < noffs; ++t) {
                        const int xx = x - offs[3 * t], yy = y - offs[3 * t + 1], zz = z - offs[3 * t + 2];
                        if (xx < 0 || yy < 0 || zz < 0 || xx >= nx || yy >= ny || zz >= nz) { out[q] = 1; break; }
                    }
                }
    }
}
