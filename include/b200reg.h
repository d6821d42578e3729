/*
 * b200reg.h -- C ABI of libb200reg.so: the B200-native (sm_100a) Demons registration, resampling and
 * label-fusion engine that replaces the SimpleITK/ITK filters on platipy's hot path.
 *
 * The reference has no C/FFI boundary on this path: its boundary is the Python call signature
 * (platipy/imaging/registration/deformable.py:190, utils.py:148,195; platipy/imaging/label/fusion.py:205,239).
 * Each entry point below replaces the ITK filter(s) behind one reference call site, cited as
 * file:line relative to the reference repository root.  `platipy_b200/_abi.py` is the ctypes binding;
 * INTEGRATION.md shows the stub a platipy maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer named d_* is DEVICE memory, h_* is HOST memory.
 *  - scalar volumes are C-order [z][y][x] (x fastest) == numpy view of a SimpleITK image.
 *  - displacement fields on the device are SoA: three contiguous f64 planes [3][z][y][x] holding
 *    (dx, dy, dz) in physical mm.  b200reg_aos_to_soa / b200reg_soa_to_aos convert from/to the AoS
 *    [z][y][x][3] layout of a sitkVectorFloat64 image.
 *  - all work is enqueued on the context's CUDA stream; calls return without synchronising unless
 *    stated.  Every call returns a b200reg_status; b200reg_last_error() gives the thread-local text.
 *  - no exceptions cross the ABI; the caller owns every buffer it passes.
 */
#ifndef B200REG_H
#define B200REG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200REG_ABI_VERSION 1
#define B200REG_MAX_TRANSFORMS 4
#define B200REG_MAX_LEVELS 8
#define B200REG_MAX_BATCH 64

typedef struct b200reg_ctx b200reg_ctx;

typedef enum {
    B200REG_OK = 0,
    B200REG_ERR_CUDA = 1,        /* CUDA runtime error (text in b200reg_last_error) */
    B200REG_ERR_ARG = 2,         /* invalid argument                              -> ValueError   */
    B200REG_ERR_UNSUPPORTED = 3, /* e.g. B-spline interpolation                   -> NotImplementedError */
    B200REG_ERR_RUNTIME = 4      /* ITK-style runtime failure (e.g. line < 4 px)  -> RuntimeError */
} b200reg_status;

/* SimpleITK pixel IDs (Linux builds) */
typedef enum {
    B200REG_I8 = 0, B200REG_U8 = 1, B200REG_I16 = 2, B200REG_U16 = 3, B200REG_I32 = 4, B200REG_U32 = 5,
    B200REG_I64 = 6, B200REG_U64 = 7, B200REG_F32 = 8, B200REG_F64 = 9
} b200reg_dtype;

/* sitkNearestNeighbor / sitkLinear (deformable.py:221-224) */
typedef enum { B200REG_INTERP_NN = 1, B200REG_INTERP_LINEAR = 2, B200REG_INTERP_BSPLINE = 3 } b200reg_interp; /* sitk enum values */

/* itk::ImageBase geometry: size (x,y,z), spacing, origin, direction cosines (row-major 3x3) */
typedef struct {
    int32_t size[3];
    double spacing[3];
    double origin[3];
    double direction[9];
} b200reg_geom;

typedef enum { B200REG_TFM_AFFINE = 0, B200REG_TFM_DVF = 1 } b200reg_tfm_kind;

/* One element of a transform chain, in APPLICATION order (first applied first; the reverse of
 * sitk.CompositeTransform's add order).  AFFINE: p' = matrix * p + offset
 * (itk::MatrixOffsetTransformBase).  DVF: p' = p + D(p), identity outside the field buffer
 * (itk::DisplacementFieldTransform; deformable.py:139,296). */
typedef struct {
    int32_t kind;
    int32_t pad;
    double matrix[9];
    double offset[3];
    const double* d_dvf;      /* device, SoA [3][z][y][x] f64 */
    b200reg_geom dvf_geom;
} b200reg_transform;

/* sitk.FastSymmetricForcesDemonsRegistrationFilter parameters as platipy leaves them
 * (deformable.py:244-257; everything it does not set keeps the SimpleITK default). */
typedef struct {
    double std_dev[3];            /* SetStandardDeviations, voxel units (deformable.py:253-257) */
    double update_std_dev[3];     /* UpdateFieldStandardDeviations, default 1.0 */
    int32_t smooth_displacement_field; /* deformable.py:250 */
    int32_t smooth_update_field;       /* deformable.py:249 */
    double max_error;             /* 0.1 */
    int32_t max_kernel_width;     /* 30 */
    int32_t number_of_iterations; /* deformable.py:143-144 */
    double max_rms_error;         /* 0.02 */
    double max_update_step_length;        /* 0.5 */
    double intensity_difference_threshold; /* 0.001 */
    double denominator_threshold;          /* 1e-9 */
    int32_t field_precision;      /* 0: Float64 fields like ITK's (parity mode, the default); 1: fast mode -- float32 fields, float32 FMA
                                   * smoothing, NOT bit-comparable with the reference; levels it cannot run stay in parity mode */
    int32_t reserved;
} b200reg_demons_params;

typedef struct {
    int32_t elapsed_iterations;   /* GetElapsedIterations() (utils.py:41) */
    int32_t voxels_lo;            /* voxels of the level grid (low 31 bits are enough for one GPU) */
    double metric;                /* GetMetric(): SSD / N of the last iteration */
    double rms_change;            /* GetRMSChange() */
    double gpu_ms;                /* device time of the level's Demons loop (CUDA events) */
} b200reg_demons_stats;

/* multiscale_demons configuration (deformable.py:31-42) */
typedef struct {
    int32_t n_levels;
    int32_t isotropic_resample;               /* resolution = voxel size in mm instead of shrink factor */
    double resolution_staging[B200REG_MAX_LEVELS];
    double smoothing_sigmas[B200REG_MAX_LEVELS];   /* mm; 0 = no smoothing (utils.py:216) */
    int32_t iteration_staging[B200REG_MAX_LEVELS];
    int32_t interp_order;                     /* b200reg_interp */
    b200reg_demons_params demons;             /* number_of_iterations is overwritten per level */
} b200reg_multires_config;

/* ---- context ------------------------------------------------------------------------------------ */
int b200reg_abi_version(void);
const char* b200reg_last_error(void);
/* device: CUDA ordinal; stream: a cudaStream_t (NULL = a new non-blocking stream owned by the ctx) */
int b200reg_create(int device, void* stream, b200reg_ctx** out);
int b200reg_destroy(b200reg_ctx* ctx);
int b200reg_set_stream(b200reg_ctx* ctx, void* stream);
int b200reg_synchronize(b200reg_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t b200reg_launch_count(b200reg_ctx* ctx);

/* Named semantic switches (process-wide).  The arithmetic this library replaces lives in ITK 5.3, which could not be re-read when
 * it was written; every recalled behaviour of medium confidence that is cheap to state both ways is a switch
 * whose default is the recalled behaviour.  The CPU oracle has the same switches under the same names (orc_set_semantic), and
 * tools/validate_against_sitk.py reports the setting that matches a real SimpleITK.  Names / values:
 *   discrete_gaussian_axis_order     0 = z, y, x    1 = x, y, z      (DiscreteGaussianImageFilter: utils.py:226, fusion.py:168,279)
 *   recursive_gaussian_axis_order    0 = z, x, y    1 = x, y, z      (SmoothingRecursiveGaussian: deformable.py:158)
 *   resample_linear_scanline         1 = scan-line continuous index for linear transforms, 0 = per voxel (utils.py:176-190)
 *   dvf_transform_interpolation      0 = weighted sum of the 8 neighbours, 1 = nested lerps (DisplacementFieldTransform)
 *   vector_resample_interpolation    0 = nested lerps, 1 = weighted sum (sitk.Resample of a vector image: deformable.py:130,137,154,185)
 *   binary_threshold_in_pixel_type   0 = bounds compared as real numbers, 1 = bounds cast to the pixel type (fusion.py:217-220)
 * b200reg_set_semantic returns B200REG_ERR_ARG for an unknown name; b200reg_get_semantic returns -1. */
int b200reg_set_semantic(const char* name, int value);
int b200reg_get_semantic(const char* name);

/* ---- memory helpers (hosts without their own CUDA allocator) --------------------------------------- */
int b200reg_malloc(b200reg_ctx* ctx, size_t bytes, void** d_ptr);
int b200reg_free(b200reg_ctx* ctx, void* d_ptr);
int b200reg_malloc_host(size_t bytes, void** h_ptr);   /* pinned */
int b200reg_free_host(void* h_ptr);
int b200reg_memcpy_h2d(b200reg_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int b200reg_memcpy_d2h(b200reg_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
int b200reg_memset(b200reg_ctx* ctx, void* d_ptr, int value, size_t bytes);

/* ---- layout / dtype helpers ------------------------------------------------------------------------- */
int b200reg_aos_to_soa(b200reg_ctx* ctx, const double* d_aos, double* d_soa, size_t nvox);
int b200reg_soa_to_aos(b200reg_ctx* ctx, const double* d_soa, double* d_aos, size_t nvox);
/* sitk.Cast (deformable.py:239,241,304; utils.py:190): C static_cast, float->int truncates toward zero */
int b200reg_cast(b200reg_ctx* ctx, const void* d_in, int in_dtype, void* d_out, int out_dtype, size_t n);
/* min / max of a scalar volume as double (deformable.py:290 CT-like test; RescaleIntensity). Synchronises. */
int b200reg_minmax(b200reg_ctx* ctx, const void* d_in, int dtype, size_t n, double* h_min, double* h_max);

/* ---- N1: itk::DiscreteGaussianImageFilter (utils.py:226; fusion.py:168,279) ------------------------- */
int b200reg_discrete_gaussian_f32(b200reg_ctx* ctx, const float* d_in, float* d_out, const b200reg_geom* geom,
                                  const double variance[3], int max_kernel_width, double max_error, int use_image_spacing);
/* smooth_and_resample (utils.py:195-267) as one call for a Float32 image: DiscreteGaussian(variance, maximumKernelWidth, useImageSpacing)
 * (utils.py:216-226) then Resample onto out_geom with an identity transform and default pixel value 0 (utils.py:257-267).  With
 * allow_restricted = 1 a level that shrinks enough for it to pay, read through a linear interpolator, is blurred only at the planes /
 * rows / columns the resampler reads; 2 takes that form whenever it is possible, 0 never (all three give the same bits). */
int b200reg_smooth_and_resample_f32(b200reg_ctx* ctx, const float* d_in, const b200reg_geom* in_geom, const double variance[3],
                                    int max_kernel_width, const b200reg_geom* out_geom, int interp, float* d_out, int allow_restricted);
/* GaussianOperator coefficients (host-side; used by tests): returns radius, fills kernel[0..2r] */
int b200reg_gaussian_operator(double variance, double max_error, int max_kernel_width, double* h_kernel, int capacity);

/* Host-only diagnostic (no context, no device work): 1 when sitk.Resample from in_geom onto out_geom through an identity transform
 * provably reads voxel i for output voxel i -- every continuous index the resampler would compute (index -> point of the output grid ->
 * continuous index of the input grid; with allow_scanline != 0 in the scan-line form of linear transforms when that semantic switch is
 * on) equals the integer index exactly.  The library then copies instead of interpolating (utils.py:257-267 with shrink factor 1,
 * deformable.py:185) and reads a displacement field that lives on the output grid per index (allow_scanline = 0).  0 is always safe. */
int b200reg_identity_resample_is_exact(const b200reg_geom* in_geom, const b200reg_geom* out_geom, int allow_scanline);

/* ---- N2/N5/N9: itk::ResampleImageFilter, scalar (utils.py:176-190,257-267; deformable.py:140,281-301) */
int b200reg_resample(b200reg_ctx* ctx, const void* d_in, int dtype, const b200reg_geom* in_geom, void* d_out,
                     const b200reg_geom* out_geom, const b200reg_transform* chain, int n_chain, int interp,
                     double default_value);
/* Batched form: n images on the same input grid through the same chain onto the same output grid
 * (multiatlas/run.py:331-345: one CT + S label masks).  The chain (and the DVF behind it) is evaluated
 * once per output voxel. */
int b200reg_resample_batch(b200reg_ctx* ctx, int n, const void* const* d_in, const int* dtypes, const b200reg_geom* in_geom,
                           void* const* d_out, const b200reg_geom* out_geom, const b200reg_transform* chain, int n_chain,
                           const int* interps, const double* default_values);
/* ---- N3/N7: itk::ResampleImageFilter on a VectorFloat64 image (deformable.py:130,137,154,185) -------- */
int b200reg_resample_vec3(b200reg_ctx* ctx, const double* d_in_soa, const b200reg_geom* in_geom, double* d_out_soa,
                          const b200reg_geom* out_geom, const b200reg_transform* chain, int n_chain, double default_value);
/* dvf_total + sitk.Resample(dvf_iter, tfm_total) (deformable.py:154), d_total updated in place.
 * d_scratch_soa: 3 planes on the same grid. */
int b200reg_compose_dvf(b200reg_ctx* ctx, double* d_total_soa, const double* d_iter_soa, const b200reg_geom* geom,
                        double* d_scratch_soa);

/* ---- N11: sitk.TransformToDisplacementField (deformable.py:101-108): D(x) = T(x) - x on the given grid -------- */
int b200reg_transform_to_dvf(b200reg_ctx* ctx, const b200reg_geom* out_geom, const b200reg_transform* chain, int n_chain,
                             double* d_out_soa);

/* ---- N6: sitk.FastSymmetricForcesDemonsRegistrationFilter.Execute (deformable.py:149) ---------------- */
/* Starts from a zero field on the fixed grid; d_out_soa receives the field.  h_stats is filled after an
 * internal stream synchronisation. */
int b200reg_demons_execute(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                           const b200reg_geom* moving_geom, const b200reg_demons_params* params, double* d_out_soa,
                           b200reg_demons_stats* h_stats);
/* IterationEvent data of the most recent b200reg_demons_execute (level 0) / b200reg_multiscale_demons call on this context
 * (deformable.py:260-264, utils.py:37-41 print GetElapsedIterations() / GetMetric() at every iteration): h_metric_rms receives
 * (metric, RMS change) pairs, iteration 1 first, at most capacity_iterations of them.  Returns the number of iterations recorded
 * for that level (0 for an unknown level).  Host data, no synchronisation. */
int b200reg_demons_trace(b200reg_ctx* ctx, int level, double* h_metric_rms, int capacity_iterations);
/* One InitializeIteration + CalculateChange (warp + ESM force), for unit parity tests: d_w (f32 warped
 * moving, FLT_MAX outside), d_u_soa (raw update); h_metric / h_rms after synchronisation. */
int b200reg_demons_force(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                         const b200reg_geom* moving_geom, const double* d_field_soa, const b200reg_demons_params* params,
                         float* d_w, double* d_u_soa, double* h_metric, double* h_rms);
/* PDEDeformableRegistrationFilter::SmoothDisplacementField on its own (x -> y -> z, voxel-unit sigmas) */
int b200reg_pde_smooth_field(b200reg_ctx* ctx, double* d_field_soa, const b200reg_geom* geom, const double std_dev[3],
                             double max_error, int max_kernel_width);

/* ---- N8: sitk.SmoothingRecursiveGaussian on a VectorFloat64 image (deformable.py:158) ----------------- */
int b200reg_recursive_gaussian_vec3(b200reg_ctx* ctx, double* d_field_soa, const b200reg_geom* geom, const double sigma[3]);

/* ---- a2: multiscale_demons (deformable.py:31-187), device resident ------------------------------------ */
/* d_initial_soa may be NULL (zero field); with initial_geom == NULL it is a field already on the fixed grid (sampled from
 * initial_transform, deformable.py:101-108).  d_out_soa: field on the fixed grid.  h_level_stats[n_levels]
 * filled after an internal synchronisation. */
int b200reg_multiscale_demons(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                              const b200reg_geom* moving_geom, const b200reg_multires_config* cfg,
                              const double* d_initial_soa, const b200reg_geom* initial_geom, double* d_out_soa,
                              b200reg_demons_stats* h_level_stats);
/* utils.py:195-267 smooth_and_resample output grid: size = int(sz/f + 0.5), align-corners spacing */
int b200reg_pyramid_geom(const b200reg_geom* in_geom, int isotropic, double resolution, b200reg_geom* out_geom);

/* ---- N12/N13: label fusion (fusion.py:56-202, 239-292) -------------------------------------------------- */
/* compute_weight_map: vote_type 0 = unweighted, 1 = global (factor / sum SSD), 2 = local (1/(G*SD + eps)) */
int b200reg_weight_map(b200reg_ctx* ctx, const float* d_target, const float* d_moving, const b200reg_geom* geom, int vote_type,
                       double factor, double sigma, double epsilon, float* d_weight);
/* vote_type "block" (fusion.py:179-200): factor * Pow(BoxMean(SSD, radius), -1) ** |gain / 2|, Float32 stages.
 * radius = blockSize per axis (x, y, z) as passed to sitk.BoxMean. */
int b200reg_weight_map_block(b200reg_ctx* ctx, const float* d_target, const float* d_moving, const b200reg_geom* geom,
                             const int32_t radius[3], double factor, double gain, float* d_weight);
/* normalise=True / normalise=<mask image> (fusion.py:171-177,196-200): weight /= max(weight) or
 * max(sitk.Mask(weight, mask)); d_mask may be NULL.  In place, no host round trip. */
int b200reg_normalise_by_max(b200reg_ctx* ctx, float* d_weight, const uint8_t* d_mask, size_t n);
/* acc_num += w * label ; acc_den += w  (fusion.py:263,269-276), f32 arithmetic in atlas order */
int b200reg_vote_accumulate(b200reg_ctx* ctx, const uint8_t* d_label, const float* d_weight, float* d_acc_num,
                            float* d_acc_den, size_t n, int first);
/* num / guarded den -> DiscreteGaussian(var) -> RescaleIntensity(0,1) -> Threshold (fusion.py:264-288).
 * d_num is overwritten; d_out may alias d_num.  Synchronises (global min/max). */
int b200reg_vote_finalize(b200reg_ctx* ctx, float* d_num, const float* d_den, const b200reg_geom* geom,
                          double smooth_variance, double threshold, float* d_out);
/* sitk.BinaryThreshold(img, lowerThreshold, upperThreshold) -> UInt8 {0,1} (fusion.py:217-220) */
int b200reg_binary_threshold(b200reg_ctx* ctx, const void* d_in, int dtype, size_t n, double lower, double upper, uint8_t* d_out);
/* STAPLE exchange: acc |= (label != 0) << bit ; out = (acc >> bit) & 1.  With disjoint bits per atlas a SUM
 * all-reduce of the int32 volume equals the bitwise OR, after which every rank holds all decisions. */
int b200reg_pack_decision(b200reg_ctx* ctx, const uint8_t* d_label, int bit, int32_t* d_packed, size_t n, int first);
int b200reg_unpack_decision(b200reg_ctx* ctx, const int32_t* d_packed, int bit, uint8_t* d_out, size_t n);
/* Compact exchange formats of the structure-sharded fusion (multiatlas/run.py:364, fusion.py:205-292):
 *  - STAPLE: one decision mask per structure, bit a = BinaryThreshold(label of atlas a, 0.5, 255) (fusion.py:217-220).
 *    packed_dtype B200REG_U8 (<= 8 atlases), B200REG_U16 (<= 16) or B200REG_U32; ranks own disjoint bits, so a SUM
 *    reduce-scatter over the structure axis is the bitwise OR.  b200reg_staple_packed runs sitk.STAPLE + RescaleIntensity +
 *    Threshold (fusion.py:223-232) straight from the reduced mask: raters = the bits of holder_mask in ascending order
 *    (the atlases that hold the structure), at most 16.  h_pq / h_elapsed may be NULL (the call then does not synchronise).
 *  - unweighted vote: u8 sum of the label values; d_flag (device int32, zeroed by the caller) is raised when a label
 *    value exceeds 1 -- the caller then falls back to the float32 accumulators of b200reg_vote_accumulate.
 *    b200reg_vote_finalize_counts = fusion.py:263-288 with num = float(count), den = n_holders. */
int b200reg_pack_label(b200reg_ctx* ctx, const uint8_t* d_label, int bit, void* d_packed, int packed_dtype, size_t n, int first);
int b200reg_staple_packed(b200reg_ctx* ctx, const void* d_packed, int packed_dtype, uint32_t holder_mask, size_t n,
                          double confidence_weight, uint32_t max_iterations, double threshold, int rescale, double* d_out,
                          double* h_pq, int32_t* h_elapsed);
int b200reg_count_accumulate(b200reg_ctx* ctx, const uint8_t* d_label, uint8_t* d_counts, size_t n, int first, int32_t* d_flag);
int b200reg_vote_finalize_counts(b200reg_ctx* ctx, const uint8_t* d_counts, int n_holders, const b200reg_geom* geom,
                                 double smooth_variance, double threshold, float* d_out);
/* ---- N14: sitk.STAPLE + RescaleIntensity + Threshold (fusion.py:217-232) --------------------------------- */
/* d_decisions: n_raters pointers (host array of device pointers) to u8 volumes already binarised
 * (>= 0.5).  d_out: f64.  h_pq (optional): 2*n_raters doubles (p then q).  Synchronises. */
int b200reg_staple(b200reg_ctx* ctx, const uint8_t* const* d_decisions, int n_raters, size_t n, double confidence_weight,
                   uint32_t max_iterations, double threshold, int rescale, double* d_out, double* h_pq, int32_t* h_elapsed);

/* ---- N15: process_probability_image (fusion.py:295-328) and the binary post-processing around it ---------- */
/* sitk.BinaryFillhole(img) (fusion.py:311; FullyConnected=False, foreground 1): background components that do not
 * reach the image border become foreground.  size = (x, y, z); d_out may alias d_in. */
int b200reg_binary_fillhole(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], int fully_connected, uint8_t* d_out);
/* sitk.ConnectedComponent -> LabelShapeStatistics -> argmax(GetNumberOfPixels) -> (labels == k) -> Cast(UInt8)
 * (fusion.py:314-328), equivalently sitk.RelabelComponent(sitk.ConnectedComponent(x)) == 1 (multiatlas/run.py:423):
 * the largest face-connected object, the first in raster order on ties.  h_n_components / h_voxels (optional) are
 * filled after a stream synchronisation; with no object the output is all zeros. */
int b200reg_largest_component(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], int fully_connected, uint8_t* d_out,
                              int64_t* h_n_components, int64_t* h_voxels);
/* The whole of process_probability_image on the device: p / max(p) -> BinaryThreshold(lower = threshold) ->
 * BinaryFillhole -> largest connected component -> UInt8.  dtype: B200REG_F32 or B200REG_F64. */
int b200reg_process_probability(b200reg_ctx* ctx, const void* d_prob, int dtype, const int32_t size[3], double threshold,
                                uint8_t* d_out, int64_t* h_n_components);

/* ---- linear_registration (linear.py:50-260): sitk.ImageRegistrationMethod metric evaluation ---------------------- */
/* MeanSquares metric with linear interpolation over every `stride`-th fixed voxel in raster order (REGULAR sampling,
 * linear.py:150-152), optional fixed / moving masks (linear.py:159-163; u8 volumes on the fixed / moving grid).
 * The point map is y = total_matrix x + total_offset (optimised transform followed by the moving-initial transform,
 * linear.py:135-138); initial_matrix is the moving-initial transform's matrix and center the optimised transform's
 * centre.  h_out (after a stream synchronisation): [0] sum of squared differences, [1] number of valid samples,
 * [2..4] s = sum w, [5..13] S = sum w (x - center)^T row-major, w = 2 (M - F) initial_matrix^T grad M:
 * d(sum sq)/d(translation) = s, d(sum sq)/d(matrix parameter k) = <dR/dp_k, S>. */
int b200reg_linreg_meansq(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                          const b200reg_geom* moving_geom, const double total_matrix[9], const double total_offset[3],
                          const double initial_matrix[9], const double center[3], const uint8_t* d_fixed_mask,
                          const uint8_t* d_moving_mask, int stride, double h_out[14]);

/* Metric "correlation" (linear.py:141-146, SetMetricAsCorrelation -> itk::CorrelationImageToImageMetricv4), same sampling and
 * arguments as b200reg_linreg_meansq.  h_out (after a stream synchronisation): [0] number of valid samples N, [1] sum F, [2] sum M,
 * [3] sum F^2, [4] sum M^2, [5] sum F M, then for each weight w in (1, F, M) twelve numbers: s_w = sum w h (3) and
 * S_w = sum w h (x - center)^T (9, row-major), h = initial_matrix^T grad M.  value = -sFM^2 / (sFF sMM) with mean-centred sums;
 * its derivative is alpha (G_F - mF G_1) + beta (G_M - mM G_1), alpha = -2 sFM / (sFF sMM), beta = 2 sFM^2 / (sFF sMM^2). */
int b200reg_linreg_correlation(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                               const b200reg_geom* moving_geom, const double total_matrix[9], const double total_offset[3],
                               const double initial_matrix[9], const double center[3], const uint8_t* d_fixed_mask,
                               const uint8_t* d_moving_mask, int stride, double h_out[42]);

/* Metric "mattes_mi" (linear.py:145-146, SetMetricAsMattesMutualInformation -> itk::MattesMutualInformationImageToImageMetricv4):
 * same sampling as the two metrics above, in two calls around a little host arithmetic.  *_bins = (bin size, normalised minimum)
 * = ((max - min) / (n_bins - 4), min / bin size - 2) of the fixed / moving image.
 *  1. histogram: h_hist[f * n_bins + m] = sum over samples of B3(m - M / moving bin size + moving normalised minimum) in the row
 *     of the sample's fixed bin (bins clamped to [2, n_bins - 3]); h_count = number of valid samples.  Deterministic (fixed-point
 *     integer atomics).  With p = hist / count:  value = - sum p log(p / (p_F p_M)).
 *  2. derivative: h_table[f * n_bins + m] = log(p(f, m) / p_M(m)) (0 where undefined); h_out = [s (3), S (9 row-major)] with
 *     w = sum_m table[f][m] B3'(m - u) (-1 / moving bin size), s = sum w h, S = sum w h (x - center)^T, h = initial_matrix^T grad M:
 *     d(value)/d(translation) = -s / count,  d(value)/d(matrix parameter k) = -<dR/dp_k, S> / count.
 * Both synchronise. */
int b200reg_linreg_mattes_histogram(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                                    const b200reg_geom* moving_geom, const double total_matrix[9], const double total_offset[3],
                                    const uint8_t* d_fixed_mask, const uint8_t* d_moving_mask, int stride, int n_bins, const double fixed_bins[2],
                                    const double moving_bins[2], double* h_hist, double* h_count);
int b200reg_linreg_mattes_derivative(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                                     const b200reg_geom* moving_geom, const double total_matrix[9], const double total_offset[3],
                                     const double initial_matrix[9], const double center[3], const uint8_t* d_fixed_mask,
                                     const uint8_t* d_moving_mask, int stride, int n_bins, const double fixed_bins[2], const double moving_bins[2],
                                     const double* h_table, double h_out[12]);

/* itk::ImageMomentsCalculator as sitk.CenteredTransformInitializer(fixed, moving, transform, MOMENTS) uses it (linear.py:40-43):
 * h_out = [sum v, sum v x, sum v y, sum v z] over all voxels, (x, y, z) the voxel's physical position.  Synchronises. */
int b200reg_image_moments(b200reg_ctx* ctx, const float* d_image, const b200reg_geom* geom, double h_out[4]);

/* ---- label utilities around the fusion step (multiatlas/run.py:200-259, 387-437) -------------------------------------- */
/* sitk.LabelStatisticsImageFilter.GetBoundingBox (utils/crop.py:44-46) of the non-zero voxels of a UInt8 mask:
 * h_bbox = (min x, min y, min z, max x, max y, max z); an empty mask gives max < min.  Synchronises. */
int b200reg_bounding_box(b200reg_ctx* ctx, const uint8_t* d_mask, const int32_t size[3], int32_t h_bbox[6]);
/* sitk.RegionOfInterest (crop_to_roi, utils/crop.py:74-76) and sitk.Paste (run.py:387-404): copy a region between two
 * volumes of the same pixel type.  A region outside either volume is an error (ITK raises). */
int b200reg_region_copy(b200reg_ctx* ctx, const void* d_src, const int32_t src_size[3], const int32_t src_index[3], void* d_dst,
                        const int32_t dst_size[3], const int32_t dst_index[3], const int32_t region_size[3], int dtype);
/* correct_volume_overlap (label/utils.py:23-58): labels in rank order; out[s] = labels[s] > 0 and no earlier label has the voxel */
int b200reg_resolve_overlap(b200reg_ctx* ctx, const uint8_t* const* d_labels_ranked, uint8_t* const* d_out, int n_labels, size_t n);
/* sitk.BinaryMorphologicalClosing(img, radius) (run.py:424; SafeBorder on, foreground 1): dilation then erosion with the
 * structuring element given as n_offsets (dx, dy, dz) triples (host memory).  Synchronises. */
int b200reg_binary_closing(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], const int32_t radius[3],
                           const int32_t* h_offsets, int n_offsets, uint8_t* d_out);

/* ---- distance maps, contours, binary morphology, masking (registration/utils.py:270-344, label/projection.py:9-92) ------- */
/* sitk.SignedMaurerDistanceMap(mask, insideIsPositive, squaredDistance, useImageSpacing) -> Float32 (utils.py:289-294,
 * projection.py:22-31,80-82; background value 0).  Distance to the nearest object voxel that has a background voxel among its
 * 26 neighbours (0 on those voxels), computed in single precision like ITK's filter does in its output pixel type. */
int b200reg_signed_maurer_distance_map(b200reg_ctx* ctx, const uint8_t* d_mask, const b200reg_geom* geom, int inside_is_positive,
                                       int squared_distance, int use_image_spacing, float* d_out);
/* sitk.LabelContour(image, fullyConnected) (projection.py:33,85; background 0): labelled voxels with a differently labelled
 * neighbour keep their value, everything else becomes 0.  Not in place. */
int b200reg_label_contour(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], int fully_connected, uint8_t* d_out);
/* sitk.LabelContour(image[:, :, k]) for every axial slice k (label/comparison.py:373-374, the added path length): the four
 * in-plane face neighbours only.  Not in place. */
int b200reg_label_contour_slicewise(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], uint8_t* d_out);
/* sitk.BinaryDilate / sitk.BinaryErode (utils.py:331, generation/dvf.py:269-287): foreground 1, background 0, structuring
 * element as n_offsets (dx, dy, dz) triples in host memory; boundary_to_foreground as in SimpleITK (default false for the
 * dilation, true for the erosion).  Not in place.  Synchronises. */
int b200reg_binary_dilate(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], const int32_t* h_offsets, int n_offsets,
                          int boundary_to_foreground, uint8_t* d_out);
int b200reg_binary_erode(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], const int32_t* h_offsets, int n_offsets,
                         int boundary_to_foreground, uint8_t* d_out);
/* mask_a | mask_b, &, + (modulo 256), ^ on UInt8 volumes (generation/dvf.py:66,249,290) */
typedef enum { B200REG_OP_OR = 0, B200REG_OP_AND = 1, B200REG_OP_ADD = 2, B200REG_OP_XOR = 3 } b200reg_u8_op;
int b200reg_u8_binary_op(b200reg_ctx* ctx, const uint8_t* d_a, const uint8_t* d_b, int op, uint8_t* d_out, size_t n);
/* sitk.Mask(image, mask, outsideValue) (utils.py:337-340; generation/dvf.py:66,121,200): `planes` planes of n voxels (3 for a
 * SoA displacement field); d_out may alias d_in. */
int b200reg_mask_image(b200reg_ctx* ctx, const void* d_in, int dtype, const uint8_t* d_mask, size_t n, int planes, double outside_value,
                       void* d_out);
/* image / constant on a Float32 or Float64 image (utils.py:297,342); d_out may alias d_in. */
int b200reg_divide_scalar(b200reg_ctx* ctx, const void* d_in, int dtype, size_t n, double divisor, void* d_out);

/* ---- field templates of the synthetic-deformation generators (generation/dvf.py:29-415) --------------------------------- */
/* The constant displacement `vector` (dx, dy, dz) inside d_mask (everywhere when d_mask is NULL), 0 outside
 * (dvf.py:54-66,114-121,187-200: np.zeros + vector, CopyInformation, sitk.Mask). */
int b200reg_constant_field(b200reg_ctx* ctx, const uint8_t* d_mask, size_t n, const double vector[3], double* d_out_soa);
/* generate_field_radial_bend (dvf.py:362-394): scale * cross(voxel index - reference_index, axis), all in (x, y, z) order,
 * inside the body mask cut by a half space through the reference voxel: clip_axis -1 none, 0 x, 1 y, 2 z; voxels with
 * index >= reference are kept when clip_keep_upper, else those with index < reference. */
int b200reg_radial_bend_field(b200reg_ctx* ctx, const uint8_t* d_mask, const int32_t size[3], const int32_t reference_index[3],
                              const double axis[3], double scale, int clip_axis, int clip_keep_upper, double* d_out_soa);

/* ---- compute_weight_map, vote_type "patch_correlation" (fusion.py:82-146) ------------------------------------------------ */
/* Pearson correlation of target and moving over the window (wx, wy, wz voxels) around every voxel, clipped to the image: output
 * voxel i covers [i - (w - 1) / 2, i + w / 2] per axis (fusion.py:96-103).  Float64 arithmetic and scipy.stats.pearsonr's special
 * cases (constant patch -> NaN -> 0 by fusion.py:124, two samples, clipping to [-1, 1]).  d_out: Float64 on the same grid. */
int b200reg_patch_correlation(b200reg_ctx* ctx, const float* d_target, const float* d_moving, const int32_t size[3], const int32_t window[3],
                              double* d_out);
/* out = (take_abs ? |in| : in) * mul + add on a Float32 / Float64 image, constants in the pixel type: the image-with-constant
 * operators a correlation_function is made of (fusion.py:138-146: x + 1, sitk.Abs(x)).  d_out may alias d_in. */
int b200reg_scale_shift(b200reg_ctx* ctx, const void* d_in, int dtype, size_t n, int take_abs, double mul, double add, void* d_out);

#ifdef __cplusplus
}
#endif
#endif /* B200REG_H */
