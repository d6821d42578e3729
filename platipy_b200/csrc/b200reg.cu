// b200reg.cu -- C ABI of libb200reg.so (see include/b200reg.h).  sm_100a only, -fmad=false.
#include <cstdlib>

#include "common.cuh"
#include "demons.cuh"
#include "deriche.cuh"
#include "fusion.cuh"
#include "morph.cuh"
#include "linreg.cuh"
#include "postproc.cuh"
#include "distmap.cuh"
#include "patchcorr_kernels.cuh"
#include "linreg_corr_kernels.cuh"
#include "linreg_mattes_kernels.cuh"
#include "moments_kernels.cuh"
#include "gauss.cuh"
#include "resample.cuh"
#include "pyramid.cuh"

using namespace b200;

#define API extern "C" __attribute__((visibility("default")))
#define REQUIRE(cond, ...)                                             \
    do {                                                               \
        if (!(cond)) return set_error(B200REG_ERR_ARG, __VA_ARGS__);   \
    } while (0)

namespace {
// tuning / A-B switches: PLATIPY_B200_<NAME> (the name SURVEY.md section 5 gives the configuration knobs), or the older B200REG_<NAME>
const char* knob(const char* name)
{
    char buf[96];
    snprintf(buf, sizeof(buf), "PLATIPY_B200_%s", name);
    if (const char* e = getenv(buf)) return e;
    snprintf(buf, sizeof(buf), "B200REG_%s", name);
    return getenv(buf);
}
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace
#define ENTER(ctx)                                      \
    REQUIRE((ctx) != nullptr, "null context");          \
    DeviceGuard _guard((ctx)->device)

// ---- context ------------------------------------------------------------------------------------------
API int b200reg_abi_version(void) { return B200REG_ABI_VERSION; }
API const char* b200reg_last_error(void) { return last_error_ref().c_str(); }

API int b200reg_create(int device, void* stream, b200reg_ctx** out)
{
    REQUIRE(out != nullptr, "null output pointer");
    int count = 0;
    B200_CUDA(cudaGetDeviceCount(&count));
    REQUIRE(device >= 0 && device < count, "CUDA device %d not present (%d visible)", device, count);
    B200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return set_error(B200REG_ERR_UNSUPPORTED, "libb200reg is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    auto* ctx = new b200reg_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        B200_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->owns_stream = true;
    }
    // keep freed stream-ordered allocations cached in the pool
    cudaMemPool_t pool;
    B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    B200_CUDA(cudaMallocHost(&ctx->h_scratch, 64 * sizeof(double)));
    if (const char* e = knob("FORCE_SEPARABLE")) ctx->force_separable = (e[0] == '1');
    if (const char* e = knob("UNFUSED_FORCE")) ctx->unfused_force = (e[0] == '1');
    if (const char* e = knob("STAPLE_VOXELWISE")) ctx->staple_voxelwise = (e[0] == '1');
    if (const char* e = knob("ZM_TMA")) ctx->zm_tma = atoi(e);
    if (const char* e = knob("ZM_TMA_L2")) ctx->zm_tma_l2 = atoi(e);
    if (const char* e = knob("ZM_REGADD")) ctx->zm_regadd = (e[0] != '0');
    if (const char* e = knob("PF_WARP")) ctx->pf_warp = atoi(e);
    if (const char* e = knob("PF_FORCE")) ctx->pf_force = atoi(e);
    if (const char* e = knob("WARP_MARCH")) ctx->warp_march = atoi(e);
    if (const char* e = knob("ZM_CHUNKS")) ctx->zm_chunks = atoi(e);
    if (const char* e = knob("ZM_TX32")) ctx->zm_tx32 = (e[0] != '0');
    if (const char* e = knob("ZM_ADDOUT")) ctx->zm_addout = (e[0] != '0');
    if (const char* e = knob("FORCE_ZM1")) ctx->force_zm1 = (e[0] == '1');
    if (const char* e = knob("PACK_LABELS")) ctx->pack_labels = (e[0] != '0');
    if (const char* e = knob("PDL")) ctx->pdl = (e[0] != '0');
    if (const char* e = knob("IDENTITY_COPY")) ctx->identity_copy = (e[0] != '0');
    if (const char* e = knob("CONV_STATIC_RADIUS")) ctx->conv_static_radius = (e[0] != '0');
    if (const char* e = knob("WARP_RESAMPLE")) ctx->warp_resample = (e[0] != '0');
    if (const char* e = knob("PYRAMID_RESTRICT")) ctx->pyramid_restrict = (e[0] != '0');
    if (const char* e = knob("PYRAMID_RESTRICT_COST")) ctx->pyramid_restrict_cost = atof(e);
    *out = ctx;
    return B200REG_OK;
}
API int b200reg_destroy(b200reg_ctx* ctx)
{
    if (!ctx) return B200REG_OK;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    if (ctx->h_trace) cudaFreeHost(ctx->h_trace);
    delete ctx;
    return B200REG_OK;
}
API int b200reg_set_stream(b200reg_ctx* ctx, void* stream)
{
    ENTER(ctx);
    if (ctx->owns_stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
        ctx->owns_stream = false;
    }
    ctx->stream = (cudaStream_t)stream;
    return B200REG_OK;
}
API int b200reg_synchronize(b200reg_ctx* ctx)
{
    ENTER(ctx);
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200REG_OK;
}
API int64_t b200reg_launch_count(b200reg_ctx* ctx) { return ctx ? ctx->launches : -1; }

// ---- named semantic switches (process-wide; see common.cuh) ---------------------------------------------------------------------
API int b200reg_set_semantic(const char* name, int value)
{
    int* slot = semantic_slot(name);
    REQUIRE(slot != nullptr, "unknown semantic switch '%s'", name ? name : "(null)");
    *slot = value;
    return B200REG_OK;
}
API int b200reg_get_semantic(const char* name)
{
    const int* slot = semantic_slot(name);
    return slot ? *slot : -1;
}

// ---- memory helpers ---------------------------------------------------------------------------------------
API int b200reg_malloc(b200reg_ctx* ctx, size_t bytes, void** d_ptr)
{
    ENTER(ctx);
    REQUIRE(d_ptr != nullptr, "null output pointer");
    B200_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 16));
    return B200REG_OK;
}
API int b200reg_free(b200reg_ctx* ctx, void* d_ptr)
{
    ENTER(ctx);
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    B200_CUDA(cudaFree(d_ptr));
    return B200REG_OK;
}
API int b200reg_malloc_host(size_t bytes, void** h_ptr)
{
    REQUIRE(h_ptr != nullptr, "null output pointer");
    B200_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 16));
    return B200REG_OK;
}
API int b200reg_free_host(void* h_ptr)
{
    B200_CUDA(cudaFreeHost(h_ptr));
    return B200REG_OK;
}
API int b200reg_memcpy_h2d(b200reg_ctx* ctx, void* d_dst, const void* h_src, size_t bytes)
{
    ENTER(ctx);
    B200_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return B200REG_OK;
}
API int b200reg_memcpy_d2h(b200reg_ctx* ctx, void* h_dst, const void* d_src, size_t bytes)
{
    ENTER(ctx);
    B200_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return B200REG_OK;
}
API int b200reg_memset(b200reg_ctx* ctx, void* d_ptr, int value, size_t bytes)
{
    ENTER(ctx);
    B200_CUDA(cudaMemsetAsync(d_ptr, value, bytes, ctx->stream));
    return B200REG_OK;
}

// ---- layout / dtype helpers ---------------------------------------------------------------------------------
API int b200reg_aos_to_soa(b200reg_ctx* ctx, const double* d_aos, double* d_soa, size_t n)
{
    ENTER(ctx);
    REQUIRE(d_aos && d_soa, "null pointer");
    aos_to_soa_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_aos, d_soa, n);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_soa_to_aos(b200reg_ctx* ctx, const double* d_soa, double* d_aos, size_t n)
{
    ENTER(ctx);
    REQUIRE(d_aos && d_soa, "null pointer");
    soa_to_aos_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_soa, d_aos, n);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_cast(b200reg_ctx* ctx, const void* d_in, int in_dtype, void* d_out, int out_dtype, size_t n)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out, "null pointer");
    const int nb = ctx->sm_count * 8;
    B200_DISPATCH_DTYPE(in_dtype, TI, {
        B200_DISPATCH_DTYPE(out_dtype, TO, { cast_kernel<TI, TO><<<nb, 256, 0, ctx->stream>>>((const TI*)d_in, (TO*)d_out, n); });
    });
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_minmax(b200reg_ctx* ctx, const void* d_in, int dtype, size_t n, double* h_min, double* h_max)
{
    ENTER(ctx);
    REQUIRE(d_in && n > 0, "empty input");
    TempBuf part, mm;
    B200_TRY(mm.alloc(ctx, 2 * sizeof(double)));
    B200_DISPATCH_DTYPE(dtype, T, { B200_TRY(minmax_device<T>(ctx, (const T*)d_in, n, mm.as<double>(), &part)); });
    B200_CUDA(small_d2h(ctx, ctx->h_scratch, mm.p, 2 * sizeof(double)));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_min) *h_min = ctx->h_scratch[0];
    if (h_max) *h_max = ctx->h_scratch[1];
    return B200REG_OK;
}

// ---- N1 -------------------------------------------------------------------------------------------------------
API int b200reg_discrete_gaussian_f32(b200reg_ctx* ctx, const float* d_in, float* d_out, const b200reg_geom* geom, const double variance[3],
                                      int max_kernel_width, double max_error, int use_image_spacing)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && variance && valid_geom(geom), "invalid argument");
    return discrete_gaussian_f32(ctx, d_in, d_out, *geom, variance, max_kernel_width, max_error, use_image_spacing);
}
API int b200reg_gaussian_operator(double variance, double max_error, int max_kernel_width, double* h_kernel, int capacity)
{
    const std::vector<double> k = gaussian_operator(variance, max_error, max_kernel_width);
    if ((int)k.size() > capacity) return -1;
    for (size_t i = 0; i < k.size(); ++i) h_kernel[i] = k[i];
    return ((int)k.size() - 1) / 2;
}

// Host-only diagnostic of the proof behind the copy / per-index shortcuts (resample.cuh: identity_resample_is_exact); no context, no device.
API int b200reg_identity_resample_is_exact(const b200reg_geom* in_geom, const b200reg_geom* out_geom, int allow_scanline)
{
    if (!valid_geom(in_geom) || !valid_geom(out_geom)) return 0;
    return identity_resample_is_exact(*in_geom, *out_geom, allow_scanline != 0) ? 1 : 0;
}

// ---- N2/N5/N9 ---------------------------------------------------------------------------------------------------
// one Float32 image, linear interpolation, one displacement field on the output grid: the Demons loop's warp kernel (demons_split.cuh)
static int resample_batch_routed(b200reg_ctx* ctx, int n, const void* const* d_in, const int* dtypes, const b200reg_geom& gin, void* const* d_out,
                                 const b200reg_geom& gout, const b200reg_transform* chain, int n_chain, const int* interps, const double* default_values)
{
    if (n == 1 && dtypes[0] == B200REG_F32 && d_in[0] && d_out[0] && d_in[0] != d_out[0]) {
        bool used = false;
        B200_TRY(resample_f32_on_grid_dvf(ctx, static_cast<const float*>(d_in[0]), gin, static_cast<float*>(d_out[0]), gout, chain, n_chain, interps[0],
                                          default_values[0], &used));
        if (used) return B200REG_OK;
    }
    if (n == 1 && d_in[0] && d_out[0] && d_in[0] != d_out[0] && interps[0] == B200REG_INTERP_NN) {
        // one label, nearest neighbour, field on the output grid (the per-structure calls of multiatlas/run.py:338-345)
        bool used = false;
        B200_TRY(resample_nn_on_grid_dvf(ctx, d_in[0], dtypes[0], gin, d_out[0], gout, chain, n_chain, interps[0], default_values[0], &used));
        if (used) return B200REG_OK;
    }
    return resample_batch(ctx, n, d_in, dtypes, gin, d_out, gout, chain, n_chain, interps, default_values);
}
API int b200reg_resample_batch(b200reg_ctx* ctx, int n, const void* const* d_in, const int* dtypes, const b200reg_geom* in_geom,
                               void* const* d_out, const b200reg_geom* out_geom, const b200reg_transform* chain, int n_chain,
                               const int* interps, const double* default_values)
{
    ENTER(ctx);
    REQUIRE(n > 0 && n <= B200REG_MAX_BATCH, "batch size %d not in [1, %d]", n, B200REG_MAX_BATCH);
    REQUIRE(d_in && d_out && dtypes && interps && default_values, "null argument");
    REQUIRE(valid_geom(in_geom) && valid_geom(out_geom), "invalid geometry");
    return resample_batch_routed(ctx, n, d_in, dtypes, *in_geom, d_out, *out_geom, chain, n_chain, interps, default_values);
}
API int b200reg_resample(b200reg_ctx* ctx, const void* d_in, int dtype, const b200reg_geom* in_geom, void* d_out, const b200reg_geom* out_geom,
                         const b200reg_transform* chain, int n_chain, int interp, double default_value)
{
    return b200reg_resample_batch(ctx, 1, &d_in, &dtype, in_geom, &d_out, out_geom, chain, n_chain, &interp, &default_value);
}
API int b200reg_resample_vec3(b200reg_ctx* ctx, const double* d_in_soa, const b200reg_geom* in_geom, double* d_out_soa,
                              const b200reg_geom* out_geom, const b200reg_transform* chain, int n_chain, double default_value)
{
    ENTER(ctx);
    REQUIRE(d_in_soa && d_out_soa && valid_geom(in_geom) && valid_geom(out_geom), "invalid argument");
    return resample_vec3(ctx, d_in_soa, *in_geom, d_out_soa, *out_geom, chain, n_chain, default_value);
}

API int b200reg_transform_to_dvf(b200reg_ctx* ctx, const b200reg_geom* out_geom, const b200reg_transform* chain, int n_chain, double* d_out_soa)
{
    ENTER(ctx);
    REQUIRE(d_out_soa && valid_geom(out_geom), "invalid argument");
    ChainD ch;
    B200_TRY(make_chain(chain, n_chain, &ch));
    const GeomD go = make_geomd(*out_geom);
    transform_to_dvf_kernel<<<grid3(go.nx, go.ny, go.nz), block3(), 0, ctx->stream>>>(d_out_soa, go, ch);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

static b200reg_transform dvf_transform(const double* d_soa, const b200reg_geom& g)
{
    b200reg_transform t;
    memset(&t, 0, sizeof(t));
    t.kind = B200REG_TFM_DVF;
    t.d_dvf = d_soa;
    t.dvf_geom = g;
    return t;
}

API int b200reg_compose_dvf(b200reg_ctx* ctx, double* d_total_soa, const double* d_iter_soa, const b200reg_geom* geom, double* d_scratch_soa)
{
    ENTER(ctx);
    REQUIRE(d_total_soa && d_iter_soa && d_scratch_soa && valid_geom(geom), "invalid argument");
    const b200reg_transform t = dvf_transform(d_total_soa, *geom);
    B200_TRY(resample_vec3(ctx, d_iter_soa, *geom, d_scratch_soa, *geom, &t, 1, 0.0, d_total_soa));
    B200_CUDA(cudaMemcpyAsync(d_total_soa, d_scratch_soa, 3 * nvox(*geom) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return B200REG_OK;
}

// ---- N6 -----------------------------------------------------------------------------------------------------------
static int check_demons_params(const b200reg_demons_params* p)
{
    REQUIRE(p != nullptr, "null Demons parameters");
    REQUIRE(p->number_of_iterations >= 0, "negative iteration count");
    for (int a = 0; a < 3; ++a) REQUIRE(p->std_dev[a] >= 0 && p->update_std_dev[a] >= 0, "negative standard deviation");
    REQUIRE(p->max_error > 0 && p->max_error < 1, "maximum error must be in (0, 1)");
    return B200REG_OK;
}

static int read_stats(b200reg_ctx* ctx, const DemonsWorkspace& ws, const b200reg_geom& g, b200reg_demons_stats* st, float ms, int n_iters = 0,
                      int level = -1)
{
    DemonsCtrl h;
    B200_CUDA(small_d2h(ctx, ctx->h_scratch, ws.ctrl.p, sizeof(DemonsCtrl)));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(&h, ctx->h_scratch, sizeof(h));
    if (level >= 0 && ws.trace.p) {
        // the iteration trace (metric, RMS change per iteration, written by demons_finish_kernel): a few hundred bytes
        if ((int)ctx->traces.size() <= level) ctx->traces.resize(level + 1);
        const int n = h.elapsed < n_iters ? h.elapsed : n_iters;
        ctx->traces[level].assign(2 * (size_t)(n > 0 ? n : 0), 0.0);
        if (n > 0) {
            if (ctx->h_trace_doubles < 2 * (size_t)n) {
                if (ctx->h_trace) cudaFreeHost(ctx->h_trace);
                ctx->h_trace = nullptr;
                ctx->h_trace_doubles = 2 * (size_t)(n < 512 ? 512 : n);
                B200_CUDA(cudaMallocHost(&ctx->h_trace, ctx->h_trace_doubles * sizeof(double)));
            }
            B200_CUDA(small_d2h(ctx, ctx->h_trace, ws.trace.p, sizeof(double) * 2 * (size_t)n));
            B200_CUDA(cudaStreamSynchronize(ctx->stream));
            memcpy(ctx->traces[level].data(), ctx->h_trace, sizeof(double) * 2 * (size_t)n);
        }
    }
    st->elapsed_iterations = h.elapsed;
    st->voxels_lo = (int32_t)(nvox(g) & 0x7fffffff);
    st->metric = h.metric;
    st->rms_change = h.rms;
    st->gpu_ms = ms;
    return B200REG_OK;
}

API int b200reg_demons_execute(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                               const b200reg_geom* moving_geom, const b200reg_demons_params* params, double* d_out_soa,
                               b200reg_demons_stats* h_stats)
{
    ENTER(ctx);
    REQUIRE(d_fixed && d_moving && d_out_soa && valid_geom(fixed_geom) && valid_geom(moving_geom), "invalid argument");
    B200_TRY(check_demons_params(params));
    DemonsWorkspace ws;
    B200_TRY(demons_prepare(ctx, *fixed_geom, params->number_of_iterations, &ws, true));
    ctx->traces.clear();
    cudaEvent_t e0, e1;
    B200_CUDA(cudaEventCreate(&e0));
    B200_CUDA(cudaEventCreate(&e1));
    B200_CUDA(cudaEventRecord(e0, ctx->stream));
    int rc = demons_enqueue(ctx, d_fixed, *fixed_geom, d_moving, *moving_geom, *params, d_out_soa, &ws);
    cudaEventRecord(e1, ctx->stream);
    float ms = 0.f;
    if (rc == B200REG_OK) {
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    B200_TRY(rc);
    b200reg_demons_stats local;
    B200_TRY(read_stats(ctx, ws, *fixed_geom, h_stats ? h_stats : &local, ms, params->number_of_iterations, 0));
    return B200REG_OK;
}

API int b200reg_demons_trace(b200reg_ctx* ctx, int level, double* h_metric_rms, int capacity_iterations)
{
    if (!ctx || level < 0 || level >= (int)ctx->traces.size()) return 0;
    const int n = (int)(ctx->traces[level].size() / 2);
    const int m = n < capacity_iterations ? n : capacity_iterations;
    if (h_metric_rms && m > 0) memcpy(h_metric_rms, ctx->traces[level].data(), sizeof(double) * 2 * (size_t)m);
    return n;
}

API int b200reg_demons_force(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                             const b200reg_geom* moving_geom, const double* d_field_soa, const b200reg_demons_params* params, float* d_w,
                             double* d_u_soa, double* h_metric, double* h_rms)
{
    ENTER(ctx);
    REQUIRE(d_fixed && d_moving && d_field_soa && d_w && d_u_soa && valid_geom(fixed_geom) && valid_geom(moving_geom), "invalid argument");
    B200_TRY(check_demons_params(params));
    DemonsWorkspace ws;
    B200_TRY(demons_prepare(ctx, *fixed_geom, 1, &ws, false));
    demons_ctrl_init_kernel<<<1, 1, 0, ctx->stream>>>(ws.ctrl.as<DemonsCtrl>(), 1);
    ctx->launches++;
    const GeomD gf = make_geomd(*fixed_geom), gm = make_geomd(*moving_geom);
    B200_TRY(demons_calculate_change(ctx, d_fixed, gf, d_moving, gm, d_field_soa, make_force_params(*fixed_geom, *params), &ws, 0, 1, true));
    const size_t n = nvox(*fixed_geom);
    B200_CUDA(cudaMemcpyAsync(d_w, ws.W.p, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    B200_CUDA(cudaMemcpyAsync(d_u_soa, ws.U.p, 3 * n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    b200reg_demons_stats st;
    B200_TRY(read_stats(ctx, ws, *fixed_geom, &st, 0.f));
    if (h_metric) *h_metric = st.metric;
    if (h_rms) *h_rms = st.rms_change;
    return B200REG_OK;
}

API int b200reg_pde_smooth_field(b200reg_ctx* ctx, double* d_field_soa, const b200reg_geom* geom, const double std_dev[3], double max_error,
                                 int max_kernel_width)
{
    ENTER(ctx);
    REQUIRE(d_field_soa && valid_geom(geom) && std_dev, "invalid argument");
    const size_t n = nvox(*geom);
    TempBuf t1, t2, t3;
    B200_TRY(t1.alloc(ctx, 3 * n * sizeof(double)));
    B200_TRY(t2.alloc(ctx, 3 * n * sizeof(double)));
    B200_TRY(t3.alloc(ctx, 3 * n * sizeof(double)));
    KernelCoeffs kc[3];
    B200_TRY(make_pde_coeffs(std_dev, max_error, max_kernel_width, kc));
    B200_TRY(pde_smooth(ctx, d_field_soa, nullptr, t1.as<double>(), t2.as<double>(), t3.as<double>(), geom->size[0], geom->size[1], geom->size[2], kc,
                        nullptr, 0));
    B200_CUDA(cudaMemcpyAsync(d_field_soa, t1.p, 3 * n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return B200REG_OK;
}

// ---- N8 --------------------------------------------------------------------------------------------------------------
API int b200reg_recursive_gaussian_vec3(b200reg_ctx* ctx, double* d_field_soa, const b200reg_geom* geom, const double sigma[3])
{
    ENTER(ctx);
    REQUIRE(d_field_soa && valid_geom(geom) && sigma, "invalid argument");
    for (int a = 0; a < 3; ++a) REQUIRE(sigma[a] > 0, "sigma must be positive");
    return recursive_gaussian_vec3(ctx, d_field_soa, *geom, sigma);
}

// ---- a2/a3: pyramid + multiscale_demons --------------------------------------------------------------------------------
API int b200reg_pyramid_geom(const b200reg_geom* in, int isotropic, double resolution, b200reg_geom* out)
{
    REQUIRE(valid_geom(in) && out, "invalid argument");
    REQUIRE(resolution > 0, "resolution must be positive");
    *out = *in;
    for (int a = 0; a < 3; ++a) {
        // utils.py:237-247: scale factor = voxel size / spacing (isotropic) or the shrink factor
        const double sf = isotropic ? (resolution * 1.0 / in->spacing[a]) : resolution;
        const int sz = (int)((double)in->size[a] / sf + 0.5);
        REQUIRE(sz >= 1, "pyramid level collapses to an empty grid along axis %d", a);
        out->size[a] = sz;
        // utils.py:252-255 (align corners); a size of 1 divides by zero in the reference as well
        REQUIRE(sz > 1, "pyramid level has a single voxel along axis %d (the reference divides by zero here)", a);
        out->spacing[a] = ((double)(in->size[a] - 1) * in->spacing[a]) / (double)(sz - 1);
    }
    return B200REG_OK;
}

// smooth_and_resample (utils.py:195-267) for one Float32 image onto a precomputed level grid
static int smooth_and_resample_f32(b200reg_ctx* ctx, const float* d_in, const b200reg_geom& gin, double sigma, const b200reg_geom& gout, int interp,
                                   float* d_out)
{
    const size_t n = nvox(gin);
    TempBuf sm;
    const float* src = d_in;
    if (sigma != 0.0) {
        const double var[3] = { sigma * sigma, sigma * sigma, sigma * sigma };
        double mw = 0.0;
        for (int a = 0; a < 3; ++a) mw = fmax(mw, 8 * var[a] * gin.spacing[a]);
        if (interp == B200REG_INTERP_LINEAR) {
            // a shrinking level: blur only what the level's interpolation reads (pyramid.cuh; bit-identical)
            bool used = false;
            B200_TRY(smooth_and_shrink_f32(ctx, d_in, gin, var, (int)mw, 0.01, gout, d_out, &used));
            if (used) return B200REG_OK;
        }
        B200_TRY(sm.alloc(ctx, n * sizeof(float)));
        B200_TRY(discrete_gaussian_f32(ctx, d_in, sm.as<float>(), gin, var, (int)mw, 0.01, 1));
        src = sm.as<float>();
    }
    const void* ins[1] = { src };
    void* outs[1] = { d_out };
    const int dt = B200REG_F32;
    const double dv = 0.0;
    return resample_batch(ctx, 1, ins, &dt, gin, outs, gout, nullptr, 0, &interp, &dv);
}

// a3 as one call: sitk.DiscreteGaussian (utils.py:216-226) followed by sitk.Resample onto the new grid (utils.py:257-267) for a Float32 image.
// allow_restricted 1: a level that shrinks enough to pay (linear interpolator) blurs only what the resampler reads (pyramid.cuh); 2: whenever
// the restricted form is possible; 0: never (the forms are bit-identical; the switch exists for tests and A/B timing).
API int b200reg_smooth_and_resample_f32(b200reg_ctx* ctx, const float* d_in, const b200reg_geom* in_geom, const double variance[3], int max_kernel_width,
                                        const b200reg_geom* out_geom, int interp, float* d_out, int allow_restricted)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && variance && valid_geom(in_geom) && valid_geom(out_geom), "invalid argument");
    REQUIRE(interp == B200REG_INTERP_NN || interp == B200REG_INTERP_LINEAR || interp == B200REG_INTERP_BSPLINE,
            "interpolator %d is not supported (nearest neighbour = 1, linear = 2, B-spline = 3)", interp);
    for (int a = 0; a < 3; ++a) REQUIRE(variance[a] >= 0.0, "variance must not be negative");
    if (allow_restricted && interp == B200REG_INTERP_LINEAR) {
        bool used = false;
        B200_TRY(smooth_and_shrink_f32(ctx, d_in, *in_geom, variance, max_kernel_width, 0.01, *out_geom, d_out, &used, allow_restricted == 2));
        if (used) return B200REG_OK;
    }
    TempBuf sm;
    B200_TRY(sm.alloc(ctx, nvox(*in_geom) * sizeof(float)));
    B200_TRY(discrete_gaussian_f32(ctx, d_in, sm.as<float>(), *in_geom, variance, max_kernel_width, 0.01, 1));
    const void* ins[1] = { sm.p };
    void* outs[1] = { d_out };
    const int dt = B200REG_F32;
    const double dv = 0.0;
    return resample_batch(ctx, 1, ins, &dt, *in_geom, outs, *out_geom, nullptr, 0, &interp, &dv);
}

API int b200reg_multiscale_demons(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                                  const b200reg_geom* moving_geom, const b200reg_multires_config* cfg, const double* d_initial_soa,
                                  const b200reg_geom* initial_geom, double* d_out_soa, b200reg_demons_stats* h_level_stats)
{
    ENTER(ctx);
    REQUIRE(d_fixed && d_moving && d_out_soa && cfg && valid_geom(fixed_geom) && valid_geom(moving_geom), "invalid argument");
    REQUIRE(cfg->n_levels >= 0 && cfg->n_levels <= B200REG_MAX_LEVELS, "number of levels %d not in [0, %d]", cfg->n_levels, B200REG_MAX_LEVELS);
    REQUIRE(cfg->interp_order == B200REG_INTERP_NN || cfg->interp_order == B200REG_INTERP_LINEAR || cfg->interp_order == B200REG_INTERP_BSPLINE,
            "interpolator %d is not supported (nearest neighbour = 1, linear = 2, B-spline = 3)", cfg->interp_order);
    B200_TRY(check_demons_params(&cfg->demons));
    const int L = cfg->n_levels;
    const b200reg_geom gF = *fixed_geom, gM = *moving_geom;
    B200_NVTX("b200reg_multiscale_demons");

    // deformable.py:67-94: both pyramids, all levels, from the ORIGINAL images
    std::vector<b200reg_geom> gfl(L), gml(L);
    std::vector<TempBuf> Fl(L), Ml(L);
    B200NvtxRange r_pyramid("pyramid (smooth_and_resample, all levels)");
    for (int l = 0; l < L; ++l) {
        const double res = cfg->resolution_staging[l];
        // utils.py:249-250: neither factor given -> image returned unchanged
        if (res == 0.0) {
            gfl[l] = gF;
            gml[l] = gM;
        } else {
            B200_TRY(b200reg_pyramid_geom(&gF, cfg->isotropic_resample, res, &gfl[l]));
            B200_TRY(b200reg_pyramid_geom(&gM, cfg->isotropic_resample, res, &gml[l]));
        }
        B200_TRY(Fl[l].alloc(ctx, nvox(gfl[l]) * sizeof(float)));
        B200_TRY(Ml[l].alloc(ctx, nvox(gml[l]) * sizeof(float)));
        if (res == 0.0) {
            // smoothed if a sigma was given (utils.py:216-226 precede the early return), never resampled
            const float* srcs[2] = { d_fixed, d_moving };
            float* dsts[2] = { Fl[l].as<float>(), Ml[l].as<float>() };
            const b200reg_geom* gs[2] = { &gF, &gM };
            for (int q = 0; q < 2; ++q) {
                const double sg = cfg->smoothing_sigmas[l];
                if (sg != 0.0) {
                    const double var[3] = { sg * sg, sg * sg, sg * sg };
                    double mw = 0.0;
                    for (int a = 0; a < 3; ++a) mw = fmax(mw, 8 * var[a] * gs[q]->spacing[a]);
                    B200_TRY(discrete_gaussian_f32(ctx, srcs[q], dsts[q], *gs[q], var, (int)mw, 0.01, 1));
                } else {
                    B200_CUDA(cudaMemcpyAsync(dsts[q], srcs[q], nvox(*gs[q]) * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
                }
            }
        } else {
            B200_TRY(smooth_and_resample_f32(ctx, d_fixed, gF, cfg->smoothing_sigmas[l], gfl[l], cfg->interp_order, Fl[l].as<float>()));
            B200_TRY(smooth_and_resample_f32(ctx, d_moving, gM, cfg->smoothing_sigmas[l], gml[l], cfg->interp_order, Ml[l].as<float>()));
        }
    }

    r_pyramid.end();
    // deformable.py:99-130: initial field on the fixed grid (zeros, or the given field re-gridded twice)
    const size_t nF = nvox(gF);
    TempBuf total, next;
    b200reg_geom g_total = gF;
    B200_TRY(total.alloc(ctx, 3 * nF * sizeof(double)));
    if (d_initial_soa && initial_geom) {
        // a given field: deformable.py:125 re-grids it onto the fixed image, then :130 once more
        REQUIRE(valid_geom(initial_geom), "invalid initial field geometry");
        TempBuf tmp;
        B200_TRY(tmp.alloc(ctx, 3 * nF * sizeof(double)));
        B200_TRY(resample_vec3(ctx, d_initial_soa, *initial_geom, tmp.as<double>(), gF, nullptr, 0, 0.0));
        B200_TRY(resample_vec3(ctx, tmp.as<double>(), gF, total.as<double>(), gF, nullptr, 0, 0.0));
    } else if (d_initial_soa) {
        // a field sampled from initial_transform on the fixed grid (deformable.py:101-108): only the :130 re-grid
        B200_TRY(resample_vec3(ctx, d_initial_soa, gF, total.as<double>(), gF, nullptr, 0, 0.0));
    } else {
        B200_CUDA(cudaMemsetAsync(total.p, 0, 3 * nF * sizeof(double), ctx->stream));
    }

    ctx->traces.clear();
    std::vector<cudaEvent_t> ev(2 * (size_t)L);
    for (auto& e : ev) B200_CUDA(cudaEventCreate(&e));
    std::vector<DemonsWorkspace> wss(L);
    int rc = B200REG_OK;
    for (int l = 0; l < L && rc == B200REG_OK; ++l) {
        const b200reg_geom& gl = gfl[l];
        const size_t nl = nvox(gl);
        auto step = [&]() -> int {
            char lname[64];
            snprintf(lname, sizeof(lname), "level %d (%d x %d x %d)", l, gl.size[0], gl.size[1], gl.size[2]);
            B200_NVTX(lname);
            B200NvtxRange r_regrid("regrid total + warp moving");
            // :137 dvf_total -> level grid
            B200_TRY(next.alloc(ctx, 3 * nl * sizeof(double)));
            B200_TRY(resample_vec3(ctx, total.as<double>(), g_total, next.as<double>(), gl, nullptr, 0, 0.0));
            std::swap(total.p, next.p);
            g_total = gl;
            // :139-140 warp the level's moving image by the running total (default pixel 0)
            const b200reg_transform tfm = dvf_transform(total.as<double>(), gl);
            TempBuf mw;
            B200_TRY(mw.alloc(ctx, nvox(gml[l]) * sizeof(float)));
            {
                const void* ins[1] = { Ml[l].p };
                void* outs[1] = { mw.p };
                const int dt = B200REG_F32, ip = cfg->interp_order;
                const double dv = 0.0;
                B200_TRY(resample_batch_routed(ctx, 1, ins, &dt, gml[l], outs, gml[l], &tfm, 1, &ip, &dv));
            }
            // :143-149 Demons from a zero field
            b200reg_demons_params p = cfg->demons;
            p.number_of_iterations = cfg->iteration_staging[l];
            TempBuf iter;
            B200_TRY(iter.alloc(ctx, 3 * nl * sizeof(double)));
            B200_TRY(demons_prepare(ctx, gl, p.number_of_iterations, &wss[l], true));
            r_regrid.end();
            B200NvtxRange r_loop("Demons iterations");
            B200_CUDA(cudaEventRecord(ev[2 * l], ctx->stream));
            B200_TRY(demons_enqueue(ctx, Fl[l].as<float>(), gl, mw.as<float>(), gml[l], p, iter.as<double>(), &wss[l]));
            B200_CUDA(cudaEventRecord(ev[2 * l + 1], ctx->stream));
            r_loop.end();
            B200_NVTX("compose + recursive Gaussian");
            // :154 dvf_total + Resample(dvf_iter, tfm_total)
            B200_TRY(next.alloc(ctx, 3 * nl * sizeof(double)));
            B200_TRY(resample_vec3(ctx, iter.as<double>(), gl, next.as<double>(), gl, &tfm, 1, 0.0, total.as<double>()));
            std::swap(total.p, next.p);
            // :157-159 recursive Gaussian with the filter's (voxel-unit) sigmas passed as physical
            B200_TRY(recursive_gaussian_vec3(ctx, total.as<double>(), gl, p.std_dev));
            return B200REG_OK;
        };
        rc = step();
    }
    // :185 back onto the fixed grid
    if (rc == B200REG_OK) rc = resample_vec3(ctx, total.as<double>(), g_total, d_out_soa, gF, nullptr, 0, 0.0);
    if (rc == B200REG_OK && h_level_stats) {
        for (int l = 0; l < L && rc == B200REG_OK; ++l) {
            float ms = 0.f;
            if (cudaEventSynchronize(ev[2 * l + 1]) == cudaSuccess) cudaEventElapsedTime(&ms, ev[2 * l], ev[2 * l + 1]);
            rc = read_stats(ctx, wss[l], gfl[l], &h_level_stats[l], ms, cfg->iteration_staging[l], l);
        }
    }
    if (rc != B200REG_OK) cudaStreamSynchronize(ctx->stream);
    for (auto& e : ev) cudaEventDestroy(e);
    return rc;
}

// ---- fusion ----------------------------------------------------------------------------------------------------------------
API int b200reg_weight_map(b200reg_ctx* ctx, const float* d_target, const float* d_moving, const b200reg_geom* geom, int vote_type, double factor,
                           double sigma, double epsilon, float* d_weight)
{
    ENTER(ctx);
    REQUIRE(d_target && d_moving && d_weight && valid_geom(geom), "invalid argument");
    return weight_map(ctx, d_target, d_moving, *geom, vote_type, factor, sigma, epsilon, d_weight);
}
API int b200reg_weight_map_block(b200reg_ctx* ctx, const float* d_target, const float* d_moving, const b200reg_geom* geom, const int32_t radius[3],
                                 double factor, double gain, float* d_weight)
{
    ENTER(ctx);
    REQUIRE(d_target && d_moving && d_weight && radius && valid_geom(geom), "invalid argument");
    REQUIRE(radius[0] >= 0 && radius[1] >= 0 && radius[2] >= 0, "negative block radius");
    return weight_map_block(ctx, d_target, d_moving, *geom, radius, factor, gain, d_weight);
}
API int b200reg_normalise_by_max(b200reg_ctx* ctx, float* d_weight, const uint8_t* d_mask, size_t n)
{
    ENTER(ctx);
    REQUIRE(d_weight && n > 0, "invalid argument");
    return normalise_by_max(ctx, d_weight, d_mask, n);
}
API int b200reg_vote_accumulate(b200reg_ctx* ctx, const uint8_t* d_label, const float* d_weight, float* d_acc_num, float* d_acc_den, size_t n,
                                int first)
{
    ENTER(ctx);
    REQUIRE(d_label && d_weight && d_acc_num, "invalid argument");
    vote_accumulate_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_label, d_weight, d_acc_num, d_acc_den, n, first);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_vote_finalize(b200reg_ctx* ctx, float* d_num, const float* d_den, const b200reg_geom* geom, double smooth_variance,
                              double threshold, float* d_out)
{
    ENTER(ctx);
    REQUIRE(d_num && d_out && valid_geom(geom), "invalid argument");
    return vote_finalize(ctx, d_num, d_den, *geom, smooth_variance, threshold, d_out);
}

API int b200reg_binary_threshold(b200reg_ctx* ctx, const void* d_in, int dtype, size_t n, double lower, double upper, uint8_t* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out, "invalid argument");
    const int nb = ctx->sm_count * 8;
    if (semantics().binary_threshold_in_pixel_type && dtype != B200REG_F32 && dtype != B200REG_F64) {
        // semantic switch: the bounds are cast to the (integer) pixel type before the comparison, clamped to its range
        B200_DISPATCH_DTYPE(dtype, T, {
            lower = (double)(T)Px<T>::cast_host(lower);
            upper = (double)(T)Px<T>::cast_host(upper);
        });
    }
    B200_DISPATCH_DTYPE(dtype, T, { binary_threshold_kernel<T><<<nb, 256, 0, ctx->stream>>>((const T*)d_in, d_out, n, lower, upper); });
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

API int b200reg_pack_decision(b200reg_ctx* ctx, const uint8_t* d_label, int bit, int32_t* d_packed, size_t n, int first)
{
    ENTER(ctx);
    REQUIRE(d_label && d_packed && bit >= 0 && bit < 31, "invalid argument");
    pack_decision_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_label, bit, d_packed, n, first);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_unpack_decision(b200reg_ctx* ctx, const int32_t* d_packed, int bit, uint8_t* d_out, size_t n)
{
    ENTER(ctx);
    REQUIRE(d_packed && d_out && bit >= 0 && bit < 31, "invalid argument");
    unpack_decision_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_packed, bit, d_out, n);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// ---- compact exchange formats of the sharded fusion -------------------------------------------------------------------------
API int b200reg_pack_label(b200reg_ctx* ctx, const uint8_t* d_label, int bit, void* d_packed, int packed_dtype, size_t n, int first)
{
    ENTER(ctx);
    REQUIRE(d_label && d_packed && bit >= 0, "invalid argument");
    const int nb = ctx->sm_count * 8;
    switch (packed_dtype) {
    case B200REG_U8:
        REQUIRE(bit < 8, "bit %d does not fit a UInt8 decision mask", bit);
        pack_label_kernel<uint8_t><<<nb, 256, 0, ctx->stream>>>(d_label, bit, (uint8_t*)d_packed, n, first);
        break;
    case B200REG_U16: case B200REG_I16:
        REQUIRE(bit < 16, "bit %d does not fit a 16-bit decision mask", bit);
        pack_label_kernel<uint16_t><<<nb, 256, 0, ctx->stream>>>(d_label, bit, (uint16_t*)d_packed, n, first);
        break;
    case B200REG_U32: case B200REG_I32:
        REQUIRE(bit < 31, "bit %d does not fit a 32-bit decision mask (bit 31 is kept clear for signed sums)", bit);
        pack_label_kernel<uint32_t><<<nb, 256, 0, ctx->stream>>>(d_label, bit, (uint32_t*)d_packed, n, first);
        break;
    default: return set_error(B200REG_ERR_ARG, "decision masks are UInt8, UInt16 or UInt32 (got pixel type %d)", packed_dtype);
    }
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

template <typename TM>
static int staple_packed_t(b200reg_ctx* ctx, const TM* d_packed, uint32_t holder_mask, int n_raters, size_t n, double confidence_weight,
                           uint32_t max_iterations, double threshold, int rescale, double* d_out, double* h_pq, int32_t* h_elapsed)
{
    const int nbins = 1 << n_raters;
    TempBuf hist, tabw, tabo, state;
    B200_TRY(hist.alloc(ctx, (size_t)nbins * sizeof(unsigned long long)));
    B200_TRY(tabw.alloc(ctx, (size_t)nbins * sizeof(double)));
    B200_TRY(tabo.alloc(ctx, (size_t)nbins * sizeof(double)));
    B200_TRY(state.alloc(ctx, sizeof(StapleState)));
    B200_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)nbins * sizeof(unsigned long long), ctx->stream));
    const size_t shbytes = n_raters <= 12 ? (size_t)nbins * sizeof(unsigned int) : 0;
    const bool dense = holder_mask == (n_raters >= 32 ? 0xffffffffu : ((1u << n_raters) - 1u));  // holders are bits 0 .. n_raters-1
    const int nb = ctx->sm_count * 8;
    if (dense) staple_hist_mask_kernel<TM, true><<<nb, 256, shbytes, ctx->stream>>>(d_packed, holder_mask, n_raters, n, hist.as<unsigned long long>());
    else staple_hist_mask_kernel<TM, false><<<nb, 256, shbytes, ctx->stream>>>(d_packed, holder_mask, n_raters, n, hist.as<unsigned long long>());
    staple_em_table_kernel<<<1, 1024, 0, ctx->stream>>>(hist.as<unsigned long long>(), n_raters, confidence_weight, max_iterations, threshold, rescale,
                                                         tabw.as<double>(), tabo.as<double>(), state.as<StapleState>());
    if (dense) staple_write_mask_kernel<TM, true><<<nb, 256, 0, ctx->stream>>>(d_packed, holder_mask, tabo.as<double>(), d_out, n);
    else staple_write_mask_kernel<TM, false><<<nb, 256, 0, ctx->stream>>>(d_packed, holder_mask, tabo.as<double>(), d_out, n);
    ctx->launches += 3;
    B200_CHECK_LAUNCH();
    if (h_pq || h_elapsed) {
        // the statistics are a convenience: without them the call stays asynchronous
        StapleState* h_state = nullptr;
        B200_CUDA(cudaMallocHost(&h_state, sizeof(StapleState)));
        cudaError_t e = cudaMemcpyAsync(h_state, state.p, sizeof(StapleState), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) {
            if (h_pq)
                for (int j = 0; j < n_raters; ++j) {
                    h_pq[j] = h_state->p[j];
                    h_pq[n_raters + j] = h_state->q[j];
                }
            if (h_elapsed) *h_elapsed = h_state->elapsed;
        }
        cudaFreeHost(h_state);
        B200_CUDA(e);
    }
    return B200REG_OK;
}

API int b200reg_staple_packed(b200reg_ctx* ctx, const void* d_packed, int packed_dtype, uint32_t holder_mask, size_t n, double confidence_weight,
                              uint32_t max_iterations, double threshold, int rescale, double* d_out, double* h_pq, int32_t* h_elapsed)
{
    ENTER(ctx);
    REQUIRE(d_packed && d_out && n > 0, "invalid argument");
    const int n_raters = __builtin_popcount(holder_mask);
    REQUIRE(n_raters >= 1 && n_raters <= STAPLE_PATTERN_MAX, "number of raters %d not in [1, %d] for the packed STAPLE", n_raters, STAPLE_PATTERN_MAX);
    switch (packed_dtype) {
    case B200REG_U8:
        REQUIRE(holder_mask < 256u, "holder mask does not fit a UInt8 decision mask");
        return staple_packed_t<uint8_t>(ctx, (const uint8_t*)d_packed, holder_mask, n_raters, n, confidence_weight, max_iterations, threshold, rescale, d_out,
                                        h_pq, h_elapsed);
    case B200REG_U16: case B200REG_I16:
        REQUIRE(holder_mask < 65536u, "holder mask does not fit a 16-bit decision mask");
        return staple_packed_t<uint16_t>(ctx, (const uint16_t*)d_packed, holder_mask, n_raters, n, confidence_weight, max_iterations, threshold, rescale,
                                         d_out, h_pq, h_elapsed);
    case B200REG_U32: case B200REG_I32:
        return staple_packed_t<uint32_t>(ctx, (const uint32_t*)d_packed, holder_mask, n_raters, n, confidence_weight, max_iterations, threshold, rescale,
                                         d_out, h_pq, h_elapsed);
    default: return set_error(B200REG_ERR_ARG, "decision masks are UInt8, UInt16 or UInt32 (got pixel type %d)", packed_dtype);
    }
}

API int b200reg_count_accumulate(b200reg_ctx* ctx, const uint8_t* d_label, uint8_t* d_counts, size_t n, int first, int32_t* d_flag)
{
    ENTER(ctx);
    REQUIRE(d_label && d_counts && d_flag, "invalid argument");
    count_accumulate_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_label, d_counts, n, first, d_flag);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

API int b200reg_vote_finalize_counts(b200reg_ctx* ctx, const uint8_t* d_counts, int n_holders, const b200reg_geom* geom, double smooth_variance,
                                     double threshold, float* d_out)
{
    ENTER(ctx);
    REQUIRE(d_counts && d_out && valid_geom(geom) && n_holders >= 0, "invalid argument");
    const size_t n = nvox(*geom);
    TempBuf num;
    B200_TRY(num.alloc(ctx, n * sizeof(float)));
    counts_to_prob_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_counts, (float)n_holders, num.as<float>(), n);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return vote_finalize(ctx, num.as<float>(), nullptr, *geom, smooth_variance, threshold, d_out);
}

API int b200reg_staple(b200reg_ctx* ctx, const uint8_t* const* d_decisions, int n_raters, size_t n, double confidence_weight,
                       uint32_t max_iterations, double threshold, int rescale, double* d_out, double* h_pq, int32_t* h_elapsed)
{
    ENTER(ctx);
    REQUIRE(d_decisions && d_out && n > 0, "invalid argument");
    REQUIRE(n_raters >= 1 && n_raters <= STAPLE_MAX_RATERS, "number of raters %d not in [1, %d]", n_raters, STAPLE_MAX_RATERS);
    StaplePtrs ptrs;
    ptrs.n = n_raters;
    for (int j = 0; j < n_raters; ++j) {
        REQUIRE(d_decisions[j] != nullptr, "null decision volume");
        ptrs.d[j] = d_decisions[j];
    }
    if (n_raters <= STAPLE_PATTERN_MAX && !ctx->staple_voxelwise) {
        // pattern-histogram EM: two passes over the volume, the whole iteration loop inside one block
        const int nbins = 1 << n_raters;
        TempBuf pat, hist, tabw, tabo, state;
        B200_TRY(pat.alloc(ctx, n * sizeof(uint32_t)));
        B200_TRY(hist.alloc(ctx, (size_t)nbins * sizeof(unsigned long long)));
        B200_TRY(tabw.alloc(ctx, (size_t)nbins * sizeof(double)));
        B200_TRY(tabo.alloc(ctx, (size_t)nbins * sizeof(double)));
        B200_TRY(state.alloc(ctx, sizeof(StapleState)));
        B200_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)nbins * sizeof(unsigned long long), ctx->stream));
        const size_t shbytes = n_raters <= 12 ? (size_t)nbins * sizeof(unsigned int) : 0;
        staple_pattern_kernel<<<ctx->sm_count * 8, 256, shbytes, ctx->stream>>>(ptrs, pat.as<uint32_t>(), n, hist.as<unsigned long long>());
        staple_em_table_kernel<<<1, 1024, 0, ctx->stream>>>(hist.as<unsigned long long>(), n_raters, confidence_weight, max_iterations, threshold, rescale,
                                                             tabw.as<double>(), tabo.as<double>(), state.as<StapleState>());
        staple_write_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(pat.as<uint32_t>(), tabo.as<double>(), d_out, n);
        ctx->launches += 3;
        B200_CHECK_LAUNCH();
        StapleState* h_state = nullptr;
        B200_CUDA(cudaMallocHost(&h_state, sizeof(StapleState)));
        cudaError_t e = cudaMemcpyAsync(h_state, state.p, sizeof(StapleState), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) {
            if (h_pq)
                for (int j = 0; j < n_raters; ++j) {
                    h_pq[j] = h_state->p[j];
                    h_pq[n_raters + j] = h_state->q[j];
                }
            if (h_elapsed) *h_elapsed = h_state->elapsed;
        }
        cudaFreeHost(h_state);
        B200_CUDA(e);
        return B200REG_OK;
    }
    const int nb = ctx->sm_count * 4;
    const int stride = 2 * n_raters + 2;
    TempBuf partial, state;
    B200_TRY(partial.alloc(ctx, sizeof(double) * (size_t)nb * stride));
    B200_TRY(state.alloc(ctx, sizeof(StapleState)));
    StapleState* st = state.as<StapleState>();
    staple_init_kernel<<<nb, 256, 0, ctx->stream>>>(ptrs, d_out, n, partial.as<double>());
    staple_g_kernel<<<1, 32, 0, ctx->stream>>>(partial.as<double>(), nb, n, confidence_weight, st);
    ctx->launches += 2;
    StapleState* h_state = nullptr;
    B200_CUDA(cudaMallocHost(&h_state, sizeof(StapleState)));
    uint32_t iter = 0;
    int rc = B200REG_OK;
    bool done = false;
    while (!done && iter < max_iterations) {
        // enqueue a burst of EM iterations, then look at the convergence flag once
        const uint32_t burst = 8;
        for (uint32_t b = 0; b < burst && iter < max_iterations; ++b, ++iter) {
            staple_mstep_kernel<<<nb, 256, 0, ctx->stream>>>(ptrs, d_out, n, partial.as<double>(), st);
            staple_pq_kernel<<<1, 256, 0, ctx->stream>>>(partial.as<double>(), nb, n_raters, st);
            staple_estep_kernel<<<nb, 256, 0, ctx->stream>>>(ptrs, d_out, n, st);
            staple_converge_kernel<<<1, 1, 0, ctx->stream>>>(st, n_raters, iter);
            ctx->launches += 4;
        }
        if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(h_state, st, sizeof(StapleState), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            rc = set_error(B200REG_ERR_CUDA, "STAPLE iteration failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        done = h_state->converged != 0;
    }
    if (rc == B200REG_OK) {
        if (iter == 0) {
            cudaMemcpyAsync(h_state, st, sizeof(StapleState), cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
        }
        if (h_pq)
            for (int j = 0; j < n_raters; ++j) {
                h_pq[j] = h_state->p[j];
                h_pq[n_raters + j] = h_state->q[j];
            }
        if (h_elapsed) *h_elapsed = done ? h_state->elapsed : (int32_t)iter;
    }
    cudaFreeHost(h_state);
    B200_TRY(rc);
    if (rescale || threshold != 0.0) {
        TempBuf part, mm;
        B200_TRY(mm.alloc(ctx, 2 * sizeof(double)));
        B200_TRY(minmax_device<double>(ctx, d_out, n, mm.as<double>(), &part));
        rescale_threshold_kernel<double><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_out, d_out, n, mm.as<double>(), threshold, DBL_EPSILON, rescale);
        ctx->launches++;
        B200_CHECK_LAUNCH();
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return B200REG_OK;
}

// ---- a12: process_probability_image building blocks (fusion.py:295-328; multiatlas/run.py:423) -------------------------
API int b200reg_binary_fillhole(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], int fully_connected, uint8_t* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && size, "invalid argument");
    if (fully_connected) return set_error(B200REG_ERR_UNSUPPORTED, "BinaryFillhole: FullyConnected=True is not implemented (the reference uses the default, False)");
    return binary_fillhole(ctx, d_in, size, 1, d_out);
}

API int b200reg_largest_component(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], int fully_connected, uint8_t* d_out,
                                  int64_t* h_n_components, int64_t* h_voxels)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && size, "invalid argument");
    if (fully_connected) return set_error(B200REG_ERR_UNSUPPORTED, "ConnectedComponent: FullyConnected=True is not implemented (the reference uses the default, False)");
    TempBuf info;
    B200_TRY(info.alloc(ctx, 2 * sizeof(unsigned long long)));
    B200_TRY(largest_component(ctx, d_in, size, d_out, info.as<unsigned long long>()));
    if (h_n_components || h_voxels) {
        unsigned long long* h = reinterpret_cast<unsigned long long*>(ctx->h_scratch);
        B200_CUDA(small_d2h(ctx, h, info.p, 2 * sizeof(unsigned long long)));
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h_voxels) *h_voxels = (int64_t)(h[0] >> 32);
        if (h_n_components) *h_n_components = (int64_t)h[1];
    }
    return B200REG_OK;
}

API int b200reg_process_probability(b200reg_ctx* ctx, const void* d_prob, int dtype, const int32_t size[3], double threshold, uint8_t* d_out,
                                    int64_t* h_n_components)
{
    ENTER(ctx);
    REQUIRE(d_prob && d_out && size, "invalid argument");
    REQUIRE(dtype == B200REG_F32 || dtype == B200REG_F64, "process_probability_image: Float32 or Float64 probability image expected");
    B200_TRY(check_ccl_size(size));
    const size_t n = (size_t)size[0] * size[1] * size[2];
    TempBuf part, mm, info;
    B200_TRY(mm.alloc(ctx, 2 * sizeof(double)));
    B200_TRY(info.alloc(ctx, 2 * sizeof(unsigned long long)));
    const int nb = ctx->sm_count * 8;
    if (dtype == B200REG_F32) {
        B200_TRY(minmax_device<float>(ctx, (const float*)d_prob, n, mm.as<double>(), &part));
        normalise_threshold_kernel<float><<<nb, 256, 0, ctx->stream>>>((const float*)d_prob, mm.as<double>(), threshold, 255.0, d_out, n);
    } else {
        B200_TRY(minmax_device<double>(ctx, (const double*)d_prob, n, mm.as<double>(), &part));
        normalise_threshold_kernel<double><<<nb, 256, 0, ctx->stream>>>((const double*)d_prob, mm.as<double>(), threshold, 255.0, d_out, n);
    }
    ctx->launches++;
    B200_CHECK_LAUNCH();
    B200_TRY(binary_fillhole(ctx, d_out, size, 1, d_out));
    // no object: the reference returns the (empty) filled image; the selection below writes zeros as well
    B200_TRY(largest_component(ctx, d_out, size, d_out, info.as<unsigned long long>()));
    if (h_n_components) {
        unsigned long long* h = reinterpret_cast<unsigned long long*>(ctx->h_scratch);
        B200_CUDA(small_d2h(ctx, h, info.p, 2 * sizeof(unsigned long long)));
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        *h_n_components = (int64_t)h[1];
    }
    return B200REG_OK;
}

// ---- f1: linear_registration (linear.py:50-260), mean-squares metric + derivative accumulators -------------------------
API int b200reg_linreg_meansq(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                              const b200reg_geom* moving_geom, const double total_matrix[9], const double total_offset[3],
                              const double initial_matrix[9], const double center[3], const uint8_t* d_fixed_mask, const uint8_t* d_moving_mask,
                              int stride, double h_out[14])
{
    ENTER(ctx);
    REQUIRE(d_fixed && d_moving && valid_geom(fixed_geom) && valid_geom(moving_geom) && total_matrix && total_offset && initial_matrix && center && h_out,
            "invalid argument");
    REQUIRE(stride >= 1, "sampling stride must be >= 1");
    LinRegPose ps;
    for (int r = 0; r < 3; ++r) {
        ps.b[r] = total_offset[r];
        ps.c[r] = center[r];
        for (int c = 0; c < 3; ++c) {
            ps.A[r * 3 + c] = total_matrix[r * 3 + c];
            ps.Bt[r * 3 + c] = initial_matrix[c * 3 + r];
        }
    }
    return linreg_meansq(ctx, d_fixed, *fixed_geom, d_moving, *moving_geom, ps, d_fixed_mask, d_moving_mask, stride, h_out);
}

// ---- label utilities of the atlas pipeline (utils/crop.py:24-76, label/utils.py:23-58, multiatlas/run.py:387-437) -----------
API int b200reg_bounding_box(b200reg_ctx* ctx, const uint8_t* d_mask, const int32_t size[3], int32_t h_bbox[6])
{
    ENTER(ctx);
    REQUIRE(d_mask && size && h_bbox && size[0] > 0 && size[1] > 0 && size[2] > 0, "invalid argument");
    TempBuf bb;
    B200_TRY(bb.alloc(ctx, 6 * sizeof(int)));
    bbox_init_kernel<<<1, 32, 0, ctx->stream>>>(bb.as<int>());
    bbox_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_mask, size[0], size[1], size[2], bb.as<int>());
    ctx->launches += 2;
    B200_CHECK_LAUNCH();
    int* h = reinterpret_cast<int*>(ctx->h_scratch);
    B200_CUDA(small_d2h(ctx, h, bb.p, 6 * sizeof(int)));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int d = 0; d < 6; ++d) h_bbox[d] = h[d];
    return B200REG_OK;
}
API int b200reg_region_copy(b200reg_ctx* ctx, const void* d_src, const int32_t src_size[3], const int32_t src_index[3], void* d_dst,
                            const int32_t dst_size[3], const int32_t dst_index[3], const int32_t region_size[3], int dtype)
{
    ENTER(ctx);
    REQUIRE(d_src && d_dst && src_size && src_index && dst_size && dst_index && region_size, "invalid argument");
    const size_t elem = dtype_size(dtype);
    REQUIRE(elem > 0, "unsupported pixel type %d", dtype);
    return region_copy(ctx, d_src, src_size, src_index, d_dst, dst_size, dst_index, region_size, elem);
}
API int b200reg_resolve_overlap(b200reg_ctx* ctx, const uint8_t* const* d_labels_ranked, uint8_t* const* d_out, int n_labels, size_t n)
{
    ENTER(ctx);
    REQUIRE(d_labels_ranked && d_out && n > 0, "invalid argument");
    REQUIRE(n_labels >= 1 && n_labels <= OVERLAP_MAX, "number of structures %d not in [1, %d]", n_labels, OVERLAP_MAX);
    LabelPtrs lp;
    lp.n = n_labels;
    for (int s = 0; s < n_labels; ++s) {
        REQUIRE(d_labels_ranked[s] && d_out[s], "null label pointer");
        lp.in[s] = d_labels_ranked[s];
        lp.out[s] = d_out[s];
    }
    overlap_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(lp, n);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_binary_closing(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], const int32_t radius[3], const int32_t* h_offsets,
                               int n_offsets, uint8_t* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && size && radius && h_offsets && n_offsets >= 1, "invalid argument");
    REQUIRE(d_in != d_out, "in-place closing is not supported");
    for (int d = 0; d < 3; ++d) REQUIRE(size[d] > 0 && radius[d] >= 0, "invalid size / radius");
    return binary_closing(ctx, d_in, size, radius, h_offsets, n_offsets, d_out);
}

// ---- distance maps, contours, binary morphology, masking (registration/utils.py:270-344, label/projection.py:9-92) ----------
API int b200reg_signed_maurer_distance_map(b200reg_ctx* ctx, const uint8_t* d_mask, const b200reg_geom* geom, int inside_is_positive,
                                           int squared_distance, int use_image_spacing, float* d_out)
{
    ENTER(ctx);
    REQUIRE(d_mask && d_out && valid_geom(geom), "invalid argument");
    return signed_maurer(ctx, d_mask, *geom, inside_is_positive, squared_distance, use_image_spacing, d_out);
}
API int b200reg_label_contour(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], int fully_connected, uint8_t* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && size && size[0] > 0 && size[1] > 0 && size[2] > 0, "invalid argument");
    REQUIRE(d_in != d_out, "in-place LabelContour is not supported");
    return label_contour(ctx, d_in, size, fully_connected, d_out);
}
API int b200reg_binary_dilate(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], const int32_t* h_offsets, int n_offsets,
                              int boundary_to_foreground, uint8_t* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && size && h_offsets && n_offsets >= 1 && size[0] > 0 && size[1] > 0 && size[2] > 0, "invalid argument");
    REQUIRE(d_in != d_out, "in-place dilation is not supported");
    return binary_morph(ctx, true, d_in, size, h_offsets, n_offsets, boundary_to_foreground, d_out);
}
API int b200reg_binary_erode(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], const int32_t* h_offsets, int n_offsets,
                             int boundary_to_foreground, uint8_t* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && size && h_offsets && n_offsets >= 1 && size[0] > 0 && size[1] > 0 && size[2] > 0, "invalid argument");
    REQUIRE(d_in != d_out, "in-place erosion is not supported");
    return binary_morph(ctx, false, d_in, size, h_offsets, n_offsets, boundary_to_foreground, d_out);
}
API int b200reg_u8_binary_op(b200reg_ctx* ctx, const uint8_t* d_a, const uint8_t* d_b, int op, uint8_t* d_out, size_t n)
{
    ENTER(ctx);
    REQUIRE(d_a && d_b && d_out && n > 0, "invalid argument");
    REQUIRE(op >= B200REG_OP_OR && op <= B200REG_OP_XOR, "unknown operation %d", op);
    u8_binary_op_kernel<<<elementwise_blocks(ctx, n, 256), 256, 0, ctx->stream>>>(d_a, d_b, op, d_out, n);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_mask_image(b200reg_ctx* ctx, const void* d_in, int dtype, const uint8_t* d_mask, size_t n, int planes, double outside_value,
                           void* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_mask && d_out && n > 0 && planes >= 1, "invalid argument");
    const int nb = elementwise_blocks(ctx, n, 256);
    B200_DISPATCH_DTYPE(dtype, T, mask_image_kernel<T><<<nb, 256, 0, ctx->stream>>>((const T*)d_in, d_mask, n, planes, (T)Px<T>::cast_host(outside_value), (T*)d_out));
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_divide_scalar(b200reg_ctx* ctx, const void* d_in, int dtype, size_t n, double divisor, void* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && n > 0, "invalid argument");
    REQUIRE(dtype == B200REG_F32 || dtype == B200REG_F64, "divide_scalar: Float32 or Float64 image expected");
    const int nb = elementwise_blocks(ctx, n, 256);
    if (dtype == B200REG_F32)
        divide_scalar_kernel<float><<<nb, 256, 0, ctx->stream>>>((const float*)d_in, (float)divisor, (float*)d_out, n);
    else
        divide_scalar_kernel<double><<<nb, 256, 0, ctx->stream>>>((const double*)d_in, divisor, (double*)d_out, n);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// ---- field templates of the synthetic-deformation generators (generation/dvf.py:29-415) ---------------------------------------
API int b200reg_constant_field(b200reg_ctx* ctx, const uint8_t* d_mask, size_t n, const double vector[3], double* d_out_soa)
{
    ENTER(ctx);
    REQUIRE(d_out_soa && vector && n > 0, "invalid argument");
    constant_field_kernel<<<elementwise_blocks(ctx, n, 256), 256, 0, ctx->stream>>>(d_mask, n, vector[0], vector[1], vector[2], d_out_soa);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_radial_bend_field(b200reg_ctx* ctx, const uint8_t* d_mask, const int32_t size[3], const int32_t reference_index[3],
                                  const double axis[3], double scale, int clip_axis, int clip_keep_upper, double* d_out_soa)
{
    ENTER(ctx);
    REQUIRE(d_mask && d_out_soa && size && reference_index && axis && size[0] > 0 && size[1] > 0 && size[2] > 0, "invalid argument");
    REQUIRE(clip_axis >= -1 && clip_axis <= 2, "clip_axis must be -1 (none), 0 (x), 1 (y) or 2 (z)");
    const size_t n = (size_t)size[0] * size[1] * size[2];
    radial_bend_kernel<<<elementwise_blocks(ctx, n, 256), 256, 0, ctx->stream>>>(d_mask, size[0], size[1], size[2], reference_index[0], reference_index[1],
                                                                                reference_index[2], axis[0], axis[1], axis[2], scale, clip_axis,
                                                                                clip_keep_upper, d_out_soa);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// ---- compute_weight_map, vote_type "patch_correlation" (label/fusion.py:82-146) -------------------------------------------------
API int b200reg_patch_correlation(b200reg_ctx* ctx, const float* d_target, const float* d_moving, const int32_t size[3], const int32_t window[3],
                                  double* d_out)
{
    ENTER(ctx);
    REQUIRE(d_target && d_moving && d_out && size && window, "invalid argument");
    for (int d = 0; d < 3; ++d) REQUIRE(size[d] > 0 && window[d] >= 1, "invalid size / window");
    // scipy.stats.pearsonr: "x and y must have length at least 2" -- only a 1 x 1 x 1 window (or a one-voxel image) can get there
    const long long largest = (long long)(window[0] < size[0] ? window[0] : size[0]) * (window[1] < size[1] ? window[1] : size[1]) *
                              (window[2] < size[2] ? window[2] : size[2]);
    REQUIRE(largest >= 2, "x and y must have length at least 2.");
    const size_t n = (size_t)size[0] * size[1] * size[2];
    patch_correlation_kernel<<<elementwise_blocks(ctx, n, 128), 128, 0, ctx->stream>>>(d_target, d_moving, size[0], size[1], size[2], window[0], window[1],
                                                                                       window[2], d_out);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
API int b200reg_scale_shift(b200reg_ctx* ctx, const void* d_in, int dtype, size_t n, int take_abs, double mul, double add, void* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && n > 0, "invalid argument");
    REQUIRE(dtype == B200REG_F32 || dtype == B200REG_F64, "scale_shift: Float32 or Float64 image expected");
    const int nb = elementwise_blocks(ctx, n, 256);
    if (dtype == B200REG_F32)
        scale_shift_kernel<float><<<nb, 256, 0, ctx->stream>>>((const float*)d_in, n, take_abs, (float)mul, (float)add, (float*)d_out);
    else
        scale_shift_kernel<double><<<nb, 256, 0, ctx->stream>>>((const double*)d_in, n, take_abs, mul, add, (double*)d_out);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// ---- linear_registration, metric "correlation" (linear.py:141-146): the sums behind CorrelationImageToImageMetricv4 -------------
namespace {
CorrGeom make_corr_geom(const b200reg_geom& g)
{
    const GeomD d = make_geomd(g);
    CorrGeom c;
    c.nx = d.nx;
    c.ny = d.ny;
    c.nz = d.nz;
    for (int r = 0; r < 3; ++r) c.origin[r] = d.origin[r];
    for (int r = 0; r < 9; ++r) {
        c.i2p[r] = d.i2p[r];
        c.p2i[r] = d.p2i[r];
    }
    return c;
}
}  // namespace
API int b200reg_linreg_correlation(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                                   const b200reg_geom* moving_geom, const double total_matrix[9], const double total_offset[3],
                                   const double initial_matrix[9], const double center[3], const uint8_t* d_fixed_mask,
                                   const uint8_t* d_moving_mask, int stride, double h_out[42])
{
    ENTER(ctx);
    REQUIRE(d_fixed && d_moving && valid_geom(fixed_geom) && valid_geom(moving_geom) && total_matrix && total_offset && initial_matrix && center && h_out,
            "invalid argument");
    REQUIRE(stride >= 1, "sampling stride must be >= 1");
    CorrPose ps;
    for (int r = 0; r < 3; ++r) {
        ps.b[r] = total_offset[r];
        ps.c[r] = center[r];
        for (int c = 0; c < 3; ++c) {
            ps.A[r * 3 + c] = total_matrix[r * 3 + c];
            ps.Bt[r * 3 + c] = initial_matrix[c * 3 + r];
        }
    }
    const size_t n = nvox(*fixed_geom);
    const size_t nsamples = (n + (size_t)stride - 1) / (size_t)stride;
    int nb = (int)((nsamples + 127) / 128);
    if (nb > ctx->sm_count * 8) nb = ctx->sm_count * 8;
    if (nb < 1) nb = 1;
    TempBuf part, out;
    B200_TRY(part.alloc(ctx, sizeof(double) * LINREG_CORR_NV * (size_t)nb));
    B200_TRY(out.alloc(ctx, sizeof(double) * LINREG_CORR_NV));
    linreg_corr_kernel<<<nb, 128, 0, ctx->stream>>>(d_fixed, d_moving, d_fixed_mask, d_moving_mask, make_corr_geom(*fixed_geom), make_corr_geom(*moving_geom), ps,
                                                    stride, nsamples, part.as<double>());
    linreg_corr_final_kernel<<<1, 64, 0, ctx->stream>>>(part.as<double>(), nb, out.as<double>());
    ctx->launches += 2;
    B200_CHECK_LAUNCH();
    B200_CUDA(small_d2h(ctx, ctx->h_scratch, out.p, sizeof(double) * LINREG_CORR_NV));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int v = 0; v < LINREG_CORR_NV; ++v) h_out[v] = ctx->h_scratch[v];
    return B200REG_OK;
}

// ---- alignment_registration(moments=True) (linear.py:23-47): itk::ImageMomentsCalculator behind CenteredTransformInitializer ----
API int b200reg_image_moments(b200reg_ctx* ctx, const float* d_image, const b200reg_geom* geom, double h_out[4])
{
    ENTER(ctx);
    REQUIRE(d_image && valid_geom(geom) && h_out, "invalid argument");
    const GeomD d = make_geomd(*geom);
    MomentsGeom g;
    g.nx = d.nx;
    g.ny = d.ny;
    g.nz = d.nz;
    for (int r = 0; r < 3; ++r) g.origin[r] = d.origin[r];
    for (int r = 0; r < 9; ++r) g.i2p[r] = d.i2p[r];
    const size_t n = nvox(*geom);
    const int nb = elementwise_blocks(ctx, n, 256);
    TempBuf part, out;
    B200_TRY(part.alloc(ctx, sizeof(double) * MOMENTS_NV * (size_t)nb));
    B200_TRY(out.alloc(ctx, sizeof(double) * MOMENTS_NV));
    image_moments_kernel<<<nb, 256, 0, ctx->stream>>>(d_image, g, part.as<double>());
    image_moments_final_kernel<<<1, 32, 0, ctx->stream>>>(part.as<double>(), nb, out.as<double>());
    ctx->launches += 2;
    B200_CHECK_LAUNCH();
    B200_CUDA(small_d2h(ctx, ctx->h_scratch, out.p, sizeof(double) * MOMENTS_NV));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int v = 0; v < MOMENTS_NV; ++v) h_out[v] = ctx->h_scratch[v];
    return B200REG_OK;
}

// ---- linear_registration, metric "mattes_mi" (linear.py:145-146): MattesMutualInformationImageToImageMetricv4 in two passes ------
namespace {
int mattes_setup(const double total_matrix[9], const double total_offset[3], const double* initial_matrix, const double* center, int n_bins,
                 const double fixed_bins[2], const double moving_bins[2], CorrPose& ps, MattesBins& mb)
{
    REQUIRE(total_matrix && total_offset && fixed_bins && moving_bins, "invalid argument");
    REQUIRE(n_bins >= 5 && n_bins <= 256, "number of histogram bins %d not in [5, 256]", n_bins);
    REQUIRE(fixed_bins[0] > 0.0 && moving_bins[0] > 0.0, "bin sizes must be positive (constant image?)");
    for (int r = 0; r < 3; ++r) {
        ps.b[r] = total_offset[r];
        ps.c[r] = center ? center[r] : 0.0;
        for (int c = 0; c < 3; ++c) {
            ps.A[r * 3 + c] = total_matrix[r * 3 + c];
            ps.Bt[r * 3 + c] = initial_matrix ? initial_matrix[c * 3 + r] : (r == c ? 1.0 : 0.0);
        }
    }
    mb.n = n_bins;
    mb.fbin = fixed_bins[0];
    mb.fmin = fixed_bins[1];
    mb.mbin = moving_bins[0];
    mb.mmin = moving_bins[1];
    return B200REG_OK;
}
}  // namespace
API int b200reg_linreg_mattes_histogram(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                                        const b200reg_geom* moving_geom, const double total_matrix[9], const double total_offset[3],
                                        const uint8_t* d_fixed_mask, const uint8_t* d_moving_mask, int stride, int n_bins, const double fixed_bins[2],
                                        const double moving_bins[2], double* h_hist, double* h_count)
{
    ENTER(ctx);
    REQUIRE(d_fixed && d_moving && valid_geom(fixed_geom) && valid_geom(moving_geom) && h_hist && h_count, "invalid argument");
    REQUIRE(stride >= 1, "sampling stride must be >= 1");
    CorrPose ps;
    MattesBins mb;
    B200_TRY(mattes_setup(total_matrix, total_offset, nullptr, nullptr, n_bins, fixed_bins, moving_bins, ps, mb));
    const size_t n = nvox(*fixed_geom), nsamples = (n + (size_t)stride - 1) / (size_t)stride;
    const size_t cells = (size_t)n_bins * n_bins;
    constexpr int REPLICAS = 32;  // histogram copies the blocks spread their atomics over (summed below)
    TempBuf hist;
    B200_TRY(hist.alloc(ctx, sizeof(unsigned long long) * (REPLICAS * cells + 1)));
    B200_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned long long) * (REPLICAS * cells + 1), ctx->stream));
    linreg_mattes_hist_kernel<<<elementwise_blocks(ctx, nsamples, 256), 256, 0, ctx->stream>>>(
        d_fixed, d_moving, d_fixed_mask, d_moving_mask, make_corr_geom(*fixed_geom), make_corr_geom(*moving_geom), ps, mb, stride, nsamples, REPLICAS,
        hist.as<unsigned long long>(), hist.as<unsigned long long>() + REPLICAS * cells);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    std::vector<unsigned long long> host(REPLICAS * cells + 1);
    B200_CUDA(cudaMemcpyAsync(host.data(), hist.p, sizeof(unsigned long long) * (REPLICAS * cells + 1), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t q = 0; q < cells; ++q) {
        unsigned long long sum = 0;
        for (int r = 0; r < REPLICAS; ++r) sum += host[(size_t)r * cells + q];
        h_hist[q] = (double)sum / MATTES_FIXED_POINT;
    }
    *h_count = (double)host[REPLICAS * cells];
    return B200REG_OK;
}
API int b200reg_linreg_mattes_derivative(b200reg_ctx* ctx, const float* d_fixed, const b200reg_geom* fixed_geom, const float* d_moving,
                                         const b200reg_geom* moving_geom, const double total_matrix[9], const double total_offset[3],
                                         const double initial_matrix[9], const double center[3], const uint8_t* d_fixed_mask,
                                         const uint8_t* d_moving_mask, int stride, int n_bins, const double fixed_bins[2], const double moving_bins[2],
                                         const double* h_table, double h_out[12])
{
    ENTER(ctx);
    REQUIRE(d_fixed && d_moving && valid_geom(fixed_geom) && valid_geom(moving_geom) && initial_matrix && center && h_table && h_out, "invalid argument");
    REQUIRE(stride >= 1, "sampling stride must be >= 1");
    CorrPose ps;
    MattesBins mb;
    B200_TRY(mattes_setup(total_matrix, total_offset, initial_matrix, center, n_bins, fixed_bins, moving_bins, ps, mb));
    const size_t n = nvox(*fixed_geom), nsamples = (n + (size_t)stride - 1) / (size_t)stride;
    const size_t cells = (size_t)n_bins * n_bins;
    int nb = (int)((nsamples + 127) / 128);
    if (nb > ctx->sm_count * 8) nb = ctx->sm_count * 8;
    if (nb < 1) nb = 1;
    TempBuf table, part, out;
    B200_TRY(table.alloc(ctx, sizeof(double) * cells));
    B200_TRY(part.alloc(ctx, sizeof(double) * MATTES_NV * (size_t)nb));
    B200_TRY(out.alloc(ctx, sizeof(double) * MATTES_NV));
    B200_CUDA(cudaMemcpyAsync(table.p, h_table, sizeof(double) * cells, cudaMemcpyHostToDevice, ctx->stream));
    linreg_mattes_deriv_kernel<<<nb, 128, 0, ctx->stream>>>(d_fixed, d_moving, d_fixed_mask, d_moving_mask, make_corr_geom(*fixed_geom),
                                                            make_corr_geom(*moving_geom), ps, mb, stride, nsamples, table.as<double>(), part.as<double>());
    linreg_mattes_final_kernel<<<1, 32, 0, ctx->stream>>>(part.as<double>(), nb, out.as<double>());
    ctx->launches += 2;
    B200_CHECK_LAUNCH();
    B200_CUDA(small_d2h(ctx, ctx->h_scratch, out.p, sizeof(double) * MATTES_NV));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));  // also: h_table is caller memory
    for (int v = 0; v < MATTES_NV; ++v) h_out[v] = ctx->h_scratch[v];
    return B200REG_OK;
}

// ---- added path length (label/comparison.py:346-387): LabelContour of every axial slice on its own -------------------------------
API int b200reg_label_contour_slicewise(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], uint8_t* d_out)
{
    ENTER(ctx);
    REQUIRE(d_in && d_out && size && size[0] > 0 && size[1] > 0 && size[2] > 0, "invalid argument");
    REQUIRE(d_in != d_out, "in-place LabelContour is not supported");
    const size_t n = (size_t)size[0] * size[1] * size[2];
    label_contour_slicewise_kernel<<<elementwise_blocks(ctx, n, 256), 256, 0, ctx->stream>>>(d_in, size[0], size[1], size[2], d_out);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}
