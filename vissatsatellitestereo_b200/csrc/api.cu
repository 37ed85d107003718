// Context, error reporting and constant tables of libvissat_b200.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "geo_chain.cuh"
#include "vs_common.cuh"

static thread_local std::string g_last_error;

void vs_set_error(const std::string& msg) { g_last_error = msg; }

int vs_cuda_fail(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return VS_ERR_CUDA;
}

// PROJ 6.2 etmerc.cpp setup() for WGS84 (a = 6378137, rf = 298.257223563) and pymap3d's Ellipsoid('wgs84')
// (a = 6378137, f = 1/298.2572235630).  Host double arithmetic in the upstream order.
VsEllipsoidConsts vs_make_ellipsoid_consts() {
    VsEllipsoidConsts c;
    memset(&c, 0, sizeof(c));
    // ---- pymap3d
    c.a = 6378137.0;
    double f_pm = 1.0 / 298.2572235630;
    c.b = c.a * (1.0 - f_pm);
    c.a2 = c.a * c.a;
    c.b2 = c.b * c.b;
    c.E = sqrt(c.a2 - c.b2);
    c.E2 = c.E * c.E;  // upstream squares E again wherever it appears
    double boa = c.b / c.a;
    c.b_over_a_sq = boa * boa;
    c.a_over_b = c.a / c.b;
    c.dg2rad = M_PI / 180.0;
    c.rad2dg = 180.0 / M_PI;
    // ---- PROJ
    c.proj_a = 6378137.0;
    double f0 = 1.0 / 298.257223563;
    double es = 2 * f0 - f0 * f0;
    double f = es / (1 + sqrt(1 - es));
    double n = f / (2 - f);
    double np = n;
    c.cgb[0] = n * (2 + n * (-2 / 3.0 + n * (-2 + n * (116 / 45.0 + n * (26 / 45.0 + n * (-2854 / 675.0))))));
    c.cbg[0] = n * (-2 + n * (2 / 3.0 + n * (4 / 3.0 + n * (-82 / 45.0 + n * (32 / 45.0 + n * (4642 / 4725.0))))));
    np *= n;
    c.cgb[1] = np * (7 / 3.0 + n * (-8 / 5.0 + n * (-227 / 45.0 + n * (2704 / 315.0 + n * (2323 / 945.0)))));
    c.cbg[1] = np * (5 / 3.0 + n * (-16 / 15.0 + n * (-13 / 9.0 + n * (904 / 315.0 + n * (-1522 / 945.0)))));
    np *= n;
    c.cgb[2] = np * (56 / 15.0 + n * (-136 / 35.0 + n * (-1262 / 105.0 + n * (73814 / 2835.0))));
    c.cbg[2] = np * (-26 / 15.0 + n * (34 / 21.0 + n * (8 / 5.0 + n * (-12686 / 2835.0))));
    np *= n;
    c.cgb[3] = np * (4279 / 630.0 + n * (-332 / 35.0 + n * (-399572 / 14175.0)));
    c.cbg[3] = np * (1237 / 630.0 + n * (-12 / 5.0 + n * (-24832 / 14175.0)));
    np *= n;
    c.cgb[4] = np * (4174 / 315.0 + n * (-144838 / 6237.0));
    c.cbg[4] = np * (-734 / 315.0 + n * (109598 / 31185.0));
    np *= n;
    c.cgb[5] = np * (601676 / 22275.0);
    c.cbg[5] = np * (444337 / 155925.0);
    np = n * n;
    c.Qn = 0.9996 / (1 + n) * (1 + np * (1 / 4.0 + np * (1 / 64.0 + np / 256.0)));
    c.utg[0] = n * (-0.5 + n * (2 / 3.0 + n * (-37 / 96.0 + n * (1 / 360.0 + n * (81 / 512.0 + n * (-96199 / 604800.0))))));
    c.gtu[0] = n * (0.5 + n * (-2 / 3.0 + n * (5 / 16.0 + n * (41 / 180.0 + n * (-127 / 288.0 + n * (7891 / 37800.0))))));
    c.utg[1] = np * (-1 / 48.0 + n * (-1 / 15.0 + n * (437 / 1440.0 + n * (-46 / 105.0 + n * (1118711 / 3870720.0)))));
    c.gtu[1] = np * (13 / 48.0 + n * (-3 / 5.0 + n * (557 / 1440.0 + n * (281 / 630.0 + n * (-1983433 / 1935360.0)))));
    np *= n;
    c.utg[2] = np * (-17 / 480.0 + n * (37 / 840.0 + n * (209 / 4480.0 + n * (-5569 / 90720.0))));
    c.gtu[2] = np * (61 / 240.0 + n * (-103 / 140.0 + n * (15061 / 26880.0 + n * (167603 / 181440.0))));
    np *= n;
    c.utg[3] = np * (-4397 / 161280.0 + n * (11 / 504.0 + n * (830251 / 7257600.0)));
    c.gtu[3] = np * (49561 / 161280.0 + n * (-179 / 168.0 + n * (6601661 / 7257600.0)));
    np *= n;
    c.utg[4] = np * (-4583 / 161280.0 + n * (108847 / 3991680.0));
    c.gtu[4] = np * (34729 / 80640.0 + n * (-3418889 / 1995840.0));
    np *= n;
    c.utg[5] = np * (-20648693 / 638668800.0);
    c.gtu[5] = np * (212378941 / 319334400.0);
    // UTM has phi0 = 0: Z = gatg(cbg, 0) = 0 and Zb = -Qn*(Z + clens(gtu, 2Z)) = -0.0
    c.Zb = -c.Qn * 0.0;
    return c;
}

extern "C" {

int vs_abi_version(void) { return VS_ABI_VERSION; }

const char* vs_last_error(void) { return g_last_error.c_str(); }

int vs_device_count(int* out) {
    VS_REQUIRE(out != nullptr, "vs_device_count: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *out = 0;
        return vs_cuda_fail(e, "cudaGetDeviceCount");
    }
    *out = n;
    return VS_OK;
}

int vs_ctx_create(int device, vs_ctx** out) {
    VS_REQUIRE(out != nullptr, "vs_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    VS_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        vs_set_error("vs_ctx_create: no such CUDA device");
        return VS_ERR_INVALID;
    }
    VsDeviceGuard guard(device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaDeviceProp prop;
    VS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        vs_set_error(std::string("vs_ctx_create: device '") + prop.name +
                     "' is not sm_100-class; this library carries sm_100a code only");
        return VS_ERR_CUDA;
    }
    vs_ctx* ctx = new (std::nothrow) vs_ctx();
    if (!ctx) {
        vs_set_error("vs_ctx_create: out of host memory");
        return VS_ERR_INVALID;
    }
    memset(&ctx->aoi, 0, sizeof(ctx->aoi));
    memset(&ctx->geo, 0, sizeof(ctx->geo));
    memset(&ctx->poly, 0, sizeof(ctx->poly));
    memset(&ctx->fit, 0, sizeof(ctx->fit));
    ctx->device = device;
    ctx->aoi_set = false;
    ctx->ambiguity_eps = 1e-7;
    ctx->launches = 0;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->d_scratch = nullptr;
    ctx->scratch_doubles = 0;
    ctx->d_exact = nullptr;
    ctx->timing = false;
    ctx->xch_on = false;
    memset(&ctx->xch, 0, sizeof(ctx->xch));
    ctx->occ = nullptr;
    ctx->occ_words = 0;
    ctx->occ_stack_base = nullptr;
    ctx->occ_view0 = 0;
    ctx->d_fuse_plan = nullptr;
    ctx->fuse_plan_ints = 0;
    for (int i = 0; i < VS_MAX_STREAMS; ++i) ctx->d_touched[i] = nullptr;
    ctx->touched_tiles = 0;
    ctx->cur_touched = nullptr;
    for (int i = 0; i < VS_MAX_STREAMS; ++i) {
        ctx->side_stream[i] = nullptr;
        ctx->join_event[i] = nullptr;
        ctx->d_keygrid_extra[i] = nullptr;
    }
    ctx->fork_event = nullptr;
    ctx->keygrid_extra_cells = 0;
    for (int i = 0; i < VS_MAX_STREAMS; ++i) {
        ctx->d_keygrid_alt[i] = nullptr;
        ctx->prio_stream[i] = nullptr;
        ctx->prio_ev_a[i] = ctx->prio_ev_b[i] = ctx->prio_join[i] = nullptr;
    }
    {
        const char* pm = getenv("VISSAT_PRIO");
        const int v = pm ? atoi(pm) : 0;
        ctx->prio_mode = (v == 1 || v == 2) ? v : 0;
    }
    ctx->keygrid_alt_cells = 0;
    {
        const char* fc = getenv("VISSAT_FOLD_CLEAR");
        ctx->fold_clear = !(fc != nullptr && atoi(fc) == 0);
    }
    for (int i = 0; i < VS_MAX_STREAMS; ++i)
        for (int k = 0; k < 3; ++k) ctx->d_keygrid_ab[i][k] = nullptr;
    ctx->keygrid_ab_cells = 0;
    ctx->d_ab_counters = nullptr;
    ctx->ab_counters_ints = 0;
    {
        const char* ab = getenv("VISSAT_AB");
        ctx->ab_on = ab != nullptr && atoi(ab) != 0;   // measured slower than the separate kernels: opt-in
        const char* e = getenv("VISSAT_AB_STREAMS");
        const int n = e ? atoi(e) : 2;
        ctx->ab_streams = n < 1 ? 1 : (n > VS_MAX_STREAMS ? VS_MAX_STREAMS : n);
    }
    {
        const char* e1 = getenv("VISSAT_STREAMS");
        int n = e1 ? atoi(e1) : 4;
        ctx->n_streams = n < 1 ? 1 : (n > VS_MAX_STREAMS ? VS_MAX_STREAMS : n);
    }
    {
        // The TMA-fed persistent variant of stage B (finalize_tma.cu) is bit-identical but measured slower than
        // the plain-load kernels on B200 (41.7 vs 34.3 us per 2048^2 view: its in-place decode is an extra
        // shared-memory pass and the kernel is ALU-bound, not load-latency-bound), so it is opt-in.
        const char* e = getenv("VISSAT_TMA");
        ctx->no_tma = !(e != nullptr && e[0] == '1');
        const char* wa = getenv("VISSAT_K1_WARPAGG");
        ctx->k1_warp_agg = wa != nullptr && wa[0] == '1';
        const char* l = getenv("VISSAT_K2_LEGACY");
        const char* kk = getenv("VISSAT_K2_KEYS");
        ctx->k2_mode = (l != nullptr && l[0] == '1') ? 1 : ((kk != nullptr && kk[0] == '1') ? 2 : 0);
    }
    ctx->ev_used = 0;
    *out = ctx;
    return VS_OK;
}

int vs_ctx_destroy(vs_ctx* ctx) {
    if (!ctx) return VS_OK;
    VsDeviceGuard guard(ctx->device);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_exact) cudaFree(ctx->d_exact);
    if (ctx->d_fuse_plan) cudaFree(ctx->d_fuse_plan);
    for (int i = 0; i < VS_MAX_STREAMS; ++i)
        if (ctx->d_touched[i]) cudaFree(ctx->d_touched[i]);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < VS_MAX_STREAMS; ++i) {
        if (ctx->side_stream[i]) cudaStreamDestroy(ctx->side_stream[i]);
        if (ctx->join_event[i]) cudaEventDestroy(ctx->join_event[i]);
        if (ctx->d_keygrid_extra[i]) cudaFree(ctx->d_keygrid_extra[i]);
        if (ctx->d_keygrid_alt[i]) cudaFree(ctx->d_keygrid_alt[i]);
        if (ctx->prio_stream[i]) cudaStreamDestroy(ctx->prio_stream[i]);
        if (ctx->prio_ev_a[i]) cudaEventDestroy(ctx->prio_ev_a[i]);
        if (ctx->prio_ev_b[i]) cudaEventDestroy(ctx->prio_ev_b[i]);
        if (ctx->prio_join[i]) cudaEventDestroy(ctx->prio_join[i]);
    }
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    for (int i = 0; i < VS_MAX_STREAMS; ++i)
        for (int k = 0; k < 3; ++k)
            if (ctx->d_keygrid_ab[i][k]) cudaFree(ctx->d_keygrid_ab[i][k]);
    if (ctx->d_ab_counters) cudaFree(ctx->d_ab_counters);
    delete ctx;
    return VS_OK;
}

int vs_set_ambiguity_eps(vs_ctx* ctx, double eps_cells) {
    VS_REQUIRE(ctx != nullptr, "vs_set_ambiguity_eps: ctx is NULL");
    VS_REQUIRE(eps_cells >= 0 && eps_cells < 0.5, "vs_set_ambiguity_eps: eps must be in [0, 0.5)");
    ctx->ambiguity_eps = eps_cells;
    if (ctx->aoi_set) return vs_upload_exact_params(ctx);
    return VS_OK;
}

int vs_launch_count(vs_ctx* ctx, uint64_t* out) {
    VS_REQUIRE(ctx != nullptr && out != nullptr, "vs_launch_count: NULL argument");
    *out = ctx->launches;
    return VS_OK;
}

}  // extern "C"

int vs_ensure_side_streams(vs_ctx* ctx, int n) {
    if (n > VS_MAX_STREAMS) n = VS_MAX_STREAMS;
    for (int i = 0; i < n; ++i) {
        if (!ctx->side_stream[i]) {
            VS_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream[i], cudaStreamNonBlocking));
            VS_CUDA(cudaEventCreateWithFlags(&ctx->join_event[i], cudaEventDisableTiming));
        }
    }
    if (!ctx->fork_event) VS_CUDA(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    return VS_OK;
}

int vs_ensure_scratch(vs_ctx* ctx, size_t doubles) {
    if (ctx->scratch_doubles >= doubles) return VS_OK;
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr;
    ctx->scratch_doubles = 0;
    VS_CUDA(cudaMalloc(&ctx->d_scratch, doubles * sizeof(double)));
    ctx->scratch_doubles = doubles;
    return VS_OK;
}
