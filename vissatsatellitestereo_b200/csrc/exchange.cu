// Peer-mapped device memory and the description of the stage-B peer stores (multi-GPU row-band exchange).
// The stores themselves are in finalize.cu (PeerSink); vs_views_to_dsm (pipeline.cu) builds one VsPeerPlan per view.
#include <string.h>

#include "vs_common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == VS_IPC_HANDLE_BYTES, "CUDA IPC handle size");

// numpy.array_split boundaries: the first H % n bands have one row more
void vs_band_rows(int H, int n, int* row0 /* n + 1 */) {
    const int q = H / n, r = H % n;
    row0[0] = 0;
    for (int j = 0; j < n; ++j) row0[j + 1] = row0[j] + q + (j < r ? 1 : 0);
}

// Plan for one output plane of vs_views_to_dsm (global view g).  Returns false if the plane is not one of local_stack's.
bool vs_peer_plan_for(const vs_ctx* ctx, const float* plane, int64_t plane_stride, VsPeerPlan* plan) {
    const vs_exchange& x = ctx->xch;
    const int H = ctx->aoi.ysize, W = ctx->aoi.xsize;
    const ptrdiff_t d = plane - x.local_stack;
    if (d < 0 || plane_stride <= 0 || d % plane_stride != 0) return false;
    const int64_t g = x.view0 + d / plane_stride;
    if (g < 0 || g >= x.n_views_total) return false;
    plan->n = x.n_ranks;
    plan->halo = x.halo;
    plan->inv = (unsigned)((((unsigned long long)x.n_ranks) << 32) / (unsigned long long)H);
    vs_band_rows(H, x.n_ranks, plan->row0);
    for (int j = 0; j < VS_MAX_RANKS; ++j) {
        plan->plane[j] = nullptr;
        plan->occ[j] = nullptr;
    }
    plan->occ_words = x.occ_words;
    plan->tiles_x = (W + VS_TILE_W - 1) / VS_TILE_W;
    plan->view = (int)g;
    for (int j = 0; j < x.n_ranks; ++j) {
        const int r0 = plan->row0[j], r1 = plan->row0[j + 1];
        if (r1 == r0) continue;
        const int h0 = r0 - x.halo > 0 ? r0 - x.halo : 0;
        const int h1 = r1 + x.halo < H ? r1 + x.halo : H;
        plan->plane[j] = x.band_stack[j] + (size_t)g * (size_t)(h1 - h0) * (size_t)W;
        if (x.occ_words > 0) plan->occ[j] = x.occ[j];
    }
    return true;
}

extern "C" {

int vs_peer_alloc(vs_ctx* ctx, uint64_t bytes, void** dptr, uint8_t* handle64) {
    VS_REQUIRE(ctx != nullptr && dptr != nullptr && handle64 != nullptr, "vs_peer_alloc: NULL argument");
    VS_REQUIRE(bytes > 0, "vs_peer_alloc: size must be positive");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    *dptr = nullptr;
    void* p = nullptr;
    VS_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return vs_cuda_fail(e, "cudaIpcGetMemHandle");
    }
    memcpy(handle64, &h, sizeof(h));
    *dptr = p;
    return VS_OK;
}

int vs_peer_open(vs_ctx* ctx, const uint8_t* handle64, void** dptr) {
    VS_REQUIRE(ctx != nullptr && dptr != nullptr && handle64 != nullptr, "vs_peer_open: NULL argument");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    *dptr = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    VS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dptr = p;
    return VS_OK;
}

int vs_peer_close(vs_ctx* ctx, void* dptr) {
    VS_REQUIRE(ctx != nullptr, "vs_peer_close: NULL context");
    if (dptr == nullptr) return VS_OK;
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    VS_CUDA(cudaIpcCloseMemHandle(dptr));
    return VS_OK;
}

int vs_peer_free(vs_ctx* ctx, void* dptr) {
    VS_REQUIRE(ctx != nullptr, "vs_peer_free: NULL context");
    if (dptr == nullptr) return VS_OK;
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    VS_CUDA(cudaFree(dptr));
    return VS_OK;
}

int vs_set_exchange(vs_ctx* ctx, const vs_exchange* ex) {
    VS_REQUIRE(ctx != nullptr, "vs_set_exchange: NULL context");
    if (ex == nullptr) {
        ctx->xch_on = false;
        return VS_OK;
    }
    if (!ctx->aoi_set) {
        vs_set_error("vs_set_exchange: call vs_set_aoi first");
        return VS_ERR_STATE;
    }
    VS_REQUIRE(ex->n_ranks >= 1 && ex->n_ranks <= VS_MAX_RANKS, "vs_set_exchange: n_ranks must be 1..16");
    VS_REQUIRE(ex->rank >= 0 && ex->rank < ex->n_ranks, "vs_set_exchange: rank out of range");
    VS_REQUIRE(ex->halo >= 0 && ex->halo <= 1, "vs_set_exchange: halo must be 0 or 1");
    VS_REQUIRE(ex->view0 >= 0 && ex->n_views_total > 0 && ex->view0 < ex->n_views_total,
               "vs_set_exchange: bad view range");
    VS_REQUIRE(ex->local_stack != nullptr, "vs_set_exchange: NULL local_stack");
    int row0[VS_MAX_RANKS + 1];
    vs_band_rows(ctx->aoi.ysize, ex->n_ranks, row0);
    for (int j = 0; j < ex->n_ranks; ++j)
        VS_REQUIRE(row0[j + 1] == row0[j] || ex->band_stack[j] != nullptr, "vs_set_exchange: NULL band_stack of a rank with rows");
    VS_REQUIRE(ex->occ_words >= 0, "vs_set_exchange: negative occ_words");
    if (ex->occ_words > 0) {
        VS_REQUIRE((int64_t)ex->occ_words * 32 >= ex->n_views_total, "vs_set_exchange: occ_words too small for n_views_total");
        for (int j = 0; j < ex->n_ranks; ++j)
            VS_REQUIRE(row0[j + 1] == row0[j] || ex->occ[j] != nullptr, "vs_set_exchange: NULL occupancy bitmap of a rank with rows");
    }
    ctx->xch = *ex;
    ctx->xch_on = true;
    return VS_OK;
}

int vs_set_occupancy(vs_ctx* ctx, uint32_t* occ, int32_t occ_words, const float* stack_base, int64_t view0) {
    VS_REQUIRE(ctx != nullptr, "vs_set_occupancy: NULL context");
    if (occ == nullptr) {
        ctx->occ = nullptr;
        ctx->occ_words = 0;
        return VS_OK;
    }
    VS_REQUIRE(occ_words > 0 && stack_base != nullptr && view0 >= 0, "vs_set_occupancy: bad argument");
    ctx->occ = occ;
    ctx->occ_words = occ_words;
    ctx->occ_stack_base = stack_base;
    ctx->occ_view0 = view0;
    return VS_OK;
}

}  // extern "C"
