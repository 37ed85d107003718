// Batched stage A + B: one C call for many views (the host language pays one FFI crossing per batch).
#include "vs_common.cuh"

bool vs_peer_plan_for(const vs_ctx* ctx, const float* plane, int64_t plane_stride, VsPeerPlan* plan);
int vs_grid_finalize_peer(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_lanes,
                          uint64_t* nan_count, const VsPeerPlan& plan, cudaStream_t stream, uint32_t* clear_other);
int vs_grid_finalize_impl(vs_ctx* ctx, const uint32_t* keygrid, int32_t xsize, int32_t ysize, float* dsm_out, int simd_lanes,
                          uint64_t* nan_count, void* stream_, uint32_t* clear_other);
int vs_grid_finalize_occ(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_lanes,
                         uint64_t* nan_count, const VsOccPlan& plan, cudaStream_t stream);
int vs_launch_clear_touched(vs_ctx* ctx, uint32_t* keygrid, unsigned char* touched, int xsize, int ysize, cudaStream_t stream);
bool vs_stage_ab_eligible(const vs_ctx* ctx, bool want_stats, bool sparse);
int vs_stage_ab_run(vs_ctx* ctx, int n_views, const float* const* depth, const int32_t* H, const int32_t* W,
                    const double* inv_proj_mats, float* dsm_stack, int64_t plane_stride, int simd_lanes,
                    uint64_t* nan_counts, int ns, cudaStream_t* streams);

// Co-scheduled path (stage_ab.cu): one kernel per view step does stage B of the previous view and stage A of the next.
static int views_to_dsm_ab(vs_ctx* ctx, int n_views, const float* const* depth, const int32_t* H, const int32_t* W,
                           const double* inv_proj_mats, float* dsm_stack, int64_t plane_stride, int simd_lanes,
                           uint64_t* nan_counts, cudaStream_t stream) {
    const size_t cells = (size_t)ctx->aoi.xsize * ctx->aoi.ysize;
    int ns = ctx->ab_streams < ctx->n_streams ? ctx->ab_streams : ctx->n_streams;
    if (ns > n_views) ns = n_views;
    int rc = vs_ensure_side_streams(ctx, ns);
    if (rc) return rc;
    if (ctx->keygrid_ab_cells < cells) {
        for (int i = 0; i < VS_MAX_STREAMS; ++i)
            for (int k = 0; k < 3; ++k) {
                if (ctx->d_keygrid_ab[i][k]) cudaFree(ctx->d_keygrid_ab[i][k]);
                ctx->d_keygrid_ab[i][k] = nullptr;
            }
        ctx->keygrid_ab_cells = 0;
    }
    for (int i = 0; i < ns; ++i)
        for (int k = 0; k < 3; ++k)
            if (!ctx->d_keygrid_ab[i][k]) VS_CUDA(cudaMalloc(&ctx->d_keygrid_ab[i][k], cells * sizeof(uint32_t)));
    ctx->keygrid_ab_cells = cells;
    const size_t need = 2 * (size_t)(n_views + ns);
    if (ctx->ab_counters_ints < need) {
        if (ctx->d_ab_counters) cudaFree(ctx->d_ab_counters);
        ctx->d_ab_counters = nullptr;
        ctx->ab_counters_ints = 0;
        VS_CUDA(cudaMalloc(&ctx->d_ab_counters, need * sizeof(int)));
        ctx->ab_counters_ints = need;
    }
    VS_CUDA(cudaMemsetAsync(ctx->d_ab_counters, 0, need * sizeof(int), stream));
    VS_CUDA(cudaEventRecord(ctx->fork_event, stream));
    cudaStream_t lanes[VS_MAX_STREAMS];
    for (int i = 0; i < ns; ++i) {
        lanes[i] = ctx->side_stream[i];
        VS_CUDA(cudaStreamWaitEvent(lanes[i], ctx->fork_event, 0));
    }
    rc = vs_stage_ab_run(ctx, n_views, depth, H, W, inv_proj_mats, dsm_stack, plane_stride, simd_lanes, nan_counts, ns, lanes);
    // join even after an error: the caller's stream stays ordered after everything enqueued here
    const std::string first_error = rc ? std::string(vs_last_error()) : std::string();
    for (int i = 0; i < ns; ++i) {
        if (cudaEventRecord(ctx->join_event[i], lanes[i]) != cudaSuccess ||
            cudaStreamWaitEvent(stream, ctx->join_event[i], 0) != cudaSuccess) {
            if (rc == VS_OK) rc = vs_cuda_fail(cudaGetLastError(), "vs_views_to_dsm: join");
        }
    }
    if (!first_error.empty()) vs_set_error(first_error);
    return rc;
}

extern "C" {

int vs_views_to_dsm(vs_ctx* ctx, int32_t n_views, const float* const* depth, const int32_t* H, const int32_t* W,
                    const double* inv_proj_mats, uint32_t* keygrid, float* dsm_stack, int64_t plane_stride,
                    int simd_lanes, uint64_t* nan_counts, uint64_t* stats, void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_views_to_dsm: NULL context");
    if (!ctx->aoi_set) {
        vs_set_error("vs_views_to_dsm: call vs_set_aoi first");
        return VS_ERR_STATE;
    }
    VS_REQUIRE(n_views >= 0, "vs_views_to_dsm: negative view count");
    if (n_views == 0) return VS_OK;
    VS_REQUIRE(depth && H && W && inv_proj_mats && keygrid && dsm_stack, "vs_views_to_dsm: NULL argument");
    const int xs = ctx->aoi.xsize, ys = ctx->aoi.ysize;
    VS_REQUIRE(plane_stride >= (int64_t)xs * ys, "vs_views_to_dsm: plane_stride smaller than a plane");
    cudaStream_t stream = (cudaStream_t)stream_;
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    {
        const bool sparse_mode = ctx->k2_mode != 1 && ctx->poly.degree > 0 &&
                                 (ctx->xch_on ? ctx->xch.occ_words > 0 : ctx->occ != nullptr);
        if (vs_stage_ab_eligible(ctx, stats != nullptr, sparse_mode))
            return views_to_dsm_ab(ctx, n_views, depth, H, W, inv_proj_mats, dsm_stack, plane_stride, simd_lanes, nan_counts,
                                   stream);
    }
    if (ctx->timing) {
        const size_t need = ctx->ev_used + 3 * (size_t)n_views;
        while (ctx->ev_pool.size() < need) {
            cudaEvent_t e;
            VS_CUDA(cudaEventCreate(&e));
            ctx->ev_pool.push_back(e);
        }
    }
    // Internal streams: view v runs on stream v % NS with its own key grid, so stage A (FP64/XU/issue-bound) of one
    // view overlaps stage B (ALU-bound) of the previous one.  Fork from / join into the caller's stream with events.
    const int NS = ctx->n_streams < n_views ? ctx->n_streams : n_views;
    const bool dual = NS >= 2;
    if (dual) {
        for (int i = 0; i < NS; ++i) {
            if (!ctx->side_stream[i]) {
                VS_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream[i], cudaStreamNonBlocking));
                VS_CUDA(cudaEventCreateWithFlags(&ctx->join_event[i], cudaEventDisableTiming));
            }
        }
        if (!ctx->fork_event) VS_CUDA(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
        if (ctx->keygrid_extra_cells < (size_t)xs * ys) {
            for (int i = 1; i < VS_MAX_STREAMS; ++i) {
                if (ctx->d_keygrid_extra[i]) cudaFree(ctx->d_keygrid_extra[i]);
                ctx->d_keygrid_extra[i] = nullptr;
            }
            ctx->keygrid_extra_cells = 0;
        }
        for (int i = 1; i < NS; ++i)
            if (!ctx->d_keygrid_extra[i]) VS_CUDA(cudaMalloc(&ctx->d_keygrid_extra[i], (size_t)xs * ys * sizeof(uint32_t)));
        ctx->keygrid_extra_cells = (size_t)xs * ys;
        VS_CUDA(cudaEventRecord(ctx->fork_event, stream));
        for (int i = 0; i < NS; ++i) VS_CUDA(cudaStreamWaitEvent(ctx->side_stream[i], ctx->fork_event, 0));
    }
    // Sparse mode (an occupancy bitmap is being written: large AOIs where a view covers part of the grid): stage A marks
    // the tiles it puts keys into, stage B looks only at tiles with a marked neighbour, and the key grid is cleared
    // tile by tile after stage B instead of by a whole-grid memset per view.  The key grids are zeroed once per call.
    const bool sparse = ctx->k2_mode != 1 && ctx->poly.degree > 0 &&
                        (ctx->xch_on ? ctx->xch.occ_words > 0 : ctx->occ != nullptr);
    const size_t n_tiles = (size_t)((xs + VS_TILE_W - 1) / VS_TILE_W) * ((ys + VS_TILE_H - 1) / VS_TILE_H);
    if (sparse) {
        if (ctx->touched_tiles < n_tiles) {
            for (int i = 0; i < VS_MAX_STREAMS; ++i) {
                if (ctx->d_touched[i]) cudaFree(ctx->d_touched[i]);
                ctx->d_touched[i] = nullptr;
            }
            ctx->touched_tiles = 0;
        }
        const int n_used = dual ? NS : 1;
        for (int i = 0; i < n_used; ++i)
            if (!ctx->d_touched[i]) VS_CUDA(cudaMalloc(&ctx->d_touched[i], n_tiles));
        ctx->touched_tiles = n_tiles;
        for (int i = 0; i < n_used; ++i) {
            cudaStream_t st = dual ? ctx->side_stream[i] : stream;
            VS_CUDA(cudaMemsetAsync(i ? ctx->d_keygrid_extra[i] : keygrid, 0, (size_t)xs * ys * sizeof(uint32_t), st));
            VS_CUDA(cudaMemsetAsync(ctx->d_touched[i], 0, n_tiles, st));
        }
    }
    // Dense mode: every internal stream alternates between TWO key grids.  Stage B of a view zeroes, tile by tile, the grid
    // the previous view of its stream used (k_grid_finalize: clear_other), which the next view's stage A then scatters
    // into -- instead of a whole-grid memset in front of every stage A (50 x 16.8 MB memset nodes per C2 step).  Both
    // grids of a stream are zeroed once per call.  (The key-space / TMA variants of stage B and the occupancy-marking launch
    // of the float-space kernel have no clear_other: those calls keep the memset per view.)
    const bool fold_clear = !sparse && ctx->occ == nullptr && !(ctx->xch_on && ctx->xch.occ_words > 0) && ctx->k2_mode != 2 &&
                            ctx->no_tma && ctx->fold_clear;
    if (fold_clear) {
        const int n_used = dual ? NS : 1;
        if (ctx->keygrid_alt_cells < (size_t)xs * ys) {
            for (int i = 0; i < VS_MAX_STREAMS; ++i) {
                if (ctx->d_keygrid_alt[i]) cudaFree(ctx->d_keygrid_alt[i]);
                ctx->d_keygrid_alt[i] = nullptr;
            }
            ctx->keygrid_alt_cells = 0;
        }
        for (int i = 0; i < n_used; ++i)
            if (!ctx->d_keygrid_alt[i]) VS_CUDA(cudaMalloc(&ctx->d_keygrid_alt[i], (size_t)xs * ys * sizeof(uint32_t)));
        ctx->keygrid_alt_cells = (size_t)xs * ys;
        for (int i = 0; i < n_used; ++i) {
            cudaStream_t st = dual ? ctx->side_stream[i] : stream;
            VS_CUDA(cudaMemsetAsync(i ? ctx->d_keygrid_extra[i] : keygrid, 0, (size_t)xs * ys * sizeof(uint32_t), st));
            VS_CUDA(cudaMemsetAsync(ctx->d_keygrid_alt[i], 0, (size_t)xs * ys * sizeof(uint32_t), st));
        }
    }
    // A/B switch VISSAT_PRIO: one of the two stages of every internal stream runs on a high-priority twin stream, ordered
    // with events, so that the block scheduler prefers that stage's CTAs whenever both kinds are pending.
    const bool prio = dual && !sparse && ctx->prio_mode != 0;
    if (prio) {
        int least = 0, greatest = 0;
        VS_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        for (int i = 0; i < NS; ++i) {
            if (!ctx->prio_stream[i]) {
                VS_CUDA(cudaStreamCreateWithPriority(&ctx->prio_stream[i], cudaStreamNonBlocking, greatest));
                VS_CUDA(cudaEventCreateWithFlags(&ctx->prio_ev_a[i], cudaEventDisableTiming));
                VS_CUDA(cudaEventCreateWithFlags(&ctx->prio_ev_b[i], cudaEventDisableTiming));
                VS_CUDA(cudaEventCreateWithFlags(&ctx->prio_join[i], cudaEventDisableTiming));
            }
            // the twin starts after everything enqueued on the internal stream so far (fork + the initial memsets)
            VS_CUDA(cudaEventRecord(ctx->prio_ev_b[i], ctx->side_stream[i]));
            VS_CUDA(cudaStreamWaitEvent(ctx->prio_stream[i], ctx->prio_ev_b[i], 0));
        }
    }
    // On a per-view error the loop stops, but the side streams are still joined into the caller's stream below: the
    // work already enqueued stays ordered before whatever the caller enqueues next (and a stream capture in progress
    // stays joinable).  Planes [0, v) of dsm_stack are then complete, plane v onwards is undefined.
    int rc = VS_OK;
    for (int v = 0; v < n_views && rc == VS_OK; ++v) {
        const int si = dual ? v % NS : 0;
        cudaStream_t st = dual ? ctx->side_stream[si] : stream;
        cudaStream_t st_b = st;                       // stream of stage B (st: stage A)
        if (prio) {
            if (ctx->prio_mode == 1) st_b = ctx->prio_stream[si];
            else st = ctx->prio_stream[si];
        }
        uint32_t* kg = si ? ctx->d_keygrid_extra[si] : keygrid;
        uint32_t* kg_other = nullptr;
        if (fold_clear) {
            const int turn = dual ? v / NS : v;       // how many views this stream has taken before
            kg_other = ctx->d_keygrid_alt[si];
            if (turn & 1) {
                uint32_t* t = kg;
                kg = kg_other;
                kg_other = t;
            }
        }
        ctx->cur_touched = sparse ? ctx->d_touched[si] : nullptr;
        if (!sparse && !fold_clear) {
            rc = vs_keygrid_clear(ctx, kg, (int64_t)xs * ys, 4, st);
            if (rc) break;
        }
        if (ctx->timing && cudaEventRecord(ctx->ev_pool[ctx->ev_used + 0], st) != cudaSuccess) { rc = vs_cuda_fail(cudaGetLastError(), "cudaEventRecord"); break; }
        rc = vs_unproject_rasterize(ctx, depth[v], H[v], W[v], inv_proj_mats + 16 * (size_t)v, kg, 0, nullptr,
                                    stats ? stats + (size_t)v * VS_NUM_STATS : nullptr, st);
        if (rc) break;
        if (ctx->timing && cudaEventRecord(ctx->ev_pool[ctx->ev_used + 1], st) != cudaSuccess) { rc = vs_cuda_fail(cudaGetLastError(), "cudaEventRecord"); break; }
        if (prio && (cudaEventRecord(ctx->prio_ev_a[si], st) != cudaSuccess ||
                     cudaStreamWaitEvent(st_b, ctx->prio_ev_a[si], 0) != cudaSuccess)) { rc = vs_cuda_fail(cudaGetLastError(), "vs_views_to_dsm: stage order"); break; }
        float* plane = dsm_stack + (size_t)v * plane_stride;
        if (ctx->xch_on) {   // stage B also stores each row into the band stacks of the ranks that fuse it
            VsPeerPlan plan;
            if (!vs_peer_plan_for(ctx, plane, plane_stride, &plan)) {
                vs_set_error("vs_views_to_dsm: dsm_stack plane is not a plane of vs_exchange.local_stack");
                rc = VS_ERR_INVALID;
                break;
            }
            rc = vs_grid_finalize_peer(ctx, kg, xs, ys, plane, simd_lanes, nan_counts ? nan_counts + v : nullptr, plan, st_b,
                                       plan.occ_words > 0 ? nullptr : kg_other);
        } else if (ctx->occ != nullptr) {   // record which tiles of this plane hold data (vs_set_occupancy)
            const ptrdiff_t d = plane - ctx->occ_stack_base;
            const int64_t g = ctx->occ_view0 + (plane_stride > 0 ? d / plane_stride : 0);
            if (d < 0 || d % plane_stride != 0 || g < 0 || g >= (int64_t)ctx->occ_words * 32) {
                vs_set_error("vs_views_to_dsm: dsm_stack plane is outside the stack given to vs_set_occupancy");
                rc = VS_ERR_INVALID;
                break;
            }
            VsOccPlan op;
            op.occ = ctx->occ;
            op.occ_words = ctx->occ_words;
            op.tiles_x = (xs + VS_TILE_W - 1) / VS_TILE_W;
            op.view = (int)g;
            rc = vs_grid_finalize_occ(ctx, kg, xs, ys, plane, simd_lanes, nan_counts ? nan_counts + v : nullptr, op, st_b);
        } else {
            rc = vs_grid_finalize_impl(ctx, kg, xs, ys, plane, simd_lanes, nan_counts ? nan_counts + v : nullptr, st_b, kg_other);
        }
        if (rc) break;
        if (sparse) {   // leave the key grid of this stream empty again: clear the touched tiles, reset their marks
            rc = vs_launch_clear_touched(ctx, kg, ctx->d_touched[si], xs, ys, st_b);
            if (rc) break;
        }
        if (ctx->timing) {
            if (cudaEventRecord(ctx->ev_pool[ctx->ev_used + 2], st_b) != cudaSuccess) { rc = vs_cuda_fail(cudaGetLastError(), "cudaEventRecord"); break; }
            ctx->ev_used += 3;
        }
        // the next stage A of this internal stream waits for this stage B (which also zeroes the grid it scatters into)
        if (prio && (cudaEventRecord(ctx->prio_ev_b[si], st_b) != cudaSuccess ||
                     cudaStreamWaitEvent(st, ctx->prio_ev_b[si], 0) != cudaSuccess)) { rc = vs_cuda_fail(cudaGetLastError(), "vs_views_to_dsm: stage order"); break; }
    }
    ctx->cur_touched = nullptr;
    if (dual) {
        const std::string first_error = rc ? std::string(vs_last_error()) : std::string();
        if (prio) {   // the twins join their internal streams first
            for (int i = 0; i < NS; ++i) {
                if (cudaEventRecord(ctx->prio_join[i], ctx->prio_stream[i]) != cudaSuccess ||
                    cudaStreamWaitEvent(ctx->side_stream[i], ctx->prio_join[i], 0) != cudaSuccess) {
                    if (rc == VS_OK) rc = vs_cuda_fail(cudaGetLastError(), "vs_views_to_dsm: join");
                }
            }
        }
        for (int i = 0; i < NS; ++i) {
            if (cudaEventRecord(ctx->join_event[i], ctx->side_stream[i]) != cudaSuccess ||
                cudaStreamWaitEvent(stream, ctx->join_event[i], 0) != cudaSuccess) {
                if (rc == VS_OK) rc = vs_cuda_fail(cudaGetLastError(), "vs_views_to_dsm: join");
            }
        }
        if (!first_error.empty()) vs_set_error(first_error);
    }
    return rc;
}

int vs_set_streams(vs_ctx* ctx, int n_streams) {
    VS_REQUIRE(ctx != nullptr, "vs_set_streams: NULL context");
    VS_REQUIRE(n_streams >= 1 && n_streams <= VS_MAX_STREAMS, "vs_set_streams: n_streams must be 1..4");
    ctx->n_streams = n_streams;
    return VS_OK;
}

int vs_set_coschedule(vs_ctx* ctx, int enable) {
    VS_REQUIRE(ctx != nullptr, "vs_set_coschedule: NULL context");
    ctx->ab_on = enable != 0;
    return VS_OK;
}

int vs_set_timing(vs_ctx* ctx, int enable) {
    VS_REQUIRE(ctx != nullptr, "vs_set_timing: NULL context");
    ctx->timing = enable != 0;
    ctx->ev_used = 0;
    return VS_OK;
}

int vs_get_timing(vs_ctx* ctx, int32_t max, float* stage_a_ms, float* stage_b_ms, int32_t* n_out) {
    VS_REQUIRE(ctx != nullptr && n_out != nullptr, "vs_get_timing: NULL argument");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    const int n = (int)(ctx->ev_used / 3);
    *n_out = n;
    for (int i = 0; i < n && i < max; ++i) {
        VS_CUDA(cudaEventSynchronize(ctx->ev_pool[3 * i + 2]));
        if (stage_a_ms) VS_CUDA(cudaEventElapsedTime(stage_a_ms + i, ctx->ev_pool[3 * i], ctx->ev_pool[3 * i + 1]));
        if (stage_b_ms) VS_CUDA(cudaEventElapsedTime(stage_b_ms + i, ctx->ev_pool[3 * i + 1], ctx->ev_pool[3 * i + 2]));
    }
    return VS_OK;
}

}  // extern "C"
