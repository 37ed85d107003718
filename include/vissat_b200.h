/*
 * vissat_b200.h — C ABI of libvissat_b200.so: the sm_100a implementation of VisSat's aggregate_2p5d hot path.
 *
 * The reference (Kai-46/VisSatSatelliteStereo) is pure Python and has no FFI layer; its boundary for this
 * path is the Python function surface of SURVEY.md §8(b).  This header is what a ctypes binding of that
 * surface calls.  Each entry point names the reference lines it replaces (paths relative to the reference
 * repository root).
 *
 * Conventions
 *   - plain C types only; every pointer marked "dev" is device memory owned by the caller
 *     (e.g. torch.Tensor.data_ptr()); "host" pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Kernels are enqueued on it and
 *     the call returns without synchronising unless stated.
 *   - return value: 0 = ok, non-zero = error; vs_last_error() returns a thread-local message.
 *     No C++ exception crosses the ABI.  There is no CPU fallback: without a CUDA device every compute
 *     entry point fails with VS_ERR_CUDA.
 *   - a vs_ctx belongs to one device and must not be used from two host threads at once.
 */
#ifndef VISSAT_B200_H
#define VISSAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_OK 0
#define VS_ERR_INVALID 1 /* bad argument */
#define VS_ERR_CUDA 2    /* CUDA runtime error (message has the cudaError string) */
#define VS_ERR_STATE 3   /* call order (e.g. rasterize before vs_set_aoi) */
#define VS_ERR_FIT 4     /* AOI polynomial could not be validated; exact mode is still available */

#define VS_ABI_VERSION 3

typedef struct vs_ctx vs_ctx;

/* AOI + grid description.
 * Built from aoi.json exactly as the reference does:
 *   lat0/lon0/alt0  coordinate_system.py:45-47   (bbox centre, alt_min)
 *   zone/south      lib/latlon_utm_converter.py:43-50 (the reference derives them from the first valid point
 *                   of every view; for an AOI inside one zone they equal aoi.json's zone_number/hemisphere)
 *   ul_e/ul_n       produce_dsm.py:51-52
 *   xsize/ysize     produce_dsm.py:54-55 (e_size, n_size)
 *   row_res/col_res lib/proj_to_grid.py:42-43: row = floor((ul_n - N)/row_res), col = floor((E - ul_e)/col_res).
 *                   NOTE the reference divides rows by its `xresolution` and columns by `yresolution`;
 *                   the caller passes them here already in that (swapped-name) order.
 *   alt_lo/alt_hi   altitude range (metres, absolute) over which the fast ENU->grid polynomial is fitted and
 *                   validated; points outside it take the exact per-point chain.  aoi.json alt_min/alt_max
 *                   widened by the caller is a good choice.
 */
typedef struct vs_aoi {
    double lat0, lon0, alt0;
    int32_t zone;
    int32_t south; /* 1 = southern hemisphere */
    double ul_e, ul_n;
    double row_res, col_res;
    int32_t xsize, ysize;
    double alt_lo, alt_hi;
} vs_aoi;

/* Diagnostics of the per-AOI polynomial that replaces pymap3d.enu2geodetic + PROJ utm inside the fused
 * rasteriser (validated against the exact device chain at vs_set_aoi time). */
typedef struct vs_fit_info {
    int32_t degree;          /* 3..5, or 0 = exact chain for every point */
    int32_t n_terms;
    int32_t mixed;           /* 0: all terms in float64; 1 or 2: only terms of total degree <= this are float64,
                                higher terms are evaluated in float32 (validated like the rest) */
    int32_t reserved;
    double max_err_cells;    /* max |poly - exact| of the fractional row/col on held-out points */
    double max_err_alt_m;    /* max |poly - exact| altitude (m) on held-out points */
    double box_center[3];    /* ENU box the polynomial covers */
    double box_half[3];
} vs_fit_info;

/* Per-call counters written by the rasterisers (device memory, 4 x uint64, zeroed by the call). */
#define VS_STAT_VALID 0     /* pixels/points with a finite position              (aggregate_2p5d_util.py:92) */
#define VS_STAT_INGRID 1    /* of those, inside the grid                          (lib/proj_to_grid.py:48)    */
#define VS_STAT_AMBIGUOUS 2 /* points whose polynomial row/col is within `ambiguity_eps` cells of a cell edge.  When
                               stats are requested they are re-evaluated with the exact chain and scattered from
                               there; what remains is the float64 chains' own ~1e-9 m noise against the CPU's      */
#define VS_STAT_EXACT 3     /* points that took the exact per-point chain (outside the fitted altitude range) */
#define VS_NUM_STATS 4

/* ---- context ------------------------------------------------------------------------------------------ */
int vs_abi_version(void);
const char* vs_last_error(void);
int vs_device_count(int* out);
int vs_ctx_create(int device, vs_ctx** out);
int vs_ctx_destroy(vs_ctx* ctx);

/* Fit + validate the ENU->(row, col, alt) polynomial for this AOI (synchronous, ~ms).
 * max_degree: 3..5; 0 forces the exact chain for every point.  info may be NULL. */
int vs_set_aoi(vs_ctx* ctx, const vs_aoi* aoi, int max_degree, vs_fit_info* info);
int vs_set_ambiguity_eps(vs_ctx* ctx, double eps_cells); /* default 1e-7 */
/* Diagnostics: evaluate the validated polynomial on n ENU points (HOST arrays; enu = n rows of e, n, u) with the
 * arithmetic K1 uses -> fractional column, fractional row (lib/proj_to_grid.py:42-43 before the floor) and altitude.
 * tests/test_geodesy_pin.py compares it with an independent 50-digit evaluation of the map. */
int vs_fit_eval(vs_ctx* ctx, const double* enu, int64_t n, double* colf, double* rowf, double* alt);

/* ---- stage A: depth map -> max-height key grid ---------------------------------------------------------
 * Replaces aggregate_2p5d_util.py:75-98 (NaN-ing, pixel grid, M*[col,row,1,depth], perspective divide,
 * local_to_global, latlon_to_eastnorh) fused with lib/proj_to_grid.py:42-61 (cell index, bounds mask,
 * per-cell nanmax).  `keygrid` holds ysize*xsize uint32 keys: 0 = empty, otherwise the order-preserving
 * image of the float32-rounded altitude (rounding is monotone, so max commutes with it).
 *   depth        dev float32 H*W, row-major (colmap/read_dense.py:36-51 layout); values <= 0 are invalid
 *   inv_proj_mat host 16 doubles, row-major 4x4 (one inv_proj_mats.txt row, aggregate_2p5d_util.py:54-61)
 *   keygrid      dev uint32 ysize*xsize; cleared by the call when clear_first != 0
 *   height_map   dev float32 H*W or NULL: ENU-up per pixel, NaN where invalid (aggregate_2p5d_util.py:91)
 *   stats        dev uint64[VS_NUM_STATS] or NULL
 */
int vs_unproject_rasterize(vs_ctx* ctx, const float* depth, int32_t H, int32_t W, const double* inv_proj_mat,
                           uint32_t* keygrid, int clear_first, float* height_map, uint64_t* stats, void* stream);

/* Reset `n_keys` 4-byte (key_bytes = 4) or 8-byte (key_bytes = 8) keys to "empty" (the NaN placeholders of
 * lib/proj_to_grid.py:53-55).  Same effect as clear_first on the rasterisers, as a separate enqueue. */
int vs_keygrid_clear(vs_ctx* ctx, void* keygrid, int64_t n_keys, int key_bytes, void* stream);

/* lib/proj_to_grid.py:42-61 for explicit points: dev float64 N*3 rows (E, N, alt) in the grid's UTM frame.
 * Exact float64 semantics: keygrid64 holds uint64 keys of the float64 altitude (0 = empty).
 * Grid geometry is passed explicitly (this is the public proj_to_grid signature, independent of vs_set_aoi). */
int vs_points_rasterize(vs_ctx* ctx, const double* points, int64_t n_points, double xoff, double yoff,
                        double xresolution, double yresolution, int32_t xsize, int32_t ysize,
                        uint64_t* keygrid64, int clear_first, uint64_t* stats, void* stream);

/* ---- stage B: key grid -> per-view DSM -------------------------------------------------------------------
 * Replaces the 3x3 NaN-hole fill of lib/proj_to_grid.py:65-79 (median of the non-NaN neighbours read from the
 * pre-fill grid, even count -> mean of the two middles) and produce_dsm.py:58 (astype(float32) +
 * cv2.medianBlur(.,3), including OpenCV's NaN behaviour: SIMD min/max semantics on columns 1..W-2 when
 * W >= simd_lanes+2, scalar semantics on the border columns; 1-D 3-tap median for single-row/column grids).
 *   dsm_out   dev float32 ysize*xsize  (what the reference writes to dsm_tif/<view>.tif, NaN = empty)
 *   nan_count dev uint64 or NULL: number of NaN cells in dsm_out (aggregate_2p5d.py:63 "empty ratio")
 */
int vs_grid_finalize(vs_ctx* ctx, const uint32_t* keygrid, int32_t xsize, int32_t ysize, float* dsm_out,
                     int simd_lanes, uint64_t* nan_count, void* stream);

/* float64 variant for the public proj_to_grid API: decode + hole fill, no blur; filled64 is float64 ysize*xsize.
 * If blurred32 != NULL also writes cv2.medianBlur(filled64.astype(float32), 3) (produce_dsm.py:58). */
int vs_grid_finalize64(vs_ctx* ctx, const uint64_t* keygrid64, int32_t xsize, int32_t ysize, double* filled64,
                       float* blurred32, int simd_lanes, void* stream);

/* ---- stages A + B for a batch of views in one call -------------------------------------------------------------
 * aggregate_2p5d_util.convert_depth_maps' loop body (:138-146 -> :75-102) for n_views views: for each view clear the
 * key grid, vs_unproject_rasterize, vs_grid_finalize into plane v of dsm_stack.  Same kernels as the single-view
 * entry points; exists so that a host language with expensive FFI calls (Python/ctypes) issues one call per batch.
 *   depth        host array of n_views device pointers (float32 H[v]*W[v] each)
 *   H, W         host arrays of n_views image sizes
 *   inv_proj_mats host n_views*16 doubles
 *   dsm_stack    dev float32; plane v at dsm_stack + v*plane_stride, ysize*xsize each
 *   nan_counts   dev uint64[n_views] or NULL (empty cells per view, aggregate_2p5d.py:63)
 *   stats        dev uint64[n_views*VS_NUM_STATS] or NULL
 */
int vs_views_to_dsm(vs_ctx* ctx, int32_t n_views, const float* const* depth, const int32_t* H, const int32_t* W,
                    const double* inv_proj_mats, uint32_t* keygrid, float* dsm_stack, int64_t plane_stride,
                    int simd_lanes, uint64_t* nan_counts, uint64_t* stats, void* stream);

/* Number of internal streams vs_views_to_dsm spreads consecutive views over (1..4, default 4 or $VISSAT_STREAMS).
 * With more than one, stage A of one view overlaps stage B of another (they are bound by different pipes). */
int vs_set_streams(vs_ctx* ctx, int n_streams);

/* Co-scheduled stages A+B (default on, or $VISSAT_AB=0): vs_views_to_dsm launches ONE kernel per view step whose CTAs
 * alternate between stage B of the previous view (lib/proj_to_grid.py:62-79 + produce_dsm.py:58) and stage A of the next
 * (aggregate_2p5d_util.py:75-98 + lib/proj_to_grid.py:42-61), so both instruction mixes share every SM.  Taken when the
 * call asks for no per-view counters and the AOI polynomial has degree 3; bit-identical to the separate kernels, which
 * enable = 0 forces (e.g. to time stage A and stage B on their own). */
int vs_set_coschedule(vs_ctx* ctx, int enable);

/* Per-kernel device timing of vs_views_to_dsm (CUDA events recorded on the launching stream around stage A and stage
 * B of every view).  vs_set_timing(ctx, 1) enables it and resets the log; vs_get_timing synchronises the events and
 * returns up to `max` (stage A ms, stage B ms) pairs in call order; *n_out = number of views logged. */
int vs_set_timing(vs_ctx* ctx, int enable);
int vs_get_timing(vs_ctx* ctx, int32_t max, float* stage_a_ms, float* stage_b_ms, int32_t* n_out);

/* ---- multi-GPU: stage B writes the row-band exchange itself (peer stores over NVLink) ----------------------------
 * views are sharded over ranks (one process per GPU), fusion is sharded by grid row bands, and the
 * transpose between the two layouts replaces the file-system hand-off of aggregate_2p5d.py:57-66.  Instead of a
 * collective after stage B, the stage-B kernel of vs_views_to_dsm stores every output row, besides the local per-view
 * DSM, straight into the (view, row-band) stack of the rank that fuses that row -- and into the 1-row halo of the
 * neighbouring bands for the final 3x3 blur -- through peer-mapped device memory.  The transfer overlaps the kernel
 * tile by tile; no pack/copy kernels, no second pass over the data.
 *
 * vs_peer_alloc / vs_peer_open: device memory that another process can map (cudaMalloc + CUDA IPC handle; the 64
 * handle bytes travel through the host language's own channel, e.g. torch.distributed.all_gather_object).
 * vs_peer_close unmaps a pointer obtained from vs_peer_open; vs_peer_free releases one from vs_peer_alloc.
 */
#define VS_MAX_RANKS 16
#define VS_IPC_HANDLE_BYTES 64
int vs_peer_alloc(vs_ctx* ctx, uint64_t bytes, void** dptr, uint8_t* handle64);
int vs_peer_open(vs_ctx* ctx, const uint8_t* handle64, void** dptr);
int vs_peer_close(vs_ctx* ctx, void* dptr);
int vs_peer_free(vs_ctx* ctx, void* dptr);

/* Row bands follow numpy.array_split(arange(ysize), n_ranks) (aggregate_2p5d_util.py:109-122 splits its work list
 * the same way): the first ysize % n_ranks bands have one row more.  band_stack[j] is rank j's float32 array
 * (n_views_total, rows_j + halo rows present, xsize): plane g holds the rows [max(r0_j - halo, 0),
 * min(r1_j + halo, ysize)) of global view g.  A plane of vs_views_to_dsm's dsm_stack at local_stack + i*plane_stride
 * is global view view0 + i.  The caller orders the ranks' use of the stacks (all stage-B kernels of a step complete
 * on every rank -- a stream-ordered barrier such as a 1-element all-reduce -- before any rank fuses its stack, and a
 * stack is not rewritten before its owner has fused it: alternate between two stacks). */
typedef struct vs_exchange {
    int32_t n_ranks;
    int32_t rank;
    int32_t halo;              /* rows of neighbouring bands kept on each side (1 for the 3x3 blur) */
    int32_t occ_words;         /* words per tile of the occupancy bitmaps below (>= ceil(n_views_total / 32)); 0 = none */
    int64_t view0;             /* global index of this rank's first view */
    int64_t n_views_total;
    const float* local_stack;  /* base of the local per-view DSM stack (plane 0) */
    float* band_stack[VS_MAX_RANKS];
    /* Sparse exchange (SURVEY.md 8(e) "skip all-NaN row-bands via an occupancy flag"), used when occ_words > 0:
     * occ[j] is rank j's occupancy bitmap (layout: see vs_set_occupancy), peer-mapped like band_stack[j].  Stage B
     * then sets the bit of (tile, view) in the bitmap of every rank whose band (+ halo) the tile touches, and does
     * NOT store all-empty tiles into the band stacks: the owner's vs_fuse_views_sparse never reads an unmarked tile. */
    uint32_t* occ[VS_MAX_RANKS];
} vs_exchange;
/* Enable (ex != NULL; the struct is copied) or disable (ex == NULL) the peer stores of vs_views_to_dsm. */
int vs_set_exchange(vs_ctx* ctx, const vs_exchange* ex);

/* ---- occupancy bitmap: which (tile, view) pairs hold data ----------------------------------------------------------
 * On large AOIs a view covers a fraction of the grid (BASELINE.json configs[2]: ~1/3), so most (tile, view) pairs of
 * the per-view DSM stack are all-NaN.  Stage B knows which: it already skips the arithmetic of tiles whose key box is
 * empty.  With a bitmap set, it records the others: occ is uint32 [ceil(ysize / VS_TILE_H)][ceil(xsize / VS_TILE_W)]
 * [occ_words]; bit (g % 32) of word (g / 32) of tile (ty, tx) is set when global view g may hold a non-NaN value in
 * grid rows [32 ty, 32 ty + 32) x columns [64 tx, 64 tx + 64).  The caller zeroes the bitmap before a pass over the
 * views; vs_fuse_views_sparse reads, per tile, only the marked planes.
 * Plane `dsm_stack + i * plane_stride` of vs_views_to_dsm is global view view0 + (that plane - stack_base) / plane_stride.
 * occ == NULL switches the marking off. */
#define VS_TILE_W 64
#define VS_TILE_H 32
int vs_set_occupancy(vs_ctx* ctx, uint32_t* occ, int32_t occ_words, const float* stack_base, int64_t view0);

/* ---- stage C: cross-view fusion --------------------------------------------------------------------------
 * Replaces aggregate_2p5d.py:65-78: per cell over V views (in the given order = sorted file order):
 * count filter (<= 2 measurements -> NaN), np.nanmedian, MAD = nanmedian(|x - med|), reject |x - med| > MAD,
 * np.nanmean of the survivors with numpy's float32 pairwise summation order.
 *   views        dev float32: plane v starts at views + v*plane_stride (elements); each plane is rows*W row-major
 *   out_mean     dev float32 rows*W
 * (The NaN<->nodata round trip of lib/dsm_util.py:69-72,128-131 is an identity for heights further than 0.1 m
 *  from -10000 and is applied by the host-side tif reader when fusing from files.)
 */
int vs_fuse_views(vs_ctx* ctx, const float* views, int64_t plane_stride, int32_t n_views, int32_t rows, int32_t W,
                  float* out_mean, void* stream);

/* The same fusion, reading only the planes the occupancy bitmap marks (an unmarked (tile, view) pair counts as NaN and
 * its memory is never touched).  `views` planes hold grid rows [row0, row0 + rows) of an xsize = W wide grid whose
 * bitmap `occ` (full-grid tile indexing as above, n_tile_cols = ceil(W / VS_TILE_W)) was filled by stage B.  The result
 * is bit-identical to vs_fuse_views on the densified stack: numpy's summation order depends on the original view
 * indices, which are kept. */
int vs_fuse_views_sparse(vs_ctx* ctx, const float* views, int64_t plane_stride, int32_t n_views, int32_t rows, int32_t W,
                         int32_t row0, const uint32_t* occ, int32_t occ_words, float* out_mean, void* stream);

/* aggregate_2p5d.py:81 (and the blur half of produce_dsm.py:58): cv2.medianBlur(float32, 3) on rows
 * [row_begin, row_end) of an image of H_total rows.  `in` points at image row `in_row0`; rows
 * max(row_begin-1,0)..min(row_end,H_total-1) must be present in it (row-band halo for multi-GPU fusion).
 * `out` points at output row row_begin.  nan_count as above (counts only the rows written). */
int vs_median3x3(vs_ctx* ctx, const float* in, int32_t in_row0, int32_t H_total, int32_t W, int32_t row_begin,
                 int32_t row_end, float* out, int simd_lanes, uint64_t* nan_count, void* stream);

/* ---- exact converters (public converter API, float64 chain, one thread per point) ------------------------
 * lib/latlonalt_enu_converter.py:42-45  (pymap3d.enu2geodetic)   */
int vs_enu_to_geodetic(vs_ctx* ctx, const double* e, const double* n, const double* u, int64_t count, double lat0,
                       double lon0, double alt0, double* lat, double* lon, double* alt, void* stream);
/* lib/latlonalt_enu_converter.py:36-39  (pymap3d.geodetic2enu)   */
int vs_geodetic_to_enu(vs_ctx* ctx, const double* lat, const double* lon, const double* alt, int64_t count,
                       double lat0, double lon0, double alt0, double* e, double* n, double* u, void* stream);
/* lib/latlon_utm_converter.py:50-51     (pyproj utm forward, PROJ etmerc order 6) */
int vs_geodetic_to_utm(vs_ctx* ctx, const double* lat, const double* lon, int64_t count, int32_t zone, int32_t south,
                       double* east, double* north, void* stream);
/* lib/latlon_utm_converter.py:61-62     (pyproj utm inverse)     */
int vs_utm_to_geodetic(vs_ctx* ctx, const double* east, const double* north, int64_t count, int32_t zone,
                       int32_t south, double* lat, double* lon, void* stream);
/* aggregate_2p5d_util.py:96-98 in one pass: ENU -> (E, N, alt) through the exact chain. */
int vs_enu_to_utm(vs_ctx* ctx, const double* e, const double* n, const double* u, int64_t count, double lat0,
                  double lon0, double alt0, int32_t zone, int32_t south, double* east, double* north, double* alt,
                  void* stream);

/* ---- introspection ---------------------------------------------------------------------------------------
 * Number of kernel launches issued by this library since the context was created (for bench.py's
 * gpu_launches claim). */
int vs_launch_count(vs_ctx* ctx, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* VISSAT_B200_H */
