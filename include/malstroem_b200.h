/*
 * malstroem_b200.h — C ABI of the B200-native (sm_100a) raster hot path of SDFIdk/malstroem.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy types.  Every entry point
 * replaces one function of the reference's `malstroem.algorithms` layer (the layer its own native
 * plugin, malstroem/algorithms/speedups/__init__.py:37-78, rebinds); the reference file:line each one
 * stands in for is cited beside it.  The Python mirror in malstroem_b200/algorithms/ binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - rasters are C-contiguous row-major, `rows x cols`, rows*cols < 2^30 per device
 *   - functions without a suffix take HOST pointers and do H2D / compute / D2H themselves (blocking);
 *     `_dev` functions take DEVICE pointers plus a cudaStream_t (passed as void*) and only enqueue work,
 *     except where a count has to come back to the host (documented per function)
 *   - return value: 0 = ok, negative = error (MS_ERR_*); ms_last_error() gives the message
 *   - there is no CPU fallback anywhere: without a CUDA device every call fails with MS_ERR_CUDA
 */
#ifndef MALSTROEM_B200_H
#define MALSTROEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MS_OK 0
#define MS_ERR_CUDA (-1)   /* CUDA runtime / driver error, or no device */
#define MS_ERR_ARG (-2)    /* bad argument (null pointer, unsupported dtype code) */
#define MS_ERR_SHAPE (-3)  /* rows/cols too small (< 3) or raster too large for 30-bit cell indices */
#define MS_ERR_LABEL (-4)  /* a label lies outside [0, nlabels] (reference: IndexError / undefined) */
#define MS_ERR_NOCONV (-5) /* an iteration hit its safety cap (cyclic flow directions) */

/* dtype codes for `const void *data` arguments */
#define MS_F32 0
#define MS_F64 1
#define MS_U8 2
#define MS_I32 3
#define MS_I64 4

/* ---- library ---------------------------------------------------------------------------------- */
int ms_version(void);
int ms_init(int device);                 /* select device, create the memory pool; idempotent */
int ms_shutdown(void);
const char *ms_last_error(void);
int ms_device_count(void);
/* launches of this library's own kernels since the last reset (bench.py's `gpu_launches`) */
int64_t ms_kernel_launches(int reset);
/* per-kernel CUDA-event timing on the launching stream: ms_profile(n >= 1) starts recording one event pair per
 * launch (n > 1 pre-creates n pairs first), ms_profile(0) stops; ms_profile_report() synchronises the device and writes one line per kernel
 * ("<name> <launches> <total ms> <total units>") into buf, then clears the records */
int ms_profile(int enable);
/* host-side accounting since the last reset: out4 = {seconds waiting in stream syncs, number of syncs,
 * seconds inside pool allocations, number of allocations} */
int ms_host_counters(double *out4, int reset);
int ms_profile_report(char *buf, int64_t cap);
/* pinned host memory for the host-pointer entry points (optional; pageable memory also works) */
void *ms_host_alloc(int64_t bytes);
int ms_host_free(void *p);

/* ---- fill ------------------------------------------------------------------------------------- */
/* fill.fill_terrain(dtm) — malstroem/algorithms/fill.py:112-171 (sweeps: speedups/_fill.pyx:28-70).
 * `depths` (= filled - dtm, malstroem/dem.py:71) may be NULL. */
int ms_fill_terrain(const float *dtm, float *filled, float *depths, int64_t rows, int64_t cols);
int ms_fill_terrain_dev(const float *dtm, float *filled, float *depths, int64_t rows, int64_t cols,
                        void *stream);

/* the data pass of fill.minimum_safe_short_and_diag(dem) — fill.py:235-250: min and max of the raster
 * (short = 1024 ulp64(max|z|), diag = short * 2**0.5 are two scalar operations done by the caller) */
int ms_minmax_f32(const float *dem, int64_t n, float *out_min, float *out_max);
int ms_minmax_f32_dev(const float *dem, int64_t n, float *out_minmax_dev2, void *stream);

/* fill.fill_terrain_no_flats(dtm, short, diag) — fill.py:174-232 (speedups/_fill.pyx:72-124).
 * `filled` (the plain fill of the same dtm) may be NULL; then it is computed internally.
 * stats (may be NULL): [0] = relaxation rounds, [1] = tile visits, [2] = seed re-verification passes */
int ms_fill_terrain_no_flats(const float *dtm, double short_eps, double diag_eps, double *out,
                             int64_t rows, int64_t cols);
int ms_fill_terrain_no_flats_dev(const float *dtm, const float *filled, double short_eps, double diag_eps,
                                 double *out, int64_t rows, int64_t cols, int64_t *stats, void *stream);

/* ---- flow ------------------------------------------------------------------------------------- */
/* flow.terrain_flowdirection(terrain, edges_flow_outward) — flow.py:142-167; stencil
 * speedups/_flow.pyx:98-176; border rule flow.py:118-139 */
int ms_flowdir(const double *terrain, uint8_t *flowdir, int64_t rows, int64_t cols, int edges_flow_outward);
int ms_flowdir_dev(const double *terrain, uint8_t *flowdir, int64_t rows, int64_t cols,
                   int edges_flow_outward, void *stream);

/* flow.accumulated_flow(flowdir) — flow.py:344-364; speedups/_flow.pyx:225-273 */
int ms_accumulated_flow(const uint8_t *flowdir, double *accum, int64_t rows, int64_t cols);
int ms_accumulated_flow_dev(const uint8_t *flowdir, double *accum, int64_t rows, int64_t cols, void *stream);

/* flow.watersheds_from_labels(flowdir, labelled, unassigned) — flow.py:398-412;
 * speedups/_flow.pyx:276-403.  In place; label_bytes = 4 (int32) or 8 (int64). */
int ms_watersheds_from_labels(const uint8_t *flowdir, void *labelled, int label_bytes, int64_t rows,
                              int64_t cols, int64_t unassigned);
int ms_watersheds_from_labels_dev(const uint8_t *flowdir, void *labelled, int label_bytes, int64_t rows,
                                  int64_t cols, int64_t unassigned, void *stream);

/* ---- label ------------------------------------------------------------------------------------ */
/* label.connected_components(data) — label.py:19-40 (scipy.ndimage.label, 8-connectivity, components
 * numbered by first cell in row-major order).  Foreground = data != 0.  *nlabels comes back to the
 * host (the _dev form synchronises the stream for it). */
int ms_connected_components(const void *data, int dtype, int32_t *labels, int64_t rows, int64_t cols,
                            int64_t *nlabels);
int ms_connected_components_dev(const void *data, int dtype, int32_t *labels, int64_t rows, int64_t cols,
                                int64_t *nlabels, void *stream);

/* max (and min) label of an int32 raster: the `nlabels = np.max(labelled)` default of label.py:57,116,150 */
int ms_label_range(const int32_t *labels, int64_t n, int32_t *out_min, int32_t *out_max);
int ms_label_range_dev(const int32_t *labels, int64_t n, int32_t *out_minmax_dev2, void *stream);

/* label.label_stats(data, labelled, nlabels) — label.py:43-75; speedups/_label.pyx:31-97.
 * data dtype MS_F32 or MS_F64; tables have nlabels+1 entries. */
int ms_label_stats(const void *data, int dtype, const int32_t *labels, int64_t n, int64_t nlabels,
                   double *out_min, double *out_max, double *out_sum, int64_t *out_count);
int ms_label_stats_dev(const void *data, int dtype, const int32_t *labels, int64_t n, int64_t nlabels,
                       double *out_min, double *out_max, double *out_sum, int64_t *out_count, void *stream);

/* label.label_min_index / label_max_index(data, labelled, nlabels) — label.py:101-166;
 * speedups/_label.pyx:99-164.  data is float64; first cell in raster order wins ties; labels never
 * seen keep value = +inf / -inf and row = col = -1. */
int ms_label_extreme_index(const double *data, const int32_t *labels, int64_t rows, int64_t cols,
                           int64_t nlabels, int want_max, double *out_value, int64_t *out_row,
                           int64_t *out_col);
int ms_label_extreme_index_dev(const double *data, const int32_t *labels, int64_t rows, int64_t cols,
                               int64_t nlabels, int want_max, double *out_value, int64_t *out_row,
                               int64_t *out_col, void *stream);

/* label.label_count(labelled) — label.py:169-180 (np.bincount); nbins = max label + 1 */
int ms_label_count(const int32_t *labels, int64_t n, int64_t nbins, int64_t *out_count);
int ms_label_count_dev(const int32_t *labels, int64_t n, int64_t nbins, int64_t *out_count, void *stream);

/* label.keep_labels(labelled, keep_label, background) — label.py:78-98: out[i] = keep[labels[i]] */
int ms_keep_labels(const int32_t *labels, int64_t n, const uint8_t *keep, int64_t nkeep, uint8_t *out);
int ms_keep_labels_dev(const int32_t *labels, int64_t n, const uint8_t *keep, int64_t nkeep, uint8_t *out,
                       void *stream);

/* ---- the whole path, device resident (what bench.py times) -------------------------------------- */
typedef struct ms_rasters {
    int64_t rows, cols;
    const float *dem;   /* in  : float32 DEM                                   (dem.py:56)            */
    float *filled;      /* out : fill_terrain                                   (dem.py:67)            */
    float *depths;      /* out : filled - dem                                   (dem.py:71)            */
    double *fnf;        /* out : fill_terrain_no_flats                          (dem.py:80)            */
    uint8_t *flowdir;   /* out : terrain_flowdirection                          (dem.py:83)            */
    double *accum;      /* out : accumulated_flow                               (dem.py:89)  may be NULL */
    int32_t *labels;    /* out : connected_components(depths)                   (bluespots.py:159)     */
    int32_t *wsheds;    /* out : watersheds_from_labels                         (bluespots.py:183-185) */
    /* per-label tables, device memory with room for table_capacity entries (>= nlabels + 1) */
    int64_t table_capacity;
    double *st_min, *st_max, *st_sum; /* label_stats(depths, labels)            (bluespots.py:160)     */
    int64_t *st_count;
    int64_t *ws_count;                /* label_count(wsheds)                    (bluespots.py:186)     */
    double *ppmin_value;              /* label_min_index(fnf, labels)           (bluespots.py:205)     */
    int64_t *ppmin_row, *ppmin_col;
    double *ppmax_value;              /* label_max_index(accum, labels)         (bluespots.py:198)  may be NULL */
    int64_t *ppmax_row, *ppmax_col;
    /* results that come back to the host */
    int64_t nlabels;
    double short_eps, diag_eps;
    int64_t stats[8]; /* [0] Boruvka rounds, [1] catchments, [2] no-flats rounds, [3] tile visits,
                         [4] seed re-verifications, [5] pointer-jump rounds (fill), [6] (wsheds) */
} ms_rasters;

/* DEM in HBM -> every raster and table of the path in HBM.  Synchronises the stream a few times for
 * convergence flags and counts. */
int ms_pipeline_dev(ms_rasters *io, void *stream);

/* synthetic fractal DEM (malstroem_b200/synth.py, bit-identical), generated in place on the device */
int ms_synth_fractal_dev(float *dem, int64_t rows, int64_t cols, int64_t row0, int64_t col0, int seed,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif
