/*
 * tt_irt1.h -- C-ABI of the B200-native linear-spline inverse Rosenblatt transform.
 *
 * Two shared libraries are built from the same sources (tt-irt_b200/Makefile):
 *   tt-irt_b200/tt_irt_py/tt_irt1_int32.so   TTIRT_INT = int        (Python ctypes caller)
 *   tt-irt_b200/lib/libtt_irt1_int64.so      TTIRT_INT = long long  (Matlab MEX caller)
 * Both are self-contained (static cudart, no BLAS) and export
 *   (1) the reference's entry point `tt_irt1` with its exact signature, and
 *   (2) the width-independent extended entry points `ttirt_*` declared below
 *       (device-resident sampling, cached models).
 *
 * All pointers are plain host or device addresses; no torch / C++ types cross this boundary.
 * There is NO CPU fallback: without a usable CUDA device every entry point fails loudly
 * (message on stderr, outputs NaN-filled, non-zero status where a status is returned).
 */
#ifndef TT_IRT1_H
#define TT_IRT1_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TTIRT_API __attribute__((visibility("default")))
#else
#define TTIRT_API
#endif

#ifndef TTIRT_INT
#define TTIRT_INT int /* the reference's `lapackint`: tt_irt1_int32.c:20 (int) / tt_irt1_int64.c:20 (long long) */
#endif

/*
 * Drop-in replacement for the reference routine
 *   python/tt_irt_py/tt_irt1_int32.c:34   void tt_irt1(lapackint d, lapackint *n, double *xs, lapackint *ttrank,
 *   matlab/utils/tt_irt1_int64.c:34                    double *ttcore, lapackint M, double *q, double *z, double *lPz)
 * bound by  python/tt_irt_py/tt_irt.py:23-25,51  (ctypes, c_int)  and  matlab/utils/tt_irt_mex.c:5,39 (mwIndex).
 *
 *   d       dimension
 *   n       mode sizes (d)
 *   xs      grid points of all dimensions stacked (sum n)
 *   ttrank  TT ranks (d+1), ttrank[0] = ttrank[d] = 1
 *   ttcore  TT cores, core k column-major r_k x n_k x r_{k+1}, stacked (sum r n r)
 *   M       number of seed points
 *   q       seeds in [0,1], column-major M x d           (host, read only)
 *   z       samples, column-major M x d                  (host, caller-allocated, fully overwritten)
 *   lPz     log sampling density at z, length M          (host, caller-allocated, fully overwritten)
 *
 * Returns nothing (as the reference).  On any failure (no device, bad shape, CUDA error) it prints
 * one line to stderr and fills z and lPz with NaN; it never aborts the host process.
 * Caller arrays in ordinary pageable memory (numpy, mxArray) go through page-locked bounce buffers filled and drained
 * by a few host threads (TTIRT_COPY_THREADS, default min(8, cores / 2); TTIRT_NO_STAGING=1 leaves the staging to the driver).
 * Environment: TTIRT_MODE=fast|strict (default fast), TTIRT_DEVICES=<count>|all|auto (default auto: one device per
 * 2^22 seed points, at most all visible ones from TTIRT_DEVICE on -- a batch of M >= 2^22 * N is sharded over N GPUs,
 * a small one stays on one), TTIRT_DEVICE=<first ordinal> (default 0), TTIRT_CHUNK=<samples per chunk>, TTIRT_CACHE=0 (free all device
 * memory before returning), TTIRT_TRACE=1 (host-side phase times on stderr), TTIRT_VERBOSE=1.
 * Shapes: any (as the reference, tt_irt1_int32.c:41-53).  Ranks <= 64 on grids <= 72 run the fused kernels, ranks and grids up
 * to 1024 the unfused FP64 tensor-core path (TTIRT_WIDE=0: off), anything larger the one-thread-per-sample strict kernel.
 */
TTIRT_API void tt_irt1(TTIRT_INT d, TTIRT_INT *n, double *xs, TTIRT_INT *ttrank, double *ttcore, TTIRT_INT M,
             double *q, double *z, double *lPz);

/* ------------------------------------------------------------------------------------------
 * Extended entry points (identical in both libraries; all sizes are int64_t).
 * ---------------------------------------------------------------------------------------- */

typedef struct ttirt_model ttirt_model; /* opaque: cores + marginal products resident on one device */

enum { TTIRT_MODE_FAST = 0, TTIRT_MODE_STRICT = 1 };

/* Upload grid and cores to `device`, run the right-to-left marginalisation sweep there
 * (reference tt_irt1_int32.c:59-82) and keep P_k = core_k x_3 C_k resident.  NULL on failure. */
TTIRT_API ttirt_model *ttirt_model_create(int64_t d, const int64_t *n, const double *xs, const int64_t *ttrank,
                                const double *ttcore, int device);
TTIRT_API void ttirt_model_destroy(ttirt_model *model);

/* Copy the sweep's results back: pk_out receives P_0..P_{d-1} stacked (sum r_k n_k, each column-major
 * r_k x n_k), marg_out the right marginals C_0..C_{d-1} stacked (sum r_{k+1}).  Either may be NULL. */
TTIRT_API int ttirt_model_get_sweep(const ttirt_model *model, double *pk_out, double *marg_out);

/* Sample with everything resident in device memory.  d_q / d_z are column-major with leading
 * dimensions ldq / ldz (>= M); d_lpz has length M; d_idx (may be NULL) receives the grid-interval
 * index chosen per (sample, dimension), int32, column-major with leading dimension ldz.
 * Work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and the
 * call returns without synchronising.  mode: TTIRT_MODE_FAST or TTIRT_MODE_STRICT.  0 on success. */
TTIRT_API int ttirt_sample_device(ttirt_model *model, int64_t M, const double *d_q, int64_t ldq, double *d_z,
                        int64_t ldz, double *d_lpz, int32_t *d_idx, int mode, void *stream);

/* Sample with host buffers (column-major, leading dimension ld >= M for q/z/idx): chunked
 * H2D -> kernels -> D2H pipeline on the model's device, rows [0, M).  h_idx may be NULL.
 * Blocks until the outputs are complete.  0 on success. */
TTIRT_API int ttirt_sample_host(ttirt_model *model, int64_t M, const double *h_q, double *h_z, double *h_lpz,
                      int32_t *h_idx, int64_t ld, int mode);

/* Whole call on host buffers: models on n_devices devices starting at first_device, one host thread and one copy /
 * compute pipeline per device, no collective.  The M rows are dealt out to the devices chunk by chunk from a shared queue
 * (a device with a faster path to host memory takes more; TTIRT_BALANCE=static: contiguous equal shards as given by
 * ttirt_shard_rows); chunk boundaries belong to the queue, so the result is bit-identical for every device count.  Grid
 * and cores are uploaded (to the first device from the host, to the others from there by peer copies; TTIRT_FANOUT=0: every
 * device from the host) and the sweep is run on every call; only device allocations -- and, for models up to 512 KB whose
 * grid and cores arrive byte-identical, the resident model -- are reused between calls (see ttirt_cache_clear).
 * This is what tt_irt1() runs.  0 on success. */
TTIRT_API int ttirt_run_host(int64_t d, const int64_t *n, const double *xs, const int64_t *ttrank,
                   const double *ttcore, int64_t M, const double *h_q, double *h_z, double *h_lpz,
                   int32_t *h_idx, int mode, int first_device, int n_devices);

/* Host-side arithmetic of the multi-device decomposition (no device needed): the contiguous row range [*m0, *m1) of
 * shard `shard` out of `n_shards` for a batch of M samples (what ttirt_run_host gives device first_device + shard; ranks
 * of a multi-process job use the same split), and the device count the drop-in call picks when TTIRT_DEVICES is unset
 * (one device per 2^22 seed points, at most `visible`).  ttirt_shard_rows returns 0, or -1 on bad arguments. */
TTIRT_API int ttirt_shard_rows(int64_t M, int n_shards, int shard, int64_t *m0, int64_t *m1);
TTIRT_API int ttirt_auto_devices(int64_t M, int visible);

/* Which kernels the fast mode uses for a TT of this shape (host-side arithmetic, no device needed; the reference serves
 * every shape on one code path, tt_irt1_int32.c:41-53 -- so does this library, through four): */
enum {
  TTIRT_PATH_STRICT = -1, /* one thread per sample, reference operation order (ranks or grids above 1024; TTIRT_MODE=strict) */
  TTIRT_PATH_FUSED16 = 0, /* sort + fused transition kernel, r <= 16 and n <= 24 */
  TTIRT_PATH_FUSED32 = 1, /* ... r <= 32 and n <= 40 */
  TTIRT_PATH_FUSED64 = 2, /* ... r <= 64 and n <= 72 (the BASELINE metric shape) */
  TTIRT_PATH_WIDE = 3,    /* sort + grouped DMMA GEMMs + streaming tail, ranks and grids up to 1024 (TTIRT_WIDE=0: strict) */
  TTIRT_PATH_WALK = 4     /* one persistent launch for all dimensions: every n = 17, ranks <= 16, d >= 2 (TTIRT_WALK=0: fused) */
};
TTIRT_API int ttirt_path_for_shape(int64_t d, const int64_t *n, const int64_t *ttrank);

/* Per-launch CUDA-event timing of the dominant kernel (the fused transition kernel) for calls to
 * ttirt_sample_device on this model: enable(1) clears and starts, read() synchronises and returns the summed
 * kernel time in ms, the number of launches timed and their algorithmic FP64 flops
 * (rows * (4 r_k r_{k+1} + 2 r_{k+1} n_{k+1}) per launch). */
TTIRT_API void ttirt_profile_enable(ttirt_model *model, int on);
TTIRT_API int ttirt_profile_read(ttirt_model *model, double *ms_total, int64_t *launches, double *flops_total);

/* Number of this library's kernels launched by the calling process so far (bench.py's gpu_launches). */
TTIRT_API int64_t ttirt_kernel_launches(void);
/* Last error message of the calling thread ("" if none). */
TTIRT_API const char *ttirt_last_error(void);
/* Number of visible CUDA devices (0 if none / driver missing). */
TTIRT_API int ttirt_device_count(void);
/* Samples per chunk used by the pipelines (0 restores the default). */
TTIRT_API void ttirt_set_chunk(int64_t samples);
/* tt_irt1 / ttirt_run_host keep their device allocations (never any results) per device for the next call of the
 * same shape; this releases them.  TTIRT_CACHE=0 in the environment disables the reuse altogether. */
TTIRT_API void ttirt_cache_clear(void);

/* ------------------------------------------------------------------------------------------
 * The steps either side of tt_irt1 on the device.  `_device` forms take
 * device pointers and a cudaStream_t (as void*); `_host` forms take host pointers, run on device TTIRT_DEVICE
 * (default 0) and block.  All return 0 on success.
 * ---------------------------------------------------------------------------------------- */

/* Seeds: shifted rank-1 lattice, reference matlab/samplers/qmcnodes.m:6-13:
 *   q[m + ldq*k] = frac(genvec[k] * ((m0 + m) / N) + shift[k]),  m in [0, M), k in [0, d)
 * N = 2^l is the size of the whole lattice, [m0, m0 + M) the slice generated here (shards generate their own slice).
 * Bit-exact against the reference arithmetic.  The device form takes the generating vector as doubles on the device. */
TTIRT_API int ttirt_seeds_lattice_device(int64_t d, int64_t M, int64_t m0, int64_t N, const double *d_genvec,
                                         const double *d_shift, double *d_q, int64_t ldq, void *stream);
TTIRT_API int ttirt_seeds_lattice_host(int64_t d, int64_t M, int64_t m0, int64_t N, const int64_t *genvec,
                                       const double *shift, double *h_q, int64_t ld);

/* Seeds: uniform pseudo-random numbers in [0, 1) (the reference draws them with rand / np.random.random,
 * python/test_shock_absorber_tt.py:147; neither is reproducible across hosts).  Philox4x32-10, counter =
 * (sample index, dimension), key = seed: any slice [m0, m0 + M) can be generated independently. */
TTIRT_API int ttirt_seeds_uniform_device(int64_t d, int64_t M, int64_t m0, uint64_t seed, double *d_q, int64_t ldq,
                                         void *stream);
TTIRT_API int ttirt_seeds_uniform_host(int64_t d, int64_t M, int64_t m0, uint64_t seed, double *h_q, int64_t ld);

/* tt_irt1 on seeds generated on the device: rows [0, M) of the result use lattice / Philox indices [m0, m0 + M), no q is
 * uploaded.  h_q may be NULL; when given it receives the seeds used (column-major M x d, leading dimension ld, as
 * h_z).  Same chunked pipeline, modes and error behaviour as ttirt_sample_host. */
TTIRT_API int ttirt_sample_lattice_host(ttirt_model *model, int64_t M, int64_t m0, int64_t N, const int64_t *genvec,
                                        const double *shift, double *h_q, double *h_z, double *h_lpz, int64_t ld, int mode);
TTIRT_API int ttirt_sample_uniform_host(ttirt_model *model, int64_t M, int64_t m0, uint64_t seed, double *h_q, double *h_z,
                                        double *h_lpz, int64_t ld, int mode);

/* The whole call (as ttirt_run_host: models on n_devices devices, contiguous row shards) with the seeds generated on the
 * devices: Philox indices [m0, m0 + M), identical results for every device count, no q upload.  h_q may be NULL. */
TTIRT_API int ttirt_run_uniform_host(int64_t d, const int64_t *n, const double *xs, const int64_t *ttrank,
                                     const double *ttcore, int64_t M, int64_t m0, uint64_t seed, double *h_q, double *h_z,
                                     double *h_lpz, int mode, int first_device, int n_devices);

/* Uniform -> truncated normal on [-sigma, sigma], reference matlab/samplers/randref.m:31-33:
 *   y = erfinv((u - 0.5) * erf(sigma / sqrt(2)) / 0.5) * sqrt(2) */
TTIRT_API int ttirt_truncnormal_map_device(int64_t n, double sigma, const double *d_u, double *d_y, void *stream);
TTIRT_API int ttirt_truncnormal_map_host(int64_t n, double sigma, const double *h_u, double *h_y);

/* Importance-weight statistics of M samples with log exact density lfex and log sampling density lfapp:
 *   out[0] isstd      relative standard deviation of the ratio            matlab/samplers/iw_prune.m:19-21,29
 *   out[1] max_ratio  largest normalised ratio                            iw_prune.m:24
 *   out[2] err1       empirical L1 error                                  iw_prune.m:25-26
 *   out[3] tau        N / ESS                                             matlab/samplers/essinv.m:12-14
 *   out[4] H          Hellinger distance                                  matlab/samplers/hellinger.m:12-16
 *   out[5] log of the importance-sampling normalisation constant          iw_prune.m:20,25
 * weights (may be NULL) receives the normalised ratios exp(lfex - lfapp) / mean (iw_prune.m:19-21), the factor the
 * reference multiplies its quantities of interest with (:28).  Blocks until `out` (host memory) is complete. */
TTIRT_API int ttirt_iw_stats_device(int64_t M, const double *d_lfex, const double *d_lfapp, double *d_weights, double *out,
                                    void *stream);
TTIRT_API int ttirt_iw_stats_host(int64_t M, const double *h_lfex, const double *h_lfapp, double *h_weights, double *out);

/* Independence Metropolis-Hastings prune, reference matlab/samplers/mcmc_prune.m:24-43 (and the loop of
 * python/test_shock_absorber_tt.py:165-171): proposal i+1 replaces the current sample c iff
 *   exp(((lfex[i+1] - lfex[c]) - lfapp[i+1]) + lfapp[c]) >= u[i],   i = 0 .. M-2, u pre-drawn uniforms (M-1 of them).
 * src[i] (int32, M entries) is the index of the sample occupying position i afterwards (the caller gathers y, lfex,
 * lfapp rows with it); num_rejects the number of rejections; rej_hist[L-1] the number of completed runs of L
 * consecutive rejections (runs longer than rej_hist_len are counted in the last bin; may be NULL with length 0).
 * Blocks until the host outputs are complete. */
TTIRT_API int ttirt_mcmc_prune_device(int64_t M, const double *d_lfex, const double *d_lfapp, const double *d_u,
                                      int32_t *d_src, int64_t *num_rejects, int64_t *rej_hist, int64_t rej_hist_len,
                                      void *stream);
TTIRT_API int ttirt_mcmc_prune_host(int64_t M, const double *h_lfex, const double *h_lfapp, const double *h_u, int32_t *h_src,
                                    int64_t *num_rejects, int64_t *rej_hist, int64_t rej_hist_len);

#ifdef __cplusplus
}
#endif
#endif /* TT_IRT1_H */
