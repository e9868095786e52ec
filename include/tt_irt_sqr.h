/*
 * tt_irt_sqr.h -- C-ABI of the B200-native squared-density inverse Rosenblatt transform.
 *
 * The reference routine is Matlab-only:
 *   matlab/samplers/tt_irt_sqr.m:1     function [xq, lFapp] = tt_irt_sqr(xsf, f, q)
 * with its two helpers
 *   matlab/utils/tracemult.c:103-112   C(:,:,i) = A(:,:,i) * B(:,:,j(i))   (batched product, used at :77, :109, :205)
 *   matlab/utils/tracemult.c:131-136   C(i) = A(i, j(i))                    (column pick, used at :139, :146-149)
 * and is what the DIRT sampler calls per layer (matlab/samplers/tt_dirt_sample.m:46,71).  It has no C entry point, so
 * the boundary here is new: the same argument meaning as tt_irt1 (include/tt_irt1.h; grid, TT2.0 cores, ranks, seeds in,
 * samples and log-density out) plus the two things tt_irt_sqr.m adds -- the grid may carry two boundary points per
 * dimension that the cores lack (:33-36, :53-60) and q may have fewer columns than the TT has dimensions (:9, :105).
 * INTEGRATION.md shows the MEX gateway a maintainer would put in front of it.
 *
 * Both shared libraries of tt-irt_b200/Makefile export these symbols (TTIRT_INT = int / long long for tt_irt_sqr, all
 * ttirt_sqr_* are width-independent).  Pointers are plain host or device addresses.  There is NO CPU fallback.
 */
#ifndef TT_IRT_SQR_H
#define TT_IRT_SQR_H

#include <stdint.h>

#include "tt_irt1.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * [xq, lFapp] = tt_irt_sqr(xsf, f, q)  -- reference matlab/samplers/tt_irt_sqr.m:1-208.
 *
 *   d       number of TT cores                                                      (:25)
 *   n       mode sizes of the cores (d)                                             (:26)
 *   nxs     number of grid points in xs: sum(n) or sum(n + 2)                       (:33-39)
 *   xs      grid points of all dimensions stacked, boundaries included              (:30-32)
 *   ttrank  TT ranks (d+1), ttrank[0] = ttrank[d] = 1                               (:27)
 *   ttcore  cores of the SQUARE ROOT of the density, core k column-major r_k x n_k x r_{k+1}, stacked
 *   M, D    q is column-major M x D, 0 < D <= d (D < d samples the marginal of the first D variables, :9, :105)
 *   z       samples, column-major M x D            (host, caller-allocated, fully overwritten)     (:184)
 *   lFapp   log of the sampling density, length M  (host, caller-allocated, fully overwritten)     (:194)
 *
 * The reference's `keyboard` debugger stops (:17-19, :151-154) have no counterpart: seeds outside [0, 1] are not
 * checked.  On any failure one line goes to stderr and z / lFapp are NaN-filled; the host process is never aborted.
 * Environment: TTIRT_DEVICE=<first ordinal> (default 0), TTIRT_DEVICES=<count>|all|auto (default auto, as for tt_irt1: one
 * device per 2^22 seed points; rows sharded contiguously over that many GPUs), TTIRT_SQR_CHUNK=<samples per chunk> (default 2^19 / 2^18 by shape class), TTIRT_CACHE=0 (no pooling of device blocks), TTIRT_TRACE=1.
 */
TTIRT_API void tt_irt_sqr(TTIRT_INT d, TTIRT_INT *n, TTIRT_INT nxs, double *xs, TTIRT_INT *ttrank, double *ttcore,
                          TTIRT_INT M, TTIRT_INT D, double *q, double *z, double *lFapp);

/*
 * [q, lFapp] = tt_rt_sqr(xsf, f, x)  -- the forward (Rosenblatt) transform, reference matlab/samplers/tt_rt_sqr.m:1-178:
 * same arguments and sweep as tt_irt_sqr, x (column-major M x D) in, q = CDF values in [0,1] and the log-density at x out.
 * The per-dimension tail is the reference's (:129-166): cell found on the grid, quadratic-spline CDF evaluated at x_k,
 * nothing clamped.  What tt_dirt_inverse.m:39,52 calls.  Same failure behaviour and environment as tt_irt_sqr.
 */
TTIRT_API void tt_rt_sqr(TTIRT_INT d, TTIRT_INT *n, TTIRT_INT nxs, double *xs, TTIRT_INT *ttrank, double *ttcore,
                         TTIRT_INT M, TTIRT_INT D, double *x, double *q, double *lFapp);

typedef struct ttirt_sqr_model ttirt_sqr_model; /* opaque: extended cores + packed semi-marginal Gram operands on one device */

/* Upload grid and cores, extrapolate the cores to the boundary when the grid has the two extra points (:53-60), run the
 * right-to-left sweep on the device (:41-82: core x R, Householder QR of the weighted unfolding, Cartesian square) and keep
 * the per-dimension operands resident.  NULL on failure.  Shapes: r <= 64, extended n <= 72. */
TTIRT_API ttirt_sqr_model *ttirt_sqr_model_create(int64_t d, const int64_t *n, int64_t nxs, const double *xs,
                                                  const int64_t *ttrank, const double *ttcore, int device);
TTIRT_API void ttirt_sqr_model_destroy(ttirt_sqr_model *model);

/* Read back the sweep of dimension k for the parity tests: gram_out (may be NULL) receives P{k} of tt_irt_sqr.m:80,
 * column-major r_k^2 x n_k (n_k after the boundary extension); rr_out (may be NULL) receives R'R of the factor to the
 * LEFT of core k (:66-72; r_k x r_k, only defined for k >= 1), the only form in which the factor enters the result. */
TTIRT_API int ttirt_sqr_model_get_sweep(const ttirt_sqr_model *model, int64_t k, double *gram_out, double *rr_out);
/* Extended mode size of dimension k (n_k or n_k + 2). */
TTIRT_API int64_t ttirt_sqr_model_mode_size(const ttirt_sqr_model *model, int64_t k);

/* Sample with everything resident in device memory; d_q / d_z column-major M x D with leading dimensions ldq / ldz,
 * d_lf length M, d_idx (may be NULL) int32 interval indices i0 (0-based), column-major with leading dimension ldz.
 * Enqueued on `stream` (cudaStream_t as void*), no synchronisation.  0 on success. */
TTIRT_API int ttirt_sqr_sample_device(ttirt_sqr_model *model, int64_t M, int64_t D, const double *d_q, int64_t ldq,
                                      double *d_z, int64_t ldz, double *d_lf, int32_t *d_idx, void *stream);
/* The same on host buffers (leading dimension ld >= M): chunked copy in, kernels, copy out.  Blocks.  0 on success. */
TTIRT_API int ttirt_sqr_sample_host(ttirt_sqr_model *model, int64_t M, int64_t D, const double *h_q, double *h_z,
                                    double *h_lf, int32_t *h_idx, int64_t ld);
/* The forward transform tt_rt_sqr on a resident model: points x in, CDF values q and log-density out (same layouts). */
TTIRT_API int ttirt_sqr_forward_device(ttirt_sqr_model *model, int64_t M, int64_t D, const double *d_x, int64_t ldx,
                                       double *d_q, int64_t ldq, double *d_lf, int32_t *d_idx, void *stream);
TTIRT_API int ttirt_sqr_forward_host(ttirt_sqr_model *model, int64_t M, int64_t D, const double *h_x, double *h_q,
                                     double *h_lf, int32_t *h_idx, int64_t ld);
/* Whole call on host buffers (model create, sample, release): what tt_irt_sqr() runs.  The M rows are cut into contiguous
 * ranges over n_devices devices starting at first_device (cores replicated, sweep redone per device, one host thread per
 * device, no collective).  0 on success. */
TTIRT_API int ttirt_sqr_run_host(int64_t d, const int64_t *n, int64_t nxs, const double *xs, const int64_t *ttrank,
                                 const double *ttcore, int64_t M, int64_t D, const double *h_q, double *h_z, double *h_lf,
                                 int first_device, int n_devices);

/* ... and what tt_rt_sqr() runs. */
TTIRT_API int ttirt_sqr_run_forward_host(int64_t d, const int64_t *n, int64_t nxs, const double *xs, const int64_t *ttrank,
                                         const double *ttcore, int64_t M, int64_t D, const double *h_x, double *h_q, double *h_lf,
                                         int first_device, int n_devices);

/* Per-launch CUDA-event timing of the dominant kernel (the conditional-pdf contraction, sqr_pdf_kernel):
 * enable(1) clears and starts; read() synchronises and returns summed kernel time, launches and their algorithmic flops
 * (rows * (r_k (r_k + 1) n_k + r_k (r_k + 1) / 2) per launch: the symmetric half of :109-112). */
TTIRT_API void ttirt_sqr_profile_enable(ttirt_sqr_model *model, int on);
TTIRT_API int ttirt_sqr_profile_read(ttirt_sqr_model *model, double *ms_total, int64_t *launches, double *flops_total);

/* The DIRT sampler loop, reference matlab/samplers/tt_dirt_sample.m:17-73 (spline interpolation, TT cross variants: the
 * branch that calls tt_irt_sqr at :46 and :71), i.e. the caller of tt_irt_sqr, with the samples resident on the device
 * between the levels:
 *   for j = nlvl .. 1:  [z = erf(z / sqrt(2)) cdf_factor + 0.5]  ->  [z, dl] = tt_irt_sqr(x, F{j}, z)  ->  lFapp += dl
 *                       [ + sum(z.^2, 2) / 2 - log(2 cdf_factor^2 / pi) d / 2 ]            (bracketed: normal reference only)
 *   level 0:            [same map]  ->  [z, dl] = tt_irt_sqr(x0, F0, z)  ->  lFapp += dl
 * models[0] is the model of F0 on x0, models[j] of F{j} on x (nlevels = nlvl + 1; all of dimension d on one device);
 * sigma <= 0: uniform reference, sigma > 0: normal reference truncated to [-sigma, sigma] ('Normal S', :21-30).
 * q, z column-major M x d.  `_device`: enqueued on `stream`, scratch kept in models[0]; `_host`: blocks.  0 on success. */
TTIRT_API int ttirt_dirt_sample_device(int64_t nlevels, ttirt_sqr_model *const *models, double sigma, int64_t M, const double *d_q,
                                       int64_t ldq, double *d_z, int64_t ldz, double *d_lf, void *stream);
TTIRT_API int ttirt_dirt_sample_host(int64_t nlevels, ttirt_sqr_model *const *models, double sigma, int64_t M, const double *h_q,
                                     double *h_z, double *h_lf, int64_t ld);

/* The inverse of the DIRT, reference matlab/samplers/tt_dirt_inverse.m:24-59: level 0 first, then levels 1..nlvl, each
 *   [lFapp += sum(q.^2, 2) / 2]  ->  [q, dl] = tt_rt_sqr(x, F{j}, q)  ->  [q = erfinv((q - 0.5) / cdf_factor) sqrt(2)]  ->  lFapp += dl
 * (bracketed: normal reference only; the reference drops the additive constant of the normal density here, :49, mirrored).
 * Same model array and sigma as ttirt_dirt_sample_*; x (M x d, points in the target space) in, reference-space q and the
 * log-density at x out. */
TTIRT_API int ttirt_dirt_inverse_device(int64_t nlevels, ttirt_sqr_model *const *models, double sigma, int64_t M, const double *d_x,
                                        int64_t ldx, double *d_q, int64_t ldq, double *d_lf, void *stream);
TTIRT_API int ttirt_dirt_inverse_host(int64_t nlevels, ttirt_sqr_model *const *models, double sigma, int64_t M, const double *h_x,
                                      double *h_q, double *h_lf, int64_t ld);

/* tracemult, reference matlab/utils/tracemult.c (real arguments; complex input is not supported), as a standalone operator.
 * Inside tt_irt_sqr both forms are fused into the path's kernels; these serve callers of the MEX itself.
 *   d_B != NULL:  C(:,:,i) = A(:,:,i) * B(:,:,j(i))   (:103-112)   A p x m x n, B m x k x s, C p x k x n, all column-major
 *   d_B == NULL:  C(i) = A(i, j(i))                     (:131-136)   A n x s, C length n (p, m, k ignored)
 * j holds 1-based indices stored as doubles, exactly what the MEX reads (:106-107).  An index outside 1..s gives NaN in
 * that slice (device form) or an error before anything is computed (host form); the MEX reads out of bounds there.
 * `_device`: device pointers, enqueued on `stream`; `_host`: host pointers, device TTIRT_DEVICE, blocks.  0 on success. */
TTIRT_API int ttirt_tracemult_device(int64_t p, int64_t m, int64_t k, int64_t n, int64_t s, const double *d_A, const double *d_j,
                                     const double *d_B, double *d_C, void *stream);
TTIRT_API int ttirt_tracemult_host(int64_t p, int64_t m, int64_t k, int64_t n, int64_t s, const double *h_A, const double *h_j,
                                   const double *h_B, double *h_C);

#ifdef __cplusplus
}
#endif
#endif /* TT_IRT_SQR_H */
