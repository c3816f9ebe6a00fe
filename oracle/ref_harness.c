/*
 * ref_harness.c — thin driver around the UNMODIFIED reference library.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h).  This file is OUR code; it is
 * compiled by oracle/Makefile together with the reference's own sources taken
 * where they lie under /root/reference (src/pmu_estimator.c,
 * libs/iniparser/{iniparser,dictionary}.c) into oracle/_ref/libref_*.so.
 * It only calls the reference's public API (src/pmu_estimator.h:119-128) and
 * reads two fields of the caller-owned context after a call:
 *   ctx->buff_params.dftbins  (raw windowed spectrum of the last channel,
 *                              written at src/pmu_estimator.c:200 and never
 *                              modified afterwards), and
 *   ctx->synch_params.*       (win_len, n_bins, norm_factor).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "pmu_estimator.h" /* the reference header (NUM_CHANLS via -D) */

#define ORC_PREFIX refh
#include "oracle_api.h"

static void to_ref_config(const orc_config *in, estimator_config *out)
{
    memset(out, 0, sizeof(*out));
    out->n_cycles = in->n_cycles;
    out->f0 = in->f0;
    out->frame_rate = in->frame_rate;
    out->fs = in->fs;
    out->n_bins = in->n_bins;
    out->P = in->P;
    out->Q = in->Q;
    out->iter_eipdft = in->iter_eipdft ? 1 : 0;
    out->interf_trig = (float_p)in->interf_trig;
    for (int i = 0; i < 3; i++) {
        out->rocof_thresh[i] = (float_p)in->rocof_thresh[i];
        out->rocof_low_pass_coeffs[i] = (float_p)in->rocof_low_pass_coeffs[i];
    }
}

void refh_info(int *float_size, int *num_chanls)
{
    *float_size = (int)sizeof(float_p);
    *num_chanls = NUM_CHANLS;
}

void *refh_open(const orc_config *cfg, unsigned n_channels)
{
    if (n_channels != NUM_CHANLS)
        return NULL;
    estimator_config rc;
    to_ref_config(cfg, &rc);
    pmu_context *ctx = calloc(1, sizeof(pmu_context));
    if (!ctx)
        return NULL;
    if (pmu_init(ctx, &rc, CONFIG_FROM_STRUCT) != 0) {
        free(ctx);
        return NULL;
    }
    return ctx;
}

void *refh_open_ini(const char *ini_path, unsigned n_channels)
{
    if (n_channels != NUM_CHANLS)
        return NULL;
    pmu_context *ctx = calloc(1, sizeof(pmu_context));
    if (!ctx)
        return NULL;
    if (pmu_init(ctx, (void *)ini_path, CONFIG_FROM_INI) != 0) {
        free(ctx);
        return NULL;
    }
    return ctx;
}

void refh_close(void *h)
{
    if (!h)
        return;
    pmu_deinit((pmu_context *)h);
    free(h);
}

unsigned refh_win_len(void *h) { return ((pmu_context *)h)->synch_params.win_len; }
unsigned refh_n_bins(void *h) { return ((pmu_context *)h)->synch_params.n_bins; }
double refh_norm_factor(void *h) { return (double)((pmu_context *)h)->synch_params.norm_factor; }

int refh_estimate(void *h, const void *windows, double mid_fracsec, void *frames)
{
    return pmu_estimate((pmu_context *)h, (float_p *)windows, (float_p)mid_fracsec,
                        (pmu_frame *)frames);
}

/* argmax over |dftbins[0..n_bins)| with "first strict maximum wins". */
int refh_last_km(void *h)
{
    pmu_context *ctx = (pmu_context *)h;
    unsigned nb = ctx->synch_params.n_bins;
    int best = 0;
    float_p bestv = cabs(ctx->buff_params.dftbins[0]);
    for (unsigned j = 1; j < nb; j++) {
        float_p v = cabs(ctx->buff_params.dftbins[j]);
        if (v > bestv) {
            bestv = v;
            best = (int)j;
        }
    }
    return best;
}

void refh_last_spectrum(void *h, double *reim)
{
    pmu_context *ctx = (pmu_context *)h;
    unsigned nb = ctx->synch_params.n_bins;
    for (unsigned j = 0; j < nb; j++) {
        reim[2 * j] = (double)creal(ctx->buff_params.dftbins[j]);
        reim[2 * j + 1] = (double)cimag(ctx->buff_params.dftbins[j]);
    }
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double refh_run(const orc_config *cfg, unsigned n_channels, long n_inst, long n_frames,
                const void *windows, const double *mid, void *frames, int *km, int n_threads)
{
    if (n_channels != NUM_CHANLS || n_inst <= 0 || n_frames <= 0)
        return -1.0;
    estimator_config rc;
    to_ref_config(cfg, &rc);
    pmu_context *ctxs = calloc((size_t)n_inst, sizeof(pmu_context));
    if (!ctxs)
        return -1.0;
    long bad = 0;
    for (long i = 0; i < n_inst; i++)
        if (pmu_init(&ctxs[i], &rc, CONFIG_FROM_STRUCT) != 0)
            bad++;
    if (bad) {
        for (long i = 0; i < n_inst; i++)
            if (ctxs[i].pmu_initialized)
                pmu_deinit(&ctxs[i]);
        free(ctxs);
        return -2.0;
    }
    const size_t N = ctxs[0].synch_params.win_len;
    const float_p *win = (const float_p *)windows;
    pmu_frame *out = (pmu_frame *)frames;
#ifdef _OPENMP
    if (n_threads > 0)
        omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
    double t0 = now_s();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n_inst; i++) {
        for (long t = 0; t < n_frames; t++) {
            size_t u = (size_t)t * (size_t)n_inst + (size_t)i;
            pmu_estimate(&ctxs[i], (float_p *)(win + u * NUM_CHANLS * N), (float_p)mid[u],
                         out + u * NUM_CHANLS);
            if (km) {
                for (int c = 0; c < NUM_CHANLS - 1; c++)
                    km[u * NUM_CHANLS + c] = -1;
                km[u * NUM_CHANLS + NUM_CHANLS - 1] = refh_last_km(&ctxs[i]);
            }
        }
    }
    double t1 = now_s();
    for (long i = 0; i < n_inst; i++)
        pmu_deinit(&ctxs[i]);
    free(ctxs);
    return t1 - t0;
}
