/*
 * pmu_capi.c — the drop-in C API of SynchroPMU on top of the CUDA estimator.
 *
 * Compiled once per ABI variant (-DNUM_CHANLS=<n>, -DPMU_FLOAT_P=<float|double>), like
 * the reference library (CMakeLists.txt:8-12, src/func_stubs.h:38).  Same four entry
 * points, argument meaning and 0 / -1 error behaviour as reference
 * src/pmu_estimator.c:92-324; all arithmetic is done on the GPU by pmu_core.cu — a
 * pmu_context here is a batch of ONE instance x NUM_CHANLS channels.
 */
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pmu_estimator.h"
#include "../../include/pmu_batch.h"

#ifndef M_PI /* strict ISO C modes do not define it */
#define M_PI 3.14159265358979323846
#endif

/* internal entry points of pmu_core.cu (not part of the public headers) */
int pmub_set_mirror(pmub_batch *b, int enable);
const double *pmub_mirror(const pmub_batch *b);

#define PMU_DTYPE ((int)sizeof(float_p))

static pmub_batch *handle_of(pmu_context *ctx) { return (pmub_batch *)(void *)ctx->buff_params.Xf; }

static void to_core_config(const estimator_config *in, pmub_config *out)
{
    memset(out, 0, sizeof *out);
    out->n_cycles = in->n_cycles;
    out->f0 = in->f0;
    out->frame_rate = in->frame_rate;
    out->fs = in->fs;
    out->n_bins = in->n_bins;
    out->P = in->P;
    out->Q = in->Q;
    out->iter_eipdft = in->iter_eipdft ? 1u : 0u;
    out->interf_trig = (double)in->interf_trig;
    for (int i = 0; i < 3; i++) {
        out->rocof_thresh[i] = (double)in->rocof_thresh[i];
        out->rocof_low_pass_coeffs[i] = (double)in->rocof_low_pass_coeffs[i];
    }
}

static int read_config(void *cfg, bool_p config_from_ini, pmub_config *out)
{
    if (!cfg) {
        fprintf(stderr, "[pmu_init] ERROR: NULL configuration\n");
        return -1;
    }
    if (config_from_ini)
        return pmub_parse_ini((const char *)cfg, out); /* cfg is a char* path, :333-341 */
    to_core_config((const estimator_config *)cfg, out);
    return 0;
}

/* reference src/pmu_estimator.c:92-167 */
int pmu_init(pmu_context *ctx, void *cfg, bool_p config_from_ini)
{
    if (!ctx) {
        fprintf(stderr, "[pmu_init] ERROR: NULL context\n");
        return -1;
    }
    if (ctx->pmu_initialized) {
        fprintf(stderr, "[pmu_init] ERROR: pmu estimator already initialized\n");
        return -1;
    }
    pmub_config c;
    if (read_config(cfg, config_from_ini, &c) || pmub_validate(&c)) {
        fprintf(stderr, "[pmu_init] ERROR: pmu estimator configuration failed\n");
        return -1;
    }
    pmub_batch *b = NULL;
    if (pmub_create(&b, &c, PMU_DTYPE, 1, NUM_CHANLS, -1) || pmub_set_mirror(b, 1)) {
        if (b)
            pmub_destroy(b);
        fprintf(stderr, "[pmu_init] ERROR: pmu estimator configuration failed\n");
        return -1;
    }
    pmub_info info;
    pmub_get_info(b, &info);

    /* fill the caller-visible context like the reference does (:344-366, :141-160) */
    SynchrophasorEstimatorParams *sp = &ctx->synch_params;
    sp->win_len = info.win_len;
    sp->n_cycles = c.n_cycles;
    sp->f0 = c.f0;
    sp->frame_rate = c.frame_rate;
    sp->fs = c.fs;
    sp->n_bins = c.n_bins;
    sp->P = c.P;
    sp->Q = c.Q;
    sp->iter_eipdft_enabled = c.iter_eipdft ? 1 : 0;
    sp->interf_trig = (float_p)c.interf_trig;
    sp->df = (float_p)info.df;
    sp->norm_factor = (float_p)info.norm_factor;
    memset(&ctx->buff_params, 0, sizeof ctx->buff_params);
    ctx->buff_params.Xf = (void *)b; /* opaque device-estimator handle */
    HanningTransformConstants *hp = &ctx->hann_params;
    hp->C0 = -0.25;
    hp->C1 = 0.5;
    hp->C2 = -0.25;
    hp->C3 = ((float_p)sp->win_len - 1) / (float_p)sp->win_len;
    hp->C4 = 1.0 / (float_p)sp->win_len;
    hp->C5 = hp->C0 * cexp(I * M_PI * hp->C3);
    hp->C6 = hp->C2 * cexp(-I * M_PI * hp->C3);
    hp->inv_norm_factor = 1.0 / sp->norm_factor;
    RocoFEstimationStates *rp = &ctx->rocof_params;
    for (int i = 0; i < 3; i++) {
        rp->thresholds[i] = (float_p)c.rocof_thresh[i];
        rp->low_pass_coeff[i] = (float_p)c.rocof_low_pass_coeffs[i];
    }
    for (int ch = 0; ch < NUM_CHANLS; ch++) {
        rp->delay_line[ch][0] = 0.0;
        rp->delay_line[ch][1] = 0.0;
        rp->freq_old[ch] = sp->f0;
        rp->state[ch] = 0;
    }
    ctx->pmu_initialized = 1;
    return 0;
}

/* reference src/pmu_estimator.c:170-271 */
int pmu_estimate(pmu_context *ctx, float_p *in_signal_windows, float_p mid_fracsec, pmu_frame *out_frame)
{
    if (!ctx || !ctx->pmu_initialized) {
        fprintf(stderr, "[pmu_estimate] ERROR: pmu estimator not initialized, first initialize with pmu_init function\n");
        return -1;
    }
    pmub_batch *b = handle_of(ctx);
    if (pmub_estimate(b, in_signal_windows, &mid_fracsec, out_frame))
        return -1;
    /* keep ctx->rocof_params observable (state itself lives in HBM) */
    const double *m = pmub_mirror(b);
    if (m) {
        RocoFEstimationStates *rp = &ctx->rocof_params;
        for (int ch = 0; ch < NUM_CHANLS; ch++) {
            rp->freq_old[ch] = (float_p)m[4 * ch + 0];
            rp->delay_line[ch][0] = (float_p)m[4 * ch + 1];
            rp->delay_line[ch][1] = (float_p)m[4 * ch + 2];
            rp->state[ch] = m[4 * ch + 3] != 0.0;
        }
    }
    return 0;
}

/* reference src/pmu_estimator.c:274-303 */
int pmu_deinit(pmu_context *ctx)
{
    if (!ctx || !ctx->pmu_initialized) {
        fprintf(stderr, "[pmu_deinit] ERROR: pmu estimator not initialized, first initialize with pmu_init function\n");
        return -1;
    }
    pmub_destroy(handle_of(ctx));
    memset(&ctx->buff_params, 0, sizeof ctx->buff_params);
    ctx->pmu_initialized = 0;
    return 0;
}

/* reference src/pmu_estimator.c:306-324 (same text format) */
int pmu_abi_float_bytes(void) { return (int)sizeof(float_p); }

int pmu_dump_frame(pmu_frame *frame, FILE *stream)
{
    if (frame == NULL || stream == NULL) {
        fprintf(stderr, "Error: NULL pointer passed as argument\n");
        return -1;
    }
    double ph = (double)frame->synchrophasor.ph;
    ph = ph - 2 * M_PI * rint(ph / (2 * M_PI));
    int written = fprintf(stream, "[Synchrophasor] amplitude: %lf, phase: %lf, frequency: %lf, rocof: %lf\n",
                          (double)frame->synchrophasor.amp, ph, (double)frame->synchrophasor.freq, (double)frame->rocof);
    if (written < 0) {
        fprintf(stderr, "Error: failed to write to stream\n");
        return -1;
    }
    return 0;
}

/* ---- typed batched entry (pmu_estimator.h) --------------------------------------------- */
int pmu_batch_init(pmu_batch **b, void *cfg, bool_p config_from_ini, uint_p n_instances, uint_p n_channels, int device)
{
    pmub_config c;
    if (!b || read_config(cfg, config_from_ini, &c))
        return -1;
    return pmub_create(b, &c, PMU_DTYPE, n_instances, n_channels, device);
}

int pmu_estimate_batch(pmu_batch *b, const float_p *windows, const float_p *mid_fracsec, pmu_frame *out)
{
    pmub_info info;
    if (!b || pmub_get_info(b, &info))
        return -1;
    if (info.dtype != PMU_DTYPE) {
        fprintf(stderr, "[pmu_estimate_batch] ERROR: batch was created for another float_p width\n");
        return -1;
    }
    return pmub_estimate(b, windows, mid_fracsec, out);
}

int pmu_batch_deinit(pmu_batch *b) { return pmub_destroy(b); }
