/*
 * pmu_oracle.c — plain-C CPU restatement of SynchroPMU's pmu_estimate() path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker ("oracle", kind "port"), not
 * the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The shipped CUDA library
 * never links, calls or falls back to anything in this file.
 *
 * Parity pinning: this restatement is validated (tests/test_oracle.py) against
 *   (1) oracle/_ref/libref_*.so — the unmodified reference compiled from
 *       /root/reference by oracle/Makefile — whenever that build is present, and
 *   (2) tests/golden/*.json — outputs of that same reference build, generated
 *       by tests/golden/make_golden.py and committed, which travel to the GPU box.
 *
 * Every routine cites the reference lines it restates (all paths are relative to
 * /root/reference).  The arithmetic keeps the reference's operation order and,
 * for the float build (ORC_REAL=float), its "float storage, double libm" typing
 * (src/func_stubs.h:33,74-142): values typed `real` round to float on store while
 * M_PI, sin, cos, cexp, cabs and carg work in double.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef ORC_REAL
#define ORC_REAL double
#endif
typedef ORC_REAL real;
typedef ORC_REAL _Complex cplx;
typedef double _Complex zplx;

#define ORC_PREFIX orc
#include "oracle_api.h"

typedef struct {
    real amp, ph, freq;
} tone_t;

typedef struct {
    /* configuration (src/pmu_estimator.h:59-74) */
    unsigned N, n_cycles, f0, frame_rate, fs, K, P, Q, iter_on, C;
    real eps, df, S;
    /* Hann-kernel constants (src/pmu_estimator.c:153-160) */
    real c1, c3, c4, inv_S;
    cplx c5, c6;
    /* ROCOF parameters and per-channel state (src/pmu_estimator.h:100-107) */
    real thr[3], lp[3];
    real *f_old, *dl0, *dl1;
    unsigned char *st;
    /* work arrays (src/pmu_estimator.c:110-138) */
    real *hann, *sw;
    cplx *X, *Xi, *Xf, *tmpK, *tmpK2;
    /* observability for the parity tests */
    int last_km, km_armed;
} orc_est;

/* ---- configuration check: src/pmu_estimator.c:386-449 ------------------- */
static int config_ok(const orc_est *e)
{
    if (e->f0 != 50 && e->f0 != 60) return 0;                       /* :389 */
    if (e->n_cycles == 0 || e->n_cycles > 50) return 0;             /* :395 */
    if (e->fs == 0 || fmod(e->fs, e->f0) != 0.0) return 0;          /* :400 */
    if (e->frame_rate == 0) return 0;                               /* :406 */
    unsigned half = (e->n_cycles * e->fs / e->f0) / 2;              /* :411 */
    if (e->K < e->n_cycles + 2 || e->K > half) return 0;
    if (e->P == 0) return 0;                                        /* :416 */
    if (e->eps <= 0 || e->eps > 1) return 0;                        /* :421 */
    for (int i = 0; i < 3; i++)                                     /* :432-446 */
        if (e->thr[i] <= 0) return 0;
    return 1;
}

/* ---- periodic Hann table and its sum: src/pmu_estimator.c:549-560 ------- */
static real build_hann(real *w, unsigned n)
{
    real sum = 0;
    for (unsigned i = 0; i < n; i++) {
        w[i] = 0.5 * (1 - cos(2 * M_PI * i / n));
        sum += w[i];
    }
    return sum;
}

void orc_info(int *float_size, int *num_chanls)
{
    *float_size = (int)sizeof(real);
    *num_chanls = 0;
}

void orc_close(void *h)
{
    orc_est *e = (orc_est *)h;
    if (!e) return;
    free(e->f_old); free(e->dl0); free(e->dl1); free(e->st);
    free(e->hann); free(e->sw);
    free(e->X); free(e->Xi); free(e->Xf); free(e->tmpK); free(e->tmpK2);
    free(e);
}

/* ---- pmu_init / config_estimator: src/pmu_estimator.c:92-167, 327-384 --- */
void *orc_open(const orc_config *cfg, unsigned n_channels)
{
    if (n_channels < 1) return NULL;
    orc_est *e = calloc(1, sizeof(orc_est));
    if (!e) return NULL;
    e->C = n_channels;
    e->n_cycles = cfg->n_cycles; e->f0 = cfg->f0; e->frame_rate = cfg->frame_rate;
    e->fs = cfg->fs; e->K = cfg->n_bins; e->P = cfg->P; e->Q = cfg->Q;
    e->iter_on = cfg->iter_eipdft ? 1 : 0;
    e->eps = (real)cfg->interf_trig;
    for (int i = 0; i < 3; i++) {
        e->thr[i] = (real)cfg->rocof_thresh[i];
        e->lp[i] = (real)cfg->rocof_low_pass_coeffs[i];
    }
    if (e->f0 == 0) { free(e); return NULL; }
    e->N = e->n_cycles * e->fs / e->f0;                             /* :365 */
    if (e->N == 0) { free(e); return NULL; }
    e->df = (real)e->fs / (real)e->N;                               /* :366 */
    if (!config_ok(e)) { free(e); return NULL; }

    e->f_old = malloc(sizeof(real) * e->C);
    e->dl0 = calloc(e->C, sizeof(real));
    e->dl1 = calloc(e->C, sizeof(real));
    e->st = calloc(e->C, 1);
    e->hann = malloc(sizeof(real) * e->N);
    e->sw = malloc(sizeof(real) * e->N);
    e->X = malloc(sizeof(cplx) * e->N);
    e->Xi = malloc(sizeof(cplx) * e->K);
    e->Xf = malloc(sizeof(cplx) * e->K);
    e->tmpK = malloc(sizeof(cplx) * e->K);
    e->tmpK2 = malloc(sizeof(cplx) * e->K);
    if (!e->f_old || !e->dl0 || !e->dl1 || !e->st || !e->hann || !e->sw || !e->X ||
        !e->Xi || !e->Xf || !e->tmpK || !e->tmpK2) {
        orc_close(e);
        return NULL;
    }
    for (unsigned c = 0; c < e->C; c++) e->f_old[c] = e->f0;        /* :141-147 */
    e->S = build_hann(e->hann, e->N);                               /* :150 */
    real c0 = -0.25, c2 = -0.25;                                    /* :153-160 */
    e->c1 = 0.5;
    e->c3 = ((real)e->N - 1) / (real)e->N;
    e->c4 = 1.0 / (real)e->N;
    e->c5 = c0 * cexp(I * M_PI * e->c3);
    e->c6 = c2 * cexp(-I * M_PI * e->c3);
    e->inv_S = 1.0 / e->S;
    e->last_km = -1;
    return e;
}

/* ini parsing is host logic outside the arithmetic path; it is checked against
 * tests/golden/ini_configs.json instead of being restated here. */
void *orc_open_ini(const char *ini_path, unsigned n_channels)
{
    (void)ini_path; (void)n_channels;
    return NULL;
}

unsigned orc_win_len(void *h) { return ((orc_est *)h)->N; }
unsigned orc_n_bins(void *h) { return ((orc_est *)h)->K; }
double orc_norm_factor(void *h) { return (double)((orc_est *)h)->S; }

/* ---- spectrum of a real sequence: src/pmu_estimator.c:487-546 ----------- */
/* Power-of-two lengths: recursive radix-2 decimation in time, one twiddle
 * cexp(-2*pi*i*k/len) per butterfly (:511-546).  `work` supplies 2*len reals and
 * 2*len complex of scratch for this level and everything below it. */
static void radix2(const real *in, cplx *out, unsigned len, real *rwork, cplx *cwork)
{
    if (len == 1) { out[0] = in[0]; return; }
    unsigned h = len / 2;
    real *ev = rwork, *od = rwork + h;
    cplx *Ev = cwork, *Od = cwork + h;
    for (unsigned i = 0; i < h; i++) { ev[i] = in[2 * i]; od[i] = in[2 * i + 1]; }
    radix2(ev, Ev, h, rwork + len, cwork + len);
    radix2(od, Od, h, rwork + len, cwork + len);
    for (unsigned k = 0; k < h; k++) {
        cplx t = cexp(-I * 2 * M_PI * k / len) * Od[k];
        out[k] = Ev[k] + t;
        out[k + h] = Ev[k] - t;
    }
}

static int spectrum(const real *in, cplx *out, unsigned len, unsigned n_bins)
{
    if ((len & (len - 1)) == 0) {                                   /* :490-495 */
        real *rw = malloc(sizeof(real) * 2 * len);
        cplx *cw = malloc(sizeof(cplx) * 2 * len);
        if (!rw || !cw) { free(rw); free(cw); return -1; }
        radix2(in, out, len, rw, cw);
        free(rw); free(cw);
    } else {                                                        /* :499-506 */
        for (unsigned k = 0; k < n_bins; ++k) {
            out[k] = 0;
            for (unsigned n = 0; n < len; ++n)
                out[k] += in[n] * cexp(-I * ((n * k * M_PI * 2 / (real)len)));
        }
    }
    return 0;
}

/* ---- Hann-window spectrum of a single tone: macros at :59-61 ------------ */
static double dirichlet(const orc_est *e, real u)                   /* sineasin :59 */
{
    return sin(M_PI * u) / sin(e->c4 * M_PI * u);
}

static zplx tone_bin(const orc_est *e, unsigned k, real f, cplx z)  /* wf :61, whDFT :60 */
{
    real x = k - (f / e->df);
    zplx rot = cexp(-I * M_PI * e->c3 * x);
    zplx lobes = e->c5 * dirichlet(e, x - 1) + e->c1 * dirichlet(e, x) + e->c6 * dirichlet(e, x + 1);
    return z * (rot * lobes) * e->inv_S;
}

/* ---- peak search: src/pmu_estimator.c:705-720 ---------------------------- */
static void peak3(const real *m, unsigned n, unsigned *km, unsigned *kl, unsigned *kr)
{
    unsigned best = 0;
    for (unsigned i = 1; i < n; i++)
        if (m[i] > m[best]) best = i;
    *km = best;
    *kl = best == 0 ? 0 : best - 1;
    *kr = best == n - 1 ? best - 1 : best + 1;
}

/* ---- 3-point interpolated DFT: src/pmu_estimator.c:579-624 --------------- */
static int interp3(orc_est *e, const cplx *Y, tone_t *t)
{
    real mag[e->K];
    for (unsigned j = 0; j < e->K; j++) mag[j] = cabs(Y[j]);        /* :589-592 */
    unsigned k1, k2, k3;
    peak3(mag, e->K, &k1, &k2, &k3);                                /* :594 */
    if (e->km_armed) { e->last_km = (int)k1; e->km_armed = 0; }
    real d = 2 * (mag[k3] - mag[k2]) / (mag[k2] + mag[k3] + 2 * mag[k1]);   /* :598 */
    t->freq = (k1 + d) * e->df;                                     /* :601 */
    if (fabs(d) <= 10e-12) {                                        /* :603-613 */
        t->amp = mag[k1];
        t->ph = carg(Y[k1]);
        return 1;
    }
    t->amp = mag[k1] * fabs((d * d - 1) * (M_PI * d) / sin(M_PI * d));      /* :616 */
    t->ph = carg(Y[k1]) - M_PI * d;                                 /* :617 */
    return 0;
}

/* ---- enhanced IpDFT (negative-image compensation): :626-664 -------------- */
static void enhance(orc_est *e, const cplx *Y, tone_t *out)
{
    tone_t t = *out;
    if (!interp3(e, Y, &t)) {                                       /* :633 */
        cplx *Ypos = e->tmpK;
        for (unsigned p = 0; p < e->P; p++) {                       /* :641 */
            cplx z = t.amp * cexp(-I * t.ph);                       /* :646 */
            for (unsigned j = 0; j < e->K; j++) {                   /* :648-652 */
                cplx img = tone_bin(e, j, -t.freq, z);
                Ypos[j] = Y[j] - img;
            }
            if (interp3(e, Ypos, &t)) break;                        /* :654-657 */
        }
    }
    *out = t;
}

/* ---- spectrum of one real tone (both images): :563-577 ------------------- */
static void synth(const orc_est *e, cplx *out, tone_t t)
{
    cplx zp = t.amp * cexp(I * t.ph);
    cplx zn = t.amp * cexp(-I * t.ph);
    for (unsigned i = 0; i < e->K; i++)
        out[i] = tone_bin(e, i, t.freq, zp) + tone_bin(e, i, -t.freq, zn);
}

/* ---- iterative interference removal: :666-703 ---------------------------- */
static void refine(orc_est *e, const cplx *X, cplx *Xi, cplx *Xf, tone_t *fund)
{
    tone_t f = *fund, it;
    memset(&it, 0, sizeof(it));
    cplx *Xi_pure = e->tmpK2;
    for (unsigned i = 0; i < e->Q; i++) {                           /* :675 */
        enhance(e, Xi, &it);                                        /* :679 */
        synth(e, Xi_pure, it);                                      /* :680 */
        for (unsigned j = 0; j < e->K; j++) Xf[j] = X[j] - Xi_pure[j];      /* :681-684 */
        enhance(e, Xf, &f);                                         /* :685 */
        if (i < e->Q - 1) {                                         /* :687 */
            synth(e, Xf, f);                                        /* :690 */
            for (unsigned j = 0; j < e->K; j++) Xi[j] = X[j] - Xf[j];       /* :692-695 */
        }
    }
    *fund = f;
}

static double wrap_pi(real rad)                                     /* wrap_angle :58 */
{
    return rad - 2 * M_PI * rint(rad / (2 * M_PI));
}

/* ---- pmu_estimate: src/pmu_estimator.c:170-271 --------------------------- */
int orc_estimate(void *h, const void *windows, double mid_fracsec_d, void *frames)
{
    orc_est *e = (orc_est *)h;
    if (!e) return -1;
    const real *in = (const real *)windows;
    real *out = (real *)frames;
    real mid_fracsec = (real)mid_fracsec_d;
    for (unsigned c = 0; c < e->C; c++) {
        const real *x = in + (size_t)c * e->N;
        for (unsigned j = 0; j < e->N; j++) e->sw[j] = x[j] * e->hann[j];   /* :184-191 */
        if (spectrum(e->sw, e->X, e->N, e->K)) return -1;          /* :200 */
        tone_t t; memset(&t, 0, sizeof(t));
        e->km_armed = 1;
        enhance(e, e->X, &t);                                       /* :205 */
        synth(e, e->Xf, t);                                         /* :206 */
        real Ed = 0, E = 0;                                         /* :208-222 */
        for (unsigned j = 0; j < e->K; j++) {
            e->Xi[j] = e->X[j] - e->Xf[j];
            if (e->iter_on) {
                Ed += cabs(e->Xi[j] * e->Xi[j]);
                E += cabs(e->X[j] * e->X[j]);
            }
        }
        if (e->iter_on && Ed > e->eps * E)                          /* :227-233 */
            refine(e, e->X, e->Xi, e->Xf, &t);

        real r = (t.freq - e->f_old[c]) * (real)e->frame_rate;      /* :236 */
        real rd = (r - e->dl0[c]) * (real)e->frame_rate;            /* :237 */
        if (!e->st[c] && (fabs(r) > e->thr[0] || fabs(rd) > e->thr[1])) e->st[c] = 1;   /* :240 */
        if (e->st[c] && fabs(r) < e->thr[2]) e->st[c] = 0;          /* :241 */
        real y;
        if (!e->st[c]) y = e->lp[1] * r + e->lp[2] * e->dl0[c] - e->lp[0] * e->dl1[c];  /* :246-248 */
        else y = r;                                                 /* :252 */
        e->f_old[c] = t.freq;                                       /* :256-260 */
        e->dl0[c] = r;
        e->dl1[c] = y;
        out[4 * c + 0] = 2 * t.amp / e->S;                          /* :264 */
        out[4 * c + 1] = wrap_pi(t.ph - 2 * M_PI * e->f0 * mid_fracsec);    /* :265 */
        out[4 * c + 2] = t.freq;                                    /* :263 */
        out[4 * c + 3] = y;
    }
    return 0;
}

int orc_last_km(void *h) { return ((orc_est *)h)->last_km; }

void orc_last_spectrum(void *h, double *reim)
{
    orc_est *e = (orc_est *)h;
    for (unsigned j = 0; j < e->K; j++) {
        reim[2 * j] = (double)creal(e->X[j]);
        reim[2 * j + 1] = (double)cimag(e->X[j]);
    }
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Batch runner.  The restatement processes channels one estimator per
 * (instance) like the reference; km is observable for every channel because
 * each channel is run through a 1-channel view of the instance state. */
double orc_run(const orc_config *cfg, unsigned n_channels, long n_inst, long n_frames,
               const void *windows, const double *mid, void *frames, int *km, int n_threads)
{
    if (n_inst <= 0 || n_frames <= 0 || n_channels < 1) return -1.0;
    /* one single-channel estimator per (instance, channel): identical arithmetic,
     * and it exposes the first-ipDFT peak bin of every channel. */
    long n_est = n_inst * (long)n_channels;
    orc_est **est = calloc((size_t)n_est, sizeof(orc_est *));
    if (!est) return -1.0;
    for (long i = 0; i < n_est; i++) {
        est[i] = orc_open(cfg, 1);
        if (!est[i]) {
            for (long j = 0; j < i; j++) orc_close(est[j]);
            free(est);
            return -2.0;
        }
    }
    const size_t N = est[0]->N;
    const real *win = (const real *)windows;
    real *out = (real *)frames;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
    double t0 = now_s();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n_inst; i++) {
        for (long t = 0; t < n_frames; t++) {
            size_t u = (size_t)t * (size_t)n_inst + (size_t)i;
            for (unsigned c = 0; c < n_channels; c++) {
                orc_est *e = est[i * (long)n_channels + c];
                orc_estimate(e, win + (u * n_channels + c) * N, mid[u], out + (u * n_channels + c) * 4);
                if (km) km[u * n_channels + c] = e->last_km;
            }
        }
    }
    double t1 = now_s();
    for (long i = 0; i < n_est; i++) orc_close(est[i]);
    free(est);
    return t1 - t0;
}
