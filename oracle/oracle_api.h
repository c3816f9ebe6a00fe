/*
 * oracle_api.h — the one C interface both checkers export.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load these libraries.  The shipped library
 * (synchropmu_b200/csrc -> libpmu_estimator*.so) never links or calls them.
 *
 * Two implementations sit behind this header:
 *   - oracle/_ref/libref_<f64|f32>_c<C>.so : the UNMODIFIED reference
 *     (src/pmu_estimator.c + libs/iniparser, compiled from /root/reference by
 *     oracle/Makefile) wrapped by ref_harness.c.  NUM_CHANLS and float_p are
 *     compile-time there, hence one .so per (float type, channel count).
 *   - oracle/_build/liboracle_<f64|f32>.so : pmu_oracle.c, our plain-C
 *     restatement of the same algorithm (runtime channel count).
 *
 * The prefix macro ORC_FN() gives the exported names: refh_* or orc_*.
 */
#ifndef ORACLE_API_H
#define ORACLE_API_H

#ifdef __cplusplus
extern "C" {
#endif

/* Configuration in a float-type independent form (always doubles).  Field
 * meaning follows estimator_config, reference src/pmu_estimator.h:30-43. */
typedef struct {
    unsigned n_cycles;
    unsigned f0;
    unsigned frame_rate;
    unsigned fs;
    unsigned n_bins;
    unsigned P;
    unsigned Q;
    unsigned iter_eipdft;
    double interf_trig;
    double rocof_thresh[3];
    double rocof_low_pass_coeffs[3];
} orc_config;

#ifndef ORC_PREFIX
#define ORC_PREFIX orc
#endif
#define ORC_CAT2(a, b) a##_##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC_FN(name) ORC_CAT(ORC_PREFIX, name)

/* sizeof(float_p) of this build (4 or 8) and its channel count
 * (0 = runtime-selectable, the restatement). */
void ORC_FN(info)(int *float_size, int *num_chanls);

/* Open one estimator instance.  n_channels is ignored (must match) by the
 * reference build.  NULL on invalid configuration / unreadable ini. */
void *ORC_FN(open)(const orc_config *cfg, unsigned n_channels);
void *ORC_FN(open_ini)(const char *ini_path, unsigned n_channels);
void ORC_FN(close)(void *h);

unsigned ORC_FN(win_len)(void *h);
unsigned ORC_FN(n_bins)(void *h);
double ORC_FN(norm_factor)(void *h);

/* One pmu_estimate call.  windows: float_p[C][win_len]; frames: float_p[C][4]
 * = {amp, ph, freq, rocof}.  Returns the library's return code. */
int ORC_FN(estimate)(void *h, const void *windows, double mid_fracsec, void *frames);

/* Peak bin (km) chosen by the FIRST ipDFT call of the last channel processed,
 * reference find3LargestIndx rule (src/pmu_estimator.c:705-720). */
int ORC_FN(last_km)(void *h);
/* Windowed spectrum (n_bins complex values as re,im doubles) of the last
 * channel processed, i.e. what pmue_fft_r produced (src/pmu_estimator.c:200). */
void ORC_FN(last_spectrum)(void *h, double *reim);

/* Batch runner: n_inst independent instances, n_frames consecutive frames each.
 *   windows : float_p[n_frames][n_inst][C][win_len]
 *   mid     : double [n_frames][n_inst]
 *   frames  : float_p[n_frames][n_inst][C][4]
 *   km      : int    [n_frames][n_inst][C]  or NULL (-1 where not observable)
 * Instances are spread over n_threads OpenMP threads (<=0: all cores).
 * Returns the wall seconds spent in the estimate loop only (init excluded),
 * or a negative value on error. */
double ORC_FN(run)(const orc_config *cfg, unsigned n_channels, long n_inst, long n_frames,
                   const void *windows, const double *mid, void *frames, int *km,
                   int n_threads);

#ifdef __cplusplus
}
#endif
#endif
