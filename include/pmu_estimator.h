/*
 * pmu_estimator.h — drop-in C API of the B200-native synchrophasor estimator.
 *
 * Source- and binary-compatible with SynchroPMU's public header
 * (reference src/pmu_estimator.h:27-128 and the type macros of src/func_stubs.h:33-47):
 * the same four functions, the same caller-owned `pmu_context`, the same struct
 * layouts byte for byte (checked by tests/test_host_logic.py against the reference's own header),
 * so existing C callers (test/test.c, examples/simple_example*.c) and the reference's
 * ctypes wrapper (python/pmu_estimator.py) link unchanged.  Behind it every estimate
 * runs the hand-written sm_100a kernel; there is no CPU implementation in this library.
 *
 * Compile-time ABI switches (as in the reference, where they are `NUM_CHANLS`,
 * CMakeLists.txt:8-12, and the `float_p` macro, src/func_stubs.h:38):
 *     -DNUM_CHANLS=<n>        channels per estimator instance   (default 1)
 *     -DPMU_FLOAT_P=float     single-precision ABI              (default double)
 * and the matching library:  libpmu_estimator.so (double, 1 channel — the one the
 * reference's Python wrapper needs), libpmu_estimator_c6.so, libpmu_estimator_f32.so,
 * libpmu_estimator_f32_c6.so.  See INTEGRATION.md.
 *
 * NOTE - default width.  The reference's func_stubs.h:38 defines float_p as `float` (and has to be edited to get a
 * double build); THIS header defaults to `double`, the width the reference's own Python wrapper hard-codes
 * (python/pmu_estimator.py:7-14).  A caller compiled against the reference's unmodified headers (float) must link
 * libpmu_estimator_f32*.so - or be recompiled with -DPMU_FLOAT_P=float against this header.  Linking it to the
 * double library would hand float windows to a double ABI; pmu_abi_float_bytes() (4 or 8) lets a caller check the
 * width of the library it actually loaded.
 */
#ifndef PMU_ESTIMATOR_H
#define PMU_ESTIMATOR_H

#include <stdio.h>
#include <math.h>

#ifndef NUM_CHANLS
#define NUM_CHANLS 1
#endif

#ifndef PMU_FLOAT_P
#define PMU_FLOAT_P double
#endif

/* the reference's type vocabulary (src/func_stubs.h:33-47) */
#ifdef __cplusplus
#define PMU_BOOL_T bool /* same size and values as C's _Bool */
#else
#include <complex.h>
#define PMU_BOOL_T _Bool
#endif
#ifndef M_PI_p
#define M_PI_p M_PI
#endif
#ifndef float_p
#define float_p PMU_FLOAT_P
#endif
#ifndef uint_p
#define uint_p unsigned int
#endif
#ifndef bool_p
#define bool_p PMU_BOOL_T
#endif
#ifndef complex_p
#define complex_p _Complex
#endif

#ifdef __cplusplus
extern "C" {
/* C++ has no _Complex: the complex members keep their size/alignment as opaque slots */
#define PMU_CPLX_PTR void *
typedef struct { float_p re, im; } pmu_complex_slot;
#else
#define PMU_CPLX_PTR float_p complex_p *
typedef float_p complex_p pmu_complex_slot;
#endif

#define CONFIG_FROM_INI 1
#define CONFIG_FROM_STRUCT 0

/* ---- user-facing types ---------------------------------------------------------- */

/* estimator parameters handed to pmu_init(..., CONFIG_FROM_STRUCT) */
typedef struct
{
    uint_p n_cycles;      /* window length in nominal cycles          */
    uint_p f0;            /* nominal frequency, 50 or 60              */
    uint_p frame_rate;    /* frames per second                        */
    uint_p fs;            /* sample rate, a multiple of f0            */
    uint_p n_bins;        /* DFT bins used by the estimator           */
    uint_p P;             /* e-IpDFT compensation passes              */
    uint_p Q;             /* interference iterations                  */
    bool_p iter_eipdft;   /* enable the interference iterations       */
    float_p interf_trig;  /* residual-energy trigger, ]0, 1]          */
    float_p rocof_thresh[3];
    float_p rocof_low_pass_coeffs[3];
} estimator_config;

typedef struct
{
    float_p amp;
    float_p ph;
    float_p freq;
} phasor;

typedef struct
{
    phasor synchrophasor;
    float_p rocof;
} pmu_frame;

/* ---- pmu_context: caller-allocated, caller-zeroed, same bytes as the reference ---- */

typedef struct
{
    uint_p win_len;
    uint_p n_cycles;
    uint_p f0;
    uint_p frame_rate;
    uint_p fs;
    uint_p n_bins;
    uint_p P;
    uint_p Q;
    bool_p iter_eipdft_enabled;
    float_p interf_trig;
    float_p df;
    float_p norm_factor;
    phasor phasor;
} SynchrophasorEstimatorParams;

/* In the reference these are malloc'ed scratch arrays.  Here nothing is computed on
 * the host: `Xf` carries the opaque handle of the device-side estimator, the other
 * slots stay NULL.  Callers never touched them (README "pmu_context" section). */
typedef struct
{
    PMU_CPLX_PTR Xf;
    PMU_CPLX_PTR Xi;
    PMU_CPLX_PTR dftbins;
    float_p *hann_window;
    float_p *signal_windows[NUM_CHANLS];
} InternalBuffers;

typedef struct
{
    float_p C0;
    float_p C1;
    float_p C2;
    float_p C3;
    float_p C4;
    pmu_complex_slot C5;
    pmu_complex_slot C6;
    float_p inv_norm_factor;
} HanningTransformConstants;

/* mirrored from HBM after every pmu_estimate so callers that inspect it keep working */
typedef struct
{
    float_p freq_old[NUM_CHANLS];
    float_p thresholds[3];
    float_p low_pass_coeff[3];
    float_p delay_line[NUM_CHANLS][2];
    bool_p state[NUM_CHANLS];
} RocoFEstimationStates;

typedef struct pmu_context
{
    SynchrophasorEstimatorParams synch_params;
    InternalBuffers buff_params;
    HanningTransformConstants hann_params;
    RocoFEstimationStates rocof_params;
    bool_p pmu_initialized;
} pmu_context;

/* ---- the reference API (src/pmu_estimator.h:119-128) ------------------------------- */

/* cfg: estimator_config* (CONFIG_FROM_STRUCT) or char* path (CONFIG_FROM_INI).
 * 0 on success, -1 if already initialised / invalid configuration / no CUDA device. */
int pmu_init(pmu_context *ctx, void *cfg, bool_p config_from_ini);

/* in_signal_windows: float_p[NUM_CHANLS][win_len] (host memory);
 * out_frame: pmu_frame[NUM_CHANLS].  0 on success, -1 if not initialised. */
int pmu_estimate(pmu_context *ctx, float_p *in_signal_windows, float_p mid_fracsec, pmu_frame *out_frame);

int pmu_deinit(pmu_context *ctx);

int pmu_dump_frame(pmu_frame *frame, FILE *stream);

/* sizeof(float_p) of the library that was actually loaded (4 or 8): lets a caller built against the reference's
 * float headers detect that it was linked to the double library (see the NOTE at the top).  Not in the reference. */
int pmu_abi_float_bytes(void);

/* ---- batched entry (new; same layouts repeated per instance) ----------------------- */
/* pmu_estimate == a batch of one instance.  Typed wrappers over pmu_batch.h. */
typedef struct pmub_batch pmu_batch;

int pmu_batch_init(pmu_batch **b, void *cfg, bool_p config_from_ini, uint_p n_instances,
                   uint_p n_channels, int device);
/* windows float_p[n_instances][n_channels][win_len], mid_fracsec float_p[n_instances],
 * out pmu_frame[n_instances][n_channels]; host or device pointers. */
int pmu_estimate_batch(pmu_batch *b, const float_p *windows, const float_p *mid_fracsec, pmu_frame *out);
int pmu_batch_deinit(pmu_batch *b);

#ifdef __cplusplus
}
#endif

#endif /* PMU_ESTIMATOR_H */
