/*
 * dropin_single.c — the reference's usage pattern (examples/simple_example2.c in SynchroPMU:
 * struct config, one window of a 50 Hz / 2 V / 1 rad tone, pmu_init -> pmu_estimate ->
 * pmu_dump_frame -> pmu_deinit) against the B200 library.  Builds with the default ABI variant:
 *
 *   gcc -Iinclude examples/dropin_single.c -Lsynchropmu_b200/lib -lpmu_estimator -lm \
 *       -Wl,-rpath,$PWD/synchropmu_b200/lib -o dropin_single
 */
#include <math.h>
#include <stdio.h>

#include "pmu_estimator.h"

int main(void)
{
    enum { FS = 25600, N = 2048 };
    static float_p window[NUM_CHANLS][N];
    for (int c = 0; c < NUM_CHANLS; c++)
        for (int i = 0; i < N; i++)
            window[c][i] = 2.0 * cos(2 * M_PI * 50.0 * i / FS + 1.0 + c * 2 * M_PI / 3);

    estimator_config cfg;
    cfg.n_cycles = 4; cfg.f0 = 50; cfg.frame_rate = 50; cfg.fs = FS; cfg.n_bins = 11;
    cfg.P = 3; cfg.Q = 3; cfg.interf_trig = 0.0033; cfg.iter_eipdft = 1;
    cfg.rocof_thresh[0] = 3; cfg.rocof_thresh[1] = 25; cfg.rocof_thresh[2] = 0.035;
    cfg.rocof_low_pass_coeffs[0] = 0.5913; cfg.rocof_low_pass_coeffs[1] = 0.2043; cfg.rocof_low_pass_coeffs[2] = 0.2043;

    pmu_context ctx = {0};
    if (pmu_init(&ctx, &cfg, CONFIG_FROM_STRUCT)) return 1;
    if (pmu_init(&ctx, &cfg, CONFIG_FROM_STRUCT) != -1) return 2;   /* second init must be refused */

    pmu_frame frame[NUM_CHANLS];
    for (int k = 0; k < 3; k++)                                     /* three consecutive calls: ROCOF state */
        if (pmu_estimate(&ctx, (float_p *)window, 0.0, frame)) return 3;
    for (int c = 0; c < NUM_CHANLS; c++) pmu_dump_frame(&frame[c], stdout);

    if (pmu_deinit(&ctx)) return 4;
    if (pmu_deinit(&ctx) != -1) return 5;
    return 0;
}
