/*
 * batch_from_ini.c — the batched entry: many independent estimator instances x channels in one
 * call, configured from the same .ini file the reference reads (config/config.ini).
 *
 *   gcc -DNUM_CHANLS=6 -Iinclude examples/batch_from_ini.c -Lsynchropmu_b200/lib -lpmu_estimator_c6 -lm \
 *       -Wl,-rpath,$PWD/synchropmu_b200/lib -o batch_from_ini && ./batch_from_ini config/config.ini
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pmu_estimator.h"

int main(int argc, char **argv)
{
    char ini[512];
    snprintf(ini, sizeof ini, "%s", argc > 1 ? argv[1] : "config/config.ini");
    enum { FS = 25600, N = 2048, INST = 64, CH = 6 };
    float_p *win = malloc(sizeof(float_p) * INST * CH * N);
    float_p mid[INST];
    pmu_frame *out = malloc(sizeof(pmu_frame) * INST * CH);
    for (int i = 0; i < INST; i++) {
        double f = 49.5 + i / 64.0;
        mid[i] = 0.04;
        for (int c = 0; c < CH; c++)
            for (int j = 0; j < N; j++)
                win[((size_t)i * CH + c) * N + j] = (c < 3 ? 325.27 : 14.14) * cos(2 * M_PI * f * j / FS + c * 2 * M_PI / 3);
    }
    pmu_batch *b = NULL;
    if (pmu_batch_init(&b, ini, CONFIG_FROM_INI, INST, CH, -1)) return 1;
    if (pmu_estimate_batch(b, win, mid, out)) return 2;
    for (int i = 0; i < INST; i += 21)
        printf("instance %2d ch0: amp %.6f ph %.6f freq %.9f rocof %.6f\n", i, (double)out[i * CH].synchrophasor.amp,
               (double)out[i * CH].synchrophasor.ph, (double)out[i * CH].synchrophasor.freq, (double)out[i * CH].rocof);
    if (pmu_batch_deinit(b)) return 3;
    free(win); free(out);
    return 0;
}
