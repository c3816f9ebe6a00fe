/*
 * stream_from_ini.c — the streaming front end: the library keeps each stream's recent samples in a ring buffer in
 * HBM; per reporting interval the caller hands over only the `hop` new samples of every stream and gets one frame per
 * stream back (what test/test.c of the reference does by cutting a full window per call, without the 4x overlap
 * crossing PCIe again).
 *
 *   gcc -Iinclude examples/stream_from_ini.c -Lsynchropmu_b200/lib -lpmu_estimator -lm \
 *       -Wl,-rpath,$PWD/synchropmu_b200/lib -o stream_from_ini && ./stream_from_ini config/config.ini
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pmu_batch.h"

int main(int argc, char **argv)
{
    const char *ini = argc > 1 ? argv[1] : "config/config.ini";
    enum { INST = 32, CH = 3, FRAMES = 50 };
    pmub_batch *b = NULL;
    if (pmub_create_ini(&b, ini, PMUB_F64, INST, CH, -1)) { fprintf(stderr, "%s\n", pmub_last_error()); return 1; }
    pmub_info info;
    pmub_get_info(b, &info);
    pmub_config cfg;
    if (pmub_parse_ini(ini, &cfg)) return 1;
    const unsigned N = info.win_len, hop = cfg.fs / cfg.frame_rate;
    const double fs = cfg.fs;

    /* sample n of stream (i, c): a 3-phase set per instance, each instance at its own frequency, ramping at 0.2 Hz/s */
#define SAMPLE(i, c, n) (100.0 * cos(2 * M_PI * ((49.8 + 0.0125 * (i)) * ((n) / fs) + 0.1 * ((n) / fs) * ((n) / fs)) - (c) * 2 * M_PI / 3))

    /* the win_len - hop samples that precede the first push */
    double *hist = malloc(sizeof(double) * INST * CH * (N - hop));
    for (int i = 0; i < INST; i++)
        for (int c = 0; c < CH; c++)
            for (unsigned n = 0; n < N - hop; n++) hist[((size_t)i * CH + c) * (N - hop) + n] = SAMPLE(i, c, (double)n);
    if (pmub_stream_begin(b, hop, 1, hist, 0)) { fprintf(stderr, "%s\n", pmub_last_error()); return 2; }

    double *blk = malloc(sizeof(double) * INST * CH * hop);
    double *frames = malloc(sizeof(double) * INST * CH * 4);
    for (int f = 0; f < FRAMES; f++) {
        const double n0 = (double)(N - hop) + (double)f * hop;
        for (int i = 0; i < INST; i++)
            for (int c = 0; c < CH; c++)
                for (unsigned n = 0; n < hop; n++) blk[((size_t)i * CH + c) * hop + n] = SAMPLE(i, c, n0 + n);
        /* mid_fracsec = NULL: derived on the device from the running sample counter */
        if (pmub_stream_push(b, 1, blk, NULL, frames)) { fprintf(stderr, "%s\n", pmub_last_error()); return 3; }
        if (f % 10 == 9)
            printf("frame %2d instance 0 ch0: amp %.6f ph %+.6f freq %.6f rocof %+.4f | instance %d ch2: freq %.6f\n", f,
                   frames[0], frames[1], frames[2], frames[3], INST - 1, frames[(((size_t)INST - 1) * CH + 2) * 4 + 2]);
    }
    /* the centre of the last window is 1.02 s into the signal: 49.8 + 0.2 * 1.02 = 50.004 Hz */
    const double f_end = frames[2];
    printf("instance 0 frequency after %d frames: %.4f Hz (started at 49.8, ramp 0.2 Hz/s), ROCOF %.4f Hz/s\n", FRAMES, f_end, frames[3]);
    free(hist); free(blk); free(frames);
    if (pmub_destroy(b)) return 4;
    return fabs(f_end - 50.004) < 0.02 ? 0 : 5;
}
