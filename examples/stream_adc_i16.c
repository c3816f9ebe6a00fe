/*
 * stream_adc_i16.c — streaming 16-bit ADC counts, pipelined: the host hands over only the `hop` new int16 samples of
 * every stream per reporting interval (2 bytes per sample over PCIe instead of 8), the device promotes them exactly
 * (sample = volts_per_count[stream] * count) into the HBM rings, and the pushes are asynchronous - while push t is
 * being estimated the host fills and submits the buffer of push t+1, and reads the frames of push t-1.
 * Also emits the IEEE C37.118.2 wire format the frames would leave in: the CFG-2 frame once, then one 16-bit polar
 * data frame per instance and reporting interval.
 *
 *   gcc -Iinclude examples/stream_adc_i16.c -Lsynchropmu_b200/lib -lpmu_estimator -lm \
 *       -Wl,-rpath,$PWD/synchropmu_b200/lib -o stream_adc_i16 && ./stream_adc_i16 config/config.ini
 */
#include <math.h>
#include <stdint.h>
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

    /* a +-400 V ADC with 16-bit resolution; sample n of stream (i, c) as the converter would report it */
    const double volts_per_count = 400.0 / 32768.0;
#define VOLTS(i, c, n) (325.0 * cos(2 * M_PI * ((49.8 + 0.0125 * (i)) * ((n) / fs) + 0.1 * ((n) / fs) * ((n) / fs)) - (c) * 2 * M_PI / 3))
#define COUNT(i, c, n) ((int16_t)lrint(VOLTS(i, c, n) / volts_per_count))

    double scale[INST * CH];
    for (int s = 0; s < INST * CH; s++) scale[s] = volts_per_count;
    double *hist = malloc(sizeof(double) * INST * CH * (N - hop));   /* history is given in the batch dtype */
    for (int i = 0; i < INST; i++)
        for (int c = 0; c < CH; c++)
            for (unsigned n = 0; n < N - hop; n++)
                hist[((size_t)i * CH + c) * (N - hop) + n] = volts_per_count * COUNT(i, c, (double)n);
    if (pmub_stream_begin(b, hop, 1, hist, 0) || pmub_stream_set_scale(b, scale)) { fprintf(stderr, "%s\n", pmub_last_error()); return 2; }

    /* the configuration frame a PDC needs before the data frames: 16-bit polar phasors, integer FREQ/DFREQ */
    pmub_c37_options opt = {0};
    unsigned phunit[CH];
    for (int c = 0; c < CH; c++) phunit[c] = 1221;                   /* 0.01221 V per count: 400 V ~ 32760 counts */
    opt.format = PMUB_C37_POLAR;
    opt.idcode0 = 100; opt.f_nominal = cfg.f0; opt.phunit = phunit;
    unsigned char cfg2[512];
    const int cfg2_len = pmub_c37118_cfg2(cfg2, sizeof cfg2, CH, &opt, 0, "SUB", NULL, 1000000, (int)cfg.frame_rate);
    if (cfg2_len < 0) return 6;
    const int wire_len = pmub_c37118_frame_bytes_ex(CH, opt.format);

    int16_t *blk[2];
    double *frames[2];
    unsigned long long ticket[2] = {0, 0};
    for (int k = 0; k < 2; k++) {
        blk[k] = malloc(sizeof(int16_t) * INST * CH * hop);
        frames[k] = malloc(sizeof(double) * INST * CH * 4);
    }
    unsigned char *wire = malloc((size_t)INST * wire_len);
    double f_end = 0, rocof_end = 0;
    for (int f = 0; f <= FRAMES; f++) {
        const int k = f & 1;
        if (f < FRAMES) {
            const double n0 = (double)(N - hop) + (double)f * hop;
            for (int i = 0; i < INST; i++)
                for (int c = 0; c < CH; c++)
                    for (unsigned n = 0; n < hop; n++) blk[k][((size_t)i * CH + c) * hop + n] = COUNT(i, c, n0 + n);
            if (pmub_stream_push_ex(b, 1, blk[k], PMUB_SAMPLES_I16, NULL, frames[k], PMUB_PUSH_ASYNC, &ticket[k])) {
                fprintf(stderr, "%s\n", pmub_last_error());
                return 3;
            }
        }
        if (f > 0) {   /* the previous push: wait for it, use its frames while the current one runs */
            const int p = (f - 1) & 1;
            if (pmub_stream_wait(b, ticket[p])) return 4;
            opt.soc = 1700000000u + (unsigned)(f - 1) / cfg.frame_rate;
            opt.fracsec = ((unsigned)(f - 1) % cfg.frame_rate) * (1000000u / cfg.frame_rate);
            if (pmub_c37118_pack_ex(PMUB_F64, frames[p], wire, INST, CH, &opt, NULL)) return 5;
            f_end = frames[p][2];
            rocof_end = frames[p][3];
            if ((f - 1) % 10 == 9)
                printf("frame %2d instance 0 ch0: amp %.4f V ph %+.6f freq %.6f rocof %+.4f | wire frame %d bytes, SYNC %02x%02x\n",
                       f - 1, frames[p][0], frames[p][1], frames[p][2], frames[p][3], wire_len, wire[0], wire[1]);
        }
    }
    printf("CFG-2 %d bytes, data frames %d bytes each; instance 0 frequency after %d frames: %.4f Hz (started at 49.8, ramp 0.2 Hz/s), ROCOF %.4f Hz/s\n",
           cfg2_len, wire_len, FRAMES, f_end, rocof_end);
    for (int k = 0; k < 2; k++) { free(blk[k]); free(frames[k]); }
    free(hist); free(wire);
    if (pmub_destroy(b)) return 7;
    return fabs(f_end - 50.004) < 0.02 ? 0 : 8;
}
