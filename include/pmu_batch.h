/*
 * pmu_batch.h — C ABI of the B200 batched synchrophasor estimator (type-generic core).
 *
 * Plain pointers and sizes only; no CUDA or torch types in any signature.  Every
 * libpmu_estimator*.so exports these entry points next to the drop-in
 * pmu_init / pmu_estimate / pmu_deinit / pmu_dump_frame of pmu_estimator.h.
 *
 * What each one replaces in the reference (chemsallioua/SynchroPMU, paths relative to the
 * reference tree):
 *   pmub_create / pmub_create_ini  <- pmu_init            src/pmu_estimator.c:92-167
 *                                     config_estimator    src/pmu_estimator.c:327-384
 *   pmub_validate                  <- check_config_validity src/pmu_estimator.c:386-449
 *   pmub_parse_ini                 <- config_from_file    src/pmu_estimator.c:451-484
 *                                     (+ libs/iniparser line rules, iniparser.c:620-700)
 *   pmub_estimate[_async]          <- pmu_estimate        src/pmu_estimator.c:170-271,
 *                                     once per (instance, channel) of the batch
 *   pmub_destroy                   <- pmu_deinit          src/pmu_estimator.c:274-303
 *   pmub_reset / get_state / set_state <- the ROCOF members of pmu_context
 *                                     (RocoFEstimationStates, src/pmu_estimator.h:100-107)
 *
 * There is NO CPU fallback: creation fails (-1) without a CUDA device, and estimation
 * only ever runs the sm_100a kernel.
 *
 * Return convention: 0 = ok, -1 = error (message on stderr and in pmub_last_error()),
 * like the reference (src/pmu_estimator.c:96-105, 176-180).
 */
#ifndef PMU_BATCH_H
#define PMU_BATCH_H

#ifdef __cplusplus
extern "C" {
#endif

#define PMUB_F32 4 /* float_p = float  : windows, mid_fracsec and frames are float  */
#define PMUB_F64 8 /* float_p = double : ... are double                              */

typedef struct pmub_batch pmub_batch;

/* estimator_config (src/pmu_estimator.h:30-43) with the float type fixed to double so
 * one struct serves both builds; values are rounded to float_p on creation. */
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
} pmub_config;

typedef struct {
    unsigned win_len;      /* n_cycles * fs / f0                                         */
    unsigned n_bins;
    unsigned n_instances;
    unsigned n_channels;
    int dtype;             /* PMUB_F32 / PMUB_F64                                        */
    int device;
    double df;             /* fs / win_len                                               */
    double norm_factor;    /* sequential Hann sum, src/pmu_estimator.c:549-560           */
    /* how the kernel was specialised for this configuration */
    int dft_radix;         /* M: in-register real-FFT codelet size; 0 = the general kernel
                              (n_bins > 23: one warp per window, chunked pruned FFT)      */
    int dft_bins;          /* KC: rectangular bins accumulated per lane                  */
    int n_stages;          /* shared-memory ring depth                                   */
    int stage_bytes;
    int slabs_per_window;
    int load_mode;         /* 0 = TMA bulk copies, 1 = per-element cp.async, 2 = plain loads
                              (general kernel)                                            */
    int grid;              /* CTAs launched for a full batch                             */
    int block;             /* threads per CTA                                            */
    int smem_bytes;        /* dynamic shared memory per CTA                              */
    double bytes_per_estimate; /* algorithmic HBM bytes per (window x channel)           */
} pmub_info;

/* ---- host logic (no GPU needed) ------------------------------------------------ */
int pmub_validate(const pmub_config *cfg);
int pmub_parse_ini(const char *ini_path, pmub_config *out);
const char *pmub_last_error(void);

/* ---- lifecycle ----------------------------------------------------------------- */
/* device < 0: use the calling thread's current CUDA device. */
int pmub_create(pmub_batch **out, const pmub_config *cfg, int dtype, unsigned n_instances,
                unsigned n_channels, int device);
int pmub_create_ini(pmub_batch **out, const char *ini_path, int dtype, unsigned n_instances,
                    unsigned n_channels, int device);
int pmub_destroy(pmub_batch *b);
int pmub_get_info(const pmub_batch *b, pmub_info *out);

/* ---- the hot path -------------------------------------------------------------- */
/* One frame for every (instance, channel):
 *   windows     float_p [n_instances][n_channels][win_len]
 *   mid_fracsec float_p [n_instances]
 *   frames      float_p [n_instances][n_channels][4] = {amp, ph, freq, rocof}
 *               (= pmu_frame[n_instances][n_channels], src/pmu_estimator.h:45-56)
 * Each pointer may be host memory or device memory (detected).  Host inputs are copied
 * in chunks overlapped with the kernel.  Synchronous: results are complete on return. */
int pmub_estimate(pmub_batch *b, const void *windows, const void *mid_fracsec, void *frames);
/* Device pointers only; enqueues on the batch's stream and returns immediately. */
int pmub_estimate_async(pmub_batch *b, const void *d_windows, const void *d_mid_fracsec,
                        void *d_frames);
/* Several consecutive frames of the same streams in one call (e.g. a replay, or a streaming front
 * end that hands over a block of frames): exactly equivalent to calling pmub_estimate n_frames
 * times, frame f reading
 *   windows     float_p [n_frames][n_instances][n_channels][win_len]
 *   mid_fracsec float_p [n_frames][n_instances]
 * and writing frames float_p [n_frames][n_instances][n_channels][4]; the ROCOF state of every
 * stream is carried from frame to frame inside the kernel, so the launch ramp-up and the tail are
 * paid once per call instead of once per frame. */
int pmub_estimate_frames(pmub_batch *b, int n_frames, const void *windows, const void *mid_fracsec,
                         void *frames);
int pmub_estimate_frames_async(pmub_batch *b, int n_frames, const void *d_windows,
                               const void *d_mid_fracsec, void *d_frames);
/* ---- streaming front end ---------------------------------------------------------
 * In the reference the caller cuts every window out of its sample stream (test/test.c:49-62)
 * and hands over win_len samples per frame although consecutive windows overlap by
 * win_len - hop samples.  Here the library keeps each stream's recent samples in a ring buffer
 * in HBM: after pmub_stream_begin, every pushed block of `hop` new samples per stream yields one
 * frame (the window = the last win_len samples), so only the new samples cross PCIe.
 *   hop                 samples between frames; 0 = fs / frame_rate
 *   max_frames_per_push upper bound for n_frames of pmub_stream_push
 *   history             float_p [n_instances][n_channels][win_len - hop] samples preceding the
 *                       first push (host or device), NULL = zeros
 *   first_sample_index  absolute index of history[0] counted from the top of a second; used for
 *                       mid_fracsec when pmub_stream_push is given mid_fracsec == NULL
 * pmub_stream_push: new_samples float_p [n_frames][n_instances][n_channels][hop],
 *   mid_fracsec float_p [n_frames][n_instances] or NULL (= centre of each window modulo one second,
 *   computed on the device), frames float_p [n_frames][n_instances][n_channels][4].
 * Results are identical to pmub_estimate on the explicitly extracted windows. */
int pmub_stream_begin(pmub_batch *b, unsigned hop, unsigned max_frames_per_push, const void *history,
                      unsigned long long first_sample_index);
int pmub_stream_push(pmub_batch *b, int n_frames, const void *new_samples, const void *mid_fracsec,
                     void *frames);
/* The general push: narrow sample formats and/or asynchronous operation.
 *   sample_format  PMUB_SAMPLES_NATIVE  float_p samples (what pmub_stream_push takes)
 *                  PMUB_SAMPLES_F32     float samples promoted on the device (for the double estimator)
 *                  PMUB_SAMPLES_I16     int16 ADC counts; sample = scale[stream] * count with the factors given to
 *                                       pmub_stream_set_scale (host double [n_instances * n_channels], NULL = 1).
 *                  Promotion is exact, so the frames equal what pmub_estimate returns for windows holding the same
 *                  promoted values; only 4 or 2 bytes per sample cross PCIe.
 *   flags          PMUB_PUSH_ASYNC: return as soon as the work is queued.  `samples`, `mid_fracsec` and `frames` must
 *                  stay valid until pmub_stream_wait(ticket) (or pmub_sync) returns; with two alternating pinned
 *                  sample buffers the host-to-device copy of push t+1 overlaps the kernels of push t.  At most four
 *                  pushes are kept in flight (a fifth waits for the oldest).
 *   ticket         (may be NULL) receives the push's sequence number for pmub_stream_wait. */
#define PMUB_SAMPLES_NATIVE 0
#define PMUB_SAMPLES_F32 1
#define PMUB_SAMPLES_I16 2
#define PMUB_PUSH_ASYNC 1u
/* PMUB_PUSH_IN_PLACE: `samples` is ignored (may be NULL) - device code of the caller has already written the n_frames * hop
 * new samples of every stream into the rings, starting at the write position pmub_stream_ring_info reports (ring of stream
 * w = base + w * pmub_stream_ring_pitch(b), in float_p; positions [0, cap_samples), wrapping at cap_samples).  No ingest
 * pass: the estimator launch is the only kernel (plans with several slabs per window keep a mirror of each ring's first
 * win_len samples behind its end, which the library brings up to date with a small kernel first). */
#define PMUB_PUSH_IN_PLACE 2u
int pmub_stream_set_scale(pmub_batch *b, const double *scale);
int pmub_stream_push_ex(pmub_batch *b, int n_frames, const void *samples, int sample_format,
                        const void *mid_fracsec, void *frames, unsigned flags, unsigned long long *ticket);
int pmub_stream_wait(pmub_batch *b, unsigned long long ticket);
int pmub_stream_ring_info(pmub_batch *b, void **base, long long *cap_samples, long long *write_pos);
/* samples between consecutive ring rows (>= cap_samples; -1 before pmub_stream_begin): row w starts at base + w * pitch */
long long pmub_stream_ring_pitch(const pmub_batch *b);
int pmub_sync(pmub_batch *b);
/* Run on the caller's CUDA stream (cudaStream_t passed as void*); NULL = own stream. */
int pmub_set_stream(pmub_batch *b, void *cuda_stream);
/* Let consecutive launches of this batch OVERLAP (off by default).  By enabling it the caller promises that nothing enqueued
 * on the batch's stream between one launch of this batch and the next writes what that next launch reads - device windows,
 * device mid_fracsec, ring samples written in place (PMUB_PUSH_IN_PLACE): they were complete before the previous launch was
 * enqueued, or their writer is ordered through an event, another stream or the host.  The estimator kernel then waits for
 * its predecessor only where it consumes that predecessor's results (the ROCOF state of the previous frame, the frame and
 * state stores) instead of at its top: on every SM the previous launch has already left, the transforms of the next one
 * start while the last CTAs of the previous one are still in their estimator tail.  Results are identical; calls that
 * copy from host memory or ingest samples through the library's own kernels are not affected.  Applies to
 * pmub_estimate_async / pmub_estimate_frames_async and to in-place pmub_stream_push_ex. */
int pmub_set_overlap(pmub_batch *b, int enable);
/* Kernels launched by this batch so far. */
long long pmub_launch_count(const pmub_batch *b);

/* ---- ROCOF state (checkpoint / resume) ------------------------------------------ */
int pmub_reset(pmub_batch *b); /* freq_old = f0, delay line = 0, state = 0 (:141-147) */
/* host arrays of n_instances * n_channels entries each */
int pmub_get_state(pmub_batch *b, double *freq_old, double *dl0, double *dl1, unsigned char *state);
int pmub_set_state(pmub_batch *b, const double *freq_old, const double *dl0, const double *dl1,
                   const unsigned char *state);

/* ---- consumers of the frame stream (not in the reference) ------
 * Both take frames float_p [n_instances][n_channels][4] in host or device memory.
 *
 * pmub_sequence_components: channels are grouped in threes (a, b, c); out is
 *   float_p [n_instances][n_channels/3][3][2] = (amp, ph) of the positive, negative and
 *   zero sequence phasors.
 * pmub_c37118_pack: one IEEE C37.118.2-2011 data frame per instance, big-endian:
 *   SYNC 0xAA01 | FRAMESIZE | IDCODE (idcode0 + instance) | SOC | FRACSEC | STAT |
 *   n_channels x {magnitude, angle[rad]} as IEEE float32 | FREQ [Hz] | DFREQ [Hz/s] | CHK
 *   (CRC-CCITT, 0xFFFF seed), pmub_c37118_frame_bytes(n_channels) = 26 + 8 n_channels bytes each.
 *   FREQ/DFREQ come from channel 0 of the instance. */
int pmub_sequence_components(int dtype, const void *frames, void *out, unsigned n_instances,
                             unsigned n_channels);
int pmub_c37118_frame_bytes(unsigned n_channels);
int pmub_c37118_pack(int dtype, const void *frames, void *out, unsigned n_instances,
                     unsigned n_channels, unsigned idcode0, unsigned soc, unsigned fracsec,
                     unsigned stat);

/* The full C37.118.2-2011 format family, selected by the FORMAT word the CFG-2 frame announces:
 *   bit 0  phasors: 0 rectangular (real, imaginary)      1 polar (magnitude, angle)
 *   bit 1  phasors: 0 16-bit integer, scaled by PHUNIT    1 IEEE float
 *   bit 3  FREQ / DFREQ: 0 16-bit integer (deviation from nominal in mHz, ROCOF in 0.01 Hz/s)   1 IEEE float (Hz, Hz/s)
 * 16-bit phasor counts = value / (phunit[c] * 1e-5) (magnitudes unsigned, rectangular parts signed, saturating;
 * polar angles = radians * 1e4).  pmub_c37118_frame_bytes_ex(n_ch, format) = 16 + n_ch * (4 | 8) + 2 * (2 | 4) + 2.
 * `cuda_stream` (may be NULL = the legacy default stream) orders the packer after the estimator on the same stream -
 * pass what was given to pmub_set_stream; with device `frames` and `out` the call does not synchronise.  Host buffers go
 * through a staging area kept between calls. */
#define PMUB_C37_POLAR 1u
#define PMUB_C37_FLOAT_PHASORS 2u
#define PMUB_C37_FLOAT_FREQ 8u
#define PMUB_C37_FLOAT_POLAR (PMUB_C37_POLAR | PMUB_C37_FLOAT_PHASORS | PMUB_C37_FLOAT_FREQ)
typedef struct {
    unsigned format;      /* FORMAT word, bits 0..3 */
    unsigned idcode0;     /* IDCODE of instance 0; instance i uses idcode0 + i */
    unsigned soc, fracsec, stat;
    unsigned f_nominal;   /* 50 or 60 (0 = 50): reference of the integer FREQ format, FNOM of CFG-2 */
    const unsigned *phunit;            /* host, [n_channels] 24-bit conversion factors (1e-5 units per count); NULL = 1 */
    const unsigned char *phunit_is_current; /* host, [n_channels] 0 = voltage, 1 = current (CFG-2 only); NULL = voltages */
    unsigned cfg_count;   /* CFGCNT of CFG-2 */
} pmub_c37_options;
int pmub_c37118_frame_bytes_ex(unsigned n_channels, unsigned format);
int pmub_c37118_pack_ex(int dtype, const void *frames, void *out, unsigned n_instances, unsigned n_channels,
                        const pmub_c37_options *opt, void *cuda_stream);
/* CFG-2 configuration frame (host side) of instance `instance`: one PMU block with its n_channels phasors, no analog
 * or digital words.  station_name NULL = "PMU<instance>", channel_names NULL = "PH00".."PHnn", time_base = FRACSEC
 * counts per second, data_rate = frames per second.  Returns the number of bytes written (<= cap) or -1. */
int pmub_c37118_cfg2(unsigned char *out, unsigned cap, unsigned n_channels, const pmub_c37_options *opt,
                     unsigned instance, const char *station_name, const char *const *channel_names,
                     unsigned time_base, int data_rate);
int pmub_sequence_components_on(int dtype, const void *frames, void *out, unsigned n_instances,
                                unsigned n_channels, void *cuda_stream);

/* ---- frames straight into another GPU's memory (optional gather, no collective) ---------
 * rank 0: pmub_peer_alloc -> device buffer + 64-byte CUDA IPC handle (ship the handle to the other
 * processes by any means); rank r: pmub_peer_open(handle) -> a pointer valid in r's address space;
 * pass (that pointer + r's offset) as `frames` to pmub_estimate*: the kernel's frame stores then go
 * over NVLink / NVSwitch directly into rank 0's HBM.  Close with pmub_peer_close, free on the owner
 * with pmub_peer_free. */
#define PMUB_IPC_HANDLE_BYTES 64
int pmub_peer_alloc(size_t bytes, int device, void **dptr, unsigned char *handle64);
int pmub_peer_open(const unsigned char *handle64, int device, void **dptr);
int pmub_peer_close(void *dptr);
int pmub_peer_read(void *host_dst, const void *dptr, size_t bytes); /* synchronous D2H of (part of) such a buffer */
int pmub_peer_free(void *dptr);

/* ---- parity / debug observability ------------------------------------------------ */
/* After pmub_set_debug(b, 1) every estimate also records, per (instance, channel):
 *   flags    int    : bits 0-15 = peak bin km of the first ipDFT, bit 16 = interference
 *                     iterations were triggered
 *   spectrum double [n_bins][2] : the Hann-windowed DFT bins (what pmue_fft_r yields,
 *                     src/pmu_estimator.c:200)
 * fetched (to host arrays, either may be NULL) by pmub_get_debug. */
/* Batched form of the reference's inner plug point pmue_fft_r(in, out, out_len, n_bins) (src/func_stubs.h:60-62):
 * out[w][k] = sum_n in[w][n] exp(-2 pi i k n / win_len), k < n_bins, for every window w of the batch (no window function:
 * pmu_estimate feeds pmue_fft_r the Hann-weighted samples, src/pmu_estimator.c:184-200).  `in` host or device
 * float_p [n_instances][n_channels][win_len]; `out` host float_p complex [n_instances][n_channels][n_bins].
 * The ROCOF state is not touched. */
int pmub_fft_r(pmub_batch *b, const void *in, void *out);
int pmub_set_debug(pmub_batch *b, int enable);
int pmub_get_debug(pmub_batch *b, int *flags, double *spectrum);

#ifdef __cplusplus
}
#endif
#endif /* PMU_BATCH_H */
