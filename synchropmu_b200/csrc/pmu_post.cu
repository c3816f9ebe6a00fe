// pmu_post.cu — what comes right after the estimator: consumers of the
// pmu_frame[] stream.  Neither exists in the reference, which stops at pmu_frame (src/pmu_estimator.h:52-56)
// and its text dump (src/pmu_estimator.c:306-324); both are tiny HBM-bound kernels over the frames that are
// already resident on the device.
//
//  * symmetrical components of each three-phase group of channels (positive / negative / zero sequence);
//  * IEEE C37.118.2-2011 data frames (floating-point polar phasors, FREQ, DFREQ, CRC-CCITT), one per
//    estimator instance (= one PMU with n_channels phasors).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/pmu_batch.h"
#include "pmu_host.hpp"

using pmu::set_error;

namespace {

// ---- symmetrical components ---------------------------------------------------------------------------
// frames: T [n_inst][n_ch][4] = amp, ph, freq, rocof; groups of three consecutive channels (a, b, c).
// out   : T [n_inst][n_ch / 3][3][2] = (amp, ph) of the positive, negative and zero sequence phasors:
//   X1 = (Xa + a Xb + a^2 Xc)/3,  X2 = (Xa + a^2 Xb + a Xc)/3,  X0 = (Xa + Xb + Xc)/3,  a = exp(i 2 pi/3).
template <typename T>
__global__ void seqcomp_kernel(const T* __restrict__ frames, T* __restrict__ out, long long n_groups)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    double re[3], im[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const double amp = (double)frames[(g * 3 + p) * 4 + 0];
        const double ph = (double)frames[(g * 3 + p) * 4 + 1];
        double s, c;
        sincos(ph, &s, &c);
        re[p] = amp * c;
        im[p] = amp * s;
    }
    const double ar = -0.5, ai = 0.86602540378443864676;  // a = exp(i 2 pi/3); a^2 = conj(a)
    const double b1r = ar * re[1] - ai * im[1], b1i = ar * im[1] + ai * re[1];   // a   * Xb
    const double b2r = ar * re[1] + ai * im[1], b2i = ar * im[1] - ai * re[1];   // a^2 * Xb
    const double c1r = ar * re[2] - ai * im[2], c1i = ar * im[2] + ai * re[2];   // a   * Xc
    const double c2r = ar * re[2] + ai * im[2], c2i = ar * im[2] - ai * re[2];   // a^2 * Xc
    const double sr[3] = {(re[0] + b1r + c2r) / 3, (re[0] + b2r + c1r) / 3, (re[0] + re[1] + re[2]) / 3};
    const double si[3] = {(im[0] + b1i + c2i) / 3, (im[0] + b2i + c1i) / 3, (im[0] + im[1] + im[2]) / 3};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        out[(g * 3 + q) * 2 + 0] = (T)sqrt(sr[q] * sr[q] + si[q] * si[q]);
        out[(g * 3 + q) * 2 + 1] = (T)atan2(si[q], sr[q]);
    }
}

// ---- C37.118.2 data frames ---------------------------------------------------------------------------------
// FORMAT word (C37.118.2-2011 table 9, also carried by the CFG-2 frame):
//   bit 0  phasors  0 = rectangular (real, imaginary)   1 = polar (magnitude, angle)
//   bit 1  phasors  0 = 16-bit integer                  1 = IEEE float
//   bit 2  analogs  (none are sent)
//   bit 3  FREQ/DFREQ  0 = 16-bit integer (deviation from nominal in mHz; ROCOF in 0.01 Hz/s)   1 = IEEE float (Hz, Hz/s)
// 16-bit phasors are scaled by PHUNIT (1e-5 units per count): integer value = X / (phunit * 1e-5); polar angles are
// radians * 1e4.  Layout (big-endian):
//   SYNC 0xAA01 | FRAMESIZE | IDCODE | SOC | FRACSEC (time quality << 24 | fraction count) | STAT |
//   n_ch x phasor | FREQ | DFREQ | CHK (CRC-CCITT)
// FREQ/DFREQ are taken from channel 0 of the instance (the PMU reports one frequency).
struct C37Opts {
    unsigned format, idcode0, soc, fracsec, stat, f_nominal;
    const double* phunit;   // device pointer [n_ch] or null (1e-5 units per count each -> PHUNIT = 1)
};

__host__ __device__ inline int c37_phasor_bytes(unsigned format) { return (format & 2u) ? 8 : 4; }
__host__ __device__ inline int c37_freq_bytes(unsigned format) { return (format & 8u) ? 4 : 2; }
__host__ __device__ inline int c37_frame_bytes(unsigned n_ch, unsigned format)
{
    return 16 + (int)n_ch * c37_phasor_bytes(format) + 2 * c37_freq_bytes(format) + 2;
}

__device__ __forceinline__ void put16(unsigned char* p, unsigned v) { p[0] = (unsigned char)(v >> 8); p[1] = (unsigned char)v; }
__device__ __forceinline__ void put32(unsigned char* p, unsigned v)
{
    p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v;
}
__device__ __forceinline__ void putf(unsigned char* p, float f) { put32(p, __float_as_uint(f)); }
__device__ __forceinline__ int sat16(double v)
{
    const double r = rint(v);
    return (int)(r < -32768.0 ? -32768.0 : (r > 32767.0 ? 32767.0 : r));
}
__device__ __forceinline__ unsigned satu16(double v)
{
    const double r = rint(v);
    return (unsigned)(r < 0.0 ? 0.0 : (r > 65535.0 ? 65535.0 : r));
}

// CRC-CCITT as C37.118.2 annex B uses it: polynomial x^16 + x^12 + x^5 + 1, initial value 0xFFFF, no final XOR,
// most significant bit first; four bits per step through a 16-entry table.
__constant__ unsigned short c37_crc_nibble[16] = {0x0000, 0x1021, 0x2042, 0x3063, 0x4084, 0x50A5, 0x60C6, 0x70E7,
                                                  0x8108, 0x9129, 0xA14A, 0xB16B, 0xC18C, 0xD1AD, 0xE1CE, 0xF1EF};
__device__ __forceinline__ unsigned crc_ccitt(const unsigned char* p, int n)
{
    unsigned crc = 0xFFFFu;
    for (int i = 0; i < n; ++i) {
        const unsigned byte = p[i];
        crc = ((crc << 4) & 0xFFFFu) ^ c37_crc_nibble[((crc >> 12) ^ (byte >> 4)) & 0xFu];
        crc = ((crc << 4) & 0xFFFFu) ^ c37_crc_nibble[((crc >> 12) ^ byte) & 0xFu];
    }
    return crc;
}

// One thread builds the frame of one instance in shared memory; the block then writes its (contiguous) frames out with
// coalesced 32-bit stores.  blockDim.x * frame_bytes is a multiple of 4 (frames are even-sized, 128 threads per block).
template <typename T>
__global__ void c37_pack_kernel(const T* __restrict__ frames, unsigned char* __restrict__ out, long long n_inst, int n_ch,
                                int frame_bytes, C37Opts o)
{
    extern __shared__ __align__(16) unsigned char c37_smem[];
    const long long i0 = (long long)blockIdx.x * blockDim.x;
    const long long i = i0 + threadIdx.x;
    if (i < n_inst) {
        unsigned char* p = c37_smem + (size_t)threadIdx.x * frame_bytes;
        put16(p + 0, 0xAA01u);
        put16(p + 2, (unsigned)frame_bytes);
        put16(p + 4, (o.idcode0 + (unsigned)i) & 0xFFFFu);
        put32(p + 6, o.soc);
        put32(p + 10, o.fracsec);
        put16(p + 14, o.stat);
        const T* f = frames + (size_t)i * n_ch * 4;
        const int pb = c37_phasor_bytes(o.format);
        unsigned char* q = p + 16;
        for (int c = 0; c < n_ch; ++c, q += pb) {
            const double amp = (double)f[4 * c + 0], ph = (double)f[4 * c + 1];
            const bool polar = o.format & 1u, flt = o.format & 2u;
            double s, co;
            sincos(ph, &s, &co);
            if (flt) {
                putf(q, polar ? (float)amp : (float)(amp * co));
                putf(q + 4, polar ? (float)ph : (float)(amp * s));
            } else {
                const double unit = (o.phunit ? o.phunit[c] : 1.0) * 1e-5;
                if (polar) {
                    put16(q, satu16(amp / unit));
                    put16(q + 2, (unsigned)sat16(ph * 1e4) & 0xFFFFu);
                } else {
                    put16(q, (unsigned)sat16(amp * co / unit) & 0xFFFFu);
                    put16(q + 2, (unsigned)sat16(amp * s / unit) & 0xFFFFu);
                }
            }
        }
        if (o.format & 8u) {
            putf(q, (float)f[2]);
            putf(q + 4, (float)f[3]);
            q += 8;
        } else {
            put16(q, (unsigned)sat16(((double)f[2] - (double)o.f_nominal) * 1000.0) & 0xFFFFu);
            put16(q + 2, (unsigned)sat16((double)f[3] * 100.0) & 0xFFFFu);
            q += 4;
        }
        put16(q, crc_ccitt(p, frame_bytes - 2));
    }
    __syncthreads();
    const long long n_here = (n_inst - i0 < (long long)blockDim.x) ? n_inst - i0 : (long long)blockDim.x;
    const size_t bytes = (size_t)n_here * frame_bytes;
    unsigned char* dst = out + (size_t)i0 * frame_bytes;
    const size_t words = bytes / 4;
    for (size_t w = threadIdx.x; w < words; w += blockDim.x)
        reinterpret_cast<unsigned*>(dst)[w] = reinterpret_cast<const unsigned*>(c37_smem)[w];
    for (size_t r = words * 4 + threadIdx.x; r < bytes; r += blockDim.x) dst[r] = c37_smem[r];   // tail of the last block
}

bool on_device(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

#define POST_OK(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            set_error(std::string("[pmu_post] CUDA ERROR: ") + cudaGetErrorString(e__) + " at " #expr); \
            return -1;                                                                                  \
        }                                                                                               \
    } while (0)

// Device staging for host-side callers, kept between calls (one pair of buffers per host thread and device).
struct Staging {
    void* in = nullptr;
    void* out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    int device = -1;
    ~Staging()
    {
        if (in) cudaFree(in);
        if (out) cudaFree(out);
    }
    int reserve(size_t in_bytes, size_t out_bytes)
    {
        int dev = 0;
        POST_OK(cudaGetDevice(&dev));
        if (dev != device) {
            if (in) cudaFree(in);
            if (out) cudaFree(out);
            in = out = nullptr;
            in_cap = out_cap = 0;
            device = dev;
        }
        if (in_cap < in_bytes) {
            if (in) cudaFree(in);
            in = nullptr; in_cap = 0;
            POST_OK(cudaMalloc(&in, in_bytes));
            in_cap = in_bytes;
        }
        if (out_cap < out_bytes) {
            if (out) cudaFree(out);
            out = nullptr; out_cap = 0;
            POST_OK(cudaMalloc(&out, out_bytes));
            out_cap = out_bytes;
        }
        return 0;
    }
};
thread_local Staging g_staging;

// Runs `launch(d_in, d_out, stream)` on `stream` with device copies of host buffers where needed; synchronous for host
// outputs, asynchronous (stream-ordered) when everything already lives on the device.
template <typename F>
int with_device_buffers(const void* in, size_t in_bytes, void* out, size_t out_bytes, cudaStream_t stream, F launch)
{
    const bool in_dev = on_device(in), out_dev = on_device(out);
    void *d_in = const_cast<void*>(in), *d_out = out;
    if (!in_dev || !out_dev) {
        if (g_staging.reserve(in_dev ? 0 : in_bytes, out_dev ? 0 : out_bytes)) return -1;
        if (!in_dev) {
            d_in = g_staging.in;
            POST_OK(cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, stream));
        }
        if (!out_dev) d_out = g_staging.out;
    }
    launch(d_in, d_out, stream);
    POST_OK(cudaGetLastError());
    if (!out_dev) {
        POST_OK(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, stream));
        POST_OK(cudaStreamSynchronize(stream));
    } else if (!in_dev) {
        POST_OK(cudaStreamSynchronize(stream));   // the staging buffer may be reused by the next call
    }
    return 0;
}

int seqcomp_on(int dtype, const void* frames, void* out, unsigned n_instances, unsigned n_channels, cudaStream_t stream)
{
    if (!frames || !out || n_instances < 1 || n_channels < 3 || n_channels % 3 != 0 || (dtype != PMUB_F32 && dtype != PMUB_F64)) {
        set_error("[pmu_post] ERROR: sequence components need frames of 3k channels per instance");
        return -1;
    }
    const long long n_groups = (long long)n_instances * (n_channels / 3);
    const size_t in_bytes = (size_t)n_instances * n_channels * 4 * dtype, out_bytes = (size_t)n_groups * 6 * dtype;
    const int block = 256, grid = (int)((n_groups + block - 1) / block);
    return with_device_buffers(frames, in_bytes, out, out_bytes, stream, [&](void* d_in, void* d_out, cudaStream_t s) {
        if (dtype == PMUB_F64) seqcomp_kernel<double><<<grid, block, 0, s>>>((const double*)d_in, (double*)d_out, n_groups);
        else seqcomp_kernel<float><<<grid, block, 0, s>>>((const float*)d_in, (float*)d_out, n_groups);
    });
}

int c37_pack_on(int dtype, const void* frames, void* out, unsigned n_instances, unsigned n_channels, const pmub_c37_options* opt,
                cudaStream_t stream)
{
    if (!frames || !out || !opt || n_instances < 1 || n_channels < 1 || (dtype != PMUB_F32 && dtype != PMUB_F64) ||
        (opt->format & ~0xFu)) {
        set_error("[pmu_post] ERROR: bad argument");
        return -1;
    }
    const int fb = c37_frame_bytes(n_channels, opt->format);
    if (fb > 0xFFFF) { set_error("[pmu_post] ERROR: C37.118.2 frames are limited to 65535 bytes"); return -1; }
    const size_t in_bytes = (size_t)n_instances * n_channels * 4 * dtype, out_bytes = (size_t)n_instances * fb;
    const int block = 128, grid = (int)((n_instances + block - 1) / block);
    const size_t smem = (size_t)block * fb;
    C37Opts o;
    o.format = opt->format; o.idcode0 = opt->idcode0; o.soc = opt->soc; o.fracsec = opt->fracsec; o.stat = opt->stat;
    o.f_nominal = opt->f_nominal ? opt->f_nominal : 50;
    o.phunit = nullptr;
    double* d_unit = nullptr;
    if (!(opt->format & 2u) && opt->phunit) {
        // PHUNIT counts (1e-5 units per bit) of the 16-bit phasor formats, one per channel
        POST_OK(cudaMallocAsync((void**)&d_unit, n_channels * sizeof(double), stream));
        std::vector<double> u(n_channels);
        for (unsigned c = 0; c < n_channels; ++c) u[c] = (double)(opt->phunit[c] & 0xFFFFFFu);
        POST_OK(cudaMemcpyAsync(d_unit, u.data(), n_channels * sizeof(double), cudaMemcpyHostToDevice, stream));
        POST_OK(cudaStreamSynchronize(stream));   // `u` goes out of scope
        o.phunit = d_unit;
    }
    int rc = -1;
    if (smem > 48 * 1024 &&
        (cudaFuncSetAttribute(c37_pack_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
         cudaFuncSetAttribute(c37_pack_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)) {
        cudaGetLastError();
        set_error("[pmu_post] ERROR: too many channels per instance for the C37.118.2 packer");
    } else {
        rc = with_device_buffers(frames, in_bytes, out, out_bytes, stream, [&](void* d_in, void* d_out, cudaStream_t s) {
            if (dtype == PMUB_F64)
                c37_pack_kernel<double><<<grid, block, smem, s>>>((const double*)d_in, (unsigned char*)d_out, n_instances,
                                                                  (int)n_channels, fb, o);
            else
                c37_pack_kernel<float><<<grid, block, smem, s>>>((const float*)d_in, (unsigned char*)d_out, n_instances,
                                                                 (int)n_channels, fb, o);
        });
    }
    if (d_unit) cudaFreeAsync(d_unit, stream);
    return rc;
}

// ---- CFG-2 (host only: a few hundred bytes, sent once) ----------------------------------------------------------
unsigned host_crc_ccitt(const unsigned char* p, size_t n)
{
    unsigned crc = 0xFFFFu;
    for (size_t i = 0; i < n; ++i) {
        crc ^= (unsigned)p[i] << 8;
        for (int b = 0; b < 8; ++b) crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) & 0xFFFFu : (crc << 1) & 0xFFFFu;
    }
    return crc;
}
void h16(std::vector<unsigned char>& v, unsigned x) { v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x); }
void h32(std::vector<unsigned char>& v, unsigned x) { h16(v, x >> 16); h16(v, x & 0xFFFFu); }
void hname(std::vector<unsigned char>& v, const char* s)
{
    size_t n = s ? strlen(s) : 0;
    for (size_t i = 0; i < 16; ++i) v.push_back(i < n ? (unsigned char)s[i] : (unsigned char)' ');
}

}  // namespace

extern "C" {

int pmub_sequence_components(int dtype, const void* frames, void* out, unsigned n_instances, unsigned n_channels)
{
    return seqcomp_on(dtype, frames, out, n_instances, n_channels, cudaStreamLegacy);
}

int pmub_sequence_components_on(int dtype, const void* frames, void* out, unsigned n_instances, unsigned n_channels,
                                void* cuda_stream)
{
    return seqcomp_on(dtype, frames, out, n_instances, n_channels, cuda_stream ? (cudaStream_t)cuda_stream : cudaStreamLegacy);
}

int pmub_c37118_frame_bytes(unsigned n_channels) { return c37_frame_bytes(n_channels, PMUB_C37_FLOAT_POLAR); }
int pmub_c37118_frame_bytes_ex(unsigned n_channels, unsigned format) { return c37_frame_bytes(n_channels, format & 0xFu); }

int pmub_c37118_pack(int dtype, const void* frames, void* out, unsigned n_instances, unsigned n_channels, unsigned idcode0,
                     unsigned soc, unsigned fracsec, unsigned stat)
{
    pmub_c37_options o;
    memset(&o, 0, sizeof o);
    o.format = PMUB_C37_FLOAT_POLAR;
    o.idcode0 = idcode0; o.soc = soc; o.fracsec = fracsec; o.stat = stat; o.f_nominal = 50;
    return c37_pack_on(dtype, frames, out, n_instances, n_channels, &o, cudaStreamLegacy);
}

int pmub_c37118_pack_ex(int dtype, const void* frames, void* out, unsigned n_instances, unsigned n_channels,
                        const pmub_c37_options* opt, void* cuda_stream)
{
    return c37_pack_on(dtype, frames, out, n_instances, n_channels, opt, cuda_stream ? (cudaStream_t)cuda_stream : cudaStreamLegacy);
}

// CFG-2 frame of instance `instance` (one PMU block: NUM_PMU = 1, its n_channels phasors, no analogs / digitals).
// Returns the frame size, or -1 when `cap` is too small / an argument is bad.
int pmub_c37118_cfg2(unsigned char* out, unsigned cap, unsigned n_channels, const pmub_c37_options* opt, unsigned instance,
                     const char* station_name, const char* const* channel_names, unsigned time_base, int data_rate)
{
    if (!out || !opt || n_channels < 1 || (opt->format & ~0xFu)) { set_error("[pmu_post] ERROR: bad argument"); return -1; }
    std::vector<unsigned char> v;
    h16(v, 0xAA31u);                              // SYNC: configuration frame 2, version 1
    h16(v, 0);                                    // FRAMESIZE (patched below)
    h16(v, (opt->idcode0 + instance) & 0xFFFFu);  // IDCODE of the data stream
    h32(v, opt->soc);
    h32(v, opt->fracsec);
    h32(v, time_base & 0xFFFFFFu);                // TIME_BASE: FRACSEC counts per second
    h16(v, 1);                                    // NUM_PMU
    char stn[32];
    snprintf(stn, sizeof stn, "%s%u", station_name ? station_name : "PMU", instance);
    hname(v, stn);
    h16(v, (opt->idcode0 + instance) & 0xFFFFu);  // IDCODE of the PMU
    h16(v, opt->format & 0xFu);                   // FORMAT
    h16(v, n_channels);                           // PHNMR
    h16(v, 0);                                    // ANNMR
    h16(v, 0);                                    // DGNMR
    for (unsigned c = 0; c < n_channels; ++c) {
        char nm[32];
        snprintf(nm, sizeof nm, "PH%02u", c);
        hname(v, channel_names && channel_names[c] ? channel_names[c] : nm);
    }
    for (unsigned c = 0; c < n_channels; ++c)     // PHUNIT: bits 31-24 = 0 voltage / 1 current, 23-0 = 1e-5 units per count
        h32(v, ((opt->phunit_is_current && opt->phunit_is_current[c]) ? 0x01000000u : 0u) |
                   ((opt->phunit ? opt->phunit[c] : 1u) & 0xFFFFFFu));
    h16(v, opt->f_nominal == 60 ? 0 : 1);         // FNOM: bit 0 = 1 -> 50 Hz, 0 -> 60 Hz
    h16(v, opt->cfg_count & 0xFFFFu);             // CFGCNT
    h16(v, (unsigned)data_rate & 0xFFFFu);        // DATA_RATE: frames per second (negative: seconds per frame)
    const unsigned size = (unsigned)v.size() + 2;
    if (size > cap || size > 0xFFFFu) { set_error("[pmu_post] ERROR: CFG-2 buffer too small"); return -1; }
    v[2] = (unsigned char)(size >> 8);
    v[3] = (unsigned char)size;
    h16(v, host_crc_ccitt(v.data(), v.size()));
    memcpy(out, v.data(), v.size());
    return (int)v.size();
}

// ---- peer-visible frame buffers (multi-GPU gather without a collective) --------------------------------
// A buffer allocated here on one GPU can be opened by the other ranks' processes (CUDA IPC); passing the
// opened pointer as `frames` to pmub_estimate* makes the kernel's 128-bit frame stores land directly in
// the owner's HBM over NVLink - the gather of pmu_frame[] to rank 0 becomes part of the kernel epilogue.
int pmub_peer_alloc(size_t bytes, int device, void** dptr, unsigned char* handle64)
{
    if (!dptr || !handle64 || bytes == 0) { set_error("[pmu_peer] ERROR: bad argument"); return -1; }
    if (device >= 0) POST_OK(cudaSetDevice(device));
    POST_OK(cudaMalloc(dptr, bytes));
    POST_OK(cudaMemset(*dptr, 0, bytes));
    cudaIpcMemHandle_t h;
    POST_OK(cudaIpcGetMemHandle(&h, *dptr));
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    memcpy(handle64, &h, sizeof h);
    return 0;
}

int pmub_peer_open(const unsigned char* handle64, int device, void** dptr)
{
    if (!dptr || !handle64) { set_error("[pmu_peer] ERROR: bad argument"); return -1; }
    if (device >= 0) POST_OK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    POST_OK(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int pmub_peer_close(void* dptr)
{
    POST_OK(cudaIpcCloseMemHandle(dptr));
    return 0;
}

int pmub_peer_read(void* host_dst, const void* dptr, size_t bytes)
{
    POST_OK(cudaDeviceSynchronize());
    POST_OK(cudaMemcpy(host_dst, dptr, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int pmub_peer_free(void* dptr)
{
    POST_OK(cudaFree(dptr));
    return 0;
}

}  // extern "C"
