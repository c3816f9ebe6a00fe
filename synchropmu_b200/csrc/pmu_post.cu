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

#include <string>

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

// ---- C37.118.2 data frame ---------------------------------------------------------------------------------
__device__ __forceinline__ void put16(unsigned char* p, unsigned v) { p[0] = (unsigned char)(v >> 8); p[1] = (unsigned char)v; }
__device__ __forceinline__ void put32(unsigned char* p, unsigned v)
{
    p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v;
}
__device__ __forceinline__ void putf(unsigned char* p, float f) { put32(p, __float_as_uint(f)); }

// CRC-CCITT as C37.118.2 annex B uses it: polynomial x^16 + x^12 + x^5 + 1, initial value 0xFFFF, no final XOR,
// most significant bit first.
__device__ __forceinline__ unsigned crc_ccitt(const unsigned char* p, int n)
{
    unsigned crc = 0xFFFFu;
    for (int i = 0; i < n; ++i) {
        crc ^= (unsigned)p[i] << 8;
#pragma unroll
        for (int b = 0; b < 8; ++b) crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) & 0xFFFFu : (crc << 1) & 0xFFFFu;
    }
    return crc;
}

// One thread per instance.  Layout (big-endian), FORMAT = 0x000F (float polar phasors, float analogs, float FREQ):
//   SYNC 0xAA01 | FRAMESIZE | IDCODE | SOC | FRACSEC (time quality << 24 | fraction count) | STAT |
//   n_ch x {magnitude f32, angle f32 rad} | FREQ f32 (Hz) | DFREQ f32 (Hz/s) | CHK
// FREQ/DFREQ are taken from channel 0 of the instance (the PMU reports one frequency).
template <typename T>
__global__ void c37_pack_kernel(const T* __restrict__ frames, unsigned char* __restrict__ out, long long n_inst, int n_ch,
                                int frame_bytes, unsigned idcode0, unsigned soc, unsigned fracsec, unsigned stat)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inst) return;
    unsigned char* p = out + (size_t)i * frame_bytes;
    put16(p + 0, 0xAA01u);
    put16(p + 2, (unsigned)frame_bytes);
    put16(p + 4, (idcode0 + (unsigned)i) & 0xFFFFu);
    put32(p + 6, soc);
    put32(p + 10, fracsec);
    put16(p + 14, stat);
    const T* f = frames + (size_t)i * n_ch * 4;
    for (int c = 0; c < n_ch; ++c) {
        putf(p + 16 + 8 * c, (float)f[4 * c + 0]);
        putf(p + 20 + 8 * c, (float)f[4 * c + 1]);
    }
    putf(p + 16 + 8 * n_ch, (float)f[2]);
    putf(p + 20 + 8 * n_ch, (float)f[3]);
    put16(p + 24 + 8 * n_ch, crc_ccitt(p, 24 + 8 * n_ch));
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

// Runs `launch(d_in, d_out)` with device copies of host buffers where needed.
template <typename F>
int with_device_buffers(const void* in, size_t in_bytes, void* out, size_t out_bytes, F launch)
{
    const bool in_dev = on_device(in), out_dev = on_device(out);
    void *d_in = const_cast<void*>(in), *d_out = out;
    if (!in_dev) {
        POST_OK(cudaMalloc(&d_in, in_bytes));
        POST_OK(cudaMemcpy(d_in, in, in_bytes, cudaMemcpyHostToDevice));
    }
    if (!out_dev) POST_OK(cudaMalloc(&d_out, out_bytes));
    launch(d_in, d_out);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && !out_dev) e = cudaMemcpy(out, d_out, out_bytes, cudaMemcpyDeviceToHost);
    if (!in_dev) cudaFree(d_in);
    if (!out_dev) cudaFree(d_out);
    if (e != cudaSuccess) {
        set_error(std::string("[pmu_post] CUDA ERROR: ") + cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace

extern "C" {

int pmub_sequence_components(int dtype, const void* frames, void* out, unsigned n_instances, unsigned n_channels)
{
    if (!frames || !out || n_instances < 1 || n_channels < 3 || n_channels % 3 != 0 || (dtype != PMUB_F32 && dtype != PMUB_F64)) {
        set_error("[pmu_post] ERROR: sequence components need frames of 3k channels per instance");
        return -1;
    }
    const long long n_groups = (long long)n_instances * (n_channels / 3);
    const size_t in_bytes = (size_t)n_instances * n_channels * 4 * dtype, out_bytes = (size_t)n_groups * 6 * dtype;
    const int block = 256, grid = (int)((n_groups + block - 1) / block);
    return with_device_buffers(frames, in_bytes, out, out_bytes, [&](void* d_in, void* d_out) {
        if (dtype == PMUB_F64) seqcomp_kernel<double><<<grid, block>>>((const double*)d_in, (double*)d_out, n_groups);
        else seqcomp_kernel<float><<<grid, block>>>((const float*)d_in, (float*)d_out, n_groups);
    });
}

int pmub_c37118_frame_bytes(unsigned n_channels) { return 26 + 8 * (int)n_channels; }

int pmub_c37118_pack(int dtype, const void* frames, void* out, unsigned n_instances, unsigned n_channels, unsigned idcode0,
                     unsigned soc, unsigned fracsec, unsigned stat)
{
    if (!frames || !out || n_instances < 1 || n_channels < 1 || (dtype != PMUB_F32 && dtype != PMUB_F64)) {
        set_error("[pmu_post] ERROR: bad argument");
        return -1;
    }
    const int fb = pmub_c37118_frame_bytes(n_channels);
    const size_t in_bytes = (size_t)n_instances * n_channels * 4 * dtype, out_bytes = (size_t)n_instances * fb;
    const int block = 128, grid = (int)((n_instances + block - 1) / block);
    return with_device_buffers(frames, in_bytes, out, out_bytes, [&](void* d_in, void* d_out) {
        if (dtype == PMUB_F64)
            c37_pack_kernel<double><<<grid, block>>>((const double*)d_in, (unsigned char*)d_out, n_instances, (int)n_channels, fb,
                                                     idcode0, soc, fracsec, stat);
        else
            c37_pack_kernel<float><<<grid, block>>>((const float*)d_in, (unsigned char*)d_out, n_instances, (int)n_channels, fb,
                                                    idcode0, soc, fracsec, stat);
    });
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
