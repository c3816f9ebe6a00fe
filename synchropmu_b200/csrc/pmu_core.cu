// pmu_core.cu — device-side estimator object and the type-generic C ABI (pmu_batch.h).
//
// Owns everything pmu_init malloc'ed in the reference (src/pmu_estimator.c:110-160) — now
// device tables and per-stream ROCOF state resident in HBM — and launches the fused
// kernel of pmu_kernel.cuh.  No CPU estimation path exists here: without a CUDA device
// pmub_create fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "pmu_host.hpp"
#include "pmu_kernel.cuh"

using pmu::set_error;

#define CUDA_OK(expr)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            set_error(std::string("[pmu] CUDA ERROR: ") + cudaGetErrorString(e__) + " at " #expr); \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

struct pmub_batch {
    pmub_config cfg;
    int dtype;
    unsigned n_inst, C;
    int device;
    pmu::Derived d;
    pmu::Plan plan;
    long long n_windows;
    int num_sms;
    // device memory
    pmu_c *dU, *dV, *dT;
    pmu_c* dU1;                 // U for one window per warp pass (small launches of a plan that packs windows)
    int n_u1, u_per_slab1;
    double *st_fold, *st_dl0, *st_dl1;
    unsigned char* st_state;
    unsigned char* d_windows;   // staging for host inputs (lazy, grows with the frame count)
    size_t d_windows_cap;
    unsigned char* d_mid;
    size_t d_mid_cap;
    unsigned char* d_out;       // frames [n][4] float_p, then export [n][4] double
    unsigned char* d_frames;    // staging for multi-frame host outputs (lazy)
    size_t d_frames_cap;
    // streaming front end (pmub_stream_*): per-stream sample rings resident in HBM
    unsigned char* d_ring;      // float_p [n_windows][ring_cap]
    long long ring_cap;         // samples per stream ring (multiple of hop, >= win_len + (max_frames-1)*hop)
    long long ring_pitch;       // samples between ring rows: ring_cap, or ring_cap + win_len (mirrored rings, see ring_mirror)
    bool ring_mirror;           // the first win_len samples of every ring are kept a second time behind its end
    long long ring_wpos;        // next write position
    int hop, stream_max_frames;
    unsigned long long stream_origin, stream_pushed;  // absolute index of sample 0; samples pushed so far
    unsigned char* d_push;      // staging of a push's new samples when they are scattered by a kernel (lazy)
    size_t d_push_cap;
    // asynchronous / narrow-format ingest (pmub_stream_push_ex): two staging slots so that the host->device copy of
    // push t+1 overlaps the kernel of push t, per-stream scale factors for integer samples
    unsigned char* d_stage[2];
    size_t d_stage_cap[2];
    cudaEvent_t ev_h2d[2], ev_ingested[2], ev_done[4];
    bool ev_ingested_valid[2], ev_done_valid[4];
    double* d_scale;
    unsigned long long push_no;
    int* d_dbg;
    double* d_spec;
    unsigned long long* d_timeline;   // profiling aid (pmub_debug_timeline)
    bool debug, mirror;
    bool overlap;   // pmub_set_overlap: launches may start before their predecessor has drained (KernelArgs::defer_wait)
    bool spec_raw;              // debug spectrum = rectangular-window bins (pmub_fft_r)
    // small-batch pinned staging
    unsigned char* h_in;
    unsigned char* h_out;
    size_t h_in_bytes, h_out_bytes;
    cudaStream_t stream[2];     // copy/compute overlap of the chunked host path
    cudaEvent_t ev_order;       // orders the two worker streams after the main stream
    cudaStream_t user_stream;
    bool has_user_stream;
    long long launches;
    KernelArgs args;
};

// ---- kernel dispatch: one translation unit per float_p width (pmu_launch.cuh) ----------
int pmu_launch_f64(int M, int KC, const KernelArgs& a, int grid, int block, int smem, cudaStream_t s);
int pmu_launch_f32(int M, int KC, const KernelArgs& a, int grid, int block, int smem, cudaStream_t s);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder()
{
    static tmap_encode_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (tmap_encode_fn)p;
    }();
    return fn;
}
static_assert(sizeof(CUtensorMap) == sizeof(PmuTensorMap), "tensor map size");

// Describe `rows` windows at `base` as the tensor [rows][M][L] with a box of one slab [1][M][W].  Returns false when the
// geometry does not qualify (then the producer issues one bulk copy per row).
static bool plan_uses_tensor_maps(const pmub_batch* b)
{
    const pmu::Plan& p = b->plan;
    static const int off = getenv("PMU_NO_TENSOR_MAP") ? atoi(getenv("PMU_NO_TENSOR_MAP")) : 0;
    const bool whole = (p.nslab == 1 && p.pitch == p.L);   // one plain bulk copy per window already
    return !(off || p.M < 2 || whole || !p.aligned16 || p.pitch > 256 || p.W > p.pitch) && tmap_encoder() != nullptr;
}

// The rings of the streaming front end as the tensor [streams][M][ring_pitch] whose rows are L samples apart: with the
// mirror behind each ring's end a window is win_len contiguous samples from any start position, and a slab of it is the
// box [1][M][W] at column start + slab * W.  (Row stride < row extent: the M rows of a window overlap in the
// descriptor; cuTensorMapEncodeTiled accepts it and the copy engine just computes addresses.)
static bool encode_ring_tensor(const pmub_batch* b, const void* base, long long rows, PmuTensorMap* out)
{
    const pmu::Plan& p = b->plan;
    if (!b->ring_mirror || !plan_uses_tensor_maps(b) || rows < 1 || rows > 0x7fffffffll) return false;
    tmap_encode_fn enc = tmap_encoder();
    const cuuint64_t s = (cuuint64_t)b->dtype;
    if (!enc || ((uintptr_t)base & 15u) || ((cuuint64_t)b->ring_pitch * s) % 16) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)b->ring_pitch, (cuuint64_t)p.M, (cuuint64_t)rows};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.L * s, (cuuint64_t)b->ring_pitch * s};
    const cuuint32_t box[3] = {(cuuint32_t)p.pitch, (cuuint32_t)p.M, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUtensorMap m;
    const CUresult r = enc(&m, b->dtype == PMUB_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                           const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    memcpy(out, &m, sizeof m);
    return true;
}

static bool encode_window_tensor(const pmub_batch* b, const void* base, long long rows, PmuTensorMap* out)
{
    const pmu::Plan& p = b->plan;
    if (!plan_uses_tensor_maps(b) || rows < 1 || rows > 0x7fffffffll) return false;
    tmap_encode_fn enc = tmap_encoder();
    if (!enc || ((uintptr_t)base & 15u)) return false;
    const cuuint64_t s = (cuuint64_t)b->dtype;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.L, (cuuint64_t)p.M, (cuuint64_t)rows};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.L * s, (cuuint64_t)b->d.N * s};
    const cuuint32_t box[3] = {(cuuint32_t)p.pitch, (cuuint32_t)p.M, 1u};   // columns past L are zero-filled
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUtensorMap m;
    const CUresult r = enc(&m, b->dtype == PMUB_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                           const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    memcpy(out, &m, sizeof m);
    return true;
}

static int grid_for(const pmub_batch* b, long long n)
{
    static const int forced = getenv("PMU_GRID") ? atoi(getenv("PMU_GRID")) : 0;
    if (forced > 0) return forced;
    long long g = (n + 7) / 8;
    if (g > b->num_sms) g = b->num_sms;
    if (g < 1) g = 1;
    return (int)g;
}

// Launch the kernel over streams [w0, w1) of the batch (w0 on an instance boundary), for
// n_frames consecutive frames laid out frame-major with the full batch as frame pitch.
// Where a launch reads its windows from when it is fed by the streaming front end.
struct RingSrc {
    long long start;            // ring position of frame 0's first sample
    unsigned long long t0;      // absolute sample index (mod fs) of that sample
};

static int launch_range(pmub_batch* b, long long w0, long long w1, int n_frames, const void* d_windows_base,
                        const void* d_mid_base, void* d_frames_base, double* d_export_base, cudaStream_t s,
                        const RingSrc* ring = nullptr, bool inputs_untouched = false)
{
    if (w1 <= w0 || n_frames < 1) return 0;
    KernelArgs a = b->args;
    const size_t sz = (size_t)b->dtype;
    if (ring) {
        a.windows = b->d_ring + (size_t)w0 * (size_t)b->ring_pitch * sz;
        a.ring_cap = b->ring_cap;
        a.ring_pitch = b->ring_pitch;
        a.ring_start = ring->start;
        a.hop = b->hop;
        a.fs = (int)b->cfg.fs;
        a.t0_samples = ring->t0;
    } else {
        a.windows = (const unsigned char*)d_windows_base + (size_t)w0 * b->d.N * sz;
    }
    a.mid = d_mid_base ? (const unsigned char*)d_mid_base + (size_t)(w0 / b->C) * sz : nullptr;
    a.frames = (unsigned char*)d_frames_base + (size_t)w0 * 4 * sz;
    a.st_fold = b->st_fold + w0;
    a.st_dl0 = b->st_dl0 + w0;
    a.st_dl1 = b->st_dl1 + w0;
    a.st_state = b->st_state + w0;
    a.n_windows = w1 - w0;
    a.n_frames = n_frames;
    // streaming front end: the frames of a group of streams back to back, so that the overlapping part of consecutive
    // windows is re-read from L2 rather than from HBM (PMU_GROUP_MAJOR=0 keeps the frame-major order, 2 forces
    // group-major for explicit windows as well)
    static const int gm_env = getenv("PMU_GROUP_MAJOR") ? atoi(getenv("PMU_GROUP_MAJOR")) : 1;
    a.group_major = (n_frames > 1 && ((ring && gm_env >= 1) || gm_env >= 2)) ? 1 : 0;
    a.frame_stride = b->n_windows;
    a.mid_frame_stride = b->n_inst;
    a.dbg = b->debug ? b->d_dbg + w0 : nullptr;
    a.spec_out = b->debug ? b->d_spec + (size_t)w0 * b->cfg.n_bins * 2 : nullptr;
    a.spec_raw = b->spec_raw ? 1 : 0;
    a.state_export = d_export_base ? d_export_base + (size_t)w0 * 4 : nullptr;
    a.timeline = b->d_timeline;
    // a handful of windows (the single-instance drop-in path): one warp per window end to end
    static const int wpw_max = getenv("PMU_WARP_PER_WINDOW_MAX") ? atoi(getenv("PMU_WARP_PER_WINDOW_MAX")) : 64;
    a.warp_per_window = (a.n_windows <= wpw_max && a.n_windows <= PMU_MAX_GROUPS && b->cfg.n_bins <= PMU_WARP) ? 1 : 0;
    if (a.warp_per_window && b->plan.wp > 1) {   // the unpacked instantiation of the same plan
        a.wp = 1;
        a.U = b->dU1;
        a.n_u = b->n_u1;
        a.u_per_slab = b->u_per_slab1;
    }
    // TMA bulk copies need 16-byte aligned sources; otherwise per-element cp.async
    bool ok16 = b->plan.aligned16 && (((uintptr_t)a.windows) % 16 == 0);
    if (ring) ok16 = ok16 && ((size_t)b->ring_cap * sz) % 16 == 0 && ((size_t)b->ring_pitch * sz) % 16 == 0 && ((size_t)b->hop * sz) % 16 == 0;
    a.load_mode = ok16 ? PMU_LOAD_BULK : PMU_LOAD_ELEM;
    static const int force_elem = getenv("PMU_FORCE_ELEM_LOAD") ? atoi(getenv("PMU_FORCE_ELEM_LOAD")) : 0;
    if (force_elem) a.load_mode = PMU_LOAD_ELEM;
    static const int no_pdl = getenv("PMU_NO_PDL") ? atoi(getenv("PMU_NO_PDL")) : 0;
    a.pdl = no_pdl ? 0 : 1;
    // inputs_untouched: the library itself enqueued nothing that writes this launch's inputs since the previous launch;
    // with the caller's promise of the same (pmub_set_overlap) the wait for the preceding grid moves to the ROCOF step
    a.defer_wait = (b->overlap && inputs_untouched && a.pdl && !b->debug && !b->d_timeline) ? 1 : 0;
    a.use_tmap = 0;
    if (!ring && a.load_mode == PMU_LOAD_BULK &&
        encode_window_tensor(b, a.windows, (long long)(n_frames - 1) * b->n_windows + (w1 - w0), &a.tmap))
        a.use_tmap = 1;
    if (ring && a.load_mode == PMU_LOAD_BULK && encode_ring_tensor(b, a.windows, w1 - w0, &a.tmap)) a.use_tmap = 1;
    int grid = grid_for(b, a.n_windows);
    if (b->plan.M == 0) {   // general kernel: one warp per window, plan.n_cons active warps per CTA
        a.warp_per_window = 0;
        static const int no_fft = getenv("PMU_WIDE_DIRECT") ? atoi(getenv("PMU_WIDE_DIRECT")) : 0;
        a.wide_fft = (!no_fft && b->d.N % 16 == 0 && b->d.N / 16 >= 32) ? 1 : 0;
        const long long want = (a.n_windows + b->plan.n_cons - 1) / b->plan.n_cons, cap = (long long)b->num_sms * 2;
        grid = (int)(want < cap ? want : cap);
    }
    if (a.warp_per_window && !getenv("PMU_GRID")) grid = (int)(a.n_windows < b->num_sms ? a.n_windows : b->num_sms);
    const int block = PMU_WARP * (1 + b->plan.n_cons);
    int rc = (b->dtype == PMUB_F64) ? pmu_launch_f64(b->plan.M, b->plan.KC, a, grid, block, b->plan.smem_bytes, s)
                                    : pmu_launch_f32(b->plan.M, b->plan.KC, a, grid, block, b->plan.smem_bytes, s);
    if (rc == 0) b->launches++;
    return rc;
}

// How many consecutive frames one launch over `n` streams may carry: the in-kernel frame
// dependency tracks groups of 32 streams per CTA in a fixed-size table, and debug recording
// keeps one frame of observations.
static int frames_per_launch(const pmub_batch* b, long long n, int n_frames)
{
    if (n_frames <= 1 || b->debug || b->plan.M == 0) return 1;
    const int grid = grid_for(b, n);
    const long long n_local = (n + grid - 1) / grid;
    if ((n_local + PMU_WARP - 1) / PMU_WARP > PMU_MAX_GROUPS) return 1;
    if (n_local * (long long)n_frames * b->plan.nslab > (1ll << 30)) return 1;
    return n_frames;
}

// launch_range for any frame count: splits into per-frame launches when one launch cannot carry them
static int launch_frames(pmub_batch* b, long long w0, long long w1, int n_frames, const void* dW, const void* dM,
                         void* dF, double* d_export, cudaStream_t s, const RingSrc* ring = nullptr, bool inputs_untouched = false)
{
    const int fpl = frames_per_launch(b, w1 - w0, n_frames);
    const size_t sz = (size_t)b->dtype;
    for (int f = 0; f < n_frames; f += fpl) {
        const int nf = (n_frames - f < fpl) ? n_frames - f : fpl;
        RingSrc r2;
        if (ring) {
            r2.start = (ring->start + (long long)f * b->hop) % b->ring_cap;
            r2.t0 = (ring->t0 + (unsigned long long)f * (unsigned)b->hop) % b->cfg.fs;
        }
        if (launch_range(b, w0, w1, nf, ring ? nullptr : (const unsigned char*)dW + (size_t)f * b->n_windows * b->d.N * sz,
                         dM ? (const unsigned char*)dM + (size_t)f * b->n_inst * sz : nullptr,
                         (unsigned char*)dF + (size_t)f * b->n_windows * 4 * sz, d_export, s, ring ? &r2 : nullptr, inputs_untouched))
            return -1;
    }
    return 0;
}

static int ensure_cap(unsigned char** p, size_t* cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) {
        set_error(std::string("[pmu_estimate] CUDA ERROR: ") + cudaGetErrorString(e) + " allocating staging memory");
        return -1;
    }
    *cap = need;
    return 0;
}

// Ingest of the streaming front end: the new samples arrive as one contiguous block (host copy or device buffer) and
// this kernel scatters them into the per-stream rings, promoting narrow sample formats on the way:
//   src [n_frames][n_streams][hop] -> ring[w][(wpos + f*hop + j) mod cap] = (RingT)(scale[w] * src)   (scale == nullptr: 1)
// HBM-bound, trivially: reads hop * sizeof(SrcT) and writes hop * sizeof(RingT) per stream and frame.
template <typename RingT, typename SrcT>
__global__ void ring_scatter_kernel(RingT* __restrict__ ring, const SrcT* __restrict__ src, const double* __restrict__ scale,
                                    long long n_streams, long long w0, long long w1, int n_frames, int hop, long long cap,
                                    long long wpos, long long pitch, int mirror_len)
{
    // one warp per (frame, stream) row of `hop` consecutive samples: coalesced on both sides, one division per row
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_w = w1 - w0;
    const long long rows = n_w * n_frames;
    for (long long row = (((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; row < rows; row += warps) {
        const int f = (int)(row / n_w);
        const long long w = w0 + (row - (long long)f * n_w);
        const long long p0 = (wpos + (long long)f * hop) % cap;
        const SrcT* __restrict__ s = src + ((long long)f * n_streams + w) * hop;
        RingT* __restrict__ d = ring + w * pitch;
        const double sc = scale ? scale[w] : 1.0;
        for (int j = lane; j < hop; j += 32) {
            long long p = p0 + j;
            if (p >= cap) p -= cap;
            const SrcT v = s[j];
            const RingT r = scale ? (RingT)(sc * (double)v) : (RingT)v;
            d[p] = r;
            if (p < mirror_len) d[cap + p] = r;     // mirrored rings: the first win_len samples once more behind the end
        }
    }
}

template <typename SrcT>
static cudaError_t launch_scatter(pmub_batch* b, const void* src, const double* scale, long long w0, long long w1, int n_frames,
                                  cudaStream_t s)
{
    const long long rows = (w1 - w0) * n_frames;          // a warp per row, 8 warps per block
    int blocks = (int)((rows + 7) / 8);
    if (blocks > b->num_sms * 16) blocks = b->num_sms * 16;
    if (blocks < 1) blocks = 1;
    if (b->dtype == PMUB_F64)
        ring_scatter_kernel<double, SrcT><<<blocks, 256, 0, s>>>((double*)b->d_ring, (const SrcT*)src, scale, b->n_windows, w0, w1,
                                                                  n_frames, b->hop, b->ring_cap, b->ring_wpos, b->ring_pitch,
                                                                  b->ring_mirror ? (int)b->d.N : 0);
    else
        ring_scatter_kernel<float, SrcT><<<blocks, 256, 0, s>>>((float*)b->d_ring, (const SrcT*)src, scale, b->n_windows, w0, w1,
                                                                 n_frames, b->hop, b->ring_cap, b->ring_wpos, b->ring_pitch,
                                                                 b->ring_mirror ? (int)b->d.N : 0);
    return cudaGetLastError();
}

// Mirrored rings: bring the mirror up to date after `len` samples were written from ring position `pos` on by something
// other than the scatter kernel (a 2-D copy, the history at pmub_stream_begin, a device-side producer writing in place).
template <typename RingT>
__global__ void ring_mirror_kernel(RingT* __restrict__ ring, long long w0, long long w1, long long cap, long long pitch,
                                   int mirror_len, long long pos, long long len)
{
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = w0 + ((((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5); w < w1; w += warps) {
        RingT* __restrict__ d = ring + w * pitch;
        for (long long j = lane; j < len; j += 32) {
            long long p = pos + j;
            if (p >= cap) p -= cap;
            if (p < mirror_len) d[cap + p] = d[p];
        }
    }
}

// returns true when a kernel was enqueued
static bool refresh_mirror(pmub_batch* b, long long w0, long long w1, long long pos, long long len, cudaStream_t s)
{
    if (!b->ring_mirror || len <= 0) return false;
    const long long N = b->d.N, cap = b->ring_cap;
    pos %= cap;
    if (pos >= N && pos + len <= cap) return false;       // nothing of [pos, pos + len) lies in the mirrored part
    int blocks = (int)((w1 - w0 + 7) / 8);
    if (blocks > b->num_sms * 16) blocks = b->num_sms * 16;
    if (blocks < 1) blocks = 1;
    if (b->dtype == PMUB_F64)
        ring_mirror_kernel<double><<<blocks, 256, 0, s>>>((double*)b->d_ring, w0, w1, cap, b->ring_pitch, (int)N, pos, len);
    else
        ring_mirror_kernel<float><<<blocks, 256, 0, s>>>((float*)b->d_ring, w0, w1, cap, b->ring_pitch, (int)N, pos, len);
    return true;
}

static bool is_device_ptr(const void* p)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Without a caller-supplied stream the work goes to the LEGACY default stream, which orders
// it after everything the caller already queued on blocking streams (e.g. the kernels that
// produced device-resident inputs) - the same implicit ordering a synchronous C library has.
static cudaStream_t main_stream(pmub_batch* b) { return b->has_user_stream ? b->user_stream : cudaStreamLegacy; }

static int free_batch(pmub_batch* b)
{
    if (!b) return 0;
    cudaSetDevice(b->device);
    cudaFree(b->dU); cudaFree(b->dV); cudaFree(b->dT); cudaFree(b->dU1);
    cudaFree(b->st_fold); cudaFree(b->st_dl0); cudaFree(b->st_dl1); cudaFree(b->st_state);
    cudaFree(b->d_windows); cudaFree(b->d_mid); cudaFree(b->d_out); cudaFree(b->d_frames); cudaFree(b->d_ring); cudaFree(b->d_push);
    cudaFree(b->d_dbg); cudaFree(b->d_spec); cudaFree(b->d_timeline);
    cudaFree(b->d_stage[0]); cudaFree(b->d_stage[1]); cudaFree(b->d_scale);
    for (int i = 0; i < 2; ++i) {
        if (b->ev_h2d[i]) cudaEventDestroy(b->ev_h2d[i]);
        if (b->ev_ingested[i]) cudaEventDestroy(b->ev_ingested[i]);
    }
    for (int i = 0; i < 4; ++i)
        if (b->ev_done[i]) cudaEventDestroy(b->ev_done[i]);
    if (b->h_in) cudaFreeHost(b->h_in);
    if (b->h_out) cudaFreeHost(b->h_out);
    for (int i = 0; i < 2; ++i)
        if (b->stream[i]) cudaStreamDestroy(b->stream[i]);
    if (b->ev_order) cudaEventDestroy(b->ev_order);
    delete b;
    return 0;
}

extern "C" {

const char* pmub_last_error(void) { return pmu::last_error(); }
int pmub_validate(const pmub_config* cfg)
{
    if (!cfg) { set_error("[pmu_init] ERROR: NULL configuration"); return -1; }
    return pmu::validate(*cfg);
}
int pmub_parse_ini(const char* ini_path, pmub_config* out)
{
    if (!ini_path || !out) { set_error("[pmu_init] ERROR: NULL argument"); return -1; }
    return pmu::parse_ini(ini_path, out);
}

int pmub_reset(pmub_batch* b)
{
    if (!b) return -1;
    CUDA_OK(cudaSetDevice(b->device));
    const size_t n = (size_t)b->n_windows;
    std::vector<double> f0(n, (double)b->cfg.f0);  // freq_old = f0, src/pmu_estimator.c:145
    cudaStream_t s = main_stream(b);
    CUDA_OK(cudaMemcpyAsync(b->st_fold, f0.data(), n * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemsetAsync(b->st_dl0, 0, n * sizeof(double), s));
    CUDA_OK(cudaMemsetAsync(b->st_dl1, 0, n * sizeof(double), s));
    CUDA_OK(cudaMemsetAsync(b->st_state, 0, n, s));
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int pmub_create(pmub_batch** out, const pmub_config* cfg, int dtype, unsigned n_instances, unsigned n_channels, int device)
{
    if (!out || !cfg) { set_error("[pmu_init] ERROR: NULL argument"); return -1; }
    *out = nullptr;
    if (dtype != PMUB_F32 && dtype != PMUB_F64) { set_error("[pmu_init] ERROR: dtype must be PMUB_F32 or PMUB_F64"); return -1; }
    if (n_instances < 1 || n_channels < 1) { set_error("[pmu_init] ERROR: number of channels/instances must be at least 1"); return -1; }
    if (pmu::validate(*cfg)) return -1;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        set_error("[pmu_init] ERROR: no CUDA device available - this library has no CPU estimation path");
        return -1;
    }
    if (device < 0) CUDA_OK(cudaGetDevice(&device));
    if (device >= ndev) { set_error("[pmu_init] ERROR: CUDA device index out of range"); return -1; }
    CUDA_OK(cudaSetDevice(device));
    int cc_major = 0, max_smem = 0, num_sms = 0;
    CUDA_OK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    if (cc_major != 10) {
        set_error("[pmu_init] ERROR: the estimator kernels are built for sm_100a (B200) only");
        return -1;
    }

    pmub_batch* b = new pmub_batch();
    memset(b, 0, sizeof *b);
    b->cfg = *cfg;
    b->dtype = dtype;
    b->n_inst = n_instances;
    b->C = n_channels;
    b->device = device;
    b->num_sms = num_sms;
    b->n_windows = (long long)n_instances * n_channels;
    b->d = pmu::derive(*cfg, dtype);
    if (pmu::make_plan(*cfg, dtype, max_smem, &b->plan)) { delete b; return -1; }

    std::vector<pmu_c> U, V, T;
    pmu::build_tables(b->plan, b->d.N, &U, &V, &T);
    auto fail = [&]() { free_batch(b); return -1; };
#define CK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_error(std::string("[pmu_init] CUDA ERROR: ") + cudaGetErrorString(e__) + " at " #expr); return fail(); } } while (0)
    CK(cudaMalloc(&b->dU, U.size() * sizeof(pmu_c)));
    CK(cudaMalloc(&b->dV, V.size() * sizeof(pmu_c)));
    CK(cudaMalloc(&b->dT, T.size() * sizeof(pmu_c)));
    CK(cudaMemcpy(b->dU, U.data(), U.size() * sizeof(pmu_c), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b->dV, V.data(), V.size() * sizeof(pmu_c), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b->dT, T.data(), T.size() * sizeof(pmu_c), cudaMemcpyHostToDevice));
    if (b->plan.wp > 1) {
        // small launches run one window per warp (lane = bin estimator): same plan, twiddles for 32 lanes per window
        pmu::Plan p1 = b->plan;
        p1.wp = 1;
        p1.u_per_slab = (p1.W + 32 * p1.pk - 1) / (32 * p1.pk);
        p1.n_u = p1.nslab * p1.u_per_slab * p1.KC;
        std::vector<pmu_c> U1, V1, T1;
        pmu::build_tables(p1, b->d.N, &U1, &V1, &T1);
        b->n_u1 = p1.n_u;
        b->u_per_slab1 = p1.u_per_slab;
        CK(cudaMalloc(&b->dU1, U1.size() * sizeof(pmu_c)));
        CK(cudaMemcpy(b->dU1, U1.data(), U1.size() * sizeof(pmu_c), cudaMemcpyHostToDevice));
    }
    const size_t n = (size_t)b->n_windows;
    CK(cudaMalloc(&b->st_fold, n * sizeof(double)));
    CK(cudaMalloc(&b->st_dl0, n * sizeof(double)));
    CK(cudaMalloc(&b->st_dl1, n * sizeof(double)));
    CK(cudaMalloc(&b->st_state, n));
    CK(cudaMalloc(&b->d_out, n * 4 * (size_t)dtype + n * 4 * sizeof(double)));
    CK(cudaMalloc(&b->d_mid, (size_t)n_instances * dtype));
    b->d_mid_cap = (size_t)n_instances * dtype;
    CK(cudaStreamCreateWithFlags(&b->stream[0], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&b->stream[1], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&b->ev_order, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&b->ev_h2d[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b->ev_ingested[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 4; ++i) CK(cudaEventCreateWithFlags(&b->ev_done[i], cudaEventDisableTiming));
    // small batches (the single-instance pmu_estimate path) go through pinned staging
    const size_t in_bytes = n * b->d.N * (size_t)dtype + (size_t)n_instances * dtype;
    if (in_bytes <= (size_t)(4 << 20)) {
        b->h_in_bytes = in_bytes;
        b->h_out_bytes = n * 4 * (size_t)dtype + n * 4 * sizeof(double);
        CK(cudaMallocHost(&b->h_in, b->h_in_bytes));
        CK(cudaMallocHost(&b->h_out, b->h_out_bytes));
        CK(cudaMalloc(&b->d_windows, n * b->d.N * (size_t)dtype));
        b->d_windows_cap = n * b->d.N * (size_t)dtype;
    }
#undef CK

    // kernel arguments that never change
    KernelArgs& a = b->args;
    memset(&a, 0, sizeof a);
    PostParams& pp = a.post;
    const pmu::Derived& d = b->d;
    pp.K = (int)cfg->n_bins; pp.P = (int)cfg->P; pp.Q = (int)cfg->Q; pp.iter_en = cfg->iter_eipdft ? 1 : 0;
    pp.jlo = b->plan.jlo; pp.jhi = b->plan.jhi; pp.C = (int)n_channels; pp.f0 = (int)cfg->f0;
    pp.df = d.df; pp.inv_df = 1.0 / d.df; pp.S = d.S; pp.invS = d.invS; pp.eps = d.eps; pp.invN = 1.0 / (double)d.N;
    {
        const long double pi = 3.14159265358979323846264338327950288L;
        pp.c3q = (double)(-0.25L * cosl(pi / (long double)d.N));  // cos(pi C3) = -cos(pi/N)
        pp.s3q = (double)(0.25L * sinl(pi / (long double)d.N));   // sin(pi C3) =  sin(pi/N)
    }
    pp.frame_rate = (double)cfg->frame_rate;
    pp.thr0 = d.thr[0]; pp.thr1 = d.thr[1]; pp.thr2 = d.thr[2];
    pp.lp0 = d.lp[0]; pp.lp1 = d.lp[1]; pp.lp2 = d.lp[2];
    pp.trig = b->dT;
    pp.trig_off = (unsigned)b->plan.off_trig;
    a.N = (int)d.N; a.L = b->plan.L; a.W = b->plan.W; a.pitch = b->plan.pitch; a.wp = b->plan.wp; a.pk = b->plan.pk; a.u_per_slab = b->plan.u_per_slab; a.nslab = b->plan.nslab; a.Kp = b->plan.Kp;
    a.n_stage = b->plan.n_stage; a.stage_bytes = b->plan.stage_bytes; a.n_cons = b->plan.n_cons;
    a.n_raw = b->plan.n_raw; a.raw_stride = b->plan.raw_stride; a.red_bytes = b->plan.red_bytes;
    a.n_u = b->plan.n_u;
    a.interleave = (b->plan.nslab > 1 && 2 * b->plan.nslab > b->plan.n_stage &&
                    !(getenv("PMU_NO_INTERLEAVE") && atoi(getenv("PMU_NO_INTERLEAVE")))) ? 1 : 0;
    a.off_ctrl = b->plan.off_ctrl; a.off_U = b->plan.off_U; a.off_V = b->plan.off_V; a.off_trig = b->plan.off_trig;
    a.off_raw = b->plan.off_raw; a.off_red = b->plan.off_red; a.off_stage = b->plan.off_stage;
    a.U = b->dU; a.V = b->dV; a.trig = b->dT;

    if (pmub_reset(b)) { free_batch(b); return -1; }
    *out = b;
    return 0;
}

int pmub_create_ini(pmub_batch** out, const char* ini_path, int dtype, unsigned n_instances, unsigned n_channels, int device)
{
    pmub_config cfg;
    if (pmub_parse_ini(ini_path, &cfg)) return -1;
    return pmub_create(out, &cfg, dtype, n_instances, n_channels, device);
}

int pmub_destroy(pmub_batch* b)
{
    if (!b) { set_error("[pmu_deinit] ERROR: NULL batch"); return -1; }
    return free_batch(b);
}

int pmub_get_info(const pmub_batch* b, pmub_info* o)
{
    if (!b || !o) return -1;
    memset(o, 0, sizeof *o);
    o->win_len = b->d.N; o->n_bins = b->cfg.n_bins; o->n_instances = b->n_inst; o->n_channels = b->C;
    o->dtype = b->dtype; o->device = b->device; o->df = b->d.df; o->norm_factor = b->d.S;
    o->dft_radix = b->plan.M; o->dft_bins = b->plan.KC; o->n_stages = b->plan.n_stage;
    o->stage_bytes = b->plan.stage_bytes; o->slabs_per_window = b->plan.nslab;
    o->load_mode = b->plan.M == 0 ? 2 : (b->plan.aligned16 ? PMU_LOAD_BULK : PMU_LOAD_ELEM);
    o->grid = grid_for(b, b->n_windows); o->block = b->plan.M == 0 ? 256 : PMU_WARP * (1 + b->plan.n_cons);
    o->smem_bytes = b->plan.smem_bytes;
    o->bytes_per_estimate = pmu::bytes_per_estimate(b->d.N, b->dtype, b->C);
    return 0;
}

int pmub_set_stream(pmub_batch* b, void* cuda_stream)
{
    if (!b) return -1;
    b->has_user_stream = cuda_stream != nullptr;
    b->user_stream = (cudaStream_t)cuda_stream;
    return 0;
}

/* See include/pmu_batch.h: the caller's promise that lets consecutive launches of this batch overlap. */
int pmub_set_overlap(pmub_batch* b, int enable)
{
    if (!b) return -1;
    b->overlap = enable != 0;
    return 0;
}

long long pmub_launch_count(const pmub_batch* b) { return b ? b->launches : 0; }

int pmub_sync(pmub_batch* b)
{
    if (!b) return -1;
    CUDA_OK(cudaSetDevice(b->device));
    CUDA_OK(cudaStreamSynchronize(main_stream(b)));
    CUDA_OK(cudaStreamSynchronize(b->stream[0]));
    CUDA_OK(cudaStreamSynchronize(b->stream[1]));
    return 0;
}

int pmub_estimate_frames_async(pmub_batch* b, int n_frames, const void* d_windows, const void* d_mid, void* d_frames)
{
    if (!b || !d_windows || !d_mid || !d_frames || n_frames < 1) { set_error("[pmu_estimate] ERROR: bad argument"); return -1; }
    CUDA_OK(cudaSetDevice(b->device));
    return launch_frames(b, 0, b->n_windows, n_frames, d_windows, d_mid, d_frames, nullptr, main_stream(b), nullptr, true);
}

int pmub_estimate_async(pmub_batch* b, const void* d_windows, const void* d_mid, void* d_frames)
{
    return pmub_estimate_frames_async(b, 1, d_windows, d_mid, d_frames);
}

int pmub_estimate_frames(pmub_batch* b, int n_frames, const void* windows, const void* mid, void* frames)
{
    if (!b || !windows || !mid || !frames || n_frames < 1) { set_error("[pmu_estimate] ERROR: bad argument"); return -1; }
    CUDA_OK(cudaSetDevice(b->device));
    const size_t sz = (size_t)b->dtype;
    const size_t n = (size_t)b->n_windows;
    const size_t win_bytes = n * b->d.N * sz, mid_bytes = (size_t)b->n_inst * sz, fr_bytes = n * 4 * sz;  // per frame
    const bool win_dev = is_device_ptr(windows), mid_dev = is_device_ptr(mid), fr_dev = is_device_ptr(frames);
    double* d_export = reinterpret_cast<double*>(b->d_out + fr_bytes);

    if (win_dev && mid_dev && fr_dev) {
        if (launch_frames(b, 0, b->n_windows, n_frames, windows, mid, frames, b->mirror ? d_export : nullptr, main_stream(b))) return -1;
        if (b->mirror)   // keep the drop-in API's view of the ROCOF state fresh whatever kind of pointers came in
            CUDA_OK(cudaMemcpyAsync(b->h_out + fr_bytes, d_export, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, main_stream(b)));
        CUDA_OK(cudaStreamSynchronize(main_stream(b)));
        return 0;
    }

    // ---- small batch, one frame, host buffers: one pinned round trip (the pmu_estimate path) ----
    if (b->h_in && n_frames == 1 && !win_dev && !mid_dev && !fr_dev) {
        cudaStream_t s = main_stream(b);
        memcpy(b->h_in, windows, win_bytes);
        memcpy(b->h_in + win_bytes, mid, mid_bytes);
        // A few windows: the kernel reads the pinned staging buffer and writes the pinned result buffer itself
        // (mapped host memory, bulk copies straight over PCIe) - one submission instead of three.
        static const size_t zc_max = getenv("PMU_ZERO_COPY_MAX") ? (size_t)atoll(getenv("PMU_ZERO_COPY_MAX")) : (size_t)(512 << 10);
        if (win_bytes <= zc_max) {
            if (launch_range(b, 0, b->n_windows, 1, b->h_in, b->h_in + win_bytes, b->h_out,
                             b->mirror ? reinterpret_cast<double*>(b->h_out + fr_bytes) : nullptr, s))
                return -1;
            CUDA_OK(cudaStreamSynchronize(s));
            memcpy(frames, b->h_out, fr_bytes);
            return 0;
        }
        CUDA_OK(cudaMemcpyAsync(b->d_windows, b->h_in, win_bytes, cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaMemcpyAsync(b->d_mid, b->h_in + win_bytes, mid_bytes, cudaMemcpyHostToDevice, s));
        if (launch_range(b, 0, b->n_windows, 1, b->d_windows, b->d_mid, b->d_out, b->mirror ? d_export : nullptr, s)) return -1;
        const size_t back = b->mirror ? b->h_out_bytes : fr_bytes;
        CUDA_OK(cudaMemcpyAsync(b->h_out, b->d_out, back, cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaStreamSynchronize(s));
        memcpy(frames, b->h_out, fr_bytes);
        return 0;
    }

    // ---- general path: chunk by instance, copies overlapped with kernels on two streams ----
    if (!win_dev && ensure_cap(&b->d_windows, &b->d_windows_cap, win_bytes * n_frames)) return -1;
    if (!mid_dev && ensure_cap(&b->d_mid, &b->d_mid_cap, mid_bytes * n_frames)) return -1;
    if (!fr_dev && n_frames > 1 && ensure_cap(&b->d_frames, &b->d_frames_cap, fr_bytes * n_frames)) return -1;
    const unsigned char* dW = win_dev ? (const unsigned char*)windows : b->d_windows;
    const unsigned char* dM = mid_dev ? (const unsigned char*)mid : b->d_mid;
    unsigned char* dF = fr_dev ? (unsigned char*)frames : (n_frames > 1 ? b->d_frames : b->d_out);
    int n_chunks = 1;
    if (!win_dev) {
        const size_t target = 32u << 20;  // ~32 MiB of samples per chunk
        n_chunks = (int)((win_bytes * n_frames + target - 1) / target);
        if (n_chunks < 1) n_chunks = 1;
        if (n_chunks > 64) n_chunks = 64;
        if ((unsigned)n_chunks > b->n_inst) n_chunks = (int)b->n_inst;
    }
    if (!mid_dev) CUDA_OK(cudaMemcpyAsync(b->d_mid, mid, mid_bytes * n_frames, cudaMemcpyHostToDevice, main_stream(b)));
    if (!b->has_user_stream) {
        // the two worker streams start after whatever is already queued on the main stream
        CUDA_OK(cudaEventRecord(b->ev_order, main_stream(b)));
        CUDA_OK(cudaStreamWaitEvent(b->stream[0], b->ev_order, 0));
        CUDA_OK(cudaStreamWaitEvent(b->stream[1], b->ev_order, 0));
    }
    for (int c = 0; c < n_chunks; ++c) {
        cudaStream_t s = b->has_user_stream ? b->user_stream : b->stream[c & 1];
        const long long i0 = (long long)b->n_inst * c / n_chunks, i1 = (long long)b->n_inst * (c + 1) / n_chunks;
        const long long w0 = i0 * b->C, w1 = i1 * b->C;
        if (!win_dev)   // rows = frames, row pitch = one full frame of the batch
            CUDA_OK(cudaMemcpy2DAsync(b->d_windows + (size_t)w0 * b->d.N * sz, win_bytes,
                                      (const unsigned char*)windows + (size_t)w0 * b->d.N * sz, win_bytes,
                                      (size_t)(w1 - w0) * b->d.N * sz, (size_t)n_frames, cudaMemcpyHostToDevice, s));
        if (launch_frames(b, w0, w1, n_frames, dW, dM, dF, b->mirror ? d_export : nullptr, s)) return -1;
        if (!fr_dev)
            CUDA_OK(cudaMemcpy2DAsync((unsigned char*)frames + (size_t)w0 * 4 * sz, fr_bytes, dF + (size_t)w0 * 4 * sz, fr_bytes,
                                      (size_t)(w1 - w0) * 4 * sz, (size_t)n_frames, cudaMemcpyDeviceToHost, s));
    }
    if (pmub_sync(b)) return -1;
    if (b->mirror) CUDA_OK(cudaMemcpy(b->h_out + fr_bytes, d_export, n * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int pmub_estimate(pmub_batch* b, const void* windows, const void* mid, void* frames)
{
    return pmub_estimate_frames(b, 1, windows, mid, frames);
}

// ---- streaming front end -----------------------------------------------------------------------
int pmub_stream_begin(pmub_batch* b, unsigned hop, unsigned max_frames_per_push, const void* history,
                      unsigned long long first_sample_index)
{
    if (!b) { set_error("[pmu_stream] ERROR: NULL batch"); return -1; }
    CUDA_OK(cudaSetDevice(b->device));
    if (hop == 0) hop = b->cfg.fs / b->cfg.frame_rate;   // one frame every 1/frame_rate seconds
    if (hop == 0 || hop > b->d.N) { set_error("[pmu_stream] ERROR: hop must be in 1..win_len"); return -1; }
    if (max_frames_per_push < 1) max_frames_per_push = 1;
    if (pmub_sync(b)) return -1;
    const size_t sz = (size_t)b->dtype;
    const long long N = b->d.N;
    long long cap = N + (long long)(max_frames_per_push - 1) * hop;
    cap = (cap + hop - 1) / hop * hop;                  // whole hops: a pushed block never wraps
    if (b->d_ring) { cudaFree(b->d_ring); b->d_ring = nullptr; }
    // plans that load their slabs as tensor-map tiles get mirrored rings (a window is then contiguous wherever it starts)
    b->ring_mirror = b->plan.M != 0 && plan_uses_tensor_maps(b) && ((size_t)hop * sz) % 16 == 0;
    long long pitch = cap;
    if (b->ring_mirror) pitch = (cap + N + 3) / 4 * 4;    // rows stay 16-byte aligned for either sample size
    CUDA_OK(cudaMalloc(&b->d_ring, (size_t)b->n_windows * (size_t)pitch * sz));
    CUDA_OK(cudaMemset(b->d_ring, 0, (size_t)b->n_windows * (size_t)pitch * sz));
    b->ring_cap = cap;
    b->ring_pitch = pitch;
    b->hop = (int)hop;
    b->stream_max_frames = (int)max_frames_per_push;
    const long long hist = N - hop;                     // samples needed before the first push completes a window
    if (history && hist > 0)
        CUDA_OK(cudaMemcpy2D(b->d_ring, (size_t)pitch * sz, history, (size_t)hist * sz, (size_t)hist * sz, (size_t)b->n_windows,
                             cudaMemcpyDefault));
    if (history && hist > 0 && refresh_mirror(b, 0, b->n_windows, 0, hist, 0)) CUDA_OK(cudaDeviceSynchronize());
    b->ring_wpos = hist % cap;
    b->stream_origin = first_sample_index;
    b->stream_pushed = (unsigned long long)hist;
    return 0;
}

int pmub_stream_push(pmub_batch* b, int n_frames, const void* new_samples, const void* mid, void* frames)
{
    if (!b || !b->d_ring) { set_error("[pmu_stream] ERROR: call pmub_stream_begin first"); return -1; }
    if (!new_samples || !frames || n_frames < 1 || n_frames > b->stream_max_frames) {
        set_error("[pmu_stream] ERROR: bad argument (n_frames must be 1..max_frames_per_push)");
        return -1;
    }
    CUDA_OK(cudaSetDevice(b->device));
    const size_t sz = (size_t)b->dtype;
    const size_t n = (size_t)b->n_windows;
    const long long cap = b->ring_cap, hop = b->hop, N = b->d.N;
    const size_t mid_bytes = (size_t)b->n_inst * sz, fr_bytes = n * 4 * sz;   // per frame
    const bool in_dev = is_device_ptr(new_samples), mid_dev = mid ? is_device_ptr(mid) : true, fr_dev = is_device_ptr(frames);
    if (mid && !mid_dev && ensure_cap(&b->d_mid, &b->d_mid_cap, mid_bytes * n_frames)) return -1;
    if (!fr_dev && ensure_cap(&b->d_frames, &b->d_frames_cap, fr_bytes * n_frames)) return -1;
    if (!in_dev && (size_t)hop * sz < 2048 && ensure_cap(&b->d_push, &b->d_push_cap, n * (size_t)hop * sz * n_frames)) return -1;
    const unsigned char* dM = mid ? (mid_dev ? (const unsigned char*)mid : b->d_mid) : nullptr;
    unsigned char* dF = fr_dev ? (unsigned char*)frames : b->d_frames;
    RingSrc ring;
    ring.start = ((b->ring_wpos + hop - N) % cap + cap) % cap;
    ring.t0 = (b->stream_origin + b->stream_pushed + (unsigned long long)hop - (unsigned long long)N) % b->cfg.fs;

    int n_chunks = 1;
    if (!in_dev) {
        const size_t target = 32u << 20;
        n_chunks = (int)((n * (size_t)hop * sz * n_frames + target - 1) / target);
        if (n_chunks < 1) n_chunks = 1;
        if (n_chunks > 64) n_chunks = 64;
        if ((unsigned)n_chunks > b->n_inst) n_chunks = (int)b->n_inst;
    }
    if (mid && !mid_dev) CUDA_OK(cudaMemcpyAsync(b->d_mid, mid, mid_bytes * n_frames, cudaMemcpyHostToDevice, main_stream(b)));
    if (!b->has_user_stream) {
        CUDA_OK(cudaEventRecord(b->ev_order, main_stream(b)));
        CUDA_OK(cudaStreamWaitEvent(b->stream[0], b->ev_order, 0));
        CUDA_OK(cudaStreamWaitEvent(b->stream[1], b->ev_order, 0));
    }
    for (int c = 0; c < n_chunks; ++c) {
        cudaStream_t s = b->has_user_stream ? b->user_stream : b->stream[c & 1];
        const long long i0 = (long long)b->n_inst * c / n_chunks, i1 = (long long)b->n_inst * (c + 1) / n_chunks;
        const long long w0 = i0 * b->C, w1 = i1 * b->C;
        // ingest: frame f's hop samples of every stream go to ring position (wpos + f*hop) mod cap
        if ((size_t)hop * sz >= 2048) {
            // wide rows: 2-D DMA straight from the caller's buffer into the rings
            // (two when the block crosses the end of the ring: the write position is a multiple of the hop only
            //  if the hop divides win_len)
            for (int f = 0; f < n_frames; ++f) {
                const long long p = (b->ring_wpos + (long long)f * hop) % cap;
                const long long n1 = (cap - p < hop) ? cap - p : hop;
                const unsigned char* src = (const unsigned char*)new_samples + ((size_t)f * n + (size_t)w0) * (size_t)hop * sz;
                const size_t rp = (size_t)b->ring_pitch;
                CUDA_OK(cudaMemcpy2DAsync(b->d_ring + ((size_t)w0 * rp + (size_t)p) * sz, rp * sz, src,
                                          (size_t)hop * sz, (size_t)n1 * sz, (size_t)(w1 - w0), cudaMemcpyDefault, s));
                if (n1 < hop)
                    CUDA_OK(cudaMemcpy2DAsync(b->d_ring + (size_t)w0 * rp * sz, rp * sz, src + (size_t)n1 * sz,
                                              (size_t)hop * sz, (size_t)(hop - n1) * sz, (size_t)(w1 - w0), cudaMemcpyDefault, s));
            }
            refresh_mirror(b, w0, w1, b->ring_wpos, (long long)n_frames * hop, s);
        } else {
            // narrow rows: contiguous copies into a staging area (host input), then a scatter kernel
            const unsigned char* src = (const unsigned char*)new_samples;
            if (!in_dev) {
                for (int f = 0; f < n_frames; ++f)
                    CUDA_OK(cudaMemcpyAsync(b->d_push + ((size_t)f * n + (size_t)w0) * (size_t)hop * sz,
                                            (const unsigned char*)new_samples + ((size_t)f * n + (size_t)w0) * (size_t)hop * sz,
                                            (size_t)(w1 - w0) * (size_t)hop * sz, cudaMemcpyHostToDevice, s));
                src = b->d_push;
            }
            if (sz == 8) CUDA_OK(launch_scatter<double>(b, src, nullptr, w0, w1, n_frames, s));
            else CUDA_OK(launch_scatter<float>(b, src, nullptr, w0, w1, n_frames, s));
            CUDA_OK(cudaGetLastError());
        }
        if (launch_frames(b, w0, w1, n_frames, nullptr, dM, dF, nullptr, s, &ring)) return -1;
        if (!fr_dev)
            CUDA_OK(cudaMemcpy2DAsync((unsigned char*)frames + (size_t)w0 * 4 * sz, fr_bytes, dF + (size_t)w0 * 4 * sz, fr_bytes,
                                      (size_t)(w1 - w0) * 4 * sz, (size_t)n_frames, cudaMemcpyDeviceToHost, s));
    }
    b->ring_wpos = (b->ring_wpos + (long long)n_frames * hop) % cap;
    b->stream_pushed += (unsigned long long)n_frames * (unsigned long long)hop;
    return pmub_sync(b);
}

// Per-stream scale factors of integer sample formats (sample = scale[stream] * count); NULL = 1.
int pmub_stream_set_scale(pmub_batch* b, const double* scale)
{
    if (!b) { set_error("[pmu_stream] ERROR: NULL batch"); return -1; }
    CUDA_OK(cudaSetDevice(b->device));
    if (pmub_sync(b)) return -1;
    if (!scale) {
        cudaFree(b->d_scale);
        b->d_scale = nullptr;
        return 0;
    }
    if (!b->d_scale) CUDA_OK(cudaMalloc(&b->d_scale, (size_t)b->n_windows * sizeof(double)));
    CUDA_OK(cudaMemcpy(b->d_scale, scale, (size_t)b->n_windows * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

static size_t sample_bytes(const pmub_batch* b, int fmt)
{
    switch (fmt) {
        case PMUB_SAMPLES_NATIVE: return (size_t)b->dtype;
        case PMUB_SAMPLES_F32: return 4;
        case PMUB_SAMPLES_I16: return 2;
    }
    return 0;
}

// The general push: any sample format, optionally asynchronous.  Host samples are copied into one of two staging
// slots on a copy stream (the copy of push t+1 overlaps the kernels of push t), scattered / promoted into the rings and
// estimated on the compute stream; the frames go back with an asynchronous copy and an event marks the push complete.
int pmub_stream_push_ex(pmub_batch* b, int n_frames, const void* samples, int sample_format, const void* mid, void* frames,
                        unsigned flags, unsigned long long* ticket)
{
    if (!b || !b->d_ring) { set_error("[pmu_stream] ERROR: call pmub_stream_begin first"); return -1; }
    const size_t esz = b ? sample_bytes(b, sample_format) : 0;
    const bool in_place = (flags & PMUB_PUSH_IN_PLACE) != 0;   // the caller's device code already wrote the samples into the rings
    if ((!samples && !in_place) || !frames || n_frames < 1 || n_frames > b->stream_max_frames || esz == 0) {
        set_error("[pmu_stream] ERROR: bad argument (n_frames must be 1..max_frames_per_push, sample_format one of PMUB_SAMPLES_*)");
        return -1;
    }
    if (sample_format == PMUB_SAMPLES_NATIVE && !(flags & (PMUB_PUSH_ASYNC | PMUB_PUSH_IN_PLACE))) {
        if (ticket) *ticket = b->push_no;
        return pmub_stream_push(b, n_frames, samples, mid, frames);
    }
    CUDA_OK(cudaSetDevice(b->device));
    const size_t sz = (size_t)b->dtype;
    const size_t n = (size_t)b->n_windows;
    const long long cap = b->ring_cap, hop = b->hop, N = b->d.N;
    const size_t mid_bytes = (size_t)b->n_inst * sz, fr_bytes = n * 4 * sz;   // per frame
    const size_t in_bytes = n * (size_t)hop * esz * n_frames;
    const bool in_dev = in_place || is_device_ptr(samples), mid_dev = mid ? is_device_ptr(mid) : true, fr_dev = is_device_ptr(frames);
    if (mid && !mid_dev && ensure_cap(&b->d_mid, &b->d_mid_cap, mid_bytes * n_frames)) return -1;
    if (!fr_dev && ensure_cap(&b->d_frames, &b->d_frames_cap, fr_bytes * n_frames)) return -1;
    const int slot = (int)(b->push_no & 1);
    const int dslot = (int)(b->push_no & 3);
    cudaStream_t s_copy = b->has_user_stream ? b->user_stream : b->stream[0];
    cudaStream_t s_comp = b->has_user_stream ? b->user_stream : b->stream[1];
    if (!b->has_user_stream) {
        // start after whatever the caller already queued on the (legacy) default stream, e.g. the producer of device samples
        CUDA_OK(cudaEventRecord(b->ev_order, main_stream(b)));
        CUDA_OK(cudaStreamWaitEvent(s_copy, b->ev_order, 0));
        CUDA_OK(cudaStreamWaitEvent(s_comp, b->ev_order, 0));
    }
    // at most four pushes in flight: the completion event about to be reused must have fired
    if (b->ev_done_valid[dslot]) CUDA_OK(cudaEventSynchronize(b->ev_done[dslot]));
    const void* src = samples;
    if (!in_dev) {
        if (b->d_stage_cap[slot] < in_bytes) {
            // growing a slot: nothing may still be reading the old allocation
            if (pmub_sync(b)) return -1;
            if (ensure_cap(&b->d_stage[slot], &b->d_stage_cap[slot], in_bytes)) return -1;
        }
        if (!b->has_user_stream && b->ev_ingested_valid[slot]) CUDA_OK(cudaStreamWaitEvent(s_copy, b->ev_ingested[slot], 0));
        CUDA_OK(cudaMemcpyAsync(b->d_stage[slot], samples, in_bytes, cudaMemcpyHostToDevice, s_copy));
        if (!b->has_user_stream) {
            CUDA_OK(cudaEventRecord(b->ev_h2d[slot], s_copy));
            CUDA_OK(cudaStreamWaitEvent(s_comp, b->ev_h2d[slot], 0));
        }
        src = b->d_stage[slot];
    }
    const unsigned char* dM = nullptr;
    if (mid) {
        if (mid_dev) dM = (const unsigned char*)mid;
        else {
            CUDA_OK(cudaMemcpyAsync(b->d_mid, mid, mid_bytes * n_frames, cudaMemcpyHostToDevice, s_comp));
            dM = b->d_mid;
        }
    }
    unsigned char* dF = fr_dev ? (unsigned char*)frames : b->d_frames;
    if (!in_place) switch (sample_format) {
        case PMUB_SAMPLES_NATIVE:
            if (sz == 8) CUDA_OK(launch_scatter<double>(b, src, nullptr, 0, (long long)n, n_frames, s_comp));
            else CUDA_OK(launch_scatter<float>(b, src, nullptr, 0, (long long)n, n_frames, s_comp));
            break;
        case PMUB_SAMPLES_F32: CUDA_OK(launch_scatter<float>(b, src, nullptr, 0, (long long)n, n_frames, s_comp)); break;
        case PMUB_SAMPLES_I16: CUDA_OK(launch_scatter<short>(b, src, b->d_scale, 0, (long long)n, n_frames, s_comp)); break;
    }
    if (!in_dev && !b->has_user_stream) {
        CUDA_OK(cudaEventRecord(b->ev_ingested[slot], s_comp));
        b->ev_ingested_valid[slot] = true;
    }
    RingSrc ring;
    ring.start = ((b->ring_wpos + hop - N) % cap + cap) % cap;
    ring.t0 = (b->stream_origin + b->stream_pushed + (unsigned long long)hop - (unsigned long long)N) % b->cfg.fs;
    // samples written in place by the caller's device code: mirrored rings need their mirror brought up to date
    bool own_kernel_before = false;
    if (in_place) own_kernel_before = refresh_mirror(b, 0, (long long)n, b->ring_wpos, (long long)n_frames * hop, s_comp);
    // (in place: no ingest kernel, no copy of mid_fracsec before the launch)
    if (launch_frames(b, 0, (long long)n, n_frames, nullptr, dM, dF, nullptr, s_comp, &ring,
                      in_place && !own_kernel_before && (!mid || mid_dev)))
        return -1;
    if (!fr_dev) CUDA_OK(cudaMemcpyAsync(frames, dF, fr_bytes * n_frames, cudaMemcpyDeviceToHost, s_comp));
    CUDA_OK(cudaEventRecord(b->ev_done[dslot], s_comp));
    b->ev_done_valid[dslot] = true;
    if (ticket) *ticket = b->push_no;
    b->push_no++;
    b->ring_wpos = (b->ring_wpos + (long long)n_frames * hop) % cap;
    b->stream_pushed += (unsigned long long)n_frames * (unsigned long long)hop;
    if (!(flags & PMUB_PUSH_ASYNC)) CUDA_OK(cudaEventSynchronize(b->ev_done[dslot]));
    return 0;
}

// Where a device-side producer writes the next samples itself (PMUB_PUSH_IN_PLACE): ring base, samples per stream ring,
// and the position (in samples, < cap) the next hop of every stream goes to; a block that crosses the end wraps to 0.
int pmub_stream_ring_info(pmub_batch* b, void** base, long long* cap_samples, long long* write_pos)
{
    if (!b || !b->d_ring) { set_error("[pmu_stream] ERROR: call pmub_stream_begin first"); return -1; }
    if (base) *base = b->d_ring;
    if (cap_samples) *cap_samples = b->ring_cap;
    if (write_pos) *write_pos = b->ring_wpos;
    return 0;
}

// Samples between consecutive ring rows (>= the ring capacity: some plans keep a mirror of each ring's head behind its end,
// maintained by the library - a producer writes positions [0, capacity) of row w at base + w * pitch only).
long long pmub_stream_ring_pitch(const pmub_batch* b) { return (b && b->d_ring) ? b->ring_pitch : -1; }

// Block until the push that returned `ticket` has delivered its frames.
int pmub_stream_wait(pmub_batch* b, unsigned long long ticket)
{
    if (!b) { set_error("[pmu_stream] ERROR: NULL batch"); return -1; }
    if (ticket >= b->push_no) { set_error("[pmu_stream] ERROR: no such push"); return -1; }
    if (ticket + 4 < b->push_no) return 0;   // its event slot has been reused, which required it to have completed
    CUDA_OK(cudaSetDevice(b->device));
    const int dslot = (int)(ticket & 3);
    if (b->ev_done_valid[dslot]) CUDA_OK(cudaEventSynchronize(b->ev_done[dslot]));
    return 0;
}

int pmub_get_state(pmub_batch* b, double* freq_old, double* dl0, double* dl1, unsigned char* state)
{
    if (!b) return -1;
    CUDA_OK(cudaSetDevice(b->device));
    if (pmub_sync(b)) return -1;
    const size_t n = (size_t)b->n_windows;
    if (freq_old) CUDA_OK(cudaMemcpy(freq_old, b->st_fold, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (dl0) CUDA_OK(cudaMemcpy(dl0, b->st_dl0, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (dl1) CUDA_OK(cudaMemcpy(dl1, b->st_dl1, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (state) CUDA_OK(cudaMemcpy(state, b->st_state, n, cudaMemcpyDeviceToHost));
    return 0;
}

int pmub_set_state(pmub_batch* b, const double* freq_old, const double* dl0, const double* dl1, const unsigned char* state)
{
    if (!b) return -1;
    CUDA_OK(cudaSetDevice(b->device));
    if (pmub_sync(b)) return -1;
    const size_t n = (size_t)b->n_windows;
    if (freq_old) CUDA_OK(cudaMemcpy(b->st_fold, freq_old, n * sizeof(double), cudaMemcpyHostToDevice));
    if (dl0) CUDA_OK(cudaMemcpy(b->st_dl0, dl0, n * sizeof(double), cudaMemcpyHostToDevice));
    if (dl1) CUDA_OK(cudaMemcpy(b->st_dl1, dl1, n * sizeof(double), cudaMemcpyHostToDevice));
    if (state) CUDA_OK(cudaMemcpy(b->st_state, state, n, cudaMemcpyHostToDevice));
    return 0;
}

int pmub_set_debug(pmub_batch* b, int enable)
{
    if (!b) return -1;
    CUDA_OK(cudaSetDevice(b->device));
    if (enable && !b->d_dbg) {
        const size_t n = (size_t)b->n_windows;
        CUDA_OK(cudaMalloc(&b->d_dbg, n * sizeof(int)));
        CUDA_OK(cudaMalloc(&b->d_spec, n * b->cfg.n_bins * 2 * sizeof(double)));
        CUDA_OK(cudaMemset(b->d_dbg, 0, n * sizeof(int)));
        CUDA_OK(cudaMemset(b->d_spec, 0, n * b->cfg.n_bins * 2 * sizeof(double)));
    }
    b->debug = enable != 0;
    return 0;
}

int pmub_get_debug(pmub_batch* b, int* flags, double* spectrum)
{
    if (!b || !b->d_dbg) { set_error("[pmu] ERROR: debug recording is not enabled"); return -1; }
    CUDA_OK(cudaSetDevice(b->device));
    if (pmub_sync(b)) return -1;
    const size_t n = (size_t)b->n_windows;
    if (flags) CUDA_OK(cudaMemcpy(flags, b->d_dbg, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (spectrum) CUDA_OK(cudaMemcpy(spectrum, b->d_spec, n * b->cfg.n_bins * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

/* The reference's inner plug point as a batched call: pmue_fft_r(in, out, out_len, n_bins) (src/func_stubs.h:60-62,
 * default dft_r, src/pmu_estimator.c:487-510) for every window of the batch - the n_bins lowest bins of the DFT of the
 * samples AS GIVEN (no window applied; pmu_estimate feeds it Hann-weighted samples, :184-200).
 *   in   float_p [n_instances][n_channels][win_len], host or device
 *   out  float_p complex (re, im interleaved) [n_instances][n_channels][n_bins], host
 * Runs the same kernel (the transform is fused into it) with spectrum recording on; the ROCOF state is left untouched. */
int pmub_fft_r(pmub_batch* b, const void* in, void* out)
{
    if (!b || !in || !out) { set_error("[pmu_fft_r] ERROR: bad argument"); return -1; }
    CUDA_OK(cudaSetDevice(b->device));
    if (pmub_sync(b)) return -1;
    const size_t n = (size_t)b->n_windows, K = b->cfg.n_bins, sz = (size_t)b->dtype;
    std::vector<double> f0(n), d0(n), d1(n);
    std::vector<unsigned char> st(n);
    if (pmub_get_state(b, f0.data(), d0.data(), d1.data(), st.data())) return -1;
    const bool was_debug = b->debug, was_mirror = b->mirror;
    if (pmub_set_debug(b, 1)) return -1;
    b->spec_raw = true;
    b->mirror = false;
    int rc = -1;
    unsigned char *d_fr = nullptr, *d_mid = nullptr;
    if (cudaMalloc(&d_fr, n * 4 * sz) == cudaSuccess && cudaMalloc(&d_mid, (size_t)b->n_inst * sz) == cudaSuccess &&
        cudaMemset(d_mid, 0, (size_t)b->n_inst * sz) == cudaSuccess)
        rc = pmub_estimate(b, in, d_mid, d_fr);
    cudaFree(d_fr);
    cudaFree(d_mid);
    b->spec_raw = false;
    b->mirror = was_mirror;
    std::vector<double> spec(n * K * 2);
    if (rc == 0) rc = pmub_get_debug(b, nullptr, spec.data());
    b->debug = was_debug;
    if (pmub_set_state(b, f0.data(), d0.data(), d1.data(), st.data())) rc = -1;
    if (rc) return -1;
    if (sz == 8) memcpy(out, spec.data(), spec.size() * sizeof(double));
    else
        for (size_t i = 0; i < spec.size(); ++i) ((float*)out)[i] = (float)spec[i];
    return 0;
}

/* Profiling aid: with enable != 0 every following launch records, per CTA and warp, four globaltimer stamps (ns):
 * consumer set-up done, first window available, tickets exhausted, warp exit.  pmub_debug_timeline copies the
 * stamps of the LAST launch to `out` ([n_ctas][warps][4]) and returns the number of CTAs x warps via the pointers. */
int pmub_debug_timeline(pmub_batch* b, int enable, unsigned long long* out, int* n_ctas, int* n_warps)
{
    if (!b) return -1;
    CUDA_OK(cudaSetDevice(b->device));
    if (pmub_sync(b)) return -1;
    const int warps = PMU_BLOCK_THREADS / PMU_WARP;
    const size_t bytes = (size_t)b->num_sms * warps * 4 * sizeof(unsigned long long);
    if (out && b->d_timeline) CUDA_OK(cudaMemcpy(out, b->d_timeline, bytes, cudaMemcpyDeviceToHost));
    if (n_ctas) *n_ctas = b->num_sms;
    if (n_warps) *n_warps = warps;
    if (enable && !b->d_timeline) {
        CUDA_OK(cudaMalloc(&b->d_timeline, bytes));
        CUDA_OK(cudaMemset(b->d_timeline, 0, bytes));
    } else if (!enable && b->d_timeline) {
        cudaFree(b->d_timeline);
        b->d_timeline = nullptr;
    }
    return 0;
}

/* Mirror of the ROCOF state for the single-instance drop-in API: after a synchronous
 * pmub_estimate with host buffers, pmub_mirror() points at [n][4] doubles
 * {freq_old, dl0, dl1, state}.  Internal to libpmu_estimator (used by pmu_capi.c). */
int pmub_set_mirror(pmub_batch* b, int enable)
{
    if (!b) return -1;
    if (enable && !b->h_out) { set_error("[pmu] ERROR: state mirroring needs a small batch"); return -1; }
    b->mirror = enable != 0;
    return 0;
}
const double* pmub_mirror(const pmub_batch* b)
{
    if (!b || !b->mirror) return nullptr;
    return reinterpret_cast<const double*>(b->h_out + (size_t)b->n_windows * 4 * (size_t)b->dtype);
}

}  // extern "C"
