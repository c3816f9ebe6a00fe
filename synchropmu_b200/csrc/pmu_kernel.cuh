// pmu_kernel.cuh — the fused, persistent sm_100a estimator kernel.
//
// One CTA per SM.  Warp 0 is the PRODUCER: one elected lane streams sample windows from
// HBM into a ring of shared-memory stages with the TMA engine (cp.async.bulk +
// mbarrier complete_tx); when windows are not 16-byte granular it falls back to per-element
// cp.async from all 32 lanes.  Warps 1..NC are CONSUMERS: each claims the next window of
// the CTA's contiguous range (shared-memory ticket), computes its pruned rectangular DFT
// from the stage (pmu_dft.cuh), sums the 32 lane-partials through a per-warp shared-memory
// scratch and drops the K+1 bins into the batch's raw buffer.  The warp that completes
// a batch of <= 32 consecutive windows becomes its POST-PROCESSOR: lane = window, Hann in
// the frequency domain, IpDFT / e-IpDFT / interference iterations (pmu_math.cuh), ROCOF
// state read-modify-write in HBM, frame store.  No block-wide barrier after start-up.
// A launch may carry several consecutive frames of the same streams: the ring then streams
// (frame, stream) items frame-major, and a batch of frame f waits for the same streams' batch of
// frame f-1 before touching the ROCOF state, so only one ramp-up and one tail are paid per launch.
//
// Replaces reference src/pmu_estimator.c:170-271 (pmu_estimate) for a whole batch.
#pragma once

#include <stdint.h>

#include "pmu_dft.cuh"

#define PMU_WARP 32
// Threads per CTA the kernel is compiled for (a multiple of 128: ptxas budgets registers per
// 4-warp allocation unit).  256 -> 1 producer + up to 7 consumer warps at <= 255 registers.
#ifndef PMU_BLOCK_THREADS
#define PMU_BLOCK_THREADS 256
#endif
#define PMU_MAX_CONSUMERS (PMU_BLOCK_THREADS / PMU_WARP - 1)
#define PMU_MAX_STAGES 16
#define PMU_PITCH16 128       // stage row pitch (elements) of the 16-point-codelet kernels
#define PMU_PITCH16_SHORT 64  // ... for 1024-sample windows, stored at their natural pitch, two per warp pass
#define PMU_PITCH16_96 96     // ... for windows of a multiple of 96 (not 128) residues: M-class 3072 = 2 x 96 x 16
#define PMU_MAX_RAW 8
#define PMU_MAX_GROUPS 256  // groups of <= 32 streams per CTA when a launch carries several frames

enum { PMU_LOAD_BULK = 0, PMU_LOAD_ELEM = 1 };

// A CUtensorMap (driver API, 128 bytes, 64-byte aligned), kept opaque here so that this header needs no <cuda.h>.
struct alignas(64) PmuTensorMap {
    unsigned long long opaque[16];
};

struct KernelArgs {
    // TMA descriptor of the launch's windows viewed as a 3-D tensor [rows][M][L] (row = frame * frame_stride + stream):
    // one cp.async.bulk.tensor per slab instead of M row copies.  Valid when use_tmap != 0.
    PmuTensorMap tmap;
    int use_tmap;
    PostParams post;
    // DFT geometry
    int N, L, W, nslab, Kp;      // window length, residues (N/M), slab width, slabs per window, n_bins+1
    int pitch;                   // stage row pitch in elements (compile-time in the 16-point-codelet kernels)
    int wp;                      // windows per warp pass (1, or 4 for short windows: 8 lanes per window)
    int load_mode;               // PMU_LOAD_BULK / PMU_LOAD_ELEM
    int n_stage, stage_bytes;
    int n_cons;                  // consumer warps
    int n_raw, raw_stride;       // raw-bin buffers in the ring, doubles per window slot (an odd number of 16-byte units)
    int red_bytes;               // per consumer warp: lane-reduction scratch, reused as the estimator's W columns
    int n_u;                     // entries of U (complex)
    // shared-memory layout (byte offsets from the dynamic smem base)
    int off_ctrl, off_U, off_V, off_trig, off_raw, off_red, off_stage;
    // tables (global memory; copied to smem by the kernel)
    const pmu_c* U;              // [ceil(L/32)][KC]   W_N^{32 i k}
    const pmu_c* V;              // [KC][32]           W_N^{lane k}
    const pmu_c* trig;           // [jlo + jhi + 1]
    // batch
    const void* windows;         // InT [n_windows][N]
    const void* mid;             // InT [n_windows / C]
    void* frames;                // InT [n_windows][4]
    double* st_fold;             // ROCOF state, one entry per window stream
    double* st_dl0;
    double* st_dl1;
    unsigned char* st_state;
    long long n_windows;         // windows (streams) handled by this launch
    // consecutive frames processed by this launch, in order, per stream (ROCOF state carried)
    int n_frames;
    // item order of a multi-frame launch: 0 = frame-major (all streams of frame 0, then frame 1, ...), 1 = group-major
    // (the n_frames windows of a group of <= 32 streams back to back: in ring mode consecutive windows of a stream
    // overlap by win_len - hop samples, which are then still in L2 when the next frame reads them)
    int group_major;
    // multi-slab windows whose slabs would fill most of the ring: interleave the slabs of n_cons consecutive windows
    // (see ItemWalker); windows of a few slabs in a deep ring (M-class: 2 slabs, 11 stages) stay window after window
    int interleave;
    long long frame_stride;      // streams per frame in the caller's arrays (>= n_windows: a launch
                                 // may cover a sub-range of the batch), in windows
    long long mid_frame_stride;  // instances per frame in `mid`
    // streaming front end: when ring_cap > 0 `windows` is a per-stream ring buffer
    // InT [n_streams][ring_cap]; frame f of a stream is the win_len samples starting at
    // (ring_start + f * hop) mod ring_cap (one wrap at most), and mid_fracsec may be derived
    // from the running sample counter instead of being read (mid == nullptr)
    long long ring_cap;
    long long ring_start;
    int hop;
    int fs;
    unsigned long long t0_samples;   // absolute index (mod fs) of the first sample of frame 0's window
    // optional debug / mirror outputs (may be null); frame-major like `frames`
    int* dbg;                    // [n_frames][frame_stride]  km | flags
    double* spec_out;            // [n_frames][frame_stride][K][2] Hann-windowed bins
    int spec_raw;                // ... or the rectangular-window bins (what pmue_fft_r returns for the same samples)
    double* state_export;        // [frame_stride][4]  f_old, dl0, dl1, state after the last frame
    // small launches (a few windows, e.g. the single-instance pmu_estimate): each warp runs the estimator of
    // the window it has just transformed, lane = bin (peak search = warp-shuffle argmax), no batching
    int warp_per_window;
    int wide_fft;                // general kernel: 16 | N, bins by the pruned FFT instead of direct summation
    int pdl;                     // launched with programmatic stream serialization (see pmu_launch.cuh)
    // profiling aid: [gridDim.x][warps][4] globaltimer stamps (setup done, first window ready,
    // tickets exhausted, warp exit); null in normal operation
    unsigned long long* timeline;
    // (appended so that the offsets of everything above stay what the double kernels were tuned with)
    int pk;                      // 2: float build, two adjacent residues per lane through the packed FP32 pipe (CT = F2); else 1
    int u_per_slab;              // rows of U per slab
    // pmub_set_overlap: the caller guarantees that nothing enqueued between this batch's previous launch and this one
    // writes what the launch READS (windows / ring samples / mid_fracsec).  The wait for the preceding grid then moves from
    // the top of the kernel to the first access that really depends on it - the ROCOF state of the previous frame and
    // the frame / state stores - so the transforms of this launch start on every SM the previous launch has already
    // left, while its last CTAs are still in their estimator tail.
    int defer_wait;
    // streaming front end: samples between consecutive ring rows (>= ring_cap).  Plans whose slabs are tensor-map tiles keep
    // a MIRROR of each ring's first win_len samples behind its end (positions cap ... cap + win_len - 1), so that every
    // window is contiguous in memory and one tiled copy per slab works for rings too (the tensor's row stride L is
    // smaller than its row extent: rows of one window overlap in the descriptor, which the copy engine does not mind).
    long long ring_pitch;
};

struct CtrlBlock {
    unsigned long long full[PMU_MAX_STAGES];
    unsigned long long empty[PMU_MAX_STAGES];
    int stage_seq[PMU_MAX_STAGES];  // how many times each stage has been released by a consumer
    int ticket;
    int batch_cnt[PMU_MAX_RAW];
    int raw_owner[PMU_MAX_RAW];
    int grp_done[PMU_MAX_GROUPS];   // per group of <= 32 streams: frames fully post-processed
};

#if defined(__CUDACC__)

// ---- PTX helpers ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity)
{
    uint32_t addr = smem_u32(bar), done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware, do not spin
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// One box [1][M][W] of the windows tensor -> shared memory (rows W elements apart); out-of-range columns of the last
// slab are zero-filled and still counted in the transaction bytes.
__device__ __forceinline__ void tma_load_slab(void* dst, const PmuTensorMap* tmap, int col, int row, void* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(col), "r"(0), "r"(row)
        : "memory");
}
// Copy `len` elements starting `off` elements into a window that begins at ring position `start`
// of a ring of `cap` elements (cap == 0: plain linear window at `base`): one bulk copy, or two when
// the segment crosses the end of the ring.  Sizes/addresses stay 16-byte granular because the wrap
// point is a whole number of hops into the window (checked on the host).
template <typename InT>
__device__ __forceinline__ void bulk_segment(unsigned char* dst, const InT* base, long long start, long long cap,
                                             int off, int len, void* bar)
{
    if (cap == 0) {
        bulk_g2s(dst, base + off, (uint32_t)(len * sizeof(InT)), bar);
        return;
    }
    long long p = start + off;
    if (p >= cap) p -= cap;
    const long long room = cap - p;
    const int n1 = room < (long long)len ? (int)room : len;
    bulk_g2s(dst, base + p, (uint32_t)(n1 * sizeof(InT)), bar);
    if (n1 < len) bulk_g2s(dst + (size_t)n1 * sizeof(InT), base, (uint32_t)((len - n1) * sizeof(InT)), bar);
}

template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(void* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_volatile_s32(const int* p)
{
    return *reinterpret_cast<const volatile int*>(p);
}

// ---- division by a run-time constant without the ~25-instruction integer-divide sequence ----
// q = umulhi(x, floor((2^32-1)/d)) is the true quotient or one less (x < 2^31): one correction.
// PLAIN: an ordinary division instead (the 24-bin kernels have no registers to spare for the magic numbers; their
// per-window work is large enough for the ~20-instruction sequence not to matter).
template <bool PLAIN>
struct FastDivT {
    unsigned d, m;
    __device__ __forceinline__ void init(unsigned dd)
    {
        d = dd;
        if constexpr (!PLAIN) m = 0xFFFFFFFFu / dd;
    }
    __device__ __forceinline__ unsigned div(unsigned x, unsigned* rem) const
    {
        if constexpr (PLAIN) {
            const unsigned q = x / d;
            *rem = x - q * d;
            return q;
        } else {
            unsigned q = __umulhi(x, m);
            unsigned r = x - q * d;
            if (r >= d) { ++q; r -= d; }
            *rem = r;
            return q;
        }
    }
};
typedef FastDivT<false> FastDiv;

// ---- window -> (batch, slot) mapping inside one CTA: balanced batches of <= 32 -----------
template <bool PLAIN>
struct BatchMapT {
    int nb, q, rem;  // batches, base size, number of batches with q+1 windows
    FastDivT<PLAIN> by_q, by_q1, by_F;
    __device__ __forceinline__ void init(int n, int cap = PMU_WARP, int n_frames = 1)
    {
        nb = (n + cap - 1) / cap;
        if (nb < 1) nb = 1;
        q = n / nb;
        rem = n - q * nb;
        by_q.init((unsigned)q);
        by_q1.init((unsigned)(q + 1));
        by_F.init((unsigned)n_frames);
    }
    // group-major item order of a launch with F frames: item -> (frame, pass).
    //   item = (group * F + frame) * size + slot  within the groups of one size  =>  item / size = group * F + frame
    // (two multiply-high divisions by the dividers the batch map keeps anyway; the two plain 32-bit divisions this used
    //  to be were ~ 55 instructions per window, 4 % of a streaming launch)
    __device__ __forceinline__ void item_group_major(int item, int F, int* f, int* t) const
    {
        const int big_items = rem * (q + 1) * F;
        unsigned slot, fr;
        if (item < big_items) {
            const unsigned u = by_q1.div((unsigned)item, &slot);
            const unsigned g = by_F.div(u, &fr);
            *f = (int)fr;
            *t = (int)g * (q + 1) + (int)slot;
        } else {
            const unsigned u = by_q.div((unsigned)(item - big_items), &slot);
            const unsigned gg = by_F.div(u, &fr);
            *f = (int)fr;
            *t = rem * (q + 1) + (int)gg * q + (int)slot;
        }
    }
    __device__ __forceinline__ void locate(int t, int* g, int* slot, int* size, int* first) const
    {
        int big = rem * (q + 1);
        unsigned r;
        if (t < big) {
            *g = (int)by_q1.div((unsigned)t, &r); *slot = (int)r; *size = q + 1; *first = *g * (q + 1);
        } else {
            int gg = (int)by_q.div((unsigned)(t - big), &r);
            *g = rem + gg; *slot = (int)r; *size = q; *first = big + gg * q;
        }
    }
};

typedef BatchMapT<false> BatchMap;

// ---- the producer's walk over the launch's slabs, in the order the consumers expect them in the ring ----------------
// Items (one warp pass = one window, or WP packed ones) are numbered in ticket order: frame-major, or group-major for
// streaming launches.  A single-slab item is one ring position.  Multi-slab windows are INTERLEAVED: the items are taken
// in groups of C = n_cons consecutive ones (one per consumer warp) and the ring carries slab 0 of every item of the
// group, then slab 1 of every item, ...  Filling the ring window after window would put all of one window's slabs into
// the stages at once (a 16384-sample window is 8 slabs) and leave the other warps with nothing to work on.
//   position of (item, slab) = gstart * nslab + slab * Cg + (item - gstart),  gstart = item / C * C, Cg = min(C, n_items - gstart)
struct ItemWalker {
    int f, t, sl;           // frame, pass and slab of the current ring position
    long long ring_pos;     // ring position of frame f's first sample (streaming front end)
    int group_major, n_packs, F, hop, nslab, C, n_items;
    long long cap, pos0;
    int g, slot, gsize, gfirst, q, rem;      // single-slab walk
    int gstart, Cg, j;                       // multi-slab walk
    __device__ __forceinline__ void set_item(int item)
    {
        if (group_major) {
            const int big_items = rem * (q + 1) * F;
            if (item < big_items) {
                const int gg = item / ((q + 1) * F), r = item - gg * (q + 1) * F;
                f = r / (q + 1);
                t = gg * (q + 1) + (r - f * (q + 1));
            } else {
                const int i2 = item - big_items;
                const int gg = i2 / (q * F), r = i2 - gg * q * F;
                f = r / q;
                t = rem * (q + 1) + gg * q + (r - f * q);
            }
        } else {
            f = item / n_packs;
            t = item - f * n_packs;
        }
        ring_pos = pos0 + (long long)f * hop;
        if (cap && ring_pos >= cap) ring_pos -= cap;
    }
    template <class Map>
    __device__ __forceinline__ void init(const Map& bm, int n_packs_, int n_frames, int group_major_, long long ring_cap,
                                         long long ring_start, int hop_, int nslab_, int n_cons, int interleave)
    {
        group_major = group_major_ && n_frames > 1;
        n_packs = n_packs_; F = n_frames; hop = hop_; cap = ring_cap; nslab = nslab_; C = interleave ? n_cons : 1;
        n_items = n_packs_ * n_frames;
        pos0 = ring_cap ? ring_start % ring_cap : 0;
        ring_pos = pos0;
        f = 0; t = 0; sl = 0; g = 0; slot = 0; gfirst = 0;
        q = bm.q; rem = bm.rem;
        gsize = rem > 0 ? q + 1 : q;
        gstart = 0; j = 0;
        Cg = C < n_items ? C : n_items;
    }
    __device__ __forceinline__ void next_frame()
    {
        ++f;
        ring_pos += hop;
        if (cap && ring_pos >= cap) ring_pos -= cap;
    }
    __device__ __forceinline__ void advance()      // to the next ring position
    {
        if (nslab > 1 && C > 1) {
            if (++j == Cg) {
                j = 0;
                if (++sl == nslab) {
                    sl = 0;
                    gstart += Cg;
                    Cg = C < n_items - gstart ? C : n_items - gstart;
                }
            }
            set_item(gstart + j);
            return;
        }
        if (++sl < nslab) return;      // window after window: the next slab of the same item
        sl = 0;
        if (!group_major) {
            if (++t == n_packs) { t = 0; next_frame(); }
            return;
        }
        if (++slot == gsize) {
            slot = 0;
            next_frame();
            if (f == F) {
                f = 0;
                ring_pos = pos0;
                gfirst += gsize;
                ++g;
                gsize = g < rem ? q + 1 : q;
            }
        }
        t = gfirst + slot;
    }
};

// ---- estimator with lane = bin: one window per warp (small launches) ---------------------------------
// Same arithmetic as pmu_math.cuh, bin j evaluated by lane j: the reference's find3LargestIndx (:705-720)
// becomes a warp-shuffle argmax that keeps the first strict maximum (and lets a NaN win only at bin 0), the
// energy sums (:219-220) become shuffle reductions; everything scalar is computed redundantly by all lanes.
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// The peak search with lane = bin.  |Y_j|^2 >= 0, and the bit patterns of non-negative doubles order like unsigned
// integers, so the maximum is two warp-wide integer reductions (REDUX: high words, then the low words of the lanes that
// tie on the high word) and the first maximum is the lowest lane of a ballot - a third of the latency of a five-round
// shuffle butterfly on (value, index) pairs.  NaNs are keyed as 0 (never win) unless at bin 0.  Every lane takes the
// square root of its own bin while the reductions are in flight, so the interpolation tail starts from magnitudes.
__device__ __forceinline__ int ipdft_warp(const PostParams& p, double yr, double yi, int j, bool valid, Tone* out, int* km_out)
{
    const double v = yr * yr + yi * yi;
    const double mag = pmu_sqrt(v);
    const unsigned long long bits = (valid && v == v) ? (unsigned long long)__double_as_longlong(v) : 0ull;
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned hi_max = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lo_max = __reduce_max_sync(0xffffffffu, hi == hi_max ? lo : 0u);
    const unsigned winners = __ballot_sync(0xffffffffu, hi == hi_max && lo == lo_max);
    const double v0 = __shfl_sync(0xffffffffu, v, 0);
    int km = __ffs((int)winners) - 1;
    km = (v0 != v0 || km >= p.K) ? 0 : km;              // a NaN wins only at bin 0; idle lanes (keyed 0) never
    const double mk = __shfl_sync(0xffffffffu, mag, km);
    const double ml = __shfl_sync(0xffffffffu, mag, km > 0 ? km - 1 : 0);                // kl = km when km == 0 (:718)
    const double mr_raw = __shfl_sync(0xffffffffu, mag, km + 1 < p.K ? km + 1 : km);
    const double pyr = __shfl_sync(0xffffffffu, yr, km);
    const double pyi = __shfl_sync(0xffffffffu, yi, km);
    const double mr = (km == p.K - 1) ? ml : mr_raw;                                     // kr = km-1 at the last bin (:719)
    if (km_out) *km_out = km;
    return ipdft_tail(p, km, mk, km == 0 ? mk : ml, mr, pyr, pyi, out);
}

// value at bin j of wf(j, f, z) for this lane (three reciprocals, no sliding window)
__device__ __forceinline__ void sweep_bin(const PostParams& p, const Sweep& sw, int j, double* wr, double* wi)
{
    const int m = j - sw.kstar;
    sw.eval(p, j, sw.inv_den(p, m - 1), sw.inv_den(p, m), sw.inv_den(p, m + 1), wr, wi);
}

// e_ipDFT (:626-664) with lane = bin; one copy of the pass for the plain and the compensated iterations (see
// ipdft_pass in pmu_math.cuh for why).  Returns the peak bin of the first pass.
__device__ __forceinline__ int e_ipdft_warp(const PostParams& p, double yr, double yi, int j, bool valid, Tone* t)
{
    Tone cur;
    cur.amp = 0; cur.cr = 1; cur.ci = 0; cur.freq = 0;
    int km_first = 0;
    for (int it = 0; it <= p.P; ++it) {
        double wr = 0.0, wi = 0.0;
        if (it > 0) {
            Sweep sw;
            sw.init(p, -cur.freq, cur.amp * cur.cr, -cur.amp * cur.ci);
            sweep_bin(p, sw, j, &wr, &wi);
        }
        Tone nxt;
        int km;
        const int done = ipdft_warp(p, yr - wr, yi - wi, j, valid, &nxt, &km);
        if (it == 0) km_first = km;
        cur = nxt;
        if (done) break;
    }
    *t = cur;
    return km_first;
}

// X - pureTone(t) at this lane's bin
__device__ __forceinline__ void residual_bin(const PostParams& p, double xr, double xi, int j, const Tone& t, double* wr_, double* wi_)
{
    Sweep sp, sn;
    sp.init(p, t.freq, t.amp * t.cr, t.amp * t.ci);
    sn.init(p, -t.freq, t.amp * t.cr, -t.amp * t.ci);
    double ar, ai, br, bi;
    sweep_bin(p, sp, j, &ar, &ai);
    sweep_bin(p, sn, j, &br, &bi);
    *wr_ = xr - (ar + br);
    *wi_ = xi - (ai + bi);
}

// ROCOF stage, frame fill and stores for one window (src/pmu_estimator.c:236-265); one thread.
template <typename InT>
__device__ __forceinline__ void finish_window(const KernelArgs& a, const Tone& f, int dbg, long long w, int frame)
{
    const PostParams& pp = a.post;
    const long long wf = (long long)frame * a.frame_stride + w;
    if (a.defer_wait) asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous launch left this stream's state
    RocofState rs;
    rs.f_old = __ldcg(&a.st_fold[w]);
    rs.dl0 = __ldcg(&a.st_dl0[w]);
    rs.dl1 = __ldcg(&a.st_dl1[w]);
    rs.st = __ldcg(&a.st_state[w]);
    const double y = sizeof(InT) == 4 ? rocof_step_f32(pp, f.freq, &rs) : rocof_step(pp, f.freq, &rs);
    a.st_fold[w] = rs.f_old;
    a.st_dl0[w] = rs.dl0;
    a.st_dl1[w] = rs.dl1;
    a.st_state[w] = (unsigned char)rs.st;
    double mid;
    if (a.mid) {
        mid = (double)reinterpret_cast<const InT*>(a.mid)[(long long)frame * a.mid_frame_stride + w / pp.C];
    } else {
        const unsigned long long c = (a.t0_samples + (unsigned long long)frame * (unsigned)a.hop + (unsigned)(a.N / 2)) % (unsigned)a.fs;
        mid = (double)(InT)((double)c / (double)a.fs);
    }
    double fr[4];
    fill_frame(pp, f, y, mid, fr);
    InT* out = reinterpret_cast<InT*>(a.frames) + (size_t)wf * 4;
    out[0] = (InT)fr[0]; out[1] = (InT)fr[1]; out[2] = (InT)fr[2]; out[3] = (InT)fr[3];
    if (a.dbg) a.dbg[wf] = dbg;
    if (a.state_export && frame == a.n_frames - 1) {
        double* e = a.state_export + (size_t)w * 4;
        e[0] = rs.f_old; e[1] = rs.dl0; e[2] = rs.dl1; e[3] = (double)rs.st;
    }
}

template <typename InT, bool PRIVATE_PP>
__device__ __noinline__ void post_window_warp(const KernelArgs& a, const double* row, long long w, int frame, int lane)
{
    // a private copy (registers): through the reference the loop-invariant parameters are re-read with generic loads
    // (~30 cycles each on the single warp's critical path) wherever the compiler cannot prove they do not alias a store
    PostParams pp_copy;
    const PostParams* pp_ptr = &a.post;
    if constexpr (PRIVATE_PP) {
        pp_copy = a.post;
        pp_ptr = &pp_copy;
    }
    const PostParams& pp = *pp_ptr;
    const int K = pp.K;
    const bool valid = lane < K;
    const int j = valid ? lane : K - 1;
    const long long wf = (long long)frame * a.frame_stride + w;
    const pmu_c xw = hann_combine(row, j);
    const double xr = xw.x, xi = xw.y;
    if (a.spec_out && valid) {
        a.spec_out[((size_t)wf * K + j) * 2] = a.spec_raw ? row[2 * j] : xr;
        a.spec_out[((size_t)wf * K + j) * 2 + 1] = a.spec_raw ? row[2 * j + 1] : xi;
    }
    // the same single loop as estimate_tone (pmu_math.cuh): "estimate a tone from Y, subtract it from X into W"
    Tone f;
    f.amp = 0; f.cr = 1; f.ci = 0; f.freq = 0;
    int dbg = 0;
    double yr = xr, yi = xi;
    const int last = pp.iter_en ? 2 * pp.Q : 0;
    for (int s = 0; s <= last; ++s) {
        Tone t;
        const int km = e_ipdft_warp(pp, yr, yi, j, valid, &t);        // :205, :679, :685
        if (s == 0) dbg = km & PMU_DBG_KM_MASK;
        if (!(s & 1)) f = t;
        if (s == last || !pp.iter_en) break;
        residual_bin(pp, xr, xi, j, t, &yr, &yi);                     // :206-215, :680-684, :690-695
        if (s == 0) {
            const double Ed = warp_sum(valid ? yr * yr + yi * yi : 0.0);   // :219
            const double E = warp_sum(valid ? xr * xr + xi * xi : 0.0);    // :220
            if (!(Ed > pp.eps * E)) break;                            // :229
            dbg |= PMU_DBG_TRIGGERED;
        }
    }
    if (lane == 0) finish_window<InT>(a, f, dbg, w, frame);
}

// ---- post-processing of one batch (lane = window) --------------------------------------
// Batch = <= 32 consecutive streams of this CTA at frame f.  w_first = index (within this
// launch) of the batch's first stream.  The estimator's two bin arrays need no scratch of their own:
// X is the batch's raw-bin buffer, Hann-combined in place (a lane-private row of 16-byte entries), and the
// residual spectrum W lives in the posting warp's lane-reduction scratch, which is idle until the warp
// transforms its next window.  Every consumer warp can therefore post a batch at the same time; the raw
// buffer goes back to the ring when the estimator is done with X.
// PRIVATE_PP: work on a register copy of the parameters (see post_window_warp).  The 24-bin kernels keep the reference:
// their transform already needs every register, and the copy would make the call site spill.
template <typename InT, bool PRIVATE_PP>
__device__ __noinline__ void post_batch(const KernelArgs& a, int raw_buf, long long w_first, int frame_group, int bsz,
                                        int batch_id)
{
    // (few scalar arguments: beyond the register-passed ones the ABI goes through local memory)
    const int lane = threadIdx.x % PMU_WARP;
    const int frame = frame_group >> 8, group = frame_group & 0xff;          // group < PMU_MAX_GROUPS = 256
    const unsigned w_off = (unsigned)a.off_red + (unsigned)(threadIdx.x / PMU_WARP - 1) * (unsigned)a.red_bytes;
    unsigned char* smem = pmu_smem;
    PostParams pp_copy;
    const PostParams* pp_ptr = &a.post;
    if constexpr (PRIVATE_PP) {
        pp_copy = a.post;
        pp_ptr = &pp_copy;
    }
    const PostParams& pp = *pp_ptr;
    CtrlBlock* ctrl = reinterpret_cast<CtrlBlock*>(smem + a.off_ctrl);
    const int K = pp.K;
    double* row = reinterpret_cast<double*>(smem + a.off_raw) + ((size_t)raw_buf * PMU_WARP + lane) * a.raw_stride;
    const Col X = Col::make(reinterpret_cast<pmu_c*>(row), 1);
    const Col Wk = Col::make(reinterpret_cast<pmu_c*>(smem + w_off) + lane, PMU_WARP);
    const bool active = lane < bsz;
    const long long w = w_first + lane;                         // stream index within the launch
    const long long wf = (long long)frame * a.frame_stride + w; // row in the frame-major arrays

    Tone t;
    int dbg = 0;
    if (active) {
        if (a.spec_out && a.spec_raw) {
            for (int k = 0; k < K; ++k) {
                const pmu_c xr_ = X.get(k);
                a.spec_out[((size_t)wf * K + k) * 2] = xr_.x;
                a.spec_out[((size_t)wf * K + k) * 2 + 1] = xr_.y;
            }
        }
        hann_combine_inplace(X, K);
        if (a.spec_out && !a.spec_raw) {
            for (int k = 0; k < K; ++k) {
                const pmu_c xw = X.get(k);
                a.spec_out[((size_t)wf * K + k) * 2] = xw.x;
                a.spec_out[((size_t)wf * K + k) * 2 + 1] = xw.y;
            }
        }
        dbg = estimate_tone(pp, X, Wk, &t);
    }
    // hand the raw buffer to batch (batch_id + n_raw)
    __syncwarp();
    if (lane == 0) {
        ctrl->batch_cnt[raw_buf] = 0;
        __threadfence_block();
        *reinterpret_cast<volatile int*>(&ctrl->raw_owner[raw_buf]) = batch_id + a.n_raw;
    }

    // Launches that carry several frames: the estimate above has no cross-frame dependency (the batches of a stream group's
    // consecutive frames run concurrently on different warps); only the ROCOF recurrence below needs the state its
    // predecessor frame leaves, so the hand-over (a shared-memory counter per group) is waited for HERE, not before the
    // estimate.
    if (a.n_frames > 1) {
        if (lane == 0)
            while (ld_volatile_s32(&ctrl->grp_done[group]) != frame) __nanosleep(64);
        __syncwarp();
        __threadfence_block();
    }
    // ... and the previous LAUNCH's, when its wait was deferred to here (pmub_set_overlap)
    if (a.defer_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (active) {
        RocofState rs;
        // the previous frame's state may have been written by another warp of this CTA
        rs.f_old = __ldcg(&a.st_fold[w]);
        rs.dl0 = __ldcg(&a.st_dl0[w]);
        rs.dl1 = __ldcg(&a.st_dl1[w]);
        rs.st = __ldcg(&a.st_state[w]);
        double y = sizeof(InT) == 4 ? rocof_step_f32(pp, t.freq, &rs) : rocof_step(pp, t.freq, &rs);
        a.st_fold[w] = rs.f_old;
        a.st_dl0[w] = rs.dl0;
        a.st_dl1[w] = rs.dl1;
        a.st_state[w] = (unsigned char)rs.st;
        double mid;
        if (a.mid) {
            mid = (double)reinterpret_cast<const InT*>(a.mid)[(long long)frame * a.mid_frame_stride + w / pp.C];
        } else {
            // centre of the window, modulo one second, from the running sample counter
            const unsigned long long c = (a.t0_samples + (unsigned long long)frame * (unsigned)a.hop + (unsigned)(a.N / 2)) % (unsigned)a.fs;
            mid = (double)(InT)((double)c / (double)a.fs);
        }
        double fr[4];
        fill_frame(pp, t, y, mid, fr);
        InT* out = reinterpret_cast<InT*>(a.frames) + (size_t)wf * 4;
        if constexpr (sizeof(InT) == 8) {
            double2* o2 = reinterpret_cast<double2*>(out);
            o2[0] = make_double2(fr[0], fr[1]);
            o2[1] = make_double2(fr[2], fr[3]);
        } else {
            *reinterpret_cast<float4*>(out) = make_float4((float)fr[0], (float)fr[1], (float)fr[2], (float)fr[3]);
        }
        if (a.dbg) a.dbg[wf] = dbg;
        if (a.state_export && frame == a.n_frames - 1) {
            double2* e2 = reinterpret_cast<double2*>(a.state_export + (size_t)w * 4);
            e2[0] = make_double2(rs.f_old, rs.dl0);
            e2[1] = make_double2(rs.dl1, (double)rs.st);
        }
    }
    __syncwarp();
    if (a.n_frames > 1 && lane == 0) {
        __threadfence_block();
        *reinterpret_cast<volatile int*>(&ctrl->grp_done[group]) = frame + 1;
    }
}

// Load the M samples of one residue from a stage (row pitch in elements).  FULL: every lane of
// the group is a valid residue, no masking.
template <typename InT, typename CT, int M, int PITCH, bool FULL>
__device__ __forceinline__ void load_residue(const InT* xs, int pitch, int a_loc, int w_s, CT* x)
{
    if constexpr (IsPacked<CT>::value) {
        // residues a_loc (even) and a_loc + 1 of the window: one 64-bit load per row when the pitch is even
        const bool v0 = FULL || a_loc < w_s, v1 = FULL || a_loc + 1 < w_s;
        const int a_c = v0 ? a_loc : 0;
#pragma unroll
        for (int b = 0; b < M; ++b) {
            float f0, f1;
            if constexpr (PITCH > 0 && PITCH % 2 == 0) {
                const float2 v = *reinterpret_cast<const float2*>(xs + b * pitch + a_c);
                f0 = v.x; f1 = v.y;
            } else {
                f0 = (float)xs[b * pitch + a_c];
                f1 = (float)xs[b * pitch + (v1 ? a_c + 1 : a_c)];
            }
            x[b] = CT(v0 ? f0 : 0.0f, v1 ? f1 : 0.0f);
        }
    } else if constexpr (FULL) {
#pragma unroll
        for (int b = 0; b < M; ++b) x[b] = (CT)xs[b * pitch + a_loc];
    } else {
        const bool valid = a_loc < w_s;
        const int a_c = valid ? a_loc : 0;
#pragma unroll
        for (int b = 0; b < M; ++b) {
            CT v = (CT)xs[b * pitch + a_c];
            x[b] = valid ? v : (CT)0;
        }
    }
}

// ---- the kernel ----------------------------------------------------------------------
// PITCH > 0: stage rows are PITCH elements apart (compile-time: the sample loads need no address
// arithmetic); PITCH == 0: row pitch = a.W at run time.
// CT: arithmetic type of the DFT stage (double; float for the float_p = float build) - everything from
// the lane reduction on is double.
// WP: windows per warp pass.  Short windows (a few KB) are processed WP = 4 (or 2) at a time, 8 (16) lanes per window: the
// per-pass costs (ticket, barrier round trip, stage release, batch bookkeeping) are paid once for four windows, no
// lane idles in a masked residue group, and the four windows arrive in one bulk copy.  Only for single-slab windows
// stored at their natural pitch (PITCH == 0 kernels).
template <typename InT, typename CT, int M, int KC, int PITCH, int WP = 1>
__global__ void __launch_bounds__(PMU_BLOCK_THREADS, 1) pmu_fused_kernel(const __grid_constant__ KernelArgs a)
{
    // packed windows lie in the stage exactly as in memory (natural pitch); the plan guarantees pitch == L for them
    // CT = F2 (float build): every lane carries two ADJACENT residues of its window through the packed FP32 pipe
    // (pmu_dft.cuh); the lane partials are scalar again after the final rotation
    constexpr int PK = IsPacked<CT>::value ? 2 : 1;
    constexpr int G = PMU_WARP / WP;        // lanes per window
    typedef typename ScalarOf<CT>::type cscal;
    typedef typename CplxOf<CT>::type ccplx;
    unsigned char* smem = pmu_smem;
    CtrlBlock* ctrl = reinterpret_cast<CtrlBlock*>(smem + a.off_ctrl);
    ccplx* sU = reinterpret_cast<ccplx*>(smem + a.off_U);
    ccplx* sV = reinterpret_cast<ccplx*>(smem + a.off_V);
    pmu_c* sT = reinterpret_cast<pmu_c*>(smem + a.off_trig);
    unsigned char* stages = smem + a.off_stage;

    const int tid = threadIdx.x;
    const int warp = tid / PMU_WARP;
    const int lane = tid % PMU_WARP;
    const int nthreads = blockDim.x;
    // let the next launch in the stream be scheduled as soon as SMs free up (it waits for this grid to finish, above)
    asm volatile("griddepcontrol.launch_dependents;");

    // contiguous, balanced range of streams for this CTA
    const long long w_begin = (a.n_windows * (long long)blockIdx.x) / gridDim.x;
    const long long w_end = (a.n_windows * (long long)(blockIdx.x + 1)) / gridDim.x;
    const int n_local = (int)(w_end - w_begin);

    // ---- one-time setup --------------------------------------------------------------
    if (tid == 0) {
        const uint32_t full_count = (a.load_mode == PMU_LOAD_BULK) ? 1u : (uint32_t)PMU_WARP;
        for (int s = 0; s < a.n_stage; ++s) {
            mbar_init(&ctrl->full[s], full_count);
            mbar_init(&ctrl->empty[s], 1u);
            ctrl->stage_seq[s] = 0;
        }
        ctrl->ticket = 0;
        for (int i = 0; i < PMU_MAX_RAW; ++i) {
            ctrl->batch_cnt[i] = 0;
            ctrl->raw_owner[i] = i;
        }
        mbar_fence_init();
    }
    for (int i = tid; i < PMU_MAX_GROUPS; i += nthreads) ctrl->grp_done[i] = 0;
    for (int i = tid; i < a.n_u; i += nthreads) {
        const pmu_c w = a.U[i];
        ccplx c; c.x = (cscal)w.x; c.y = (cscal)w.y;
        sU[i] = c;
    }
    if constexpr (PK == 2) {
        // lane l rotates its two residues 2l, 2l+1: pairs of twiddles from the 64-lane table
        RPair<CT>* sV2 = reinterpret_cast<RPair<CT>*>(sV);
        for (int i = tid; i < KC * PMU_WARP; i += nthreads) {
            const int k = i / PMU_WARP, l = i % PMU_WARP;
            const pmu_c w0 = a.V[k * 2 * PMU_WARP + 2 * l], w1 = a.V[k * 2 * PMU_WARP + 2 * l + 1];
            RPair<CT> c;
            c.x = CT((float)w0.x, (float)w1.x);
            c.y = CT((float)w0.y, (float)w1.y);
            sV2[i] = c;
        }
    } else {
        for (int i = tid; i < KC * PMU_WARP; i += nthreads) {
            const pmu_c w = a.V[i];
            ccplx c; c.x = (cscal)w.x; c.y = (cscal)w.y;
            sV[i] = c;
        }
    }
    for (int i = tid; i < a.post.jlo + a.post.jhi + 1; i += nthreads) sT[i] = a.trig[i];
    // Everything above reads only this estimator's own tables.  From here on the kernel touches the caller's windows,
    // the frames and the ROCOF state: wait for whatever precedes this launch in the stream (a no-op without PDL).
    if (!a.defer_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    if (n_local <= 0) return;

    const size_t win_bytes = (size_t)a.N * sizeof(InT);
    const int n_packs = (n_local + WP - 1) / WP;   // warp passes per frame (= streams when WP == 1)
    const int n_items = n_packs * a.n_frames;      // (frame, pass) pairs, frame-major
    const int total_slabs = n_items * a.nslab;
    const int pitch = PITCH > 0 ? PITCH : a.W;

    if (warp == 0) {
        // ================= PRODUCER =================
        if (a.load_mode == PMU_LOAD_BULK) {
            // A window that is already laid out at the stage pitch is one bulk copy (lane 0); WP such windows are
            // consecutive in the caller's array and still one copy.  A slab is one tiled TMA copy, or M row copies:
            // lane b issues row b, so the address arithmetic of the rows runs in parallel instead of as one thread's
            // serial chain (which capped multi-slab windows at ~13 M copies/s per SM).
            const bool whole = (a.nslab == 1 && pitch == a.L);
            int st = 0;                            // stage of ring position i
            uint32_t par = 0;
            BatchMap pbm;
            pbm.init(n_packs, PMU_WARP / WP);
            ItemWalker wk;                         // (frame, pass, slab) at ring position i
            wk.init(pbm, n_packs, a.n_frames, a.group_major, a.ring_cap, a.ring_start, a.hop, a.nslab, a.n_cons, a.interleave);
            for (int i = 0; i < total_slabs; ++i) {
                const int t = wk.t, f = wk.f, sl = wk.sl;
                const long long ring_pos = wk.ring_pos;
                if (lane == 0) mbar_wait(&ctrl->empty[st], par ^ 1u);
                __syncwarp();
                unsigned char* dst = stages + (size_t)st * a.stage_bytes;
                const int tw = t * WP;                                   // first stream of the pass
                const int cnt = (WP == 1) ? 1 : min(WP, n_local - tw);   // streams in it
                const InT* src;
                long long start = 0;
                if (a.ring_cap == 0) {
                    src = reinterpret_cast<const InT*>(a.windows) + ((size_t)f * a.frame_stride + (size_t)(w_begin + tw)) * a.N;
                } else {
                    src = reinterpret_cast<const InT*>(a.windows) + (size_t)(w_begin + tw) * a.ring_pitch;
                    start = ring_pos;
                }
                if (whole) {
                    if (lane == 0) {
                        mbar_expect_tx(&ctrl->full[st], (uint32_t)(win_bytes * cnt));
                        if (a.ring_cap == 0) {
                            bulk_g2s(dst, src, (uint32_t)(win_bytes * cnt), &ctrl->full[st]);
                        } else {
                            for (int c = 0; c < cnt; ++c)
                                bulk_segment<InT>(dst + (size_t)c * win_bytes, src + (size_t)c * a.ring_pitch, start, a.ring_cap, 0, a.N,
                                                  &ctrl->full[st]);
                        }
                    }
                } else if (WP == 1 && a.use_tmap) {
                    // TMA-tiled: the copy engine walks the M rows of the slab itself
                    if (lane == 0) {
                        mbar_expect_tx(&ctrl->full[st], (uint32_t)(pitch * sizeof(InT)) * M);
                        if (a.ring_cap == 0)
                            tma_load_slab(dst, &a.tmap, sl * a.W, (int)((long long)f * a.frame_stride + w_begin + t), &ctrl->full[st]);
                        else   // mirrored ring: the window is contiguous from its start position
                            tma_load_slab(dst, &a.tmap, (int)ring_pos + sl * a.W, (int)(w_begin + t), &ctrl->full[st]);
                    }
                } else if (WP == 1) {
                    const int a_lo = sl * a.W;
                    const int w_s = min(a.W, a.L - a_lo);
                    if (lane == 0) mbar_expect_tx(&ctrl->full[st], (uint32_t)(w_s * sizeof(InT)) * M);
                    __syncwarp();
                    for (int b = lane; b < M; b += PMU_WARP)
                        bulk_segment<InT>(dst + (size_t)b * pitch * sizeof(InT), src, start, a.ring_cap, a_lo + a.L * b, w_s,
                                          &ctrl->full[st]);
                }
                if (++st == a.n_stage) { st = 0; par ^= 1u; }
                wk.advance();
            }
        } else {
            int st = 0;
            uint32_t par = 0;
            BatchMap pbm;
            pbm.init(n_packs, PMU_WARP / WP);
            ItemWalker wk;
            wk.init(pbm, n_packs, a.n_frames, a.group_major, a.ring_cap, a.ring_start, a.hop, a.nslab, a.n_cons, a.interleave);
            for (int i = 0; i < total_slabs; ++i) {
                const int t = wk.t, f = wk.f, sl = wk.sl;
                const long long ring_pos = wk.ring_pos;
                mbar_wait(&ctrl->empty[st], par ^ 1u);
                InT* dst = reinterpret_cast<InT*>(stages + (size_t)st * a.stage_bytes);
                const int tw = t * WP;
                const int cnt = (WP == 1) ? 1 : min(WP, n_local - tw);
                const InT* src;
                long long start = 0;
                if (a.ring_cap == 0) {
                    src = reinterpret_cast<const InT*>(a.windows) + ((size_t)f * a.frame_stride + (size_t)(w_begin + tw)) * a.N;
                } else {
                    src = reinterpret_cast<const InT*>(a.windows) + (size_t)(w_begin + tw) * a.ring_pitch;
                    start = ring_pos;
                }
                if constexpr (WP == 1) {
                    const int a_lo = sl * a.W;
                    const int w_s = min(a.W, a.L - a_lo);
                    for (int b = 0; b < M; ++b)
                        for (int e = lane; e < w_s; e += PMU_WARP) {
                            long long p = start + a_lo + (long long)a.L * b + e;
                            if (a.ring_cap != 0 && p >= a.ring_cap) p -= a.ring_cap;
                            cp_async_elem<(int)sizeof(InT)>(dst + (size_t)b * pitch + e, src + p);
                        }
                } else {
                    // natural pitch: the stage holds the cnt windows back to back, exactly as they lie in memory
                    const size_t wstride = a.ring_cap == 0 ? (size_t)a.N : (size_t)a.ring_pitch;
                    for (int c = 0; c < cnt; ++c)
                        for (int e = lane; e < a.N; e += PMU_WARP) {
                            long long p = start + e;
                            if (a.ring_cap != 0 && p >= a.ring_cap) p -= a.ring_cap;
                            cp_async_elem<(int)sizeof(InT)>(dst + (size_t)c * a.N + e, src + (size_t)c * wstride + p);
                        }
                }
                cp_async_mbar_arrive(&ctrl->full[st]);
                if (++st == a.n_stage) { st = 0; par ^= 1u; }
                wk.advance();
            }
        }
        return;
    }
    if (warp > a.n_cons) return;

    // ================= CONSUMERS =================
    unsigned long long* tl = a.timeline ? a.timeline + ((size_t)blockIdx.x * (PMU_BLOCK_THREADS / PMU_WARP) + warp) * 4 : nullptr;
    if (tl && lane == 0) { tl[0] = global_ns(); tl[1] = 0; }
    BatchMapT<(KC >= 24)> bm;             // batches of <= 32 streams = <= 32 / WP passes
    bm.init(n_packs, PMU_WARP / WP, a.n_frames);
    FastDivT<(KC >= 24)> by_nlocal, by_nstage;
    by_nlocal.init((unsigned)n_packs);
    const int sub = lane / G;             // which stream of the pass this lane works on (0 when WP == 1)
    const int gl = lane % G;              // its lane index within that stream's group
    by_nstage.init((unsigned)a.n_stage);
    const int w_per_it = PK == 1 ? a.W / PMU_WARP : a.u_per_slab;   // rows of U per slab (W % 32 == 0 when an unpacked plan has several slabs)
    FastDivT<(KC >= 24)> by_nraw;
    by_nraw.init((unsigned)a.n_raw);
    const unsigned red_off = (unsigned)a.off_red + (unsigned)(warp - 1) * (unsigned)a.red_bytes;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&ctrl->ticket, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) {
            if (tl && lane == 0) tl[2] = global_ns();
            break;
        }
        int f = 0, t = item;
        if (a.n_frames > 1) {
            if (a.group_major) {
                bm.item_group_major(item, a.n_frames, &f, &t);
            } else {
                unsigned r;
                f = (int)by_nlocal.div((unsigned)item, &r);
                t = (int)r;
            }
        }

        // ---- pruned rectangular DFT of pass t (streams t*WP ... t*WP + WP-1 of this CTA), frame f ----
        CT acc[2 * KC];  // ar = acc[0..KC), ai = acc[KC..2KC)
        CT* ar = acc;
        CT* ai = acc + KC;
        bool first = true;
        // ring position of the item's slabs (interleaved within groups of n_cons items: see ItemWalker)
        int idx0 = item * a.nslab, idx_step = 1;
        if (WP == 1 && a.nslab > 1 && a.interleave) {   // packed windows are single-slab by construction
            const int gstart = item / a.n_cons * a.n_cons;
            idx_step = min(a.n_cons, n_items - gstart);
            idx0 = gstart * a.nslab + (item - gstart);
        }
        for (int sl = 0; sl < a.nslab; ++sl) {
            const int idx = idx0 + sl * idx_step;
            unsigned st_u;
            const int use = (int)by_nstage.div((unsigned)idx, &st_u);
            const int st = (int)st_u;
            // Tickets are claimed out of order w.r.t. the ring: a parity wait is only unambiguous
            // once the stage's previous occupant (slab idx - n_stage) has been released.
            if (lane == 0)
                while (ld_volatile_s32(&ctrl->stage_seq[st]) != use) __nanosleep(32);
            __syncwarp();
            mbar_wait(&ctrl->full[st], (uint32_t)(use & 1));
            if (tl && lane == 0 && tl[1] == 0) tl[1] = global_ns();
            const InT* xs = reinterpret_cast<const InT*>(stages + (size_t)st * a.stage_bytes) + (WP == 1 ? 0 : sub * a.N);
            const int a_lo = sl * a.W;
            const int w_s = min(a.W, a.L - a_lo);
            const int n_it = (w_s + PK * G - 1) / (PK * G);
            const bool all_full = (w_s % (PK * G)) == 0;
            auto residue_group = [&](int ii) {
                const int a_loc = PK * (gl + G * ii);
                CT x[M];
                // only the last residue group of a slab can be partly out of range
                if (all_full || ii + 1 < n_it) load_residue<InT, CT, M, PITCH, true>(xs, pitch, a_loc, w_s, x);
                else load_residue<InT, CT, M, PITCH, false>(xs, pitch, a_loc, w_s, x);
                const ccplx* u = sU + (size_t)(sl * w_per_it + ii) * KC;
                if (first) {
                    accumulate_residue<M, KC, true, CT>(x, u, ar, ai);
                    first = false;
                } else {
                    accumulate_residue<M, KC, false, CT>(x, u, ar, ai);
                }
            };
            for (int ii = 0; ii < n_it; ++ii) residue_group(ii);
            __syncwarp();
            if (lane == 0) {
                *reinterpret_cast<volatile int*>(&ctrl->stage_seq[st]) = use + 1;
                mbar_arrive(&ctrl->empty[st]);
            }
        }
        if constexpr (PK == 2) rotate_lane_pairs<KC, CT>(reinterpret_cast<const RPair<CT>*>(sV) + gl, PMU_WARP, ar, ai);
        else rotate_lane<KC, CT>(sV + gl, PMU_WARP, ar, ai);

        // ---- sum the 32 lane partials through shared memory: every lane drops its KC complex
        //      partials into its row of the warp's scratch (128-bit stores, conflict-free row
        //      stride), then lane c adds up column c of the 32 rows and deposits the total ----
        constexpr int RS = 2 * KC + 2;  // row stride in scalar elements: an odd number of complex slots
        cscal* red = reinterpret_cast<cscal*>(smem + red_off);
        {
            RPair<cscal>* my = reinterpret_cast<RPair<cscal>*>(red + (size_t)lane * RS);
#pragma unroll
            for (int k = 0; k < KC; ++k)
                if (k < a.Kp) {
                    RPair<cscal> c;
                    c.x = hsum(ar[k]);      // packed kernels: the two residues' partials are added here
                    c.y = hsum(ai[k]);
                    my[k] = c;
                }
        }
        if (WP == 1 && a.warp_per_window) {
            // small launch: this warp finishes its own window right away, lane = bin
            double* myrow = reinterpret_cast<double*>(smem + a.off_raw) + (size_t)(warp - 1) * a.raw_stride;
            __syncwarp();
            for (int c = lane; c < 2 * a.Kp; c += PMU_WARP) {
                const cscal* col = red + c;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int l = 0; l < PMU_WARP; l += 4) {
                    s0 += ct_to_double(col[(l + 0) * RS]);
                    s1 += ct_to_double(col[(l + 1) * RS]);
                    s2 += ct_to_double(col[(l + 2) * RS]);
                    s3 += ct_to_double(col[(l + 3) * RS]);
                }
                myrow[c] = (s0 + s1) + (s2 + s3);
            }
            __syncwarp();
            // frames of one stream are claimed by tickets in frame order, but by different warps: frame f waits
            // for the ROCOF state written by frame f-1 of the same stream
            if (a.n_frames > 1) {
                if (lane == 0)
                    while (ld_volatile_s32(&ctrl->grp_done[t]) != f) __nanosleep(64);
                __syncwarp();
                __threadfence_block();
            }
            post_window_warp<InT, (KC < 24)>(a, myrow, w_begin + t, f, lane);
            if (a.n_frames > 1) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    *reinterpret_cast<volatile int*>(&ctrl->grp_done[t]) = f + 1;
                }
            }
            continue;
        }
        int g, slot, bsz, first_t;                 // in passes; converted to streams below when WP > 1
        bm.locate(t, &g, &slot, &bsz, &first_t);
        const int gb = a.group_major ? g * a.n_frames + f : f * bm.nb + g;   // batch id over the whole launch, in ticket order
        unsigned rb_u;
        by_nraw.div((unsigned)gb, &rb_u);
        const int rb = (int)rb_u;
        if (lane == 0) {
            while (ld_volatile_s32(&ctrl->raw_owner[rb]) != gb) __nanosleep(64);
        }
        __syncwarp();
        __threadfence_block();
        double* row = reinterpret_cast<double*>(smem + a.off_raw) + ((size_t)rb * PMU_WARP + slot * WP) * a.raw_stride;
        {
#pragma unroll
        for (int s2 = 0; s2 < WP; ++s2) {          // stream s2 of the pass: rows s2*G ... s2*G + G-1 of the scratch
            if (WP > 1 && t * WP + s2 >= n_local) break;   // the CTA's last pass may be short
            for (int c = lane; c < 2 * a.Kp; c += PMU_WARP) {
                const cscal* col = red + (size_t)s2 * G * RS + c;
                double s0 = 0.0, s1 = 0.0, s2_ = 0.0, s3 = 0.0;   // the column sums are always double
#pragma unroll
                for (int l = 0; l < G; l += 4) {
                    s0 += ct_to_double(col[(l + 0) * RS]);
                    s1 += ct_to_double(col[(l + 1) * RS]);
                    s2_ += ct_to_double(col[(l + 2) * RS]);
                    s3 += ct_to_double(col[(l + 3) * RS]);
                }
                row[(size_t)s2 * a.raw_stride + c] = (s0 + s1) + (s2_ + s3);
            }
        }
        }
        __syncwarp();
        int done = 0;
        if (lane == 0) {
            __threadfence_block();
            done = atomicAdd(&ctrl->batch_cnt[rb], 1) + 1;
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done == bsz) {
            if constexpr (WP > 1) {                // passes -> streams
                first_t *= WP;
                bsz = min(bsz * WP, n_local - first_t);
            }
            // ---- this warp post-processes batch (f, g) ----
            __threadfence_block();
            post_batch<InT, (KC < 24)>(a, rb, w_begin + first_t, (f << 8) | g, bsz, gb);
        }
    }
    if (tl && lane == 0) tl[3] = global_ns();
}

#endif  // __CUDACC__
