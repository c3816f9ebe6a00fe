// pmu_wide.cuh — the general estimator kernel: any window length, any number of bins.
//
// The fused kernel (pmu_kernel.cuh) keeps a lane's partial spectrum in registers and therefore stops at
// n_bins = 23.  The reference accepts every n_cycles + 2 <= n_bins <= N/2 (check_config_validity,
// src/pmu_estimator.c:386-449); those remaining configurations (long analysis windows of >= 22 cycles, or a caller
// that simply asks for many bins) run here: one warp per window (own scratch in shared memory),
//   1. the K+1 lowest rectangular-window bins: when 16 | N, the fused kernel's pruned FFT (16-point real codelets per
//      residue, pmu_dft.cuh) run once per chunk of 16 bins, a warp per chunk, twiddles from the full table W_N^m;
//      otherwise by direct summation (dft_r's non-power-of-two branch, :532-545, for the bins the estimator reads
//      only), a warp per bin;
//   2. Hann window in the frequency domain (pmu_dft.cuh);
//   3. the estimator with the bins strided over the lanes: peak search = per-lane scan + warp-shuffle argmax.
// This is the coverage path, not the roofline path (samples come through L1, the transform is repeated per chunk).
#pragma once

#include "pmu_kernel.cuh"

#if defined(__CUDACC__)

#define PMU_WIDE_THREADS 256

// y = base[j] - (sw ? wf_sw(j) : 0)
__device__ __forceinline__ pmu_c wide_bin(const PostParams& p, const pmu_c* base, const Sweep* sw, int j)
{
    pmu_c y = base[j];
    if (sw) {
        double wr, wi;
        sweep_bin(p, *sw, j, &wr, &wi);
        y.x -= wr;
        y.y -= wi;
    }
    return y;
}

// One ipDFT (:579-623) over base[] (- the image described by sw): find3LargestIndx (:705-720) = first strict maximum,
// NaN only wins at bin 0.
__device__ __forceinline__ int ipdft_wide(const PostParams& p, const pmu_c* base, const Sweep* sw, int lane, Tone* out, int* km_out)
{
    double key = -INFINITY;
    int idx = 0x7fffffff;
    int nan0 = 0;
    for (int j = lane; j < p.K; j += PMU_WARP) {
        const pmu_c y = wide_bin(p, base, sw, j);
        const double v = y.x * y.x + y.y * y.y;
        if (j == 0 && v != v) nan0 = 1;
        if (v > key) { key = v; idx = j; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ok = __shfl_xor_sync(0xffffffffu, key, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (ok > key || (ok == key && oi < idx)) { key = ok; idx = oi; }
    }
    nan0 = __shfl_sync(0xffffffffu, nan0, 0);
    const int km = (nan0 || idx >= p.K) ? 0 : idx;
    const pmu_c yk = wide_bin(p, base, sw, km);
    Peak pk;
    pk.km = km;
    pk.best = yk.x * yk.x + yk.y * yk.y;
    pk.yr = yk.x;
    pk.yi = yk.y;
    pk.left = pk.best;
    pk.right = pk.best;
    pk.prev = 0.0;
    if (km > 0) {
        const pmu_c yl = wide_bin(p, base, sw, km - 1);
        pk.left = yl.x * yl.x + yl.y * yl.y;
    }
    if (km + 1 < p.K) {
        const pmu_c yr = wide_bin(p, base, sw, km + 1);
        pk.right = yr.x * yr.x + yr.y * yr.y;
    }
    if (km_out) *km_out = km;
    return ipdft_finish(p, pk, out);
}

// e_ipDFT (:626-664)
__device__ __forceinline__ void e_ipdft_wide(const PostParams& p, const pmu_c* base, int lane, Tone* t, int* km_first)
{
    Tone cur = *t;
    if (!ipdft_wide(p, base, nullptr, lane, &cur, km_first)) {
        for (int it = 0; it < p.P; ++it) {
            Sweep sw;
            sw.init(p, -cur.freq, cur.amp * cur.cr, -cur.amp * cur.ci);
            Tone nxt;
            const int done = ipdft_wide(p, base, &sw, lane, &nxt, nullptr);
            cur = nxt;
            if (done) break;
        }
    }
    *t = cur;
}

// W = X - pureTone(t) (:563-577), optionally the two energies of :219-220
__device__ __forceinline__ void residual_wide(const PostParams& p, const pmu_c* X, const Tone& t, pmu_c* W, int lane, double* Ed, double* E)
{
    Sweep sp, sn;
    sp.init(p, t.freq, t.amp * t.cr, t.amp * t.ci);
    sn.init(p, -t.freq, t.amp * t.cr, -t.amp * t.ci);
    double ed = 0.0, e = 0.0;
    __syncwarp();
    for (int j = lane; j < p.K; j += PMU_WARP) {
        double ar, ai, br, bi;
        sweep_bin(p, sp, j, &ar, &ai);
        sweep_bin(p, sn, j, &br, &bi);
        const pmu_c x = X[j];
        pmu_c w;
        w.x = x.x - (ar + br);
        w.y = x.y - (ai + bi);
        W[j] = w;
        ed += w.x * w.x + w.y * w.y;
        e += x.x * x.x + x.y * x.y;
    }
    __syncwarp();
    if (Ed) {
        *Ed = warp_sum(ed);
        *E = warp_sum(e);
    }
}

// Bins 16*chunk ... 16*chunk + 15 of one window by the pruned FFT: lane l owns residues l + 32 i (pmu_dft.cuh), the
// 32 lane partials of each bin are summed by a halving butterfly that leaves bin j's total on lane j (re) / j + 16 (im).
template <typename InT>
__device__ __forceinline__ void wide_fft_chunk(const KernelArgs& a, const InT* src, long long start, int chunk, int lane, double* rect)
{
    constexpr int M = 16, KC = 16;
    const int N = a.N, L = N / M, k0 = KC * chunk;
    double ar[KC], ai[KC];
    const int n_it = (L + PMU_WARP - 1) / PMU_WARP;
    for (int ii = 0; ii < n_it; ++ii) {
        const int a_loc = lane + PMU_WARP * ii;
        const bool valid = a_loc < L;
        double x[M];
#pragma unroll
        for (int b = 0; b < M; ++b) {
            long long p = start + (valid ? a_loc : 0) + (long long)L * b;
            if (a.ring_cap != 0 && p >= a.ring_cap) p -= a.ring_cap;
            const double v = (double)__ldg(src + p);
            x[b] = valid ? v : 0.0;
        }
        if (ii == 0) {
            accumulate_residue<M, KC, true, double, false>(x, nullptr, ar, ai);
        } else {
            pmu_c u[KC];   // W_N^(32 ii (k0 + k)), lane-uniform
            const unsigned step = (unsigned)(((unsigned long long)PMU_WARP * ii) % (unsigned)N);
            unsigned idx = (unsigned)(((unsigned long long)step * (unsigned)k0) % (unsigned)N);
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                u[k] = __ldg(a.U + idx);
                idx += step;
                if (idx >= (unsigned)N) idx -= (unsigned)N;
            }
            accumulate_residue<M, KC, false, double, false>(x, u, ar, ai);
        }
    }
    {   // rotate by the lane twiddle W_N^(lane (k0 + k))
        unsigned idx = (unsigned)(((unsigned long long)lane * (unsigned)k0) % (unsigned)N);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const pmu_c w = __ldg(a.U + idx);
            const double r = ar[k] * w.x - ai[k] * w.y;
            const double i = ar[k] * w.y + ai[k] * w.x;
            ar[k] = r;
            ai[k] = i;
            idx += (unsigned)lane;
            if (idx >= (unsigned)N) idx -= (unsigned)N;
        }
    }
    // 32 values per lane (v[j] = re of bin j, v[16 + j] = im) -> lane j holds the sum over lanes of v[j]
    double v[2 * KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) { v[k] = ar[k]; v[KC + k] = ai[k]; }
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            // keep the half of the values this lane is responsible for, hand the other half to the partner lane
            const double keep = upper ? v[j + half] : v[j];
            const double give = upper ? v[j] : v[j + half];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, give, half);
        }
    }
    const int k = k0 + (lane & (KC - 1));
    if (k < a.Kp) rect[2 * k + (lane >> 4)] = v[0];
}

// One warp per window, start to finish (no block-wide barrier after the table load): a.n_cons warps of the CTA are
// active, each with its own scratch (rectangular bins, X, W) of a.stage_bytes bytes at a.off_raw.
template <typename InT>
__global__ void __launch_bounds__(PMU_WIDE_THREADS) pmu_wide_kernel(const __grid_constant__ KernelArgs a)
{
    const PostParams& pp = a.post;
    const int K = pp.K, Kp = a.Kp, N = a.N;
    pmu_c* sT = reinterpret_cast<pmu_c*>(pmu_smem + a.off_trig);
    const int tid = threadIdx.x, lane = tid % PMU_WARP, warp = tid / PMU_WARP;
    const int n_act = a.n_cons;   // active warps per CTA

    for (int i = tid; i < pp.jlo + pp.jhi + 1; i += PMU_WIDE_THREADS) sT[i] = a.trig[i];
    __syncthreads();
    if (warp >= n_act) return;
    unsigned char* mine = pmu_smem + a.off_raw + (size_t)warp * a.stage_bytes;
    double* rect = reinterpret_cast<double*>(mine);                                   // [2 Kp]
    pmu_c* X = reinterpret_cast<pmu_c*>(mine + (((size_t)2 * Kp * 8 + 15) / 16) * 16);   // [K]
    pmu_c* W = X + K;                                                                 // [K]

    for (long long w = (long long)blockIdx.x * n_act + warp; w < a.n_windows; w += (long long)gridDim.x * n_act) {
        // where this window's samples live: explicit windows, or a stretch of the stream's ring
        const InT* src;
        long long start = 0;
        if (a.ring_cap == 0) {
            src = reinterpret_cast<const InT*>(a.windows) + (size_t)w * N;
        } else {
            src = reinterpret_cast<const InT*>(a.windows) + (size_t)w * a.ring_pitch;
            start = a.ring_start % a.ring_cap;
        }
        __syncwarp();
        // ---- 1. rectangular bins 0..K ----
        if (a.wide_fft) {   // 16 | N: pruned FFT, one chunk of 16 bins after the other
            for (int c = 0; c * 16 < Kp; ++c) wide_fft_chunk<InT>(a, src, start, c, lane, rect);
        } else {            // direct summation, four partial sums per lane
            for (int k = 0; k < Kp; ++k) {
                const unsigned step = (unsigned)(((unsigned long long)PMU_WARP * (unsigned)k) % (unsigned)N);
                unsigned idx = (unsigned)(((unsigned long long)lane * (unsigned)k) % (unsigned)N);
                double sr[4] = {0.0, 0.0, 0.0, 0.0}, si[4] = {0.0, 0.0, 0.0, 0.0};
                for (int n0 = lane; n0 < N; n0 += 4 * PMU_WARP) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int n = n0 + u * PMU_WARP;
                        if (n < N) {
                            long long p = start + n;
                            if (a.ring_cap != 0 && p >= a.ring_cap) p -= a.ring_cap;
                            const double x = (double)__ldg(src + p);
                            const pmu_c tw = __ldg(a.U + idx);       // W_N^(k n mod N)
                            sr[u] += x * tw.x;
                            si[u] += x * tw.y;
                            idx += step;
                            if (idx >= (unsigned)N) idx -= (unsigned)N;
                        }
                    }
                }
                const double r = warp_sum((sr[0] + sr[1]) + (sr[2] + sr[3]));
                const double i = warp_sum((si[0] + si[1]) + (si[2] + si[3]));
                if (lane == 0) {
                    rect[2 * k] = r;
                    rect[2 * k + 1] = i;
                }
            }
        }
        __syncwarp();
        // ---- 2. Hann window in the frequency domain ----
        for (int j = lane; j < K; j += PMU_WARP) {
            const pmu_c x = hann_combine(rect, j);
            X[j] = x;
            if (a.spec_out) {
                a.spec_out[((size_t)w * K + j) * 2] = a.spec_raw ? rect[2 * j] : x.x;
                a.spec_out[((size_t)w * K + j) * 2 + 1] = a.spec_raw ? rect[2 * j + 1] : x.y;
            }
        }
        __syncwarp();
        // ---- 3. estimator (:205-233, iter_e_ipDFT :666-703), bins strided over the lanes ----
        Tone f;
        f.amp = 0; f.cr = 1; f.ci = 0; f.freq = 0;
        int km = 0;
        e_ipdft_wide(pp, X, lane, &f, &km);
        int dbg = km & PMU_DBG_KM_MASK;
        if (pp.iter_en) {
            double Ed, E;
            residual_wide(pp, X, f, W, lane, &Ed, &E);
            if (Ed > pp.eps * E) {
                dbg |= PMU_DBG_TRIGGERED;
                Tone it;
                it.amp = 0; it.cr = 1; it.ci = 0; it.freq = 0;
                for (int i = 0; i < pp.Q; ++i) {
                    e_ipdft_wide(pp, W, lane, &it, nullptr);
                    residual_wide(pp, X, it, W, lane, nullptr, nullptr);
                    e_ipdft_wide(pp, W, lane, &f, nullptr);
                    if (i + 1 < pp.Q) residual_wide(pp, X, f, W, lane, nullptr, nullptr);
                }
            }
        }
        if (lane == 0) finish_window<InT>(a, f, dbg, w, 0);
    }
}

#endif  // __CUDACC__
