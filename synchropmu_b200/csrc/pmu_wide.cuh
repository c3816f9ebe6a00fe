// pmu_wide.cuh — the general estimator kernel: any window length, any number of bins.
//
// The fused kernel (pmu_kernel.cuh) keeps a lane's partial spectrum in registers and therefore stops at
// n_bins = 23.  The reference accepts every n_cycles + 2 <= n_bins <= N/2 (check_config_validity,
// src/pmu_estimator.c:386-449); those remaining configurations (long analysis windows of >= 22 cycles, or a caller
// that simply asks for many bins) run here: one CTA per window,
//   1. the K+1 lowest rectangular-window bins by direct summation (dft_r's non-power-of-two branch, :532-545, for the
//      bins the estimator reads only), a warp per bin, exact table twiddles W_N^(k n mod N);
//   2. Hann window in the frequency domain (pmu_dft.cuh);
//   3. the estimator on warp 0, bins strided over the lanes: peak search = per-lane scan + warp-shuffle argmax.
// Throughput is O(N * n_bins) per window - this is the coverage path, not the roofline path.
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

template <typename InT>
__global__ void __launch_bounds__(PMU_WIDE_THREADS) pmu_wide_kernel(const __grid_constant__ KernelArgs a)
{
    const PostParams& pp = a.post;
    const int K = pp.K, Kp = a.Kp, N = a.N;
    pmu_c* sT = reinterpret_cast<pmu_c*>(pmu_smem + a.off_trig);
    double* rect = reinterpret_cast<double*>(pmu_smem + a.off_raw);
    pmu_c* X = reinterpret_cast<pmu_c*>(pmu_smem + a.off_post);
    pmu_c* W = X + K;
    const int tid = threadIdx.x, lane = tid % PMU_WARP, warp = tid / PMU_WARP;
    constexpr int NW = PMU_WIDE_THREADS / PMU_WARP;

    for (int i = tid; i < pp.jlo + pp.jhi + 1; i += PMU_WIDE_THREADS) sT[i] = a.trig[i];

    for (long long w = blockIdx.x; w < a.n_windows; w += gridDim.x) {
        // where this window's samples live: explicit windows, or a stretch of the stream's ring
        const InT* src;
        long long start = 0;
        if (a.ring_cap == 0) {
            src = reinterpret_cast<const InT*>(a.windows) + (size_t)w * N;
        } else {
            src = reinterpret_cast<const InT*>(a.windows) + (size_t)w * a.ring_cap;
            start = a.ring_start % a.ring_cap;
        }
        __syncthreads();   // previous window's estimator is done with rect / X / W (and the table is in place)
        // ---- 1. rectangular bins 0..K, one warp per bin, four partial sums per lane ----
        for (int k = warp; k < Kp; k += NW) {
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
        __syncthreads();
        // ---- 2. Hann window in the frequency domain ----
        for (int j = tid; j < K; j += PMU_WIDE_THREADS) {
            const pmu_c x = hann_combine(rect, j);
            X[j] = x;
            if (a.spec_out) {
                a.spec_out[((size_t)w * K + j) * 2] = x.x;
                a.spec_out[((size_t)w * K + j) * 2 + 1] = x.y;
            }
        }
        __syncthreads();
        // ---- 3. estimator (:205-233, iter_e_ipDFT :666-703) on warp 0 ----
        if (warp == 0) {
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
}

#endif  // __CUDACC__
