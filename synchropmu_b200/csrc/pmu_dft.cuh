// pmu_dft.cuh — output-pruned real DFT of one sample window, lane-parallel.
//
// Replaces the reference's pmue_fft_r -> dft_r (src/pmu_estimator.c:487-546: a
// recursive radix-2 FFT of ALL N bins for power-of-two N, a direct n_bins x N DFT
// otherwise) and the time-domain Hann multiply (:184-191).  Only bins 0..n_bins-1 are
// ever read by the estimator (:589-592, :212-221), so with N = L*M (M a small power of
// two) we compute, for the KC lowest bins of the RECTANGULAR-window spectrum,
//
//     X[k] = sum_{a<L} W_N^{a k} * Y_a[k mod M],      Y_a = real DFT_M( x[a + L b], b<M )
//
// Lane l of a warp owns the residues a = l + 32 i.  Per residue it runs an M-point real
// FFT codelet in registers (compile-time twiddles), multiplies by the lane-uniform
// twiddle W_N^{32 i k} (shared-memory broadcast) and accumulates; after the last residue
// the sum is rotated by the lane twiddle W_N^{l k}.  The 32 partial spectra are then
// summed column-wise through shared memory (pmu_kernel.cuh).  The Hann window is applied
// afterwards in the frequency domain, exact for the periodic Hann of :556:
//     Xw[k] = X[k]/2 - (X[k-1] + X[k+1])/4,   X[-1] = conj X[1],
// which is why KC must cover n_bins + 1 bins.  The same code is the mixed-radix path for
// non-power-of-two windows (3072 = 192*16, 600 = 75*8, odd N: M = 1 = direct DFT).
#pragma once

#include "pmu_math.cuh"

// ---- compute type of the DFT stage -----------------------------------------------
// double for the float_p = double build.  The float build may run the DFT stage in FP32 (inputs are
// floats, the reference's float build rounds every intermediate to float anyway); everything after
// the lane reduction stays in double.
struct pmu_cf {
    float x, y;
};

// Two FP32 values that travel together through Blackwell's packed FP32 pipe (FADD2 / FMUL2 / FFMA2, sm_100): the float
// build transforms TWO windows per lane, the same residue of both in one instruction stream, so its DFT stage issues
// half the instructions of the scalar FP32 version.  Twiddles stay scalar (the hardware broadcasts a 32-bit operand to
// both halves, negations fold into operand modifiers).  Each half is rounded exactly like the scalar operation.
struct F2 {
    float lo, hi;
    PMU_HD F2() {}
    PMU_HD F2(float a, float b) : lo(a), hi(b) {}
    PMU_HD explicit F2(float a) : lo(a), hi(a) {}
    PMU_HD explicit F2(double a) : lo((float)a), hi((float)a) {}
    PMU_HD explicit F2(int a) : lo((float)a), hi((float)a) {}
};
#if defined(__CUDA_ARCH__)
#define PMU_F2(v) make_float2((v).lo, (v).hi)
PMU_HD F2 pmu_f2(float2 v) { return F2(v.x, v.y); }
PMU_HD F2 operator+(F2 a, F2 b) { return pmu_f2(__fadd2_rn(PMU_F2(a), PMU_F2(b))); }
PMU_HD F2 operator-(F2 a, F2 b) { return pmu_f2(__fadd2_rn(PMU_F2(a), make_float2(-b.lo, -b.hi))); }
PMU_HD F2 operator*(F2 a, F2 b) { return pmu_f2(__fmul2_rn(PMU_F2(a), PMU_F2(b))); }
PMU_HD F2 operator*(float w, F2 b) { return pmu_f2(__fmul2_rn(make_float2(w, w), PMU_F2(b))); }
PMU_HD F2 pmu_fma(float w, F2 b, F2 c) { return pmu_f2(__ffma2_rn(make_float2(w, w), PMU_F2(b), PMU_F2(c))); }
PMU_HD F2 pmu_fma(F2 a, F2 b, F2 c) { return pmu_f2(__ffma2_rn(PMU_F2(a), PMU_F2(b), PMU_F2(c))); }
#else
PMU_HD F2 operator+(F2 a, F2 b) { return F2(a.lo + b.lo, a.hi + b.hi); }
PMU_HD F2 operator-(F2 a, F2 b) { return F2(a.lo - b.lo, a.hi - b.hi); }
PMU_HD F2 operator*(F2 a, F2 b) { return F2(a.lo * b.lo, a.hi * b.hi); }
PMU_HD F2 operator*(float w, F2 b) { return F2(w * b.lo, w * b.hi); }
PMU_HD F2 pmu_fma(float w, F2 b, F2 c) { return F2(fmaf(w, b.lo, c.lo), fmaf(w, b.hi, c.hi)); }
PMU_HD F2 pmu_fma(F2 a, F2 b, F2 c) { return F2(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }
#endif
PMU_HD F2 operator-(F2 a) { return F2(-a.lo, -a.hi); }
PMU_HD F2& operator+=(F2& a, F2 b) { a = a + b; return a; }

template <typename R> struct CplxOf;
template <> struct CplxOf<double> { typedef pmu_c type; };
template <> struct CplxOf<float> { typedef pmu_cf type; };
template <> struct CplxOf<F2> { typedef pmu_cf type; };          // scalar twiddles, broadcast by the hardware
template <typename R> struct ScalarOf { typedef R type; };      // type of a twiddle component
template <> struct ScalarOf<F2> { typedef float type; };
template <typename R> struct IsPacked { static constexpr bool value = false; };
template <> struct IsPacked<F2> { static constexpr bool value = true; };
template <typename R> struct alignas(2 * sizeof(R)) RPair { R x, y; };   // one complex partial in the reduction scratch
PMU_HD double ct_to_double(double v) { return v; }
PMU_HD double ct_to_double(float v) { return (double)v; }
PMU_HD double hsum(double v) { return v; }
PMU_HD float hsum(float v) { return v; }
PMU_HD float hsum(F2 v) { return v.lo + v.hi; }
template <typename R> struct IsDouble { static constexpr bool value = false; };
template <> struct IsDouble<double> { static constexpr bool value = true; };

// w_re * y_re - w_im * y_im, w_re * y_im + w_im * y_re and their accumulating forms.  The scalar versions are the plain
// expressions (the compiler contracts them as it always did); the packed versions spell out the FFMA2s.
template <typename S, typename R> PMU_HD R cmul_re(S wr, S wi, R yr, R yi) { return wr * yr - wi * yi; }
template <typename S, typename R> PMU_HD R cmul_im(S wr, S wi, R yr, R yi) { return wr * yi + wi * yr; }
template <typename S, typename R> PMU_HD void cmac(R& ar, R& ai, S wr, S wi, R yr, R yi)
{
    ar += wr * yr - wi * yi;
    ai += wr * yi + wi * yr;
}
template <typename S, typename R> PMU_HD void rmac(R& ar, R& ai, S wr, S wi, R yr)   // y real
{
    ar += wr * yr;
    ai += wi * yr;
}
PMU_HD F2 cmul_re(float wr, float wi, F2 yr, F2 yi) { return pmu_fma(wr, yr, -(wi * yi)); }
PMU_HD F2 cmul_im(float wr, float wi, F2 yr, F2 yi) { return pmu_fma(wr, yi, wi * yr); }
PMU_HD void cmac(F2& ar, F2& ai, float wr, float wi, F2 yr, F2 yi)
{
    ar = pmu_fma(-wi, yi, pmu_fma(wr, yr, ar));
    ai = pmu_fma(wi, yr, pmu_fma(wr, yi, ai));
}
PMU_HD void rmac(F2& ar, F2& ai, float wr, float wi, F2 yr)
{
    ar = pmu_fma(wr, yr, ar);
    ai = pmu_fma(wi, yr, ai);
}

// ---- compile-time loop ---------------------------------------------------------
template <int I>
struct IC {
    static constexpr int value = I;
};
template <int B, int E, class F>
PMU_HD void static_for(F&& f)
{
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}

// cos(2 pi k/16), k = 0..15
PMU_HD constexpr double cos16(int k)
{
    constexpr double c1 = 0.92387953251128675613, c2 = 0.70710678118654752440, c3 = 0.38268343236508977173;
    switch (k & 15) {
        case 0: return 1.0;
        case 1: case 15: return c1;
        case 2: case 14: return c2;
        case 3: case 13: return c3;
        case 4: case 12: return 0.0;
        case 5: case 11: return -c3;
        case 6: case 10: return -c2;
        case 7: case 9: return -c1;
        default: return -1.0;
    }
}
PMU_HD constexpr double sin16(int k) { return cos16(k - 4); }

// On the device the three non-trivial twiddle values are read from the constant bank: an FP64
// instruction takes a c[bank][offset] operand for free, whereas a 64-bit literal has to be
// materialised with two (uniform) moves every time the compiler rematerialises it.
#if defined(__CUDACC__)
static __constant__ double pmu_tw16[3] = {0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173};
#endif
template <int K16, typename R = double>
PMU_HD R cosv16()
{
    constexpr int k = K16 & 15;
#if defined(__CUDA_ARCH__)
    if constexpr (IsDouble<R>::value) {
        if constexpr (k == 1 || k == 15) return pmu_tw16[0];
        else if constexpr (k == 2 || k == 14) return pmu_tw16[1];
        else if constexpr (k == 3 || k == 13) return pmu_tw16[2];
        else if constexpr (k == 5 || k == 11) return -pmu_tw16[2];
        else if constexpr (k == 6 || k == 10) return -pmu_tw16[1];
        else if constexpr (k == 7 || k == 9) return -pmu_tw16[0];
        else return cos16(k);
    } else {
        return (R)cos16(k);   // 32-bit literals are instruction immediates
    }
#else
    return (R)cos16(k);
#endif
}
template <int K16, typename R = double>
PMU_HD R sinv16() { return cosv16<K16 - 4, R>(); }

// (tr, ti) = W_M^K * (orr, oi),  W_M = exp(-2 pi i / M), M | 16; trivial twiddles cost nothing.
template <int K, int M, typename R>
PMU_HD void tw_mul(R orr, R oi, R* tr, R* ti)
{
    constexpr int k16 = (K * (16 / M)) & 15;
    if constexpr (k16 == 0) { *tr = orr; *ti = oi; }
    else if constexpr (k16 == 4) { *tr = oi; *ti = -orr; }        // -i
    else if constexpr (k16 == 8) { *tr = -orr; *ti = -oi; }       // -1
    else if constexpr (k16 == 12) { *tr = -oi; *ti = orr; }       // +i
    else {
        typedef typename ScalarOf<R>::type S;
        const S wr = cosv16<k16, S>(), wi = -sinv16<k16, S>();
        *tr = cmul_re(wr, wi, orr, oi);
        *ti = cmul_im(wr, wi, orr, oi);
    }
}

// Real-input FFT codelet: x[0..M) -> bins 0..M/2 (re[], im[]).  Radix-2 decimation in
// time on real data, conjugate symmetry of the half-size transforms used throughout.
template <int M, typename R = double>
struct RFFT {
    static_assert(M == 1 || M == 2 || M == 4 || M == 8 || M == 16, "codelet sizes");
    PMU_HD static void run(const R* x, R* re, R* im)
    {
        if constexpr (M == 1) {
            re[0] = x[0]; im[0] = (R)0;
        } else if constexpr (M == 2) {
            re[0] = x[0] + x[1]; im[0] = (R)0;
            re[1] = x[0] - x[1]; im[1] = (R)0;
        } else {
            constexpr int H = M / 2;
            R ev[H], od[H];
            static_for<0, H>([&](auto i) {
                ev[i.value] = x[2 * i.value];
                od[i.value] = x[2 * i.value + 1];
            });
            R er[H / 2 + 1], ei[H / 2 + 1], qr[H / 2 + 1], qi[H / 2 + 1];
            RFFT<H, R>::run(ev, er, ei);
            RFFT<H, R>::run(od, qr, qi);
            // X[k] = E[k] + W^k O[k] and X[H-k] = conj(E[k] - W^k O[k]) share the twiddle product; E[0], E[H/2],
            // O[0], O[H/2] are real by symmetry
            re[0] = er[0] + qr[0]; im[0] = (R)0;
            re[H] = er[0] - qr[0]; im[H] = (R)0;
            re[H / 2] = er[H / 2]; im[H / 2] = -qr[H / 2];          // W^(H/2) = -i
            static_for<1, H / 2>([&](auto kc) {
                constexpr int k = kc.value;
                R tr, ti;
                tw_mul<k, M, R>(qr[k], qi[k], &tr, &ti);
                re[k] = er[k] + tr;     im[k] = ei[k] + ti;
                re[H - k] = er[k] - tr; im[H - k] = ti - ei[k];
            });
        }
    }
};

// Bin k (compile-time) of the M-periodic, conjugate-symmetric codelet output.
template <int K, int M, typename R>
PMU_HD void ysel(const R* re, const R* im, R* yr, R* yi, bool* is_real)
{
    constexpr int r = K % M;
    constexpr int idx = (r <= M / 2) ? r : M - r;
    constexpr bool cj = (r > M / 2);
    *yr = re[idx];
    *yi = cj ? -im[idx] : im[idx];
    *is_real = (idx == 0) || (2 * idx == M);
}

// Accumulate one residue: acc[k] (+)= U[k] * Y[k mod M], k < KC.
//   FIRST : U == 1 for every k (i == 0), acc is initialised instead of accumulated.
//   u     : lane-uniform twiddles W_N^{32 i k} (broadcast loads), ignored when FIRST.
//   BIN0  : the first of the KC bins is bin 0 of the spectrum (twiddle 1, real codelet output: one add).  The general
//           kernel accumulates later chunks of 16 bins (bins 16 c ... 16 c + 15, same k mod M) with BIN0 = false.
template <int M, int KC, bool FIRST, typename R, bool BIN0 = true>
PMU_HD void accumulate_residue(const R* x, const typename CplxOf<R>::type* u, R* ar, R* ai)
{
    R re[M / 2 + 1], im[M / 2 + 1];
    RFFT<M, R>::run(x, re, im);
    static_for<0, KC>([&](auto kc) {
        constexpr int k = kc.value;
        R yr, yi;
        bool rl;
        ysel<k, M, R>(re, im, &yr, &yi, &rl);
        constexpr int r = k % M;
        constexpr bool is_real = (r == 0) || (2 * r == M) || (M == 1);
        if constexpr (FIRST) {
            ar[k] = yr;
            ai[k] = is_real ? (R)0 : yi;
        } else if constexpr (k == 0 && BIN0) {
            ar[0] += yr;  // W^0 = 1, Y[0] real
        } else {
            const typename CplxOf<R>::type w = u[k];
            if constexpr (is_real) rmac(ar[k], ai[k], w.x, w.y, yr);
            else cmac(ar[k], ai[k], w.x, w.y, yr, yi);
        }
    });
}

// Rotate the lane's partial spectrum by its own twiddle V[k] = W_N^{lane k}.
template <int KC, typename R>
PMU_HD void rotate_lane(const typename CplxOf<R>::type* v, int v_stride, R* ar, R* ai)
{
    static_for<1, KC>([&](auto kc) {
        constexpr int k = kc.value;
        const typename CplxOf<R>::type w = v[k * v_stride];
        R r = cmul_re(w.x, w.y, ar[k], ai[k]);
        R i = cmul_im(w.x, w.y, ar[k], ai[k]);
        ar[k] = r;
        ai[k] = i;
    });
}

// Hann window in the frequency domain: windowed bin k from rectangular bins k-1, k, k+1
// (rect[-1] = conj rect[1]).  raw = interleaved re,im of the rectangular bins 0..K.
PMU_HD pmu_c hann_combine(const double* raw, int k)
{
    double lr, li;
    if (k == 0) { lr = raw[2]; li = -raw[3]; }
    else { lr = raw[2 * (k - 1)]; li = raw[2 * (k - 1) + 1]; }
    pmu_c o;
    o.x = 0.5 * raw[2 * k] - 0.25 * (lr + raw[2 * (k + 1)]);
    o.y = 0.5 * raw[2 * k + 1] - 0.25 * (li + raw[2 * (k + 1) + 1]);
    return o;
}

// The same combination over a whole lane-private row, in place: X holds the K+1 rectangular bins on entry and the K
// Hann-windowed bins on return (entry K keeps its rectangular value).  Same operations, in the same order, as K calls
// of hann_combine.
PMU_HD void hann_combine_inplace(const Col& X, int K)
{
    pmu_c cur = X.get(0);
    pmu_c left = X.get(1);
    left.y = -left.y;   // rect[-1] = conj rect[1]
    for (int k = 0; k < K; ++k) {
        const pmu_c nxt = X.get(k + 1);
        X.put(k, 0.5 * cur.x - 0.25 * (left.x + nxt.x), 0.5 * cur.y - 0.25 * (left.y + nxt.y));
        left = cur;
        cur = nxt;
    }
}

// ... and for packed kernels, whose two halves are different residues: pairs of twiddles (W_N^{2l k}, W_N^{(2l+1) k}).
template <int KC, typename R>
PMU_HD void rotate_lane_pairs(const RPair<R>* v, int v_stride, R* ar, R* ai)
{
    static_for<0, KC>([&](auto kc) {
        constexpr int k = kc.value;
        const RPair<R> w = v[k * v_stride];
        R r = pmu_fma(ar[k], w.x, -(ai[k] * w.y));
        R i = pmu_fma(ar[k], w.y, ai[k] * w.x);
        ar[k] = r;
        ai[k] = i;
    });
}
