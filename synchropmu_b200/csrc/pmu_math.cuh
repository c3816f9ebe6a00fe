// pmu_math.cuh — per-window estimator arithmetic (one window per thread/lane).
//
// Everything after the DFT in SynchroPMU's pmu_estimate() (reference
// src/pmu_estimator.c:205-265) restated for a GPU lane: IpDFT, e-IpDFT negative-image
// compensation, pure-tone synthesis, the interference energy test, the iterative
// interharmonic estimation, the two-state ROCOF filter and the frame fill.
//
// The formulation is NOT the reference's: the Hann-window tone spectrum wf()
// (macros at src/pmu_estimator.c:59-61, 1 cexp + 6 sin per bin) is evaluated as a
// sweep over bins with two sincospi() per sweep:
//
//   x_k = k - f/df = m + r,  kstar = rint(f/df),  r = kstar - f/df (|r| <= 1/2),  m = k - kstar
//   sin(pi (x_k + e))   = (-1)^(m+e) sin(pi r)                        e in {-1,0,1}
//   sin(pi (x_k + e)/N) = sin(pi r/N) cos((m+e) pi/N) + cos(pi r/N) sin((m+e) pi/N)
//   exp(-i pi C3 x_k) (-1)^m = exp(-i pi C3 r) exp(i pi m/N)            (C3 = 1 - 1/N)
//   C5 = -exp(+i pi C3)/4, C6 = conj(C5), exp(i pi C3) = -cos(pi/N) + i sin(pi/N)
//
// so that wf(k, f, z) = zz * E[m] * ( c3q (I[m-1] + I[m+1]) + I[m]/2 + i s3q (I[m-1] - I[m+1]) )
// with zz = z * inv_norm * sin(pi r) * exp(-i pi C3 r), E[j] = exp(i pi j/N) from a table
// and I[j] = 1/sin(pi (r+j)/N) (one new reciprocal per bin).  Anchoring the angle
// addition at the nearest bin keeps every Dirichlet ratio accurate to a few ulp, also
// when the tone sits (almost) on a bin; r == 0 gives 0 * inf = NaN exactly where the
// reference gets sin(0)/sin(0) = NaN.
//
// All functions are __host__ __device__ so tests/host_emul can run this exact code on
// the CPU against the oracle; the shipped library only ever runs it on the GPU.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define PMU_HD __host__ __device__ __forceinline__
#include <cuda_runtime.h>
typedef double2 pmu_c;
#else
#define PMU_HD inline
struct alignas(16) pmu_c {
    double x, y;
};
#endif

#ifndef PMU_PI
#define PMU_PI 3.14159265358979323846 /* M_PI, reference src/func_stubs.h:33 */
#endif

#if defined(__CUDACC__)
// the general-range routine, out of line: it is only reached when an interpolated peak is more than half a bin away
// from its bin, and inlined into the estimator's loops its 75 instructions are pure instruction-cache ballast
static __device__ __noinline__ double2 pmu_sincospi_general(double x)
{
    double s, c;
    sincospi(x, &s, &c);
    return make_double2(s, c);
}
#endif
PMU_HD void pmu_sincospi(double x, double* s, double* c)
{
#if defined(__CUDA_ARCH__)
    const double2 sc = pmu_sincospi_general(x);
    *s = sc.x;
    *c = sc.y;
#else
    // host emulation only: |x| <= ~0.5 here, argument rounding costs < 1 ulp of pi*x
    *s = sin(PMU_PI * x);
    *c = cos(PMU_PI * x);
#endif
}

#ifndef PMU_FAST_TRIG
#define PMU_FAST_TRIG 1
#endif
// unroll factors of the estimator's bin loops (independent reciprocal / sweep chains per bin: the loops are latency-,
// not issue-bound, so several bins in flight per lane pay)
// (measured again on the one-copy estimator, profiles/r2_ab_variants.jsonl: NOT unrolling is best - the smaller the
//  estimator's code, the fewer instruction-cache misses with seven warps at different places in it: 1 / 1 against 2 / 2
//  gives +3 % triggered, +1.5 % config 5, +5 % float build)
#ifndef PMU_UNROLL_COMP
#define PMU_UNROLL_COMP 1
#endif
#ifndef PMU_UNROLL_RES
#define PMU_UNROLL_RES 1
#endif
#ifndef PMU_FAST_SQRT
#define PMU_FAST_SQRT 1
#endif
#define PMU_PRAGMA(x) _Pragma(#x)
#define PMU_UNROLL(n) PMU_PRAGMA(unroll n)

// sin(pi x), cos(pi x) for |x| <= 1/2 (the Dirichlet anchor r of a sweep): Taylor polynomials of the half angle
// (|pi x / 2| <= pi/4: truncation < 5e-17) and one angle doubling, 21 FMAs in two independent chains instead of
// the general-range library routine.  x == 0 gives sin = 0 exactly, NaN propagates.
PMU_HD void pmu_sincospi_half(double x, double* s, double* c)
{
#if PMU_FAST_TRIG
    const double t = (0.5 * PMU_PI) * x;
    const double u = t * t;
    // Estrin's scheme: the two degree-7 polynomials in u are evaluated as trees of depth 4 instead of Horner chains of
    // depth 8 (same coefficients; two more multiplies each) - the estimator is bound by dependent-issue latency
    const double u2 = u * u, u4 = u2 * u2;
    const double s01 = fma(8.3333333333333333333e-03, u, -1.6666666666666666667e-01);   //  1/5! u - 1/3!
    const double s23 = fma(2.7557319223985890653e-06, u, -1.9841269841269841270e-04);   //  1/9! u - 1/7!
    const double s45 = fma(1.6059043836821614599e-10, u, -2.5052108385441718775e-08);   // 1/13! u - 1/11!
    const double s6 = -7.6471637318198164759e-13;                                       // -1/15!
    const double sp = fma(fma(s6, u2, s45), u4, fma(s23, u2, s01));
    const double sh = fma(sp * u, t, t);
    const double c01 = fma(4.1666666666666666667e-02, u, -0.5);                          //  1/4! u - 1/2!
    const double c23 = fma(2.4801587301587301587e-05, u, -1.3888888888888888889e-03);   //  1/8! u - 1/6!
    const double c45 = fma(2.0876756987868098979e-09, u, -2.7557319223985890653e-07);   // 1/12! u - 1/10!
    const double c67 = fma(4.7794773323873852974e-14, u, -1.1470745597729724714e-11);   // 1/16! u - 1/14!
    const double cp = fma(fma(c67, u2, c45), u4, fma(c23, u2, c01));
    const double ch = fma(cp, u, 1.0);
    *s = 2.0 * sh * ch;
    *c = (ch - sh) * (ch + sh);
#else
    pmu_sincospi(x, s, c);
#endif
}

// sin(pi x), cos(pi x) for |pi x| <= 0.1 (the r/N arguments of windows of >= 16 samples): truncation < 3e-17.
PMU_HD void pmu_sincospi_small(double x, double* s, double* c)
{
#if PMU_FAST_TRIG
    const double t = PMU_PI * x;
    const double u = t * t;
    double sp = 2.7557319223985890653e-06;
    sp = fma(sp, u, -1.9841269841269841270e-04);
    sp = fma(sp, u, 8.3333333333333333333e-03);
    sp = fma(sp, u, -1.6666666666666666667e-01);
    *s = fma(sp * u, t, t);
    double cp = -2.7557319223985890653e-07;
    cp = fma(cp, u, 2.4801587301587301587e-05);
    cp = fma(cp, u, -1.3888888888888888889e-03);
    cp = fma(cp, u, 4.1666666666666666667e-02);
    cp = fma(cp, u, -0.5);
    *c = fma(cp, u, 1.0);
#else
    pmu_sincospi(x, s, c);
#endif
}

// 1/x: hardware seed (~20 bits, full exponent range) + two Newton steps; 1/0 and 1/NaN give NaN here, which is what
// the callers turn an infinite Dirichlet term into anyway (0 * inf).
PMU_HD double pmu_rcp(double x)
{
#if PMU_FAST_TRIG && defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}

// sqrt(x) for x = |Y|^2 >= 0: hardware reciprocal-square-root seed + two coupled (Goldschmidt) iterations and a final
// residual correction, within an ulp of the library routine without its slow-path branch.  0 -> 0; denormal or
// infinite arguments (outside the documented |Y| range) give NaN.
PMU_HD double pmu_sqrt(double x)
{
#if PMU_FAST_SQRT && defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    g = fma(fma(-g, g, x), h, g);
    return x == 0.0 ? 0.0 : g;
#else
    return sqrt(x);
#endif
}

// Read-only parameters of the post-DFT stage (uniform over a batch).
struct PostParams {
    int K;         // n_bins
    int P, Q;      // e-IpDFT passes, interference iterations
    int iter_en;   // iter_eipdft_enabled
    int jlo, jhi;  // trig table covers j in [-jlo, jhi]
    int C;         // channels per instance (mid_fracsec is per instance)
    int f0;
    double df, inv_df;
    double S, invS;       // norm_factor and 1/norm_factor (src/pmu_estimator.c:150,160)
    double eps;           // interf_trig
    double invN;
    double c3q, s3q;      // cos(pi C3)/4, sin(pi C3)/4
    double frame_rate;
    double thr0, thr1, thr2;  // ROCOF thresholds
    double lp0, lp1, lp2;     // ROCOF low-pass coefficients
    const pmu_c* trig;    // trig[j + jlo] = (cos(j pi/N), sin(j pi/N)) (host emulation; global copy on device)
    unsigned trig_off;    // device: byte offset of the table's shared-memory copy from the dynamic smem base
};

#if defined(__CUDACC__)
// The kernel's dynamic shared memory.  Tables and per-lane columns are addressed as 32-bit offsets
// from this base, so the compiler emits LDS/STS with 32-bit address arithmetic.
extern __shared__ __align__(128) unsigned char pmu_smem[];
#endif

// Table read: on the device the table sits in shared memory (the assumption lets the compiler
// emit LDS instead of a generic LD).
PMU_HD pmu_c pmu_trig(const PostParams& p, int idx)
{
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const pmu_c*>(pmu_smem + p.trig_off + (unsigned)idx * 16u);
#else
    return p.trig[idx];
#endif
}

// A tone estimate.  The phase is carried as the unit phasor (cr, ci) = exp(i ph): the
// compensation / synthesis steps only ever need exp(+-i ph) (src/pmu_estimator.c:567-568, :646),
// so no atan2 / sincos round trip is needed between passes; the angle itself is taken once,
// when the frame is written (:265).
struct Tone {
    double amp, cr, ci, freq;
};

// One sweep of wf(k, f, z) over the bins for a fixed tone (f, z).
struct Sweep {
    double zr, zi;  // zz
    double sr, cr;  // sin(pi r/N), cos(pi r/N)
    int kstar;

    PMU_HD void init(const PostParams& p, double f, double z_re, double z_im)
    {
#if PMU_FAST_TRIG
        double q = f * p.inv_df;  // x = k - f/df (src/pmu_estimator.c:61), to within an ulp of q
#else
        double q = f / p.df;
#endif
        double ks = rint(q);
        const double lim = (double)(p.K + 1);
        if (!(ks >= -lim)) ks = -lim;  // also catches NaN: keeps table indices in range
        if (!(ks <= lim)) ks = lim;
        kstar = (int)ks;
        double r = ks - q;
        double spr, cpr;
        pmu_sincospi_half(r, &spr, &cpr);
        if (p.invN <= 1.0 / 16.0) pmu_sincospi_small(r * p.invN, &sr, &cr);
        else pmu_sincospi_half(r * p.invN, &sr, &cr);
        // exp(-i pi C3 r) = exp(-i pi r) exp(+i pi r/N)
        double er = cpr * cr + spr * sr;
        double ei = cpr * sr - spr * cr;
        double g = p.invS * spr;
        zr = (z_re * er - z_im * ei) * g;
        zi = (z_re * ei + z_im * er) * g;
    }

    // 1 / sin(pi (r + j)/N), t = trig[j] = (cos(j pi/N), sin(j pi/N))
    PMU_HD double inv_den_t(const pmu_c& t) const { return pmu_rcp(sr * t.x + cr * t.y); }
    PMU_HD double inv_den(const PostParams& p, int j) const { return inv_den_t(pmu_trig(p, j + p.jlo)); }

    // value at bin k = kstar + m given e = trig[m] and I[m-1], I[m], I[m+1]
    PMU_HD void eval_e(const PostParams& p, const pmu_c& e, double im1, double i0, double ip1, double* wr, double* wi) const
    {
        double br = p.c3q * (im1 + ip1) + 0.5 * i0;
        double bi = p.s3q * (im1 - ip1);
        double tr = e.x * br - e.y * bi;
        double ti = e.x * bi + e.y * br;
        *wr = zr * tr - zi * ti;
        *wi = zr * ti + zi * tr;
    }
    PMU_HD void eval(const PostParams& p, int k, double im1, double i0, double ip1, double* wr, double* wi) const
    {
        eval_e(p, pmu_trig(p, (k - kstar) + p.jlo), im1, i0, ip1, wr, wi);
    }
};

// A sweep walked bin by bin from bin 0: keeps the sliding window of the three reciprocals and the table entry of the
// current bin (the entry loaded for I[m+1] is the e of the next bin), so every bin costs one table read and one
// reciprocal.
struct SweepWalk {
    double im1, i0;
    pmu_c e;      // trig[m] of the bin about to be evaluated
    int idx;      // table index of trig[m + 1]

    PMU_HD void start(const PostParams& p, const Sweep& sw)
    {
        idx = -sw.kstar + p.jlo;          // trig[m] for bin 0
        im1 = sw.inv_den_t(pmu_trig(p, idx - 1));
        e = pmu_trig(p, idx);
        i0 = sw.inv_den_t(e);
        ++idx;
    }
    PMU_HD void step(const PostParams& p, const Sweep& sw, double* wr, double* wi)
    {
        const pmu_c tn = pmu_trig(p, idx);
        const double ip1 = sw.inv_den_t(tn);
        sw.eval_e(p, e, im1, i0, ip1, wr, wi);
        im1 = i0;
        i0 = ip1;
        e = tn;
        ++idx;
    }
};

// Running "first strict maximum + neighbours" over |Y_j|^2
// (find3LargestIndx, src/pmu_estimator.c:705-720; NaN never wins unless at index 0).
struct Peak {
    double best, left, right, prev, yr, yi;
    int km;

    PMU_HD void first(double v, double re, double im)
    {
        best = v; left = v; right = v; prev = v; yr = re; yi = im; km = 0;
    }
    PMU_HD void next(int j, double v, double re, double im)
    {
        bool up = v > best;
        if (!up && j == km + 1) right = v;
        if (up) { left = prev; km = j; best = v; yr = re; yi = im; }
        prev = v;
    }
};

// ipDFT tail (src/pmu_estimator.c:594-623) from the tracked peak.  Returns 1 when
// |delta| <= 1e-11 (the reference's literal 10e-12, :603).
//   reference: amp = |Y| |(d^2-1) pi d / sin(pi d)|,  ph = arg(Y) - pi d
//   here     : exp(i ph) = (Y/|Y|) exp(-i pi d), sin(pi d) from the same sincospi.
PMU_HD int ipdft_tail(const PostParams& p, int km, double mk, double ml, double mr, double yr, double yi, Tone* t)
{
    double d = 2 * (mr - ml) * pmu_rcp(ml + mr + 2 * mk);       // :598
    t->freq = (km + d) * p.df;                                  // :601
    double inv = pmu_rcp(mk);
    double ur = yr * inv, ui = yi * inv;                        // exp(i arg Y)
    if (fabs(d) <= 10e-12) {                                    // :603-613
        t->amp = mk;
        t->cr = ur;
        t->ci = ui;
        return 1;
    }
    double sd, cd;
    if (fabs(d) <= 0.5) pmu_sincospi_half(d, &sd, &cd);         // the usual case: the peak bin is the nearest one
    else pmu_sincospi(d, &sd, &cd);
    t->amp = mk * fabs((d * d - 1) * (PMU_PI * d) * pmu_rcp(sd));  // :616
    t->cr = ur * cd + ui * sd;                                  // :617  exp(i(arg Y - pi d))
    t->ci = ui * cd - ur * sd;
    return 0;
}

PMU_HD int ipdft_finish(const PostParams& p, const Peak& pk, Tone* t)
{
    double mk = pmu_sqrt(pk.best);
    double ml = (pk.km == 0) ? mk : pmu_sqrt(pk.left);          // kl = km when km == 0 (:718)
    double mr = (pk.km == p.K - 1) ? ml : pmu_sqrt(pk.right);   // kr = km-1 at the last bin (:719)
    return ipdft_tail(p, pk.km, mk, ml, mr, pk.yr, pk.yi, t);
}

// Column view of a [K][32] complex array (lane = window); shared memory on the device.
struct Col {
#if defined(__CUDA_ARCH__)
    unsigned off;     // byte offset (from the dynamic smem base) of bin 0 of this lane's column
    unsigned stride;  // bytes between consecutive bins
    __device__ __forceinline__ static Col make(pmu_c* ptr, int stride_elems)
    {
        Col c;
        c.off = (unsigned)(reinterpret_cast<unsigned char*>(ptr) - pmu_smem);
        c.stride = (unsigned)stride_elems * 16u;
        return c;
    }
    __device__ __forceinline__ pmu_c get(int j) const
    {
        return *reinterpret_cast<const pmu_c*>(pmu_smem + off + (unsigned)j * stride);
    }
    __device__ __forceinline__ void put(int j, double re, double im) const
    {
        pmu_c v; v.x = re; v.y = im;
        *reinterpret_cast<pmu_c*>(pmu_smem + off + (unsigned)j * stride) = v;
    }
#else
    pmu_c* base;
    int stride;
    static Col make(pmu_c* ptr, int stride_elems)
    {
        Col c;
        c.base = ptr;
        c.stride = stride_elems;
        return c;
    }
    pmu_c get(int j) const { return base[j * stride]; }
    void put(int j, double re, double im) const
    {
        pmu_c v; v.x = re; v.y = im;
        base[j * stride] = v;
    }
#endif
};

// One ipDFT over Y[j] (comp == false) or over Y[j] - wf(j, -f, amp e^{-i ph}) (comp == true,
// the e_ipDFT inner loop :646-654).  km_out receives the chosen peak bin.
//
// The peak search (find3LargestIndx, :705-720: first strict maximum, a NaN only wins at bin 0) is a running scan
// written so that every bin costs one compare and a handful of predicated moves: `up` is forced at bin 0 (so a NaN
// there sticks, as every later compare against it is false), the right neighbour is whatever follows the bin that
// took the lead last, the left neighbour is the bin before it.
//
// `comp` is a run-time flag on purpose: the plain and the compensated pass share ONE copy of the loop and of the
// interpolation tail.  The estimator's working set of instructions is what bounds it when the interference
// iterations run (ncu: `no_instruction` was the top stall with every call site inlined separately), so every routine
// below exists exactly once in the kernel.
PMU_HD int ipdft_pass(const PostParams& p, const Col& Y, bool comp, const Tone& prev, Tone* out, int* km_out)
{
    Sweep sw;
    SweepWalk wk;
    if (comp) {
        sw.init(p, -prev.freq, prev.amp * prev.cr, -prev.amp * prev.ci);  // amp e^{-i ph}, :646
        wk.start(p, sw);
    }
    Peak pk;
    pk.best = 0; pk.left = 0; pk.right = 0; pk.prev = 0; pk.yr = 0; pk.yi = 0; pk.km = 0;
    bool was_up = false;
    PMU_UNROLL(PMU_UNROLL_COMP)
    for (int j = 0; j < p.K; ++j) {
        pmu_c y = Y.get(j);
        if (comp) {
            double wr, wi;
            wk.step(p, sw, &wr, &wi);
            y.x -= wr;
            y.y -= wi;
        }
        const double v = y.x * y.x + y.y * y.y;
        pk.right = was_up ? v : pk.right;
        const bool up = (v > pk.best) || (j == 0);
        pk.left = up ? pk.prev : pk.left;
        pk.best = up ? v : pk.best;
        pk.km = up ? j : pk.km;
        pk.yr = up ? y.x : pk.yr;
        pk.yi = up ? y.y : pk.yi;
        pk.prev = v;
        was_up = up;
    }
    if (km_out) *km_out = pk.km;
    return ipdft_finish(p, pk, out);
}

// e_ipDFT (src/pmu_estimator.c:626-664): the plain ipDFT, then up to P compensated ones; stops as soon as one of them
// reports |delta| <= 1e-11.  Returns the peak bin of the first pass.
PMU_HD int e_ipdft(const PostParams& p, const Col& Y, Tone* t)
{
    Tone cur;
    cur.amp = 0; cur.cr = 1; cur.ci = 0; cur.freq = 0;
    int km_first = 0;
    for (int it = 0; it <= p.P; ++it) {
        Tone nxt;
        int km;
        const int done = ipdft_pass(p, Y, it > 0, cur, &nxt, &km);
        if (it == 0) km_first = km;
        cur = nxt;
        if (done) break;
    }
    *t = cur;
    return km_first;
}

// W[j] = X[j] - pureTone(t)[j]   (pureTone :563-577; Xi/Xf updates :215, :681-684, :692-695)
// and Ed = sum |W[j]|^2, E = sum |X[j]|^2  (:219-220 compute cabs(z*z) = |z|^2; only the first call's sums are used).
PMU_HD void residual(const PostParams& p, const Col& X, const Tone& t, const Col& W, double* Ed, double* E)
{
    Sweep sp, sn;
    sp.init(p, t.freq, t.amp * t.cr, t.amp * t.ci);    // phsr_0 = amp e^{+i ph}
    sn.init(p, -t.freq, t.amp * t.cr, -t.amp * t.ci);  // phsr_1 = amp e^{-i ph}
    SweepWalk wp, wn;
    wp.start(p, sp);
    wn.start(p, sn);
    double ed = 0, e = 0;
    PMU_UNROLL(PMU_UNROLL_RES)
    for (int j = 0; j < p.K; ++j) {
        double ar, ai, br, bi;
        wp.step(p, sp, &ar, &ai);
        wn.step(p, sn, &br, &bi);
        pmu_c x = X.get(j);
        double wr = x.x - (ar + br);
        double wi = x.y - (ai + bi);
        W.put(j, wr, wi);
        ed += wr * wr + wi * wi;
        e += x.x * x.x + x.y * x.y;
    }
    *Ed = ed;
    *E = e;
}

// Bits of the per-window debug word (exported only through the debug entry point).
#define PMU_DBG_KM_MASK 0xffff
#define PMU_DBG_TRIGGERED 0x10000

// Everything between the DFT and the ROCOF stage for one window:
// src/pmu_estimator.c:205-233 (+ iter_e_ipDFT :666-703).  X holds the Hann-windowed
// bins, W is scratch of the same shape.  Returns the debug word.
//
// The reference's sequence
//     f  = e_ipDFT(X);  Xi = X - pure(f);  trigger test;
//     Q x { it = e_ipDFT(Xi);  Xf = X - pure(it);  f = e_ipDFT(Xf);  [Xi = X - pure(f)] }
// is one alternation  "estimate a tone from Y, subtract it from X into W"  with the roles of the two tones swapping, so
// it runs here as a single loop of 2Q + 1 steps around one copy of e_ipdft and one copy of residual.
PMU_HD int estimate_tone(const PostParams& p, const Col& X, const Col& W, Tone* out)
{
    Tone f;
    f.amp = 0; f.cr = 1; f.ci = 0; f.freq = 0;
    int dbg = 0;
    const int last = p.iter_en ? 2 * p.Q : 0;
    for (int s = 0; s <= last; ++s) {
        Tone t;
        const int km = e_ipdft(p, s == 0 ? X : W, &t);       // :205 (s = 0), :679 (odd s: interference), :685 (even s)
        if (s == 0) dbg = km & PMU_DBG_KM_MASK;
        if (!(s & 1)) f = t;                                 // even steps estimate the fundamental
        if (s == last || !p.iter_en) break;
        double Ed, E;
        residual(p, X, t, W, &Ed, &E);                       // :206-215, :680-684, :690-695
        if (s == 0) {
            if (!(Ed > p.eps * E)) break;                    // :227-229 (NaN compares false)
            dbg |= PMU_DBG_TRIGGERED;
        }
    }
    *out = f;
    return dbg;
}

struct RocofState {
    double f_old, dl0, dl1;
    int st;
};

// Two-state ROCOF estimator + first-order low-pass (src/pmu_estimator.c:236-260).  Written with explicitly rounded
// multiplies and adds (no FMA contraction): the kernel contains several inlined copies of this recurrence (single-frame
// batches, the multi-frame ROCOF pass, the lane = bin path) and they must agree bit for bit - and this is also what the
// reference's x86 build computes.
#if defined(__CUDA_ARCH__)
#define PMU_DMUL(a, b) __dmul_rn((a), (b))
#define PMU_DADD(a, b) __dadd_rn((a), (b))
#define PMU_DSUB(a, b) __dsub_rn((a), (b))
#else
PMU_HD double pmu_dround(double v) { volatile double t = v; return t; }
#define PMU_DMUL(a, b) pmu_dround((a) * (b))
#define PMU_DADD(a, b) pmu_dround((a) + (b))
#define PMU_DSUB(a, b) pmu_dround((a) - (b))
#endif
PMU_HD double rocof_step(const PostParams& p, double freq, RocofState* s)
{
    double r = PMU_DMUL(PMU_DSUB(freq, s->f_old), p.frame_rate);                   // :236
    double rd = PMU_DMUL(PMU_DSUB(r, s->dl0), p.frame_rate);                       // :237
    if (!s->st && (fabs(r) > p.thr0 || fabs(rd) > p.thr1)) s->st = 1;              // :240
    if (s->st && fabs(r) < p.thr2) s->st = 0;                                      // :241
    double y = r;                                                                  // :244-253, left to right
    if (!s->st) y = PMU_DSUB(PMU_DADD(PMU_DMUL(p.lp1, r), PMU_DMUL(p.lp2, s->dl0)), PMU_DMUL(p.lp0, s->dl1));
    s->f_old = freq;                                                               // :256-260
    s->dl0 = r;
    s->dl1 = y;
    return y;
}

// The same stage as the reference's float_p = float build executes it: every operand there is a float (phasor.freq,
// freq_old, delay_line, thresholds, coefficients, (float)frame_rate), so each operation rounds to float and the state
// it keeps is float.  The strict threshold tests of :240-241 sit on those rounded values - computing them in double
// flips the state for a stream whose |ROCOF| lands within a float ulp of a threshold.  The state arrays stay double
// in memory (floats are exact in them).
#if defined(__CUDA_ARCH__)
#define PMU_FMUL(a, b) __fmul_rn((a), (b))
#define PMU_FADD(a, b) __fadd_rn((a), (b))
#define PMU_FSUB(a, b) __fsub_rn((a), (b))
#else
PMU_HD float pmu_fround(float v) { volatile float t = v; return t; }   // host emulation: no contraction, no excess precision
#define PMU_FMUL(a, b) pmu_fround((a) * (b))
#define PMU_FADD(a, b) pmu_fround((a) + (b))
#define PMU_FSUB(a, b) pmu_fround((a) - (b))
#endif
PMU_HD double rocof_step_f32(const PostParams& p, double freq, RocofState* s)
{
    const float f = (float)freq, f_old = (float)s->f_old, dl0 = (float)s->dl0, dl1 = (float)s->dl1;
    const float fr = (float)p.frame_rate;
    const float r = PMU_FMUL(PMU_FSUB(f, f_old), fr);                                        // :236
    const float rd = PMU_FMUL(PMU_FSUB(r, dl0), fr);                                         // :237
    if (!s->st && (fabsf(r) > (float)p.thr0 || fabsf(rd) > (float)p.thr1)) s->st = 1;        // :240
    if (s->st && fabsf(r) < (float)p.thr2) s->st = 0;                                        // :241
    float y = r;
    if (!s->st)                                                                              // :246-248, left to right
        y = PMU_FSUB(PMU_FADD(PMU_FMUL((float)p.lp1, r), PMU_FMUL((float)p.lp2, dl0)), PMU_FMUL((float)p.lp0, dl1));
    s->f_old = (double)f;                                                                    // :256-260
    s->dl0 = (double)r;
    s->dl1 = (double)y;
    return (double)y;
}

// wrap_angle (src/pmu_estimator.c:58): ties-to-even rint, range [-pi, pi].
PMU_HD double wrap_angle_pi(double rad)
{
    return rad - 2 * PMU_PI * rint(rad / (2 * PMU_PI));
}

// Frame fill (src/pmu_estimator.c:263-265).
PMU_HD void fill_frame(const PostParams& p, const Tone& t, double rocof, double mid_fracsec, double out[4])
{
    out[0] = 2 * t.amp / p.S;
    out[1] = wrap_angle_pi(atan2(t.ci, t.cr) - 2 * PMU_PI * p.f0 * mid_fracsec);
    out[2] = t.freq;
    out[3] = rocof;
}
