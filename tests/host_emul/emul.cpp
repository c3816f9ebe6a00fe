// emul.cpp — CPU emulation of the CUDA kernel's per-lane arithmetic (TEST INFRASTRUCTURE).
//
// Compiles the SAME headers the kernel uses (pmu_math.cuh, pmu_dft.cuh, pmu_host.cpp's
// tables and plan) with g++ and walks the lane/slab/residue loops of pmu_kernel.cuh on
// the host, so the restructured DFT and sweep arithmetic can be checked against the
// oracle in the GPU-less dev container.  It is never linked into libpmu_estimator*.so
// and is not a fallback: the product only runs the arithmetic on the GPU.
#include <string.h>

#include <vector>

#include "../../synchropmu_b200/csrc/pmu_host.hpp"

static int g_fp32_dft = 1;   // the float build runs its DFT stage in FP32 (like the kernel)

template <typename InT, int M, int KC, typename R>
static void window_bins_r(const pmu::Plan& p, const std::vector<pmu_c>& Ud, const std::vector<pmu_c>& Vd,
                          const InT* x, double* raw /* 2*KC */)
{
    typedef typename CplxOf<R>::type cplx;
    std::vector<cplx> U(Ud.size()), V(Vd.size());
    for (size_t i = 0; i < Ud.size(); ++i) { U[i].x = (R)Ud[i].x; U[i].y = (R)Ud[i].y; }
    for (size_t i = 0; i < Vd.size(); ++i) { V[i].x = (R)Vd[i].x; V[i].y = (R)Vd[i].y; }
    double sum_r[KC], sum_i[KC];
    for (int k = 0; k < KC; ++k) sum_r[k] = sum_i[k] = 0.0;
    const int w_per_it = p.u_per_slab;
    // residues per iteration = virtual lanes per window: 32 / wp lanes, each carrying pk adjacent residues (float build: 2)
    const int G = (32 / p.wp) * (p.pk > 1 ? p.pk : 1);
    for (int lane = 0; lane < G; ++lane) {
        R ar[KC], ai[KC];
        bool first = true;
        for (int sl = 0; sl < p.nslab; ++sl) {
            const int a_lo = sl * p.W;
            const int w_s = (p.W < p.L - a_lo) ? p.W : p.L - a_lo;
            const int n_it = (w_s + G - 1) / G;
            for (int ii = 0; ii < n_it; ++ii) {
                const int a_loc = lane + G * ii;
                const bool valid = a_loc < w_s;
                R xv[M];
                for (int b = 0; b < M; ++b) xv[b] = valid ? (R)x[(size_t)(a_lo + a_loc) + (size_t)p.L * b] : (R)0;
                const cplx* u = U.data() + (size_t)(sl * w_per_it + ii) * KC;
                if (first) { accumulate_residue<M, KC, true, R>(xv, u, ar, ai); first = false; }
                else accumulate_residue<M, KC, false, R>(xv, u, ar, ai);
            }
        }
        rotate_lane<KC, R>(V.data() + lane, 32 * (p.pk > 1 ? p.pk : 1), ar, ai);   // the lane table covers every residue of one iteration
        for (int k = 0; k < KC; ++k) { sum_r[k] += (double)ar[k]; sum_i[k] += (double)ai[k]; }   // column sums in double
    }
    for (int k = 0; k < KC; ++k) { raw[2 * k] = sum_r[k]; raw[2 * k + 1] = sum_i[k]; }
}

template <typename InT, int M, int KC>
static void window_bins(const pmu::Plan& p, const std::vector<pmu_c>& U, const std::vector<pmu_c>& V, const InT* x, double* raw)
{
    if (g_fp32_dft && sizeof(InT) == 4) window_bins_r<InT, M, KC, float>(p, U, V, x, raw);
    else window_bins_r<InT, M, KC, double>(p, U, V, x, raw);
}

template <typename InT, int M>
static void window_bins_kc(int KC, const pmu::Plan& p, const std::vector<pmu_c>& U, const std::vector<pmu_c>& V, const InT* x, double* raw)
{
    switch (KC) {
        case 8: window_bins<InT, M, 8>(p, U, V, x, raw); break;
        case 10: window_bins<InT, M, 10>(p, U, V, x, raw); break;
        case 12: window_bins<InT, M, 12>(p, U, V, x, raw); break;
        case 16: window_bins<InT, M, 16>(p, U, V, x, raw); break;
        case 24: window_bins<InT, M, 24>(p, U, V, x, raw); break;
    }
}
// The general kernel's bins (pmu_wide.cuh): a warp per bin, lane l sums the samples n = l + 32 i into four partial
// sums with the table twiddle W_N^(k n mod N), then a butterfly over the lanes.
template <typename InT>
static void window_bins_wide(const pmu::Plan& p, const std::vector<pmu_c>& U, const InT* x, double* raw)
{
    const unsigned N = (unsigned)p.L;
    for (int k = 0; k < p.Kp; ++k) {
        double lr[32], li[32];
        for (unsigned l = 0; l < 32; ++l) {
            double sr[4] = {0, 0, 0, 0}, si[4] = {0, 0, 0, 0};
            unsigned i = 0;
            for (unsigned n = l; n < N; n += 32, ++i) {
                const pmu_c tw = U[(size_t)(((unsigned long long)k * n) % N)];
                sr[i & 3] += (double)x[n] * tw.x;
                si[i & 3] += (double)x[n] * tw.y;
            }
            lr[l] = (sr[0] + sr[1]) + (sr[2] + sr[3]);
            li[l] = (si[0] + si[1]) + (si[2] + si[3]);
        }
        for (int off = 16; off >= 1; off >>= 1) {
            double nr[32], ni[32];
            for (int l = 0; l < 32; ++l) { nr[l] = lr[l] + lr[l ^ off]; ni[l] = li[l] + li[l ^ off]; }
            memcpy(lr, nr, sizeof lr);
            memcpy(li, ni, sizeof li);
        }
        raw[2 * k] = lr[0];
        raw[2 * k + 1] = li[0];
    }
}

template <typename InT>
static void window_bins_m(const pmu::Plan& p, const std::vector<pmu_c>& U, const std::vector<pmu_c>& V, const InT* x, double* raw)
{
    switch (p.M) {
        case 0: window_bins_wide<InT>(p, U, x, raw); break;
        case 16: window_bins_kc<InT, 16>(p.KC, p, U, V, x, raw); break;
        case 8: window_bins_kc<InT, 8>(p.KC, p, U, V, x, raw); break;
        case 4: window_bins_kc<InT, 4>(p.KC, p, U, V, x, raw); break;
        case 1: window_bins_kc<InT, 1>(p.KC, p, U, V, x, raw); break;
    }
}

extern "C" {

// The launch plan for a configuration, field by field, so that its invariants can be checked without a GPU.
// out: M, KC, L, W, nslab, pitch, wp, pk, Kp, n_stage, stage_bytes, n_cons, n_raw, raw_stride, red_bytes, off_raw, off_red,
//      off_stage, smem_bytes, aligned16   (20 ints).  Returns 0 / -1.
int emul_plan(const pmub_config* cfg, int dtype, int max_smem, int* out)
{
    if (pmu::validate(*cfg)) return -1;
    pmu::Plan p;
    if (pmu::make_plan(*cfg, dtype, max_smem, &p)) return -1;
    const int v[20] = {p.M, p.KC, p.L, p.W, p.nslab, p.pitch, p.wp, p.pk, p.Kp, p.n_stage, p.stage_bytes, p.n_cons, p.n_raw,
                       p.raw_stride, p.red_bytes, p.off_raw, p.off_red, p.off_stage, p.smem_bytes, p.aligned16 ? 1 : 0};
    for (int i = 0; i < 20; ++i) out[i] = v[i];
    return 0;
}

// windows [T][S][N] (dtype), mid [T][S/C] (double), frames [T][S][4] (double),
// flags [T][S], spectrum [T][S][K][2].  State starts fresh.  Returns 0 / -1.
int emul_run(const pmub_config* cfg, int dtype, int n_channels, long n_streams, long n_frames, const void* windows,
             const double* mid, double* frames, int* flags, double* spectrum, int* plan_out /* M, KC, nslab, W, wp */)
{
    if (pmu::validate(*cfg)) return -1;
    pmu::Plan p;
    if (pmu::make_plan(*cfg, dtype, 232448, &p)) return -1;
    pmu::Derived d = pmu::derive(*cfg, dtype);
    std::vector<pmu_c> U, V, T;
    pmu::build_tables(p, d.N, &U, &V, &T);
    if (plan_out) { plan_out[0] = p.M; plan_out[1] = p.KC; plan_out[2] = p.nslab; plan_out[3] = p.W; plan_out[4] = p.wp; }

    PostParams pp;
    memset(&pp, 0, sizeof pp);
    pp.K = (int)cfg->n_bins; pp.P = (int)cfg->P; pp.Q = (int)cfg->Q; pp.iter_en = cfg->iter_eipdft ? 1 : 0;
    pp.jlo = p.jlo; pp.jhi = p.jhi; pp.C = n_channels; pp.f0 = (int)cfg->f0;
    pp.df = d.df; pp.inv_df = 1.0 / d.df; pp.S = d.S; pp.invS = d.invS; pp.eps = d.eps; pp.invN = 1.0 / (double)d.N;
    const long double pi = 3.14159265358979323846264338327950288L;
    pp.c3q = (double)(-0.25L * cosl(pi / (long double)d.N));
    pp.s3q = (double)(0.25L * sinl(pi / (long double)d.N));
    pp.frame_rate = (double)cfg->frame_rate;
    pp.thr0 = d.thr[0]; pp.thr1 = d.thr[1]; pp.thr2 = d.thr[2];
    pp.lp0 = d.lp[0]; pp.lp1 = d.lp[1]; pp.lp2 = d.lp[2];
    pp.trig = T.data();

    const int K = pp.K;
    std::vector<RocofState> st((size_t)n_streams);
    for (auto& s : st) { s.f_old = (double)cfg->f0; s.dl0 = 0; s.dl1 = 0; s.st = 0; }
    std::vector<double> raw((size_t)2 * p.KC + 2);
    std::vector<pmu_c> X((size_t)K), W((size_t)K);
    for (long t = 0; t < n_frames; ++t) {
        for (long s = 0; s < n_streams; ++s) {
            const size_t u = (size_t)t * n_streams + s;
            if (dtype == PMUB_F64) window_bins_m<double>(p, U, V, (const double*)windows + u * d.N, raw.data());
            else window_bins_m<float>(p, U, V, (const float*)windows + u * d.N, raw.data());
            const Col cx = Col::make(X.data(), 1), cw = Col::make(W.data(), 1);
            for (int k = 0; k < K; ++k) {
                pmu_c xw = hann_combine(raw.data(), k);
                cx.put(k, xw.x, xw.y);
                if (spectrum) { spectrum[(u * K + k) * 2] = xw.x; spectrum[(u * K + k) * 2 + 1] = xw.y; }
            }
            Tone tone;
            int dbg = estimate_tone(pp, cx, cw, &tone);
            double y = dtype == PMUB_F32 ? rocof_step_f32(pp, tone.freq, &st[(size_t)s]) : rocof_step(pp, tone.freq, &st[(size_t)s]);
            double m = mid[(size_t)t * (n_streams / n_channels) + s / n_channels];
            if (dtype == PMUB_F32) m = (double)(float)m;
            fill_frame(pp, tone, y, m, frames + u * 4);
            if (flags) flags[u] = dbg;
        }
    }
    return 0;
}

void emul_set_fp32_dft(int on) { g_fp32_dft = on; }
int emul_parse_ini(const char* path, pmub_config* out) { return pmu::parse_ini(path, out); }
int emul_validate(const pmub_config* c) { return pmu::validate(*c); }

}  // extern "C"
