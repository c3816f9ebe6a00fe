// pmu_host.cpp — see pmu_host.hpp.
#include "pmu_host.hpp"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>

namespace pmu {

static thread_local std::string g_err;

void set_error(const std::string& msg)
{
    g_err = msg;
    // LOGGING_LEVEL ERROR behaviour of the reference (src/pmu_estimator.c:38-42)
    fprintf(stderr, "%s\n", msg.c_str());
}
const char* last_error() { return g_err.c_str(); }

// ---- check_config_validity, src/pmu_estimator.c:386-449 (same order of checks) ----
int validate(const pmub_config& c)
{
    char buf[320];
    if (c.f0 != 50 && c.f0 != 60) {
        snprintf(buf, sizeof buf, "[pmu_init] ERROR: nominal frequency: %u not correctly set, (allowed values 50 or 60)", c.f0);
        set_error(buf);
        return -1;
    }
    if (c.n_cycles == 0 || c.n_cycles > 50) {
        snprintf(buf, sizeof buf, "[pmu_init] ERROR: window length: %u cycles is not correctly set, must be in 1..50", c.n_cycles);
        set_error(buf);
        return -1;
    }
    if (c.fs == 0 || fmod((double)c.fs, (double)c.f0) != 0.0) {
        snprintf(buf, sizeof buf, "[pmu_init] ERROR: sample rate: %u is not correctly set, must be a positive multiple of f0: %u", c.fs, c.f0);
        set_error(buf);
        return -1;
    }
    if (c.frame_rate == 0) {
        snprintf(buf, sizeof buf, "[pmu_init] ERROR: frame rate: %u is not correctly set, must be positive", c.frame_rate);
        set_error(buf);
        return -1;
    }
    const unsigned half = (c.n_cycles * c.fs / c.f0) / 2;  // unsigned arithmetic, :411
    if (c.n_bins < c.n_cycles + 2 || c.n_bins > half) {
        snprintf(buf, sizeof buf,
                 "[pmu_init] ERROR: number of dft bins: %u is not correctly set, must be >= (number of cycles + 2): %u and <= half the window length: %u",
                 c.n_bins, c.n_cycles + 2, half);
        set_error(buf);
        return -1;
    }
    if (c.P == 0) {
        set_error("[pmu_init] ERROR: ipdft iterations: 0 is not correctly set, must be positive");
        return -1;
    }
    if (!(c.interf_trig > 0) || c.interf_trig > 1) {
        snprintf(buf, sizeof buf, "[pmu_init] ERROR: interference threshold: %f is not correctly set, must be in ]0,1]", c.interf_trig);
        set_error(buf);
        return -1;
    }
    for (int i = 0; i < 3; ++i) {
        if (!(c.rocof_thresh[i] > 0)) {
            snprintf(buf, sizeof buf, "[pmu_init] ERROR: rocof threshold %d: %f is not set, must be positive", i + 1, c.rocof_thresh[i]);
            set_error(buf);
            return -1;
        }
    }
    return 0;
}

// ---- .ini reader with the line rules of the reference's vendored iniparser ------------
// (libs/iniparser/iniparser.c:620-700 iniparser_line, :714-820 iniparser_load,
//  :491-566 getint/getdouble/getboolean).  Any syntax error rejects the whole file.
static std::string strip(const std::string& s)
{
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b])) ++b;
    while (e > b && isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}
static std::string lower(std::string s)
{
    for (auto& ch : s) ch = (char)tolower((unsigned char)ch);
    return s;
}

static int load_ini(const char* path, std::map<std::string, std::string>* dict)
{
    FILE* in = fopen(path, "r");
    if (!in) {
        set_error(std::string("[pmu_init] ERROR: cannot open ini file: ") + path);
        return -1;
    }
    std::string section, pending;
    char raw[2048];
    int lineno = 0, errs = 0;
    while (fgets(raw, sizeof raw, in)) {
        ++lineno;
        std::string line = pending + raw;
        pending.clear();
        while (!line.empty() && (line.back() == '\n' || isspace((unsigned char)line.back()))) line.pop_back();
        if (!line.empty() && line.back() == '\\') {  // multi-line value
            line.pop_back();
            pending = line;
            continue;
        }
        line = strip(line);
        if (line.empty() || line[0] == '#' || line[0] == ';') continue;
        if (line[0] == '[' && line.back() == ']') {
            size_t close = line.find(']');
            section = lower(strip(line.substr(1, close - 1)));
            continue;
        }
        size_t eq = line.find('=');
        if (eq == std::string::npos || eq == 0) {
            fprintf(stderr, "ini: syntax error in %s (%d):\n-> %s\n", path, lineno, line.c_str());
            ++errs;
            continue;
        }
        std::string key = lower(strip(line.substr(0, eq)));
        std::string rest = line.substr(eq + 1);
        size_t p = 0;
        while (p < rest.size() && isspace((unsigned char)rest[p])) ++p;
        std::string val;
        if (p < rest.size() && (rest[p] == '"' || rest[p] == '\'')) {
            const char qc = rest[p];
            size_t q = rest.find(qc, p + 1);
            if (q != std::string::npos && q > p + 1) {
                val = rest.substr(p + 1, q - p - 1);  // quoted: kept verbatim
            } else {
                // "" / '' or an unterminated quote: handled like an unquoted value
                size_t stop = rest.find_first_of(";#", p);
                val = strip(rest.substr(p, stop == std::string::npos ? std::string::npos : stop - p));
                if (val == "\"\"" || val == "''") val.clear();
            }
        } else {
            size_t stop = rest.find_first_of(";#", p);
            val = strip(rest.substr(p, stop == std::string::npos ? std::string::npos : stop - p));
        }
        (*dict)[section + ":" + key] = val;
    }
    fclose(in);
    if (errs) {
        set_error(std::string("[pmu_init] ERROR: cannot parse file: ") + path);
        return -1;
    }
    return 0;
}

static long ini_int(const std::map<std::string, std::string>& d, const char* k)
{
    auto it = d.find(k);
    return it == d.end() ? 0 : strtol(it->second.c_str(), nullptr, 0);
}
static double ini_double(const std::map<std::string, std::string>& d, const char* k)
{
    auto it = d.find(k);
    return it == d.end() ? 0.0 : atof(it->second.c_str());
}
static int ini_bool(const std::map<std::string, std::string>& d, const char* k)
{
    auto it = d.find(k);
    if (it == d.end() || it->second.empty()) return 0;
    const char ch = it->second[0];
    return (ch == 'y' || ch == 'Y' || ch == '1' || ch == 't' || ch == 'T') ? 1 : 0;
}

// config_from_file, src/pmu_estimator.c:451-484 (missing keys read as 0 and are then
// rejected by validate()).
int parse_ini(const char* path, pmub_config* out)
{
    std::map<std::string, std::string> d;
    if (load_ini(path, &d)) return -1;
    memset(out, 0, sizeof *out);
    out->n_cycles = (unsigned)(int)ini_int(d, "signal:n_cycles");
    out->fs = (unsigned)(int)ini_int(d, "signal:sample_rate");
    out->f0 = (unsigned)(int)ini_int(d, "signal:nominal_freq");
    out->frame_rate = (unsigned)(int)ini_int(d, "synchrophasor:frame_rate");
    out->n_bins = (unsigned)(int)ini_int(d, "synchrophasor:number_of_dft_bins");
    out->P = (unsigned)(int)ini_int(d, "synchrophasor:ipdft_iterations");
    out->Q = (unsigned)(int)ini_int(d, "synchrophasor:iter_e_ipdft_iterations");
    out->interf_trig = ini_double(d, "synchrophasor:interference_threshold");
    out->iter_eipdft = (unsigned)ini_bool(d, "synchrophasor:iter_e_ipdft_enable");
    out->rocof_thresh[0] = ini_double(d, "rocof:threshold_1");
    out->rocof_thresh[1] = ini_double(d, "rocof:threshold_2");
    out->rocof_thresh[2] = ini_double(d, "rocof:threshold_3");
    out->rocof_low_pass_coeffs[0] = ini_double(d, "rocof:low_pass_filter_1");
    out->rocof_low_pass_coeffs[1] = ini_double(d, "rocof:low_pass_filter_2");
    out->rocof_low_pass_coeffs[2] = ini_double(d, "rocof:low_pass_filter_3");
    return 0;
}

// ---- derived constants -----------------------------------------------------------------
template <typename T>
static double hann_sum(unsigned n)  // hann(), src/pmu_estimator.c:549-560: sequential sum in float_p
{
    T sum = 0;
    for (unsigned i = 0; i < n; ++i) {
        T w = (T)(0.5 * (1 - cos(2 * M_PI * i / n)));
        sum += w;
    }
    return (double)sum;
}

Derived derive(const pmub_config& c, int dtype)
{
    Derived d;
    d.N = win_len(c);
    if (dtype == PMUB_F32) {
        d.df = (double)((float)c.fs / (float)d.N);             // :366 in float_p
        d.S = hann_sum<float>(d.N);
        d.invS = (double)(float)(1.0 / d.S);                   // :160
        d.eps = (double)(float)c.interf_trig;
        for (int i = 0; i < 3; ++i) {
            d.thr[i] = (double)(float)c.rocof_thresh[i];
            d.lp[i] = (double)(float)c.rocof_low_pass_coeffs[i];
        }
    } else {
        d.df = (double)c.fs / (double)d.N;
        d.S = hann_sum<double>(d.N);
        d.invS = 1.0 / d.S;
        d.eps = c.interf_trig;
        for (int i = 0; i < 3; ++i) {
            d.thr[i] = c.rocof_thresh[i];
            d.lp[i] = c.rocof_low_pass_coeffs[i];
        }
    }
    return d;
}

// ---- launch plan -------------------------------------------------------------------------
static int align_up(int v, int a) { return (v + a - 1) / a * a; }

static int env_int(const char* name, int dflt)
{
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Plan of the general kernel (pmu_wide.cuh, M == 0): one warp per window, the K+1 lowest bins by chunks of 16 (pruned
// FFT) or by direct summation with a full table of W_N^m, estimator scratch per warp in shared memory.  Takes every configuration the fused kernel
// cannot (n_bins > 23), as long as the per-window scratch fits.
static int make_wide_plan(const pmub_config& c, int max_smem, Plan* out)
{
    Plan p;
    memset(&p, 0, sizeof p);
    const int N = (int)win_len(c);
    const int K = (int)c.n_bins;
    p.M = 0;
    p.Kp = K + 1;
    p.KC = p.Kp;
    p.L = N; p.W = N; p.nslab = 1; p.pitch = N;
    p.n_u = N;                  // U = W_N^m, m < N
    p.jlo = K + 2;
    p.jhi = 2 * K + 1;
    p.raw_stride = 2 * p.Kp;
    // shared memory: the trig table, then one scratch region per active warp (rectangular bins, X, W)
    const int per_warp = align_up(align_up(2 * p.Kp * 8, 16) + 2 * K * 16, 128);
    int off = 0;
    p.off_ctrl = off;
    p.off_U = off; p.off_V = off;
    p.off_trig = off; off += align_up((p.jlo + p.jhi + 1) * 16, 128);
    p.off_raw = off;
    p.off_red = off; p.off_stage = off;
    int warps = (max_smem - off) / per_warp;
    if (warps > 8) warps = 8;
    if (warps < 1) {
        set_error("[pmu_init] ERROR: number of dft bins too large for the CUDA estimator's per-window scratch (about 2300 bins)");
        return -1;
    }
    p.n_cons = warps;           // active warps per 256-thread CTA, one window each
    p.stage_bytes = per_warp;
    p.smem_bytes = off + warps * per_warp;
    *out = p;
    return 0;
}

int make_plan(const pmub_config& c, int dtype, int max_smem, Plan* out)
{
    Plan p;
    memset(&p, 0, sizeof p);
    const int N = (int)win_len(c);
    const int K = (int)c.n_bins;
    const int s = dtype;
    p.Kp = K + 1;
    if (p.Kp <= 8) p.KC = 8;
    else if (p.Kp <= 10) p.KC = 10;
    else if (p.Kp <= 12) p.KC = 12;
    else if (p.Kp <= 16) p.KC = 16;
    else if (p.Kp <= 24) p.KC = 24;
    else return make_wide_plan(c, max_smem, out);
    // codelet size: largest supported power of two dividing N that still leaves >= 32 residues
    const int cand[4] = {16, 8, 4, 1};
    p.M = 1;
    for (int m : cand)
        if (N % m == 0 && (N / m >= 32 || m == 1)) { p.M = m; break; }
    p.L = N / p.M;
    if (p.M == 1 && p.KC == 24 && s == 8) return make_wide_plan(c, max_smem, out);   // see launch_kc (pmu_launch.cuh)

    if (p.M == 16) {
        // fixed row pitch: slabs are at most PMU_PITCH16 residues wide, rows PMU_PITCH16 elements apart; 1024-sample
        // windows (64 residues) are kept at their natural pitch so that two of them share a warp pass (below)
        p.pitch = (p.L == PMU_PITCH16_SHORT && env_int("PMU_PACK", 4) >= 2) ? PMU_PITCH16_SHORT : PMU_PITCH16;
        // a multiple of 96 residues that is not one of 128 (M-class: 3072 = 192 x 16): 96-wide slabs fill every stage
        if (p.L % PMU_PITCH16_96 == 0 && p.L % PMU_PITCH16 != 0 && env_int("PMU_PITCH96", 1)) p.pitch = PMU_PITCH16_96;
        p.W = p.L < p.pitch ? p.L : p.pitch;
        p.nslab = (p.L + p.W - 1) / p.W;
        p.stage_bytes = align_up(p.M * p.pitch * s, 128);
    } else {
        const int stage_target = env_int("PMU_STAGE_BYTES", 16384);
        if (N * s <= stage_target) {
            p.W = p.L;
            p.nslab = 1;
        } else {
            int w = stage_target / (p.M * s) / 32 * 32;
            if (w < 32) w = 32;
            if (w >= p.L) w = p.L;
            p.W = w;
            p.nslab = (p.L + w - 1) / w;
        }
        p.pitch = p.W;
        p.stage_bytes = align_up(p.M * p.W * s, 128);
    }
    // Short windows: four per warp pass (8 lanes each) - one bulk copy and one round of bookkeeping for four
    // windows, no idle lanes in the last residue group.  Needs the natural-pitch layout of the 8-point-codelet
    // kernels and a stage that still leaves the ring deep enough.
    p.wp = 1;
    if (p.M == 8 && p.nslab == 1 && p.pitch == p.L && env_int("PMU_PACK", 4) >= 4 && 4 * N * s <= 20 * 1024) {
        p.wp = 4;
        p.stage_bytes = align_up(p.wp * N * s, 128);
    } else if (p.M == 16 && p.nslab == 1 && p.pitch == p.L && (p.pitch == PMU_PITCH16_SHORT || (p.pitch == PMU_PITCH16 && s == 4)) &&
               env_int("PMU_PACK", 4) >= 2 && 2 * N * s <= 16 * 1024) {
        p.wp = 2;   // 1024-sample windows, and 2048-sample windows of the float build
        p.stage_bytes = align_up(p.wp * N * s, 128);
    }
    // float build: two adjacent residues per lane on the packed FP32 pipe (FADD2 / FMUL2 / FFMA2)
#ifndef PMU_F32X2
#define PMU_F32X2 1
#endif
    p.pk = (s == 4 && PMU_F32X2) ? 2 : 1;
    const int lanes_per_window = (32 / p.wp) * p.pk;   // residues covered per iteration
    // TMA bulk copies: whole window when its rows are already contiguous at the stage pitch,
    // otherwise one copy per row; both need 16-byte granularity
    const bool whole = (p.nslab == 1 && p.pitch == p.L);
    p.aligned16 = whole ? ((N * s) % 16 == 0)
                        : ((N * s) % 16 == 0 && (p.L * s) % 16 == 0 && (p.W * s) % 16 == 0 &&
                           ((p.L - (p.nslab - 1) * p.W) * s) % 16 == 0 && (p.pitch * s) % 16 == 0);

    // U[slab][group][k] = W_N^{(slab * W + group * residues per iteration) k}: rectangular, the last slab may use fewer groups
    p.u_per_slab = (p.W + lanes_per_window - 1) / lanes_per_window;
    p.n_u = p.nslab * p.u_per_slab * p.KC;
    p.jlo = K + 2;
    p.jhi = 2 * K + 1;
    // one window's K+1 rectangular bins as 16-byte complex entries, rows an odd number of entries apart: the lane-per-
    // window 128-bit reads of the estimator (which Hann-combines the row in place and uses it as X) stay conflict-free
    p.raw_stride = 2 * (p.Kp | 1);
    // per-lane row of the reduction scratch: >= 2*Kp elements, an odd number of 16-byte units so that the
    // 128-bit stores of a quarter warp hit distinct banks
    p.red_stride = 2 * p.KC + 2;  // compile-time in the kernel (immediate LDS offsets)
    // ... and the same region holds the estimator's residual spectrum W[K][32] while the warp posts a batch
    const int dft_elem = (s == 4) ? 4 : 8;   // the float build reduces float partials
    p.red_bytes = align_up(32 * p.red_stride * dft_elem, 128);
    if (p.red_bytes < K * 32 * 16) p.red_bytes = K * 32 * 16;

    int want_cons = env_int("PMU_CONSUMER_WARPS", PMU_MAX_CONSUMERS);
    if (want_cons < 1) want_cons = 1;
    if (want_cons > PMU_MAX_CONSUMERS) want_cons = PMU_MAX_CONSUMERS;
    int ns_cap = env_int("PMU_STAGES", PMU_MAX_STAGES);
    if (ns_cap > PMU_MAX_STAGES) ns_cap = PMU_MAX_STAGES;

    auto layout = [&](int n_cons, int n_raw) {
        p.n_cons = n_cons; p.n_raw = n_raw;
        int off = 0;
        p.off_ctrl = off; off += align_up((int)sizeof(CtrlBlock), 128);
        p.off_U = off;    off += align_up(p.n_u * 16, 128);
        p.off_V = off;    off += align_up(p.KC * 32 * 16, 128);   // (pairs of float twiddles per lane when pk == 2: same size)
        p.off_trig = off; off += align_up((p.jlo + p.jhi + 1) * 16, 128);
        p.off_raw = off;  off += align_up(p.n_raw * 32 * p.raw_stride * 8, 128);
        p.off_red = off;  off += p.n_cons * p.red_bytes;
        p.off_stage = off;
        int ns = (max_smem - off) / p.stage_bytes;
        if (ns > ns_cap) ns = ns_cap;
        p.n_stage = ns;
        p.smem_bytes = off + (ns > 0 ? ns : 0) * p.stage_bytes;
        return ns;
    };
    // Shared-memory budget.  A raw-bin buffer is held by its batch until the estimator has finished with it, so the
    // ring of raw buffers bounds how many batches can be in the estimator at once (plus the one or two being filled).
    // Ring depth is still worth more than anything else (profiles/r1_sweep_smem.jsonl): take as many raw buffers as fit
    // WITHOUT costing a stage relative to the minimum of 4; only the interference iterations (iter_e_ipDFT, :666-703)
    // make a batch stay long enough for more than 4 to matter.  When bins/windows are large, give up raw buffers,
    // then consumer warps, until at least `min_stages` ring stages fit.
    const int forced_raw = env_int("PMU_RAW_BUFFERS", 0);
    const int cons_tries[] = {want_cons, 5, 3, 2, 1};
    bool ok = false;
    for (int pass = 0; pass < 2 && !ok; ++pass) {
        const int min_stages = pass == 0 ? 4 : 2;
        for (int nc0 : cons_tries) {
            const int nc = nc0 < want_cons ? nc0 : want_cons;
            if (forced_raw >= 2) {
                if (layout(nc, forced_raw > PMU_MAX_RAW ? PMU_MAX_RAW : forced_raw) >= min_stages) { ok = true; break; }
                continue;
            }
            const int base_raw = nc >= 3 ? 4 : 3;
            const int ns_ref = layout(nc, base_raw);
            if (ns_ref < min_stages) {
                if (layout(nc, 2) >= min_stages) { ok = true; break; }
                continue;
            }
            int nr = base_raw;
            const int nr_max = c.iter_eipdft ? (nc + 1 < PMU_MAX_RAW ? nc + 1 : PMU_MAX_RAW) : base_raw;
            // with the interference iterations enabled, one stage beyond the eighth is worth less than the extra
            // batches in flight (the ninth stage adds < 1 % to the streaming rate)
            int ns_keep = ns_ref;
            if (c.iter_eipdft && ns_ref > 8 && env_int("PMU_TRADE_STAGE", 1)) ns_keep = ns_ref - 1;
            for (int cand_raw = nr_max; cand_raw > base_raw; --cand_raw)
                if (layout(nc, cand_raw) >= ns_keep) { nr = cand_raw; break; }
            layout(nc, nr);
            // Few stages (wide bin accumulators and/or multi-slab windows whose scratch eats the shared memory): a
            // consumer warp holds a stage for as long as it transforms it, so warps beyond n_stage - 3 only starve the
            // loads in flight.  Trade consumer warps (and raw buffers) for ring stages.
            if (p.n_stage < 7 && forced_raw < 2 && env_int("PMU_REBALANCE", 1)) {
                int best_nc = nc, best_nr = nr, best_eff = (nc < p.n_stage - 3 ? nc : p.n_stage - 3), best_ns = p.n_stage;
                for (int c2 = nc; c2 >= 3; --c2)
                    for (int r2 = nr; r2 >= 3; --r2) {
                        const int ns = layout(c2, r2);
                        const int eff = c2 < ns - 3 ? c2 : ns - 3;
                        if (eff > best_eff || (eff == best_eff && ns > best_ns && c2 >= best_nc)) {
                            best_eff = eff; best_nc = c2; best_nr = r2; best_ns = ns;
                        }
                    }
                layout(best_nc, best_nr);
            }
            ok = true;
            break;
        }
    }
    if (!ok) return make_wide_plan(c, max_smem, out);   // e.g. very long windows with many bins
    *out = p;
    return 0;
}

void build_tables(const Plan& p, unsigned N, std::vector<pmu_c>* U, std::vector<pmu_c>* V, std::vector<pmu_c>* trig)
{
    const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
    auto w = [&](unsigned long long idx) {  // W_N^idx = exp(-2 pi i idx / N), index reduced exactly
        unsigned long long r = idx % N;
        long double ang = two_pi * (long double)r / (long double)N;
        pmu_c o;
        o.x = (double)cosl(ang);
        o.y = (double)(-sinl(ang));
        return o;
    };
    if (p.M == 0) {   // general kernel: the full twiddle table, no lane table
        U->resize((size_t)N);
        for (unsigned m = 0; m < N; ++m) (*U)[m] = w(m);
        V->assign(1, w(0));
    }
    const int pk = p.pk > 1 ? p.pk : 1;
    const int per_it = (32 / (p.wp > 0 ? p.wp : 1)) * pk;
    if (p.M != 0) {
        U->resize((size_t)p.n_u);
        for (int sl = 0; sl < p.nslab; ++sl)
            for (int g = 0; g < p.u_per_slab; ++g)
                for (int k = 0; k < p.KC; ++k)
                    (*U)[((size_t)sl * p.u_per_slab + g) * p.KC + k] =
                        w(((unsigned long long)sl * p.W + (unsigned long long)g * per_it) * (unsigned long long)k);
        // lane twiddles W_N^{a k} for the 32 * pk residues a of one iteration
        V->resize((size_t)p.KC * 32 * pk);
        for (int k = 0; k < p.KC; ++k)
            for (int l = 0; l < 32 * pk; ++l) (*V)[(size_t)k * 32 * pk + l] = w((unsigned long long)l * (unsigned long long)k);
    }
    trig->resize((size_t)(p.jlo + p.jhi + 1));
    for (int j = -p.jlo; j <= p.jhi; ++j) {
        long double ang = 3.14159265358979323846264338327950288L * (long double)j / (long double)N;
        pmu_c o;
        o.x = (double)cosl(ang);
        o.y = (double)sinl(ang);
        (*trig)[(size_t)(j + p.jlo)] = o;
    }
}

double bytes_per_estimate(unsigned N, int dtype, unsigned n_channels)
{
    const double s = dtype;
    return N * s + 4 * s + 2 * (3 * s + 1) + s / (double)n_channels;
}

}  // namespace pmu
