// pmu_host.hpp — host-side logic of the estimator: configuration, validation, .ini
// parsing, derived constants, twiddle tables and the kernel launch plan.  Pure C++
// (no CUDA calls) so it can be unit-tested without a GPU.
//
// Mirrors, on the host, reference src/pmu_estimator.c:92-167 (pmu_init),
// :327-384 (config_estimator), :386-449 (check_config_validity), :451-484
// (config_from_file) and :549-560 (hann).
#pragma once

#include <string>
#include <vector>

#include "../../include/pmu_batch.h"
#include "pmu_kernel.cuh"

namespace pmu {

void set_error(const std::string& msg);  // stores for pmub_last_error() and prints to stderr
const char* last_error();

// check_config_validity + the win_len computation.  Returns 0 / -1 (message set).
int validate(const pmub_config& c);
// iniparser-compatible reader for the 15 keys of config_from_file.  0 / -1.
int parse_ini(const char* path, pmub_config* out);

inline unsigned win_len(const pmub_config& c) { return c.n_cycles * c.fs / c.f0; }

// Constants derived at init, rounded through float_p like the reference does.
struct Derived {
    unsigned N;
    double df, S, invS, eps, thr[3], lp[3];
};
Derived derive(const pmub_config& c, int dtype);

// How the kernel is specialised and laid out for one configuration.
struct Plan {
    int M, KC;          // codelet size, bins per lane
    int L, W, nslab;    // residues, slab width, slabs per window
    int pitch;          // stage row pitch in elements (PMU_PITCH16 for the 16-point codelet, else W)
    int wp;             // windows per warp pass: 4 for short single-slab windows (8 lanes per window), else 1
    int pk;             // 2 in the float build: every lane carries two ADJACENT residues of its window through the packed
                        // FP32 pipe (FADD2 / FMUL2 / FFMA2), so a pass covers (32 / wp) * pk residues per iteration
    int u_per_slab;     // residue groups (U rows) per slab: ceil(W / ((32 / wp) * pk))
    int Kp;             // n_bins + 1
    int n_stage, stage_bytes;
    int n_cons, n_raw, raw_stride;
    int red_bytes;      // per consumer warp: lane-reduction scratch = the estimator's W columns while the warp posts a batch
    int n_u;
    int jlo, jhi;
    int off_ctrl, off_U, off_V, off_trig, off_raw, off_red, off_stage;
    int red_stride;     // doubles per lane row of the per-warp reduction scratch
    int smem_bytes;
    bool aligned16;     // window pitch and row pitch allow TMA bulk copies
};
// max_smem: opt-in dynamic shared memory per block of the device (232448 on B200).
int make_plan(const pmub_config& c, int dtype, int max_smem, Plan* out);

// Twiddle / trig tables (host copies, uploaded once at creation).
void build_tables(const Plan& p, unsigned N, std::vector<pmu_c>* U, std::vector<pmu_c>* V,
                  std::vector<pmu_c>* trig);

double bytes_per_estimate(unsigned N, int dtype, unsigned n_channels);

}  // namespace pmu
