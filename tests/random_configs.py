"""Random estimator configurations that the reference's check_config_validity (src/pmu_estimator.c:386-449) accepts."""
import numpy as np

from synchropmu_b200 import api


def random_valid_configs(n, seed):
    """Configurations check_config_validity (:386-449) accepts, drawn to hit every kernel plan: 16/8/4/1-point codelets,
    single- and multi-slab windows, packed short windows, 96-wide slabs, the general kernel (n_bins > 23)."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        f0 = int(rng.choice([50, 60]))
        fs = f0 * int(rng.choice([20, 25, 32, 37, 48, 64, 75, 96, 100, 128, 171, 200, 256, 320, 512, 1024]))
        n_cycles = int(rng.integers(1, 13))
        N = n_cycles * fs // f0
        if N < 16 or N > 20000:
            continue
        hi = min(N // 2, n_cycles + 2 + int(rng.choice([0, 1, 3, 8, 20])))
        if hi < n_cycles + 2:
            continue
        cfg = dict(n_cycles=n_cycles, f0=f0, frame_rate=int(rng.choice([f0, f0 // 2, 10])), fs=fs,
                   n_bins=int(rng.integers(n_cycles + 2, hi + 1)), P=int(rng.integers(1, 4)), Q=int(rng.integers(1, 4)),
                   interf_trig=float(rng.choice([0.0033, 0.02])), rocof_thresh=[3.0, 25.0, 0.035],
                   rocof_low_pass_coeffs=[0.5913, 0.2043, 0.2043], iter_eipdft=int(rng.integers(0, 2)))
        if api.validate(cfg):
            out.append(cfg)
    return out
