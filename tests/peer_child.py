"""Child process of test_frames_land_in_another_process_buffer: opens the parent's CUDA-IPC buffer and lets the
estimator kernel write its frames straight into it (what rank r does with rank 0's buffer over NVLink)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from synchropmu_b200 import BatchEstimator, api, configs, signals  # noqa: E402

handle = bytes.fromhex(sys.argv[1])
offset = int(sys.argv[2])
n_inst = int(sys.argv[3])
cfg = configs.DEFAULT_INI
import torch  # noqa: E402

prm = signals.three_phase_params(n_inst, seed=12)
x = signals.steady(cfg, 0, prm["amp"], prm["freq"], prm["phase"], ramp=prm["ramp"]).reshape(n_inst, 6, -1)
base = api.peer_open(handle, 0)
with BatchEstimator(cfg, n_inst, 6, device=0) as est:
    dw = torch.from_numpy(x).cuda()
    dm = torch.zeros(n_inst, dtype=torch.float64, device="cuda")
    est.estimate(dw, dm, out=api.RawDevicePtr(base + offset))
api.peer_close(base)
print("child done")
