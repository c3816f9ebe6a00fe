"""B200-native batched synchrophasor estimator with SynchroPMU's C API as a drop-in.

Only the pmu_estimate() hot path of chemsallioua/SynchroPMU is implemented here, as a
hand-written sm_100a CUDA kernel behind a C ABI (include/pmu_estimator.h,
include/pmu_batch.h).  This package holds the kernel sources (csrc/), the built libraries
(lib/), and thin ctypes bindings:

* :class:`synchropmu_b200.api.BatchEstimator` — the batched entry (many instances x channels);
* :class:`synchropmu_b200.pmu_estimator.PMUEstimator` — mirror of the reference's
  single-instance Python wrapper.
"""
from . import configs, signals  # noqa: F401
from .api import BatchEstimator, build_library, load_library, shard_instances  # noqa: F401
from .pmu_estimator import EstimatorConfig, PMUEstimator  # noqa: F401

__all__ = ["BatchEstimator", "PMUEstimator", "EstimatorConfig", "build_library", "load_library",
           "shard_instances", "configs", "signals"]
