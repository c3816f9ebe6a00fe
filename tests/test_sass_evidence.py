"""SASS evidence for the shipped library, checked on the CPU (cuobjdump needs no GPU): the fused kernels keep their
TMA / bulk-copy / mbarrier / programmatic-dependent-launch instructions and contain no local-memory (spill) traffic, no
tensor-core instructions (north_star: the path is HBM/latency-bound FP64, not a dense contraction)."""
import shutil
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
LIB = ROOT / "synchropmu_b200" / "lib" / "libpmu_estimator.so"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not LIB.exists(),
                                reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def table():
    import sass_opcodes

    return sass_opcodes.opcode_table(LIB)


def test_no_fused_kernel_touches_local_memory(table):
    fused = {k: v for k, v in table.items() if "pmu_fused_kernel" in k}
    assert len(fused) >= 80
    bad = {k: (v["LDL"], v["STL"]) for k, v in fused.items() if v["LDL"] or v["STL"]}
    assert not bad, f"register spills in {bad}"


def test_headline_kernels_keep_their_copy_engine_and_barrier_instructions(table):
    for name in ("pmu_fused_kernel<double, double, 16, 12, 128, 1>",     # config 3: one bulk copy per window
                 "pmu_fused_kernel<double, double, 16, 12, 96, 1>",      # M-class: one tiled TMA copy per slab
                 "pmu_fused_kernel<double, double, 8, 10, 0, 4>",        # config 5: four windows per bulk copy
                 "pmu_fused_kernel<float, F2, 16, 12, 128, 2>"):       # float build: packed FP32 transform
        k = table[name]
        assert k["UBLKCP"] > 0, name          # cp.async.bulk.shared::cluster.global
        assert k["SYNCS"] > 0, name           # mbarrier arrive / try_wait
        assert k["ACQBULK"] > 0, name         # griddepcontrol.wait (programmatic dependent launch)
        assert k["LDGSTS"] > 0, name          # the per-element cp.async fallback loader is still compiled in
        assert k["UTCMMA"] == 0 and k["LDTM"] == 0, name
        assert k["DFMA"] > 300, name
    assert table["pmu_fused_kernel<double, double, 16, 12, 96, 1>"]["UTMALDG"] > 0   # cp.async.bulk.tensor
    # every kernel of the float build runs its transform on the packed FP32 pipe (two adjacent residues per lane)
    assert not [k for k in table if "pmu_fused_kernel<float, float" in k]
    for name in ("pmu_fused_kernel<float, F2, 16, 12, 128, 2>", "pmu_fused_kernel<float, F2, 8, 10, 0, 4>",
                 "pmu_fused_kernel<float, F2, 16, 12, 96, 1>"):
        k = table[name]
        assert k["FFMA2"] > 20 and k["FADD2"] > 20 and k["FMUL2"] > 0, name
        assert k["UBLKCP"] > 0 and k["LDL"] == 0 and k["STL"] == 0, name
    # the lane = bin peak search of small launches is two integer warp reductions
    assert table["pmu_fused_kernel<double, double, 16, 12, 128, 1>"]["CREDUX"] + table["pmu_fused_kernel<double, double, 16, 12, 128, 1>"]["REDUX"] >= 2
