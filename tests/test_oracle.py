"""The CPU checker (oracle/pmu_oracle.c) pinned against the reference.

Pins: (1) tests/golden/* — outputs of the unmodified reference build, committed;
(2) oracle/_ref/libref_*.so itself when it is present (dev container and, because the
built files travel with the snapshot, the GPU box); (3) the values recorded in
SURVEY.md section 4.  None of these tests needs a GPU.
"""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
from oracle_lib import CpuChecker
from parity import assert_frames_close, golden_floats
from stream_cases import STREAM_CASES, stream_case_windows
from synchropmu_b200 import configs, signals

GOLD = Path(__file__).resolve().parent / "golden"
SINGLE = json.loads((GOLD / "single_calls.json").read_text())
NPZ = np.load(GOLD / "single_calls.npz")
STREAMS = json.loads((GOLD / "streams.json").read_text())
DT = {"f64": np.float64, "f32": np.float32}


@pytest.fixture(scope="session", autouse=True)
def _built():
    oracle_lib.build_oracle()


@pytest.mark.parametrize("case", SINGLE, ids=[c["key"] for c in SINGLE])
def test_port_matches_golden_single_call(case):
    dt = DT[case["dtype"]]
    ck = CpuChecker("port", dt, case["C"])
    h = ck.open(case["cfg"])
    assert h
    assert ck.win_len(h) == case["win_len"]
    assert ck.norm_factor(h) == case["norm_factor"]          # sequential Hann sum, bit-exact
    frames, rc = ck.estimate(h, NPZ[case["key"] + "__win"], case["mid"])
    assert rc == 0
    want = golden_floats(case["frames"]).reshape(case["C"], 4)
    assert_frames_close(frames, want, "tight" if case["dtype"] == "f64" else "f32", case["key"])
    assert ck.last_km(h) == case["km_last"]
    spec = ck.last_spectrum(h)
    want_spec = golden_floats(case["spectrum_re"]) + 1j * golden_floats(case["spectrum_im"])
    scale = max(np.max(np.abs(want_spec)), 1e-300)
    rel = 1e-12 if case["dtype"] == "f64" else 1e-5
    assert np.max(np.abs(spec - want_spec)) <= rel * scale
    ck.close(h)


@pytest.mark.parametrize("run", STREAMS, ids=[f"{r['case']}__{r['dtype']}" for r in STREAMS])
def test_port_matches_golden_streams(run):
    sc = next(s for s in STREAM_CASES if s["name"] == run["case"])
    dt = DT[run["dtype"]]
    win, mid = stream_case_windows(sc, dtype=dt)
    ck = CpuChecker("port", dt, 1)
    frames, km, _ = ck.run(configs.NAMED[sc["cfg"]], win, mid, n_threads=2)
    want = golden_floats(run["frames"]).reshape(run["shape"])
    # inputs are regenerated (libm cos may differ by an ulp between hosts): 1e-10-level gate
    tol = dict(amp_rel=1e-10, ph=1e-10, freq=1e-9, rocof=1e-6) if run["dtype"] == "f64" else "f32"
    assert_frames_close(frames, want, tol, run["case"])
    if run["dtype"] == "f64":
        assert km.reshape(-1).tolist() == run["km"]


def test_survey_section4_pinned_values():
    """SURVEY.md section 4 table, double build, fresh context, mid_fracsec = 0."""
    ck = CpuChecker("port", np.float64, 1)
    cfg = configs.DEFAULT_INI
    h = ck.open(cfg)
    x = signals.steady(cfg, 0, [2.0], [51.0], [0.0], harmonics=[(0.1, 75.0, 0.0)])
    fr, _ = ck.estimate(h, x, 0.0)
    assert abs(fr[0, 0] - 2.0010855143623831) < 1e-12
    assert abs(fr[0, 1] - (-0.0047041137141090039)) < 1e-12
    assert abs(fr[0, 2] - 51.018312664516209) < 1e-11
    assert abs(fr[0, 3] - 50.915633225810453) < 1e-9
    ck.close(h)
    h = ck.open(cfg)
    x = signals.steady(cfg, 0, [2.0], [50.0], [1.0])
    fr, _ = ck.estimate(h, x, 0.01)
    assert abs(fr[0, 0] - 2.0) < 1e-12 and fr[0, 2] == 50.0
    assert abs(fr[0, 1] - (-2.141592653589793)) < 1e-12
    ck.close(h)


@pytest.mark.parametrize("bad", [
    dict(f0=55), dict(n_cycles=0), dict(n_cycles=51), dict(fs=25601), dict(frame_rate=0),
    dict(n_bins=5), dict(n_bins=1025), dict(P=0), dict(interf_trig=0.0), dict(interf_trig=1.5),
    dict(rocof_thresh=[0.0, 25.0, 0.035]), dict(rocof_thresh=[3.0, 25.0, -1.0]),
])
def test_port_rejects_invalid_config_like_reference(bad):
    """check_config_validity, reference src/pmu_estimator.c:386-449."""
    cfg = dict(configs.DEFAULT_INI, **bad)
    port = CpuChecker("port", np.float64, 1)
    assert not port.open(cfg)
    if oracle_lib.have("ref"):
        assert not CpuChecker("ref", np.float64, 1).open(cfg)


@pytest.mark.skipif(not oracle_lib.have("ref"), reason="oracle/_ref not built on this host")
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("cfgname", ["config.ini", "m_class_config.ini", "cfg5_600", "p_class_config.ini"])
def test_port_matches_reference_build_random_streams(cfgname, dtype):
    """Seeded random three-phase streams, 6-channel reference build vs the port."""
    cfg = configs.NAMED[cfgname]
    dt = DT[dtype]
    n_inst, T = 5, 5
    prm = signals.three_phase_params(n_inst, seed=7)
    rng = np.random.default_rng(11)
    win = np.zeros((T, n_inst, 6, configs.win_len(cfg)))
    mid = np.zeros((T, n_inst))
    for t in range(T):
        x = signals.steady(cfg, t, prm["amp"], prm["freq"], prm["phase"], ramp=prm["ramp"],
                           harmonics=[(0.05, 150.3, 0.2)])
        x = x + 1e-3 * prm["amp"][:, None] * rng.standard_normal(x.shape)
        win[t] = x.reshape(n_inst, 6, -1)
        mid[t] = signals.mid_fracsec(cfg, t)
    ref = CpuChecker("ref", dt, 6)
    port = CpuChecker("port", dt, 6)
    fr_r, km_r, _ = ref.run(cfg, win, mid, n_threads=1)
    fr_p, km_p, _ = port.run(cfg, win, mid, n_threads=2)
    assert_frames_close(fr_p, fr_r, "tight" if dtype == "f64" else "f32", cfgname)
    assert np.array_equal(km_r[..., 5], km_p[..., 5])      # reference exposes the last channel only
