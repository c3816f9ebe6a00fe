"""CPU-only tests of the host side: the C ABI loads and exports every declared symbol, struct
layouts match the reference's ABI, .ini parsing / validation behave like the reference
(golden from its own iniparser + check_config_validity), the launch plan is sane, and the
kernel's lane arithmetic — compiled for the CPU by tests/host_emul — matches the oracle."""
import ctypes as C
import json
import math
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import emul_lib
import oracle_lib
from oracle_lib import CpuChecker
from parity import assert_frames_close, golden_floats
from stream_cases import STREAM_CASES, stream_case_windows
from synchropmu_b200 import api, configs, signals

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


@pytest.fixture(scope="session", autouse=True)
def _built():
    api.build_library()
    oracle_lib.build_oracle()


def _declared(header: str):
    txt = (ROOT / "include" / header).read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pmub?_[a-z_]+)\s*\(", txt)))


@pytest.mark.parametrize("variant", [(np.float64, 1), (np.float64, 6), (np.float32, 1), (np.float32, 6)])
def test_library_loads_and_exports_every_declared_symbol(variant):
    lib = C.CDLL(str(api.lib_path(*variant)))
    names = _declared("pmu_estimator.h") + _declared("pmu_batch.h")
    assert {"pmu_init", "pmu_estimate", "pmu_deinit", "pmu_dump_frame", "pmu_batch_init", "pmu_estimate_batch",
            "pmu_batch_deinit", "pmub_create", "pmub_estimate", "pmub_estimate_async"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layouts_match_reference_abi(tmp_path):
    """sizeof/offsetof table of SURVEY.md appendix B (printed there from the real header)."""
    src = tmp_path / "abi.c"
    src.write_text('#include <stddef.h>\n#include "pmu_estimator.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(pmu_context),offsetof(pmu_context,pmu_initialized),sizeof(estimator_config),sizeof(pmu_frame),'
                   'offsetof(pmu_context,buff_params),offsetof(pmu_context,hann_params),offsetof(pmu_context,rocof_params));return 0;}')
    want = {("double", 1): (296, 288, 88, 32, 88, 128, 208), ("double", 6): (456, 448, 88, 32),
            ("float", 1): (192, 184, 60, 16), ("float", 6): (296, 288, 60, 16)}
    for (ft, c), exp in want.items():
        exe = tmp_path / f"abi_{ft}_{c}"
        subprocess.run(["gcc", f"-I{ROOT / 'include'}", f"-DNUM_CHANLS={c}", f"-DPMU_FLOAT_P={ft}", str(src), "-o", str(exe)], check=True)
        got = tuple(int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
        assert got[:len(exp)] == exp, (ft, c, got)


@pytest.mark.skipif(not (Path("/root/reference/src/pmu_estimator.h").exists()), reason="reference tree not mounted")
def test_struct_layouts_identical_to_reference_header(tmp_path):
    """Byte-for-byte: every member offset of our header == the reference header's."""
    members = ["synch_params.win_len", "synch_params.iter_eipdft_enabled", "synch_params.interf_trig", "synch_params.df",
               "synch_params.norm_factor", "synch_params.phasor", "buff_params.Xf", "buff_params.hann_window",
               "buff_params.signal_windows", "hann_params.C3", "hann_params.C5", "hann_params.C6",
               "hann_params.inv_norm_factor", "rocof_params.freq_old", "rocof_params.thresholds",
               "rocof_params.low_pass_coeff", "rocof_params.delay_line", "rocof_params.state", "pmu_initialized"]
    body = "".join(f'printf("%zu ", offsetof(pmu_context, {m}));' for m in members)
    cfgm = "".join(f'printf("%zu ", offsetof(estimator_config, {m}));' for m in
                   ["n_cycles", "Q", "iter_eipdft", "interf_trig", "rocof_thresh", "rocof_low_pass_coeffs"])
    src = tmp_path / "off.c"
    src.write_text(f'#include <stddef.h>\n#include "pmu_estimator.h"\nint main(){{{body}{cfgm}'
                   'printf("%zu %zu %zu\\n", sizeof(pmu_context), sizeof(estimator_config), sizeof(pmu_frame));return 0;}')
    for c in (1, 6):
        ours = tmp_path / f"ours{c}"
        subprocess.run(["gcc", f"-I{ROOT / 'include'}", f"-DNUM_CHANLS={c}", "-DPMU_FLOAT_P=float", str(src), "-o", str(ours)], check=True)
        ref = tmp_path / f"ref{c}"
        subprocess.run(["gcc", "-I/root/reference/src", f"-DNUM_CHANLS={c}", str(src), "-o", str(ref)], check=True)   # reference default: float
        a = subprocess.run([str(ours)], capture_output=True, text=True, check=True).stdout
        b = subprocess.run([str(ref)], capture_output=True, text=True, check=True).stdout
        assert a == b


@pytest.mark.skipif(not Path("/root/reference/examples/simple_example2.c").exists(), reason="reference tree not mounted")
def test_reference_c_callers_compile_and_link_unchanged(tmp_path):
    """Source compatibility: the reference's own drivers build against OUR header and link OUR
    library without edits (they are only run on the GPU box, via examples/ of this repo)."""
    for src, flags in (("examples/simple_example2.c", []), ("examples/simple_example.c", []),
                       ("test/test.c", ["-DPMU_FLOAT_P=float"]), ("test/test2.c", [])):
        lib = "pmu_estimator_f32" if flags else "pmu_estimator"
        exe = tmp_path / (Path(src).stem + ".out")
        r = subprocess.run(["gcc", "-O1", *flags, f"-I{ROOT / 'include'}", f"/root/reference/{src}", "-o", str(exe),
                            f"-L{api.LIB_DIR}", f"-l{lib}", "-lm", f"-Wl,-rpath,{api.LIB_DIR}"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_ini_parsing_matches_reference_golden():
    gold = json.loads((GOLD / "ini_configs.json").read_text())
    assert gold
    for rel, g in gold.items():
        cfg = api.parse_ini(str(ROOT / rel))
        accepted = cfg is not None and api.validate(cfg)
        assert accepted == g["accepted"], rel
        if accepted:
            assert configs.win_len(cfg) == g["win_len"] and cfg["n_bins"] == g["n_bins"]
    for name in ("config.ini", "p_class_config.ini", "m_class_config.ini"):
        assert api.parse_ini(str(ROOT / "config" / name)) == configs.NAMED[name]
    assert api.parse_ini(str(ROOT / "config" / "does_not_exist.ini")) is None


@pytest.mark.parametrize("bad", [
    dict(f0=55), dict(n_cycles=0), dict(n_cycles=51), dict(fs=25601), dict(frame_rate=0), dict(n_bins=5),
    dict(n_bins=1025), dict(P=0), dict(interf_trig=0.0), dict(interf_trig=1.5), dict(rocof_thresh=[0.0, 25.0, 0.035]),
])
def test_validation_rejects_what_the_reference_rejects(bad):
    assert api.validate(configs.DEFAULT_INI)
    assert not api.validate(dict(configs.DEFAULT_INI, **bad))
    assert "ERROR" in api.last_error()


REF_WRAPPER = Path("/root/reference/python/pmu_estimator.py")


@pytest.mark.skipif(not REF_WRAPPER.exists(), reason="the reference tree is not on this box")
def test_reference_python_wrapper_binds_our_library_unchanged():
    """The reference's OWN ctypes wrapper (python/pmu_estimator.py, imported from where it lies, unmodified) against
    libpmu_estimator.so: it must load, find the three symbols it declares argtypes for, its (oversized) context buffer
    must cover our pmu_context, and on this GPU-less box pmu_init must refuse with -1 and a message instead of crashing
    or silently estimating on the CPU; pmu_estimate on the unconfigured context returns None like the reference's
    "not initialized" path.  (On the GPU box the same structs are exercised by the emulation in
    tests/test_gpu_parity.py::test_reference_python_wrapper_layout_is_safe, since /root/reference does not travel.)"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_pmu_estimator", REF_WRAPPER)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    lib = api.lib_path(np.float64, 1)
    est = ref.PMUEstimator(lib_path=str(lib))
    for sym in ("pmu_init", "pmu_estimate", "pmu_deinit"):
        assert hasattr(est.lib, sym)
    # the wrapper's context mirror is sized for 2 channels of doubles: it must be at least as large as the real one
    from synchropmu_b200.pmu_estimator import _structs

    assert C.sizeof(ref.PmuContext) >= C.sizeof(_structs(C.c_double, 1)[1])
    d = configs.DEFAULT_INI
    cfg = ref.EstimatorConfig(d["n_cycles"], d["f0"], d["frame_rate"], d["fs"], d["n_bins"], d["P"], d["Q"], d["interf_trig"],
                              d["rocof_thresh"], d["rocof_low_pass_coeffs"], bool(d["iter_eipdft"]))
    torch = pytest.importorskip("torch")
    rc = est.configure_from_class(cfg)
    if torch.cuda.is_available():
        assert rc == 0
        t = np.arange(configs.win_len(d)) / d["fs"]
        got = est.estimate(list(2 * np.cos(2 * np.pi * 50 * t + 1)), 0.0)
        assert abs(got["amp"] - 2) < 1e-9 and abs(got["freq"] - 50) < 1e-7 and abs(got["ph"] - 1) < 1e-9
        assert est.deinit() == 0
    else:
        assert rc == -1                                   # no CUDA device: loud refusal, no CPU path
        assert est.estimate([0.0] * configs.win_len(d), 0.0) is None
        assert est.deinit() == -1                          # "not initialized", as in the reference (:280-284)


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: on a box without a CUDA device creation must raise, never estimate."""
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        api.BatchEstimator(configs.DEFAULT_INI, 2, 1)


# ---- the kernel's arithmetic on the CPU (tests/host_emul) vs the oracle ---------------------
@pytest.mark.parametrize("sc", STREAM_CASES, ids=[s["name"] for s in STREAM_CASES])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_kernel_arithmetic_matches_oracle_on_sweep(sc, dtype):
    dt = np.float64 if dtype == "f64" else np.float32
    cfg = configs.NAMED[sc["cfg"]]
    win, mid = stream_case_windows(sc, dtype=dt)
    want, km, _ = CpuChecker("port", dt, 1).run(cfg, win, mid)
    got, flags, _spec, _plan = emul_lib.run(cfg, win[:, :, 0, :], mid, dtype=dt)
    assert_frames_close(got, want[:, :, 0, :], dtype, sc["name"])
    assert np.array_equal(flags & 0xFFFF, km[:, :, 0])


def test_kernel_arithmetic_golden_single_calls_and_spectrum():
    meta = json.loads((GOLD / "single_calls.json").read_text())
    npz = np.load(GOLD / "single_calls.npz")
    for case in meta:
        dt = np.float64 if case["dtype"] == "f64" else np.float32
        w = npz[case["key"] + "__win"]
        got, flags, spec, _ = emul_lib.run(case["cfg"], w[None, :, :], np.full((1, 1), case["mid"]), dtype=dt,
                                           n_channels=case["C"])
        want = golden_floats(case["frames"]).reshape(case["C"], 4)
        assert_frames_close(got[0], want, case["dtype"], case["key"])
        want_spec = golden_floats(case["spectrum_re"]) + 1j * golden_floats(case["spectrum_im"])
        scale = max(float(np.max(np.abs(want_spec))), 1e-300)
        assert np.max(np.abs(spec[0, -1] - want_spec)) <= (1e-11 if dt == np.float64 else 2e-5) * scale
        if not np.isnan(want).any():
            assert int(flags[0, -1] & 0xFFFF) == case["km_last"]


@pytest.mark.parametrize("cfg,expect", [
    (configs.DEFAULT_INI, dict(M=16, KC=12, nslab=1)),
    (configs.M_CLASS_INI, dict(M=16, KC=12, nslab=2, W=96)),          # 192 residues: two full 96-wide slabs
    (configs.CFG5_600, dict(M=8, KC=10, nslab=1, wp=4)),             # short windows: four per warp pass
    (configs.DEFAULT_INI, dict(wp=1)),
    (configs.P_CLASS_INI, dict(M=16, KC=8, nslab=1, W=64, wp=2)),       # 1024 samples: two windows per warp pass
    (dict(configs.DEFAULT_INI, fs=4750, n_cycles=3, n_bins=7), dict(M=1, KC=8, nslab=1)),
    (dict(configs.DEFAULT_INI, fs=51200, n_cycles=16, n_bins=18), dict(M=16, KC=24, nslab=8)),
    (dict(configs.DEFAULT_INI, fs=51200, n_cycles=8, n_bins=18), dict(M=16, KC=24, nslab=4)),
    (dict(configs.DEFAULT_INI, f0=60, fs=4800, frame_rate=60, n_cycles=5, n_bins=9), dict(M=8, KC=10, nslab=1, wp=4)),
    # more than 23 bins: the general kernel's plan (M = 0, direct bins with the full twiddle table)
    (dict(configs.DEFAULT_INI, n_bins=24), dict(M=0, KC=25, nslab=1)),
    (dict(configs.DEFAULT_INI, n_cycles=30, n_bins=40, Q=2), dict(M=0, KC=41, nslab=1)),
    (dict(configs.DEFAULT_INI, fs=4750, n_cycles=9, n_bins=33), dict(M=0, KC=34, nslab=1)),
])
def test_plan_and_arithmetic_for_other_window_lengths(cfg, expect):
    N = configs.win_len(cfg)
    S, T = 5, 3
    rng = np.random.default_rng(N)
    win = np.zeros((T, S, 1, N))
    mid = np.zeros((T, S))
    f = cfg["f0"] + rng.uniform(-2, 2, S)
    for t in range(T):
        win[t, :, 0, :] = signals.steady(cfg, t, 1 + rng.uniform(0, 1, S), f, rng.uniform(-3, 3, S),
                                         harmonics=[(0.1, 1.5 * cfg["f0"], 0.0)])
        mid[t] = signals.mid_fracsec(cfg, t)
    want, km, _ = CpuChecker("port", np.float64, 1).run(cfg, win, mid)
    got, flags, _spec, plan = emul_lib.run(cfg, win[:, :, 0, :], mid)
    for k, v in expect.items():
        assert plan[k] == v, (plan, expect)
    assert_frames_close(got, want[:, :, 0, :], "f64", f"N={N}")
    assert np.array_equal(flags & 0xFFFF, km[:, :, 0])


# ---------------------------------------------------------------------------------------
# .ini parsing against the reference's own iniparser, on randomly mangled files
# ---------------------------------------------------------------------------------------
_INI_BASE = [("signal","n_cycles","4"),("signal","sample_rate","25600"),("signal","nominal_freq","50"),
        ("synchrophasor","frame_rate","50"),("synchrophasor","number_of_dft_bins","11"),("synchrophasor","ipdft_iterations","3"),
        ("synchrophasor","iter_e_ipdft_enable","1"),("synchrophasor","iter_e_ipdft_iterations","3"),("synchrophasor","interference_threshold","0.0033"),
        ("rocof","threshold_1","3"),("rocof","threshold_2","25"),("rocof","threshold_3","0.035"),
        ("rocof","low_pass_filter_1","0.5913"),("rocof","low_pass_filter_2","0.2043"),("rocof","low_pass_filter_3","0.2043")]

def _ini_variant(rng):
    vals = dict(((s,k),v) for s,k,v in _INI_BASE)
    # value mutations
    def setv(s,k,choices): vals[(s,k)] = rng.choice(choices)
    setv("signal","n_cycles",["4","2","6","04","0x4","4.0","4 ","+4"])
    setv("signal","sample_rate",["25600","12800","25600.0","0x6400","2.56e4","25600;c"])
    setv("signal","nominal_freq",["50","60","50.0"])
    setv("synchrophasor","frame_rate",["50","25","100","50.5"])
    setv("synchrophasor","number_of_dft_bins",["11","8","9","13","011"])
    setv("synchrophasor","ipdft_iterations",["3","1","2"])
    setv("synchrophasor","iter_e_ipdft_enable",["1","0","true","false","TRUE","yes","no","Y","n","T","F","2","on"])
    setv("synchrophasor","iter_e_ipdft_iterations",["3","1","5"])
    setv("synchrophasor","interference_threshold",["0.0033","3.3e-3",".0033","0.02","1","0.0033f"])
    setv("rocof","threshold_3",["0.035","3.5e-2","0.05"])
    lines=[]
    secs=["signal","synchrophasor","rocof"]
    if rng.random()<0.3: rng.shuffle(secs)
    for s in secs:
        sname = s.upper() if rng.random()<0.3 else (s.capitalize() if rng.random()<0.3 else s)
        lines.append(rng.choice(["[%s]","  [%s]","[%s]   ","[ %s ]","[%s] ; comment"]) % sname)
        keys=[(ss,k) for (ss,k) in vals if ss==s]
        if rng.random()<0.3: rng.shuffle(keys)
        for (ss,k) in keys:
            if rng.random()<0.03 and k != "nominal_freq": continue   # missing key (a missing nominal_freq makes the reference divide by zero)
            kn = k.upper() if rng.random()<0.2 else k
            v = vals[(ss,k)]
            if rng.random()<0.15: v = rng.choice(['"%s"',"'%s'"]) % v
            sep = rng.choice([" = ","=","  =  ","\t=\t"," =","= "])
            tail = rng.choice(["",""," ; note"," # note","   ",";x","#x"])
            lead = rng.choice([""," ","\t","    "])
            lines.append(f"{lead}{kn}{sep}{v}{tail}")
            if rng.random()<0.1: lines.append(rng.choice(["","; full comment","# another","   "]))
            if rng.random()<0.04: lines.append(f"{k} = {vals[(ss,k)]}")  # duplicate
    if rng.random()<0.05: lines.insert(rng.randrange(len(lines)), "this is not valid")
    return "\n".join(lines)+"\n"



@pytest.mark.parametrize("seed", [1, 2])
def test_ini_parser_agrees_with_the_reference_on_mangled_files(tmp_path, seed):
    """250 .ini files per seed with the quirks iniparser tolerates (case, spacing, quotes, inline comments, duplicate
    and missing keys, hexadecimal / octal / exponent numbers, every spelling of a boolean, stray lines): the unmodified
    reference (oracle/_ref) and pmub_parse_ini + pmub_validate must accept the same files, and for the accepted ones
    a probe estimate through the oracle port configured with OUR parse must equal, bit for bit, the reference's
    estimate configured from ITS parse - so every parameter was read identically."""
    import random
    from oracle_lib import have
    if not have("ref", np.float64, 1):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rng = random.Random(seed)
    ref, port = CpuChecker("ref", np.float64, 1), CpuChecker("port", np.float64, 1)
    accepted = 0
    for i in range(250):
        path = tmp_path / f"f{i}.ini"
        path.write_text(_ini_variant(rng))
        h = ref.open_ini(str(path))
        cfg = api.parse_ini(str(path))
        ours_ok = cfg is not None and api.validate(cfg)
        assert bool(h) == bool(ours_ok), (i, path.read_text(), cfg)
        if not h:
            continue
        accepted += 1
        n = ref.win_len(h)
        assert n == configs.win_len(cfg) and ref.n_bins(h) == cfg["n_bins"], (i, cfg)
        t = np.arange(n) / float(cfg["fs"])
        x = (2.0 * np.cos(2 * math.pi * 51.0 * t) + 0.2 * np.cos(2 * math.pi * 75.0 * t)).reshape(1, -1)
        want = [ref.estimate(h, x, m)[0] for m in (0.0, 0.02)]
        ref.close(h)
        hp = port.open(cfg)
        assert hp, (i, cfg)
        got = [port.estimate(hp, x, m)[0] for m in (0.0, 0.02)]
        port.close(hp)
        for g, w in zip(got, want):
            assert np.array_equal(g, w, equal_nan=True), (i, cfg, path.read_text())
    assert accepted > 40        # the rest were rejected by both, for the same files


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_config_validation_agrees_with_the_reference_on_random_structs(seed):
    """check_config_validity (:386-449) against pmub_validate on 400 random estimator_config structs per seed, most of
    them near a boundary of some rule (f0 = 0 is left out: the reference divides by it before validating)."""
    import random
    from oracle_lib import have
    if not have("ref", np.float64, 1):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rng = random.Random(seed)
    ref = CpuChecker("ref", np.float64, 1)
    n_ok = 0
    for i in range(400):
        f0 = rng.choice([50, 60])
        n_cycles = rng.choice([1, 2, 4, 6, 10, 50])
        fs = f0 * rng.choice([20, 64, 128, 256, 512])
        N = n_cycles * fs // f0
        cfg = dict(n_cycles=n_cycles, f0=f0, frame_rate=rng.choice([1, 25, 50, 60]), fs=fs,
                   n_bins=rng.choice([n_cycles + 2, n_cycles + 3, N // 2]), P=rng.choice([1, 3]), Q=rng.choice([0, 1, 3]),
                   interf_trig=rng.choice([1e-9, 0.0033, 1.0]), rocof_thresh=[3.0, 25.0, 0.035],
                   rocof_low_pass_coeffs=[rng.choice([0.5913, 0.0, -2.0]), 0.2043, rng.choice([0.2043, 1e9])],
                   iter_eipdft=rng.choice([0, 1]))
        for _ in range(rng.choice([0, 1, 1, 2])):      # push up to two fields onto / across a boundary of their rule
            k = rng.choice(["f0", "fs", "n_cycles", "n_bins", "frame_rate", "P", "interf_trig", "thr"])
            if k == "f0": cfg["f0"] = rng.choice([49, 55, 1, 61])
            elif k == "fs": cfg["fs"] = rng.choice([0, cfg["fs"] + 1, 4799, cfg["f0"]])
            elif k == "n_cycles": cfg["n_cycles"] = rng.choice([0, 51, 50, 1])
            elif k == "n_bins": cfg["n_bins"] = rng.choice([cfg["n_cycles"] + 1, N // 2 + 1, 0, N // 2])
            elif k == "frame_rate": cfg["frame_rate"] = 0
            elif k == "P": cfg["P"] = 0
            elif k == "interf_trig": cfg["interf_trig"] = rng.choice([-0.1, 0.0, 1.0001])
            else: cfg["rocof_thresh"][rng.randrange(3)] = rng.choice([0.0, -1.0])
        h = ref.open(cfg)
        ours = api.validate(cfg)
        assert bool(h) == bool(ours), (i, cfg)
        if h:
            n_ok += 1
            ref.close(h)
    assert 50 < n_ok < 350


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_valid_configurations_on_the_host_emulation(seed):
    """The kernel's own arithmetic headers compiled for the CPU (tests/host_emul), on randomly drawn valid
    configurations: every launch plan (16/8/4/1-point codelets, slabs, packed short windows, the general kernel's
    chunks) against the oracle, peak bin exact - the GPU-less twin of the randomized GPU test."""
    from random_configs import random_valid_configs
    plans = set()
    for cfg in random_valid_configs(10, seed):
        N = configs.win_len(cfg)
        S, T = 6, 3
        rng = np.random.default_rng(N + seed)
        f = cfg["f0"] + rng.uniform(-1.5, 1.5, S)
        win = np.zeros((T, S, 1, N))
        mid = np.zeros((T, S))
        for t in range(T):
            harm = [(0.1, 1.5 * cfg["f0"], 0.0)] if t == 1 else [(0.003, 3.0 * cfg["f0"], 0.5)]
            win[t, :, 0, :] = signals.steady(cfg, t, 1 + rng.uniform(0, 1, S), f, rng.uniform(-3, 3, S), harmonics=harm)
            mid[t] = signals.mid_fracsec(cfg, t)
        win += 1e-4 * rng.standard_normal(win.shape)
        want, km, _ = CpuChecker("port", np.float64, 1).run(cfg, win, mid)
        got, flags, _spec, plan = emul_lib.run(cfg, win[:, :, 0, :], mid)
        plans.add((plan["M"], plan["nslab"] > 1, plan["wp"]))
        assert_frames_close(got, want[:, :, 0, :], "f64", str(cfg))
        ok = ~np.isnan(want[:, :, 0, :]).any(axis=-1)
        assert np.array_equal((flags & 0xFFFF)[ok], km[:, :, 0][ok]), cfg
    assert len(plans) >= 3, plans


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_launch_plan_invariants_on_random_configurations(seed):
    """What the kernel silently relies on, for every plan make_plan() can produce (both float widths): the shared-memory
    layout fits and is ordered, every consumer warp's reduction scratch can also hold the estimator's W columns, raw-bin
    rows are 16-byte entries an odd number apart, packed windows are single-slab, the small-launch rows fit the raw
    region, a slab fits its stage."""
    import emul_lib
    from random_configs import random_valid_configs

    named = [configs.NAMED[k] for k in configs.NAMED]
    seen = set()
    for cfg in named + list(random_valid_configs(60, seed)):
        for dt in (np.float64, np.float32):
            p = emul_lib.plan(cfg, dt)
            assert p is not None, cfg
            s = np.dtype(dt).itemsize
            N, K = configs.win_len(cfg), cfg["n_bins"]
            assert p["smem_bytes"] <= 232448 and p["Kp"] == K + 1
            if p["M"] == 0:                                   # the general kernel has its own layout
                assert K + 1 > 24 or p["n_cons"] >= 1
                continue
            seen.add((p["M"], p["KC"], p["wp"], p["pk"], p["nslab"] > 1))
            assert p["M"] * p["L"] == N and p["KC"] >= p["Kp"] and p["KC"] in (8, 10, 12, 16, 24)
            assert p["n_stage"] >= 2 and 1 <= p["n_cons"] <= 7 and 2 <= p["n_raw"] <= 8
            assert p["off_raw"] % 128 == 0 and p["off_red"] % 128 == 0 and p["off_stage"] % 128 == 0
            assert p["off_raw"] + p["n_raw"] * 32 * p["raw_stride"] * 8 <= p["off_red"]
            assert p["off_red"] + p["n_cons"] * p["red_bytes"] <= p["off_stage"]
            assert p["off_stage"] + p["n_stage"] * p["stage_bytes"] == p["smem_bytes"]
            # X in place: rows of 16-byte complex entries, an odd number of entries apart, K + 1 entries used
            assert p["raw_stride"] % 2 == 0 and (p["raw_stride"] // 2) % 2 == 1 and p["raw_stride"] >= 2 * p["Kp"]
            # the lane-reduction scratch doubles as the estimator's W[K][32] complex columns
            elem = s
            assert p["red_bytes"] >= 32 * (2 * p["KC"] + 2) * elem and p["red_bytes"] >= K * 32 * 16
            # small launches use one raw row per consumer warp
            assert p["n_raw"] * 32 >= p["n_cons"]
            assert p["wp"] in (1, 2, 4) and p["pk"] == (2 if s == 4 else 1)
            if p["wp"] > 1:
                assert p["nslab"] == 1 and p["pitch"] == p["L"] and p["wp"] * N * s <= p["stage_bytes"]
            else:
                assert p["M"] * p["pitch"] * s <= p["stage_bytes"] and p["W"] <= p["pitch"]
            assert p["nslab"] == -(-p["L"] // p["W"])
    assert len(seen) >= 8, seen


def test_signal_record_spans_agree_with_frames():
    """A stretch of the record cut with ``span=`` is the same samples the per-frame generator yields: the streaming bench
    legs push hops of ONE continuous record (a seam between recycled blocks would put a phase jump into every window across
    it and trigger the interference loop there)."""
    cfg = configs.DEFAULT_INI
    N, hop = configs.win_len(cfg), configs.hop(cfg)
    prm = signals.three_phase_params(2)
    kw = dict(ramp=prm["ramp"], harmonics=[(0.1, 75.0, 0.0)])
    frames = [signals.steady(cfg, t, prm["amp"], prm["freq"], prm["phase"], **kw) for t in range(4)]
    hist = signals.steady(cfg, 0, prm["amp"], prm["freq"], prm["phase"], span=(0, N - hop), **kw)
    np.testing.assert_array_equal(hist, frames[0][:, :N - hop])
    for t in range(4):
        new = signals.steady(cfg, 0, prm["amp"], prm["freq"], prm["phase"], span=(N - hop + t * hop, hop), **kw)
        np.testing.assert_array_equal(new, frames[t][:, N - hop:])
    # ... and consecutive frames overlap exactly (no seam inside the record)
    np.testing.assert_array_equal(frames[1][:, :N - hop], frames[0][:, hop:])
