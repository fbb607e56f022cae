"""Host-side pieces of bench.py that need no GPU: the clock-sample parser, the defaults the driver relies on, and the
reference arm's JSON line (the compiled reference on a small sample volume)."""
import importlib.util
import json
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_clock_summary_parses_samples_and_reasons():
    b = _bench()
    s = b.ClockSampler([0], enabled=False)
    t0 = time.time()
    ok = "0, 1965, 1965, 640.1, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active"
    cap = "0, 1650, 1965, 995.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active"
    s.rows = [(t0 + 0.1, ok), (t0 + 0.2, cap), (t0 + 0.3, ok), (t0 + 9.0, cap), (t0 + 0.25, "garbage")]
    out = s.summary([(t0, t0 + 1.0)])
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    assert s.summary([(t0 + 20, t0 + 21)])["samples"] == 0


def test_defaults_meet_the_timing_rules(monkeypatch):
    b = _bench()
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = b.parse()
    assert a.gpus == 1 and a.warmup >= 3 and a.steps >= 10 and a.size == 512 and a.impl == "b200"


def test_reference_arm_prints_the_contract_line():
    from oracle import ref as O
    if not O.have_ref() and not os.path.isdir("/root/reference/3DSIFT"):
        pytest.skip("compiled reference not available")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "48"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-1000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(line) == 1, "exactly one JSON line on stdout"
    d = json.loads(line[0])
    assert d["impl"] == "reference" and d["unit"] == "Mvoxels/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["n_gpus"] == 1
