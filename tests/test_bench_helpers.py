"""bench.py helpers that need no GPU: the nvidia-smi sampler (started before the warm-up, armed at the first timed
step) against a stub `nvidia-smi`, and the CPU arm's bookkeeping."""
import importlib.util
import os
import stat
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


STUB = """#!/bin/bash
i=0
while true; do
  echo "0, $((1900 + i)), 1965, 300.5, 0x0000000000000000, Not Active, Not Active, Not Active, %s"
  i=$((i+1)); sleep 0.05
done
"""


def _install_stub(tmp_path, monkeypatch, power_cap="Not Active"):
    exe = tmp_path / "nvidia-smi"
    exe.write_text(STUB % power_cap)
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])


def test_clock_sampler_counts_from_the_armed_point(tmp_path, monkeypatch):
    _install_stub(tmp_path, monkeypatch)
    b = _bench()
    s = b.ClockSampler(0)
    s.start()
    time.sleep(0.45)                    # "warm-up": these samples are not reported
    s.arm()
    time.sleep(0.3)
    out = s.stop()
    assert 4 <= out["samples"] <= 9, out            # 0.3 s at 50 ms + the one just before arm()
    assert out["sm_mhz"] >= 1905 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == []


def test_clock_sampler_reports_power_cap_and_survives_a_missing_tool(tmp_path, monkeypatch):
    _install_stub(tmp_path, monkeypatch, power_cap="Active")
    b = _bench()
    s = b.ClockSampler(0)
    s.arm()                             # arm() alone starts the process
    time.sleep(0.25)
    assert s.stop()["reasons"] == ["sw_power_cap"]
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    s2 = b.ClockSampler(0)
    s2.start()
    s2.arm()
    assert s2.stop()["reasons"] == ["nvidia-smi unavailable"]
