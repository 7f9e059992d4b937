"""bench.py contract (CPU side): the reference arm prints ONE JSON line with the agreed keys; the product arm refuses to
run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-nel", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "quad-pts/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["dtype"] == "f64" and d["vs_baseline"] is None


def test_product_arm_needs_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_clock_sampler_nvml_and_fallback(monkeypatch):
    """The clocks line: in-process NVML polling yields several samples inside a 70 ms region and decodes the event-reason
    bits; without an NVML binding and without nvidia-smi the sampler reports no samples instead of failing."""
    import time
    import types
    sys.path.insert(0, ROOT)
    import bench
    fake = types.ModuleType("pynvml")
    fake.NVML_CLOCK_SM = 1
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByIndex = lambda i: i
    fake.nvmlDeviceGetMaxClockInfo = lambda h, t: 1965
    fake.nvmlDeviceGetClockInfo = lambda h, t: 1950
    fake.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4 | 0x40
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    with bench.ClockSampler(0) as c:
        time.sleep(0.07)
    s = c.summary()
    assert s["samples"] >= 3 and s["sm_mhz"] == 1950.0 and s["sm_max_mhz"] == 1965.0
    assert s["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]

    def broken():
        raise RuntimeError("no NVML")
    fake.nvmlInit = broken
    monkeypatch.setenv("PATH", "/nonexistent")
    with bench.ClockSampler(0) as c:
        time.sleep(0.02)
    assert c.summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
