"""Host logic of bench.py that needs no GPU: the nvidia-smi sampler's timed-region window and a dry run of the whole
single-GPU bench control flow against stand-ins for the device pieces (the JSON contract of the line it prints)."""
import importlib.util
import json
import os
import stat
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FAKE_SMI = """#!/bin/bash
sleep 0.2
i=0
while true; do echo "$((1600+i)), 1965, 900.0, Not Active, Not Active, Not Active, Active"; i=$((i+1)); sleep 0.05; done
"""


def _fake_smi_dir(tmp_path):
    d = os.path.join(tmp_path, "bin")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, "nvidia-smi")
    with open(p, "w") as fh:
        fh.write(FAKE_SMI)
    os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    return d


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_reports_only_the_timed_region(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", _fake_smi_dir(tmp_path) + os.pathsep + os.environ.get("PATH", ""))
    b = _load_bench()
    s = b.ClockSampler(0)
    s.start()
    s.wait_ready()
    assert s.lines, "the sampler must be producing before the timed region starts"
    time.sleep(0.2)
    s.mark_begin()
    time.sleep(0.3)
    s.mark_end()
    time.sleep(0.15)
    first, last = s.begin, s.end
    out = s.stop()
    assert 0 < first < last <= len(s.lines)
    assert out["samples"] == last - first and out["reasons"] == ["sw_power_cap"] and out["sm_max_mhz"] == 1965.0
    assert 1600 + first <= out["sm_mhz"] <= 1600 + last            # the median of the window, not of the whole run
    # no nvidia-smi at all: reported, not raised
    monkeypatch.setenv("PATH", os.path.join(tmp_path, "nowhere"))
    s = b.ClockSampler(0)
    s.start(); s.wait_ready(0.1); s.mark_begin(); s.mark_end()
    assert s.stop()["sm_mhz"] is None


DRY_RUN = r"""
import sys, ctypes
sys.path.insert(0, %r)
import torch
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
class FakeEvent:
    def __init__(self, enable_timing=False): pass
    def record(self): pass
    def elapsed_time(self, other): return 130.0
torch.cuda.Event = FakeEvent
import nlpscratch_b200
from nlpscratch_b200 import _lib, array
class FakeArr:
    ptr = 4096; bounds = None
    def __init__(self, shape=(1,)): self.shape = shape
array.DeviceArray.from_host = staticmethod(lambda a, *k: FakeArr(a.shape))
array.DeviceArray.empty = staticmethod(lambda shape, dt: FakeArr(shape))
_lib.call = lambda *a, **k: 0
class FakeModel:
    def __init__(self, *a, **k): self.n = 0
    def train_step(self, *a, **k): self.n += 1; return 9.0
    def launch_count(self, reset=False): return 174 * self.n
    def close(self): pass
    def synchronize(self): pass
nlpscratch_b200.Transformer = FakeModel
class FakeSlot:
    def __init__(self, m): self.v = None
    def fetch(self, l): self.v = float(l)
    def value(self): return self.v
nlpscratch_b200.HostScalarSlot = FakeSlot
ctypes.memmove = lambda *a: None
import importlib.util
spec = importlib.util.spec_from_file_location("bench", %r); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
sys.argv = ["bench.py", "--steps", "4", "--warmup", "3", "--no-cpu-baseline", "--no-other"]
b.main()
"""


def test_bench_line_contract_dry_run(tmp_path):
    """Everything device-side replaced by stand-ins: the control flow of the single-GPU arm runs to its JSON line and the
    line carries every key of the contract."""
    env = dict(os.environ, PATH=_fake_smi_dir(tmp_path) + os.pathsep + os.environ.get("PATH", ""))
    res = subprocess.run([sys.executable, "-c", DRY_RUN % (ROOT, os.path.join(ROOT, "bench.py"))], capture_output=True,
                         text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stderr[-3000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["steps"] == 4 and line["warmup"] == 3 and line["n_gpus"] == 1 and line["scaling"] == "weak"
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["clocks"]["reasons"] == ["sw_power_cap"]
    assert "FALLBACK" in line["roofline"]["peak_note"]          # the stand-in library measured nothing: said, not hidden


def test_reference_arm_runs_the_real_reference_with_all_host_threads(tmp_path):
    """`bench.py --impl reference` under torchrun's OMP_NUM_THREADS=1: the arm must still use every host core, must run the
    unmodified reference when it is installed, and must state the batch it really ran (VERDICT r01 bench hygiene)."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stderr[-3000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["config"]["batch_per_gpu"] == 4 and line["config"]["sample_of_batch_per_gpu"] == 4
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from install_reference import reference_path
    assert line["cpu_baseline"]["kind"] == ("reference" if reference_path() else "port")
