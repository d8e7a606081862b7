"""CPU suite: the parts of bench.py's contract that do not need a GPU -- the reference arm's JSON line and the refusal of the
product arm to run (or fall back) without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--genes", "400", "--samples", "24", "--cpu-sample-genes", "200")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "genes/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("genes/sec") and line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "genes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["scaling"] == "weak" and line["dtype"] == "f64"


def test_product_arm_fails_loudly_without_a_gpu():
    import ctypes

    try:  # a visible CUDA device means this is the GPU box: nothing to check here
        if ctypes.CDLL("libcuda.so.1").cuInit(0) == 0:
            n = ctypes.c_int(0)
            ctypes.CDLL("libcuda.so.1").cuDeviceGetCount(ctypes.byref(n))
            if n.value > 0:
                return
    except OSError:
        pass
    r = _run("--steps", "1", "--warmup", "1", "--genes", "200", "--samples", "12", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and not r.stdout.strip().startswith("{")
