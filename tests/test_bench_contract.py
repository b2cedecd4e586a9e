"""bench.py's output contract, checked without a GPU:
  * the product arm refuses to run without a CUDA device (no CPU fallback, non-zero exit, no JSON line);
  * the reference arm (`--impl reference`: the reference's own composed attention on its numba CPU backend, from the
    overlay tree when it is present, else the oracle port) prints ONE JSON line with the keys the driver reads;
  * the bench lines committed under profiles/ (measured on B200s by the builder) carry every key of the contract, with
    internally consistent values (frac = achieved / peak, e2e slower than the device-resident value, launches counted).
"""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _json_lines(text):
    return [json.loads(ln) for ln in text.splitlines() if ln.startswith("{")]


def test_product_arm_fails_loudly_without_a_gpu():
    import flashattn_b200 as fb
    lib = fb._lib.load("flashattention_kernel")
    if lib.fa_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not _json_lines(r.stdout)


@pytest.mark.timeout(900)
def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, NUMBA_DISABLE_CUDA="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, timeout=850, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference"
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*.json"))))
def test_committed_bench_lines_carry_the_contract(path):
    lines = [d for d in _json_lines(open(path).read()) if "metric" in d]
    assert len(lines) == 1, path
    d = lines[0]
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "attention fwd+bwd TFLOP/s" and d["unit"] == "TFLOP/s" and d["data"] == "synthetic"
    assert d["dtype"] == "bf16" and d["scaling"] == "weak" and d["warmup"] >= 3
    assert "workload" in d["config"]
    assert d["gpu_launches"] >= 2 * d["steps"]            # at least one forward and one backward kernel per step
    rf = d["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and rf["unit"] in ("GB/s", "TFLOP/s")
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-6
    assert 0.0 < rf["frac"] < 1.0
    assert rf["traffic"] is None or rf["traffic"] > 0
    ck = d["clocks"]
    assert ck["sm_mhz"] > 0 and ck["sm_max_mhz"] >= ck["sm_mhz"]
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(ck["reasons"])
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                    # host buffers + PCIe inside the timed region
    if d["n_gpus"] == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
        assert cb["value"] < e["value"]
    else:
        assert len(d["per_rank_ms"]["step"]) == d["n_gpus"]
        assert abs(max(d["per_rank_ms"]["step"]) - d["ms_per_step"]) < 1e-6      # max over ranks


@pytest.mark.timeout(900)
def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """The driver launches the reference arm like the product arm (torchrun, one process per GPU): rank 0 alone
    measures and prints, the other ranks exit 0 without work."""
    env = dict(os.environ, NUMBA_DISABLE_CUDA="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=850, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    assert lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2 and lines[0]["value"] > 0
