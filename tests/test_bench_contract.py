"""The bench.py contract on the CPU: the reference arm runs without a GPU and prints one JSON line with the agreed keys;
the committed GPU bench line (profiles/r1_bench_C2.json) carries every key the contract names."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def last_json_line(text):
    lines = [l for l in text.splitlines() if l.startswith("{")]
    assert lines, text[-2000:]
    return json.loads(lines[-1])


def test_reference_arm_runs_on_the_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-points", "3000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    j = last_json_line(out.stdout)
    assert BASE_KEYS <= set(j) and j["impl"] == "reference" and j["metric"] == "voronoi_vertices_per_sec"
    assert j["value"] > 0 and j["unit"] == "vertices/s" and j["higher_is_better"] is True
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]
    assert "no julia binary" in j["cpu_baseline"]["note"]           # the probe ran and found none in this container


def test_row_checksum_is_order_independent_and_bit_sensitive():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import numpy as np
    rng = np.random.default_rng(0)
    sig = np.sort(rng.integers(1, 1000, size=(500, 4)), axis=1).astype(np.int64)
    r = rng.random((500, 3))
    a = bench.checksum(sig, r)
    perm = rng.permutation(500)
    assert bench.checksum(sig[perm], r[perm]) == a
    r2 = r.copy(); r2[17, 1] = np.nextafter(r2[17, 1], 2.0)
    assert bench.checksum(sig, r2) != a
    s2 = sig.copy(); s2[3, 0] += 1
    assert bench.checksum(s2, r) != a


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_gpu_line_carries_the_contract_keys():
    j = last_json_line(open(os.path.join(ROOT, "profiles", "r1_bench_C2.json")).read())
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches"} <= set(j)
    assert j["metric"] == "voronoi_vertices_per_sec" and j["n_gpus"] == 1 and j["warmup"] >= 3 and j["dtype"] == "f64"
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(j["e2e"]) and j["e2e"]["h2d_bytes_per_step"] > 0
    assert j["gpu_launches"] > 0 and j["clocks"]["sm_mhz"] and not j["clocks"]["reasons"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(j["cpu_baseline"])
    assert "workload" in j["config"] and j["vs_baseline"] is None


def test_gpus_must_match_the_launch():
    """an N-GPU line is never printed from a different number of ranks"""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "4", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode != 0 and "WORLD_SIZE=2" in out.stderr and out.stdout.strip() == ""


def run_mock(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mock_backend.py"), "--no-products"] + list(args),
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                  # ONE JSON line
    return json.loads(lines[0])


def test_own_arm_python_logic_against_a_mock_backend():
    """bench.py's step loop, workload bookkeeping and JSON line without a GPU: `torch` and the `hvb200` host API are stood in for
    by tests/mock_backend.py (the oracle behind the API's names); the numbers mean nothing (data = MOCK), the keys do"""
    j = run_mock("--steps", "2", "--warmup", "3", "--cpu-points", "1500")
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "e2e", "cpu_baseline", "workloads"} <= set(j)
    assert j["data"] == "MOCK" and j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] == 3 and j["vs_baseline"] is None
    assert j["config"]["workload"].startswith("C2:") and j["config"]["parallelism"] == "slab1" and "model" not in j["config"]
    r = j["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["bytes_per_vertex"] == 429.0
    assert j["e2e"]["h2d_bytes_per_step"] == 1500 * 3 * 8 and j["e2e"]["d2h_bytes_per_step"] > 0
    assert set(j["workloads"]) == {"C4", "C3", "D4"}
    pb = j["workloads"]["D4"]["published_by_the_reference"]            # docs/src/index.md:93: 841 395.0 vertices in 14.37 s
    assert abs(pb["vertices_per_s"] - 58557.6) < 0.1 and pb["value_over_published"] > 0 and pb["e2e_over_published"] > 0
    for w in j["workloads"].values():
        assert {"value", "e2e", "roofline", "config"} <= set(w)
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] > 0


def test_single_process_mode_python_logic_against_a_mock_backend():
    j = run_mock("--gpus", "2", "--single-process", "--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--extra", "C3")
    assert j["n_gpus"] == 2 and "hvb_create_multi" in j["config"]["parallelism"] and "(3000 total)" in j["config"]["workload"]
    assert j["scaling"] == "weak" and j["workloads"]["C3"]["scaling"] == "strong"        # C3 keeps its total (BASELINE configs[2])
    assert j["roofline"]["traffic"] is None
