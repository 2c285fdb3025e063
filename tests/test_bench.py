"""bench.py's output contract: one JSON line with the metric, the roofline and cpu_baseline objects, the end-to-end arm and
the launch count; the reference arm runs on the CPU (rank 0 only under torchrun); without a GPU the product arm fails loudly."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
SMALL = ["--rays", "65536", "--cells", "48", "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.5"]
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(args, env=None):
    e = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, env=e)


def test_reference_arm_prints_the_contract_line(built):
    p = run_bench(["--impl", "reference"] + SMALL)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert BASE_KEYS <= set(rec) and rec["impl"] == "reference"
    assert rec["unit"] == "Mrays/s" and rec["higher_is_better"] is True and rec["value"] > 0
    assert rec["e2e"] == {"value": rec["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    base = rec["cpu_baseline"]
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] == rec["value"] and "sample" in base
    assert rec["gpu_launches"] == 0 and "workload" in rec["config"]
    # both arms print the same config object (the driver compares them)
    import bench
    args = bench.parse.__globals__["argparse"].Namespace(workload="s1m", rays=65536, cells=48)
    assert set(rec["config"]) == set(bench.config_dict(args, 1))


def test_reference_arm_other_ranks_exit_quietly(built):
    p = run_bench(["--impl", "reference", "--gpus", "2"] + SMALL, env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback(built):
    from libyafaray_b200 import rt
    if rt.device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = run_bench(SMALL)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)
    assert not any(l.startswith("{") for l in p.stdout.splitlines()), "a bench line was printed without a GPU"


@pytest.mark.gpu
def test_product_arm_line_on_the_gpu(built):
    p = run_bench(["--rays", str(1 << 20), "--cells", "128", "--steps", "3", "--warmup", "3", "--cpu-seconds", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert BASE_KEYS <= set(rec) and "impl" not in rec
    assert rec["n_gpus"] == 1 and rec["scaling"] == "weak" and rec["dtype"] == "f32" and rec["vs_baseline"] is None
    assert rec["gpu_launches"] == 4 * rec["steps"], "a setup pass and a queue-fed traversal launch per query per step"
    roof = rec["roofline"]
    assert "traceKernel<0,false,true,false>" in roof["kernel"] and "setupKernel<0>" in roof["kernel"]
    assert roof["bound"] in ("hbm", "issue") and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert roof["hbm"]["unit"] == "GB/s" and roof["hbm"]["peak"] > 1000 and abs(roof["hbm"]["frac"] - roof["hbm"]["achieved"] / roof["hbm"]["peak"]) < 1e-9
    assert roof["traffic"] is None, "a reduced-size run must not carry the full-size ncu traffic figure"
    par = rec["parity"]
    assert par["rays_compared"] > 0 and par["non_tie_mismatches"] == 0 and par["shadow_bool_mismatches"] == 0 and par["tuv_bit_identical_where_ids_equal"]
    assert rec["tshadow"]["mrays_per_gpu"] > 0 and rec["tshadow"]["mean_transparent_casters_on_lit_rays"] > 0
    e2e = rec["e2e"]
    assert e2e["h2d_bytes_per_step"] == 2 * (1 << 20) * 32 and e2e["d2h_bytes_per_step"] == (1 << 20) * 20 and 0 < e2e["value"] < rec["value"]
    assert rec["cpu_baseline"]["value"] > 0 and rec["cpu_baseline"]["kind"] in ("reference", "port")
    assert rec["clocks"]["sm_mhz"] and not set(rec["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_photon_gather_roofline_block():
    """tools/pm_bench.py: the roofline of the gather kernel is attached to the default workload only, names the tighter bound and is
    arithmetic on the committed counters and the live duration."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pm_bench", os.path.join(ROOT, "tools", "pm_bench.py"))
    pm_bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pm_bench)
    roof = pm_bench.gather_roofline(7.0, 1_000_000, 1_000_000, 100, "surfaces", 2.5e-4, {})
    assert roof["bound"] == "issue" and 0.4 < roof["frac"] < 0.7 and roof["issue"]["frac"] == roof["frac"]
    assert abs(roof["hbm"]["achieved"] - 1_000_000 * 820 / 7.0e-3 / 1e9) < 1e-6 and roof["traffic"] > 8 * 1_000_000 * 820
    assert pm_bench.gather_roofline(7.0, 1_000_000, 1_000_000, 8, "surfaces", 2.5e-4, {}) is None
    assert pm_bench.gather_roofline(7.0, 1_000_000, 1_000_000, 100, "surfaces", 2.5e-4, {"B200PM_KERNEL": "plain"}) is None
