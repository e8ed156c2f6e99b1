"""bench.py contract, the part that runs without a GPU: the reference arm (the reference's algorithm on the host CPU)
prints ONE JSON line with the keys the driver reads, on the same metric / unit / workload name as the CUDA arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_frame_updates_per_sec" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] >= 1 and d["value"] > 0
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - 1 * 8 * 256 * 6) < 1e-3 * 8 * 256 * 6     # value == sample units / step time
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "N=256" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.workload_name(1)          # the CUDA arm's workload name, verbatim
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_clock_merge_reports_the_slowest_gpu():
    sys.path.insert(0, ROOT)
    import bench
    a = {"sm_mhz": 1700.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 4, "source": "nvml"}
    b = {"sm_mhz": 1500.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3, "source": "nvml"}
    m = bench.merge_clocks([a, b])
    assert m["sm_mhz"] == 1500.0 and m["per_gpu_sm_mhz"] == [1700.0, 1500.0] and m["reasons"] == ["sw_power_cap"] and m["samples"] == 7
    assert bench.merge_clocks([a])["sm_mhz"] == 1700.0 and "per_gpu_sm_mhz" not in bench.merge_clocks([a])
    assert bench.merge_clocks([None])["sm_mhz"] is None
