"""Summarise `ncu --set full` captures (gpurun_out/<tag>_full_*.ncu-rep) into profiles/<round>_ncu_full_summary.json and
profiles/ncu_traffic.json (DRAM bytes per launch of the dominant kernels, tagged with the digest of the kernel sources so
that bench.py only reports them for the library they were measured on).  Runs without a GPU (ncu -i).
usage: python tools/ncu_traffic.py <tag> [round]"""
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pips_b200 import _build  # noqa: E402

tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_MB",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
}
UNIT = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}
out = {"source_digest": _build.source_digest(), "source": f"ncu --set full --clock-control none, {tag}, one Pips.forward at BASELINE cfg2", "kernels": []}
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_full_*.ncu-rep"))):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"capture": os.path.basename(rep)}
        for h, u, v in zip(hdr, units, r):
            if h == "Kernel Name":
                d["kernel"] = v.split("(")[0]
            if h in WANT and v:
                try:
                    d[WANT[h]] = round(float(v.replace(",", "")) * UNIT.get(u, 1.0), 3)
                except ValueError:
                    pass
        out["kernels"].append(d)
json.dump(out, open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_full_summary.json"), "w"), indent=1)


def traffic(pred):
    ks = [k for k in out["kernels"] if pred(k)]
    return round(sum((k.get("dram_read_MB", 0) + k.get("dram_write_MB", 0)) * 1e6 for k in ks) / len(ks)) if ks else None


tr = {"source_digest": out["source_digest"], "source": f"profiles/{rnd}_ncu_full_summary.json ({out['source']})",
      "gemm_fc_bytes_per_launch": traffic(lambda k: "gemm_tc2" in k.get("kernel", "")),
      "corr_gather_bytes_per_launch": traffic(lambda k: "corr_gather" in k.get("kernel", "")),
      "tokenmix_bytes_per_launch": traffic(lambda k: "tokenmix" in k.get("kernel", ""))}
json.dump(tr, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
for k in out["kernels"]:
    print({a: k.get(a) for a in ("kernel", "duration_us", "dram_read_MB", "dram_write_MB", "tensor_pipe_pct", "issue_active_pct", "l2_hit_pct")})
print(tr)
