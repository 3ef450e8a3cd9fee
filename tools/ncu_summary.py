"""Summarises an `ncu --set full` report as JSON (the numbers bench.py and DESIGN.md quote):
python tools/ncu_summary.py <report.ncu-rep> <shots> > profiles/<name>.json"""
import csv, io, json, subprocess, sys

rep, shots = sys.argv[1], int(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(txt)))
h, u, v = r[0], r[1], r[2]
col = {n: (v[i], u[i]) for i, n in enumerate(h)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}


def num(name):
    val, unit = col[name]
    return float(val) * scale.get(unit, 1)


out = {
    "report": rep, "kernel": col["Kernel Name"][0], "shots": shots,
    "grid": col["launch__grid_size"][0], "block": col["launch__block_size"][0],
    "registers_per_thread": col["launch__registers_per_thread"][0],
    "dynamic_smem_bytes": num("launch__shared_mem_per_block_dynamic"),
    "duration_ms_under_ncu": num("gpu__time_duration.sum"),
    "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
    "warp_instructions": float(col["smsp__inst_executed.sum"][0]),
    "issue_active_pct": float(col["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
    "sm_throughput_pct": float(col["sm__throughput.avg.pct_of_peak_sustained_elapsed"][0]),
    "dram_throughput_pct": float(col.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", col.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", ("nan", "")))[0]),
    "l1_lsu_wavefronts_shared": float(col["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0]),
    "warps_active_pct": float(col["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
}
out["dram_bytes_per_shot"] = (out["dram_bytes_read"] + out["dram_bytes_write"]) / shots
stalls = {n.split("issue_stalled_")[1]: float(val) for n, (val, _) in col.items()
          if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")}
tot = sum(stalls.values()) or 1
out["stall_share_pct"] = {k: round(100 * x / tot, 1) for k, x in sorted(stalls.items(), key=lambda kv: -kv[1]) if x / tot > 0.01}
print(json.dumps(out, indent=1))
