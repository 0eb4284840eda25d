"""Summarise an ncu report (one launch per kernel of interest) into profiles/*.json.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full_summary.json [profiles/traffic.json]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
NAMES = {"k_huffman": "k1_huffman", "k_hybrid": "k_hybrid", "k_synth": "k_synth"}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    kernels, traffic = {}, {}
    for r in rows[2:]:
        kname = r[col["Kernel Name"]]
        key = next((v for k, v in NAMES.items() if k in kname), None)
        if key is None or key in kernels:
            continue
        d, u = {}, {}
        for m in KEEP:
            if m in col and r[col[m]] != "":
                d[m] = float(r[col[m]].replace(",", ""))
                u[m] = units[col[m]]
        stalls = {}
        for n, i in col.items():
            if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and r[i] != "":
                v = float(r[i].replace(",", ""))
                if v >= 0.05:
                    stalls[n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
        d["units"] = u
        d["stalls_per_issue"] = stalls
        kernels[key] = d
        traffic[key] = (d["dram__bytes_read.sum"] * SCALE[u["dram__bytes_read.sum"]] +
                        d["dram__bytes_write.sum"] * SCALE[u["dram__bytes_write.sum"]])
    json.dump({"report": rep, "note": "one launch per kernel; per-launch values", "kernels": kernels}, open(out, "w"), indent=1)
    if len(sys.argv) > 3:
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)
    for k, d in kernels.items():
        print(k, d["gpu__time_duration.sum"], d["units"]["gpu__time_duration.sum"], "traffic", traffic[k])


if __name__ == "__main__":
    main()
