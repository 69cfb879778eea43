#!/usr/bin/env python
"""
Turn an `ncu --set full` capture into the committed evidence bench.py reads:
    python tools/ncu_traffic.py <report.ncu-rep> <traffic key> <name under profiles/>
writes profiles/<name>_details.csv (ncu --page details), profiles/<name>_summary.txt (the metrics the roofline cites) and
adds {"dram_bytes_read", "dram_bytes_write", "duration_us", "source"} under <traffic key> in profiles/traffic.json
(bench.py: roofline.traffic).  Needs the ncu CLI, no GPU.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3,
         "second": 1e6, "s": 1e6, "nsecond": 1e-3, "ns": 1e-3}


def main():
    rep, key, name = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, vals = rows[0], rows[1], rows[-1]
    d = {h: (u, v) for h, u, v in zip(head, units, vals)}

    def get(metric):
        u, v = d[metric]
        return float(v.replace(",", "")) * UNITS.get(u, 1.0)
    rd, wr, dur = get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__time_duration.sum")
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(ROOT, "profiles", name + "_details.csv"), "w") as f:
        f.write(det)
    keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__cluster_dim_x", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]
    with open(os.path.join(ROOT, "profiles", name + "_summary.txt"), "w") as f:
        f.write(f"# {os.path.basename(rep)}: kernel {d.get('Kernel Name', ('', '?'))[1]}\n")
        for k in keep:
            if k in d:
                f.write(f"{k} = {d[k][1]} {d[k][0]}\n")
        for k in sorted(d):
            if "tensor" in k and k not in keep and k.endswith("pct_of_peak_sustained_active"):
                f.write(f"{k} = {d[k][1]} {d[k][0]}\n")
    p = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(p)) if os.path.exists(p) else {}
    t[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "duration_us": dur, "source": f"profiles/{name}_summary.txt"}
    with open(p, "w") as f:
        json.dump(t, f, indent=1, sort_keys=True)
    print(key, "read %.3f GB write %.3f GB in %.1f us" % (rd / 1e9, wr / 1e9, dur))


if __name__ == "__main__":
    main()
