"""Reads .ncu-rep files (ncu --set full captures) and prints / stores the handful of numbers the bench line and
DESIGN.md quote: duration, DRAM bytes per launch, warp instructions, active lanes, issue utilisation, occupancy."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_inst",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "msecond": 1e-3, "usecond": 1e-6, "second": 1, "nsecond": 1e-9}


def summarize(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    res = {"kernel": vals[hdr.index("Kernel Name")][:80]}
    for k, name in KEYS.items():
        if k in hdr:
            i = hdr.index(k)
            v = float(vals[i].replace(",", ""))
            res[name] = v * UNIT.get(units[i], 1)
    if "dram_read" in res:
        res["dram_total"] = res["dram_read"] + res["dram_write"]
    return res


if __name__ == "__main__":
    allr = {}
    for p in sys.argv[1:]:
        allr[p] = summarize(p)
    print(json.dumps(allr, indent=1))
