"""Top source lines of an ncu report by stall samples / executed instructions (development aid).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_lines.py x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = dict(zip(hdr[4:], r[4:]))
        lines.append((cur, int(r[0]), r[1].strip(), d))
tot_s = sum(int(l[3]["# Samples"]) for l in lines) or 1
tot_i = sum(int(l[3]["Instructions Executed"]) for l in lines) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
for key in ("# Samples", "Instructions Executed"):
    print("---- by", key)
    for f, no, src, d in sorted(lines, key=lambda l: -int(l[3][key]))[:top]:
        s, i = int(d["# Samples"]), int(d["Instructions Executed"])
        thr = float(d.get("Avg. Threads Executed", 0) or 0)
        stalls = sorted(((int(v), k) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()), reverse=True)[:3]
        print("%5.1f%% smp %5.1f%% inst thr %4.1f  %s:%d  %s   [%s]" % (100.0 * s / tot_s, 100.0 * i / tot_i, thr, f, no, src[:90],
                                                                 " ".join("%s=%d" % (k[6:], v) for v, k in stalls if v)))
