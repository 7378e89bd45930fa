#!/usr/bin/env python
"""profiles/r2_k1f_traffic.json from an `ncu --set full` capture of ONE 2^20-state launch of the validity kernel:
dram__bytes_read.sum + dram__bytes_write.sum (what bench.py reports as roofline.traffic, with its source named).
usage: tools/ncu_traffic.py report.ncu-rep "<library tag printed by sqmp_version()>" """
import csv, io, json, os, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2]


def val(name):
    i = hdr.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return float(data[i]) * scale


out = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
       "dram_bytes_read": int(val("dram__bytes_read.sum")), "dram_bytes_write": int(val("dram__bytes_write.sum")),
       "kernel": data[hdr.index("Kernel Name")], "library": tag,
       "source": "profiles/%s (ncu --set full --clock-control none, one 2^20-state launch, config 2)" % os.path.basename(rep).replace(".ncu-rep", "_ncu_summary.txt")}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_k1f_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(out)
