"""profiles/traffic.json from `ncu --set full` raw pages: DRAM bytes (read + write) per launch of
every kernel bench.py reports a roofline for.   python tools/make_traffic.py RAW.csv [RAW2.csv ...]"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
meta = {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v, u = float(r[ix[name]].replace(",", "")), units[ix[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[ix["Kernel Name"]]).split("(")[0].split("<")[0]
        b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        # keep the last launch of each kernel (steady state)
        out[name] = int(b)
        meta[name] = {"source": os.path.basename(path), "duration_us": float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
                      * {"us": 1, "ms": 1e3, "ns": 1e-3}[units[ix["gpu__time_duration.sum"]]]}
out["_meta"] = meta
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
