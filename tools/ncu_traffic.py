#!/usr/bin/env python
"""DRAM bytes per kernel launch from an `ncu --set full` report -> profiles/r02_traffic.json (what bench.py reports as
roofline.traffic). Run here, no GPU: python tools/ncu_traffic.py gpurun_out/<tag>/frame.ncu-rep <commit>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, commit = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
i_name, i_rd, i_wr, i_t = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                  "gpu__time_duration.sum"))
kernels = {}
for r in rows[2:]:
    name = r[i_name].replace("void ", "").replace("pfcu::", "").replace("(bool)", "").replace("(int)", "")
    name = name[:name.index("(")] if "(" in name else name  # k_composite<1>, k_scan<0>, k_fill, ...
    kernels[name] = {"dram_bytes_read": int(float(r[i_rd]) * scale[units[i_rd]]),
                     "dram_bytes_write": int(float(r[i_wr]) * scale[units[i_wr]]),
                     "duration_us": float(r[i_t]) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[i_t], 1)}
doc = {"source": "ncu --set full --clock-control none, one warm tiger.svg@4096^2 frame, %s (%s)" % (os.path.basename(rep), commit),
       "note": "no L2 flush under ncu: most of the 64 MiB framebuffer's write-back falls outside the kernel's window, so the "
               "tile kernel's DRAM bytes are BELOW its algorithmic bytes",
       "kernels": kernels}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "profiles", "r02_traffic.json"), "w") as fp:
    json.dump(doc, fp, indent=1, sort_keys=True)
print(json.dumps(kernels, indent=1))
