"""profiles/traffic.json (DRAM bytes per launch per kernel class) from an `ncu --set full` report:
  python tools/traffic_json.py gpurun_out/r01t_full.ncu-rep profiles/traffic.json
bench.py copies the figure of its dominant kernel class into roofline.traffic."""
import csv
import json
import subprocess
import sys
from collections import defaultdict

CLASS = {"hydro_stage_vec": "hydro_stage", "hydro_stage_tiled": "hydro_stage", "hydro_stage_kernel": "hydro_stage",
         "mag_pair_vec": "magnetic_pair", "mag_iter_vec": "magnetic_iter", "mag_iter_tiled": "magnetic_iter",
         "mag_final_kernel": "magnetic_final"}
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                           "gpu__time_duration.sum"))
ur, uw = rows[1][ri], rows[1][wi]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = defaultdict(list)
for r in data:
    name = next((c for k, c in CLASS.items() if k in r[ki]), None)
    if name is None or float(r[ti].replace(",", "")) < 20.0:  # skip launches that exited at once (speculative slack)
        continue
    acc[name].append(float(r[ri].replace(",", "")) * scale[ur] + float(r[wi].replace(",", "")) * scale[uw])
res = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v), "source": sys.argv[1]}
       for k, v in acc.items()}
res["source"] = sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res, indent=1))
