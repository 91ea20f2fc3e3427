"""Opcode mix (executed warp instructions) from an `ncu --page source --csv` dump."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, ei = hdr.index("Source"), hdr.index("Instructions Executed")
tot = defaultdict(int)
total = 0
for r in rows[2:]:
    if len(r) <= ei: continue
    try: n = int(r[ei])
    except ValueError: continue
    op = r[si].split()
    if not op: continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = o.split(".")[0]
    tot[o] += n; total += n
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{k:12s} {v:12d} {v/total:6.3f}" + (f"  {v*32/cells:7.1f}/cell" if cells else ""))
print("total", total, (total*32/cells if cells else ""))
